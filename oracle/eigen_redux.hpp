// TEST INFRASTRUCTURE ONLY -- part of the CPU oracle (see oracle/README.md).
//
// Restatement of the summation ORDER of Eigen 3.3.7's fixed-size block reductions
// (Eigen/src/Core/Redux.h: redux_impl<Func, Derived, DefaultTraversal, CompleteUnrolling>
// -> redux_novec_unroller<Func, Derived, Start, Length>).  Eigen is NOT vendored under
// /root/reference (CMakeLists.txt:41 find_package(Eigen3), version unpinned; Ubuntu
// 20.04 / ROS Noetic ships 3.3.7).  Call sites in the reference:
//   GroundSegmentation.cpp:268-269  ggp.block<3,3>(..).sum()
//   GroundSegmentation.cpp:359      pointsBlock.sum()                       (S = 3 | 5)
//   GroundSegmentation.cpp:374-375  (pointsBlock o varblock).sum(), (pointsBlock o minblock).sum()
//   GroundSegmentation.cpp:457-458  gvlblock.sum(), (gvlblock o gglblock).sum()
//
// The unroller is a binary split:  redux(start, len) = len == 1 ? coeff(start)
//     : redux(start, len/2) + redux(start + len/2, len - len/2)
// over the block's coefficients in column-major order of the BLOCK
// (coeff k = block(k % S, k / S), i.e. row index fastest).
// Restated from the published Eigen 3.3.7 sources (Eigen >= 3.4 may choose a different traversal for the 5x5 case,
// SURVEY.md App. A.0); oracle/ref_shim/Eigen/Core, which the unmodified reference is compiled against, states the same order.
#pragma once

namespace ggo {

template <int Start, int Len>
struct TreeSum {
    template <typename F>
    static inline float run(const F& coeff) {
        return TreeSum<Start, Len / 2>::run(coeff) + TreeSum<Start + Len / 2, Len - Len / 2>::run(coeff);
    }
};
template <int Start>
struct TreeSum<Start, 1> {
    template <typename F>
    static inline float run(const F& coeff) {
        return coeff(Start);
    }
};

// Sum of an S x S block whose top-left element is (r0, c0) of a column-major n x n matrix.
template <int S>
inline float block_sum(const float* m, int n, int r0, int c0) {
    return TreeSum<0, S * S>::run([&](int k) { return m[(r0 + k % S) + (c0 + k / S) * n]; });
}

// Sum of the coefficient-wise product of two co-located S x S blocks.
template <int S>
inline float block_dot(const float* a, const float* b, int n, int r0, int c0) {
    return TreeSum<0, S * S>::run([&](int k) {
        const int idx = (r0 + k % S) + (c0 + k / S) * n;
        return a[idx] * b[idx];
    });
}

template <int S>
inline float block_min(const float* m, int n, int r0, int c0) {
    float best = m[r0 + c0 * n];
    for (int c = 0; c < S; ++c)
        for (int r = 0; r < S; ++r) {
            const float v = m[(r0 + r) + (c0 + c) * n];
            if (v < best) best = v;
        }
    return best;
}

}  // namespace ggo

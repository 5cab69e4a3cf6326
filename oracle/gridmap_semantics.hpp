// TEST INFRASTRUCTURE ONLY -- part of the CPU oracle (see oracle/README.md).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may use anything under oracle/.  The product path never includes this file.
//
// Restatement of the grid_map_core geometry arithmetic the GroundGrid hot path relies
// on.  grid_map_core (ANYbotics/grid_map, ROS Noetic 1.6.x; version UNPINNED in the
// reference's package.xml:29) is NOT vendored under /root/reference and is not
// installed here, so its published algorithm (GridMapMath.cpp / GridMap.cpp) is
// restated below and anchored on the reference's call sites:
//   setGeometry   <- src/GroundGrid.cpp:58
//   move          <- src/GroundGrid.cpp:67,97
//   getPosition   <- src/GroundGrid.cpp:125
//   getIndex      <- src/GroundSegmentation.cpp:228,261
//   isInside      <- src/GroundSegmentation.cpp:230
// tests/test_oracle_vs_ref.py checks these against the CPU GridMap the unmodified reference is compiled with
// (oracle/ref_shim/grid_map_core/GridMap.hpp, restated from the same published sources).
#pragma once
#include <cmath>
#include <cstdint>

namespace ggo {

struct MapGeometry {
    int n = 0;             // cells per side (size = round(length / resolution))
    double res = 0.0;      // resolution widened from float (0.33f -> 0.33000001311302185)
    double len = 0.0;      // n * res (fp64)
    double px = 0.0, py = 0.0;  // map centre position
};

// grid_map::GridMap::setGeometry(length, resolution, position)
//   size = round(length / resolution) per axis ; length = size * resolution ; start index 0.
inline void set_geometry(MapGeometry& g, double length_m, double resolution, double px, double py) {
    g.n = static_cast<int>(std::round(length_m / resolution));
    g.res = resolution;
    g.len = static_cast<double>(g.n) * resolution;
    g.px = px;
    g.py = py;
}

// getIndexFromPosition(): v = (p - 0.5*length - mapPosition) / resolution, each axis;
// index = -(int)v  (cast<int>() truncates toward zero, then the -Identity
// map-frame -> buffer-order transform).  Start index is (0,0) on this path
// (GroundGrid.cpp:143 normalises it), so no wrap is applied.  The index is returned even
// when the position is outside the map (the ray-march at GroundSegmentation.cpp:261-265
// relies on that and range-checks it itself).
inline void get_index(const MapGeometry& g, double x, double y, int& ix, int& iy) {
    const double half = 0.5 * g.len;
    const double vx = (x - half - g.px) / g.res;
    const double vy = (y - half - g.py) / g.res;
    // cast<int>() of an out-of-range / NaN double is UB in C++; clamp so the oracle is
    // defined everywhere (such indices are rejected by every caller anyway).
    auto to_int = [](double v) -> int {
        if (!(v == v)) return 1000000000;      // NaN -> far outside
        if (v > 1.0e9) return 1000000000;
        if (v < -1.0e9) return -1000000000;
        return static_cast<int>(v);            // truncation toward zero
    };
    ix = -to_int(vx);
    iy = -to_int(vy);
}

// checkIfPositionWithinMap(): t = -(p - mapPosition - 0.5*length); inside iff
// 0 <= t < length on both axes (NaN => false).  Note the subtraction order differs from
// get_index (position - mapPosition - offset here, position - offset - mapPosition there).
inline bool is_inside(const MapGeometry& g, double x, double y) {
    const double half = 0.5 * g.len;
    const double tx = -(x - g.px - half);
    const double ty = -(y - g.py - half);
    return tx >= 0.0 && ty >= 0.0 && tx < g.len && ty < g.len;
}

// getPositionFromIndex() with start index (0,0):
//   position = mapPosition + (0.5*length - 0.5*resolution) + resolution * (-(double)index)
inline void get_position(const MapGeometry& g, int ix, int iy, double& x, double& y) {
    const double off = 0.5 * g.len - 0.5 * g.res;
    x = (g.px + off) + g.res * static_cast<double>(-ix);
    y = (g.py + off) + g.res * static_cast<double>(-iy);
}

// GridMap::move(): getIndexShiftFromPositionShift rounds half away from zero,
//   cells[k] = (int)(d[k]/res + 0.5*sign(d[k])), index shift = -cells ;
// the map position advances by the ALIGNED shift cells*res.  Returns the index shift
// (buffer-order, i.e. new(r) = old(r + shift) after convertToDefaultStartIndex()).
inline void move_shift(MapGeometry& g, double nx, double ny, int& shift_i, int& shift_j) {
    const double dx = nx - g.px, dy = ny - g.py;
    const double tx = dx / g.res, ty = dy / g.res;
    const int cx = static_cast<int>(tx + 0.5 * (tx > 0 ? 1 : -1));
    const int cy = static_cast<int>(ty + 0.5 * (ty > 0 ? 1 : -1));
    shift_i = -cx;
    shift_j = -cy;
    // getPositionShiftFromIndexShift: (-Identity * indexShift).cast<double>() * resolution
    g.px += static_cast<double>(cx) * g.res;
    g.py += static_cast<double>(cy) * g.res;
}

}  // namespace ggo

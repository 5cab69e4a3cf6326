// TEST INFRASTRUCTURE ONLY (oracle/_ref build): a CPU grid_map::GridMap implementing the subset of
// grid_map_core 1.6.x (ANYbotics/grid_map, the ROS Noetic release; the reference's package.xml:29 does
// not pin a version) that the UNMODIFIED reference sources call.  grid_map is not installed in this
// image and not vendored under /root/reference, so the arithmetic of GridMapMath.cpp / GridMap.cpp is
// restated here from the published 1.6.x sources, function by function, keeping their operation order:
//
//   setGeometry()                 <- GroundGrid.cpp:58      size = round(length / resolution), length = size * resolution
//   move(position, newRegions)    <- GroundGrid.cpp:67,97   getIndexShiftFromPositionShift / getPositionShiftFromIndexShift,
//                                                           dropped rows/cols set to NaN in every layer, circular start index
//   getPosition(index, position)  <- GroundGrid.cpp:125     getPositionFromIndex
//   SubmapIterator(map, region)   <- GroundGrid.cpp:122
//   at(layer, index)              <- GroundGrid.cpp:130-131
//   convertToDefaultStartIndex()  <- GroundGrid.cpp:143
//   add(layer, value)             <- GroundSegmentation.cpp:61-67,75   (overwrites an existing layer in place)
//   operator[](layer)             <- GroundSegmentation.cpp:70-78,203-213,...
//   getIndex(position, index)     <- GroundSegmentation.cpp:228,261    getIndexFromPosition
//   isInside(position)            <- GroundSegmentation.cpp:230        checkIfPositionWithinMap
//   getSize / getResolution / getLength
//
// Layer matrices live in std::unordered_map nodes, so the function-local static references the
// reference binds on first use (GroundSegmentation.cpp:76-78, 203-213, ...) stay valid, like with the
// real library.  Never included by the product (groundgrid_b200/).
#pragma once
#include <Eigen/Core>

#include <cmath>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace grid_map {

typedef Eigen::MatrixXf Matrix;
typedef Matrix::Scalar DataType;
typedef Eigen::Vector2d Position;
typedef Eigen::Vector2d Vector;
typedef Eigen::Array2i Index;
typedef Eigen::Array2i Size;
typedef Eigen::Array2d Length;

class BufferRegion {
  public:
    enum class Quadrant { Undefined, TopLeft, TopRight, BottomLeft, BottomRight };
    BufferRegion() : quadrant_(Quadrant::Undefined) {}
    BufferRegion(const Index& startIndex, const Size& size, const Quadrant& quadrant) : startIndex_(startIndex), size_(size), quadrant_(quadrant) {}
    const Index& getStartIndex() const { return startIndex_; }
    const Size& getSize() const { return size_; }
    Quadrant getQuadrant() const { return quadrant_; }

  private:
    Index startIndex_;
    Size size_;
    Quadrant quadrant_;
};

// ---- GridMapMath.cpp (1.6.x) -------------------------------------------------------------------
namespace math {
// wrapIndexToRange(int& index, int bufferSize)
inline void wrapIndexToRange(int& index, int bufferSize) {
    if (index < bufferSize) {
        if (index >= 0) return;
        if (index >= -bufferSize) { index += bufferSize; return; }
        index = index % bufferSize;
        index += bufferSize;
    } else if (index < bufferSize * 2) {
        index -= bufferSize;
    } else {
        index = index % bufferSize;
    }
}
// getVectorToOrigin: 0.5 * mapLength (map frame -> buffer order is -Identity, applied by the callers)
inline void getVectorToOrigin(Vector& v, const Length& mapLength) { v(0) = 0.5 * mapLength(0); v(1) = 0.5 * mapLength(1); }
inline bool checkIfIndexInRange(const Index& index, const Size& bufferSize) {
    return index(0) >= 0 && index(1) >= 0 && index(0) < bufferSize(0) && index(1) < bufferSize(1);
}
// checkIfPositionWithinMap: positionTransformed = -Identity * (position - mapPosition - offset)
inline bool checkIfPositionWithinMap(const Position& position, const Length& mapLength, const Position& mapPosition) {
    Vector offset;
    getVectorToOrigin(offset, mapLength);
    const double tx = -1.0 * ((position(0) - mapPosition(0)) - offset(0)) + 0.0;  // row of -Identity: (-1)*a + (-0)*b
    const double ty = -1.0 * ((position(1) - mapPosition(1)) - offset(1)) + 0.0;
    return tx >= 0.0 && ty >= 0.0 && tx < mapLength(0) && ty < mapLength(1);
}
// getBufferIndexFromIndex / getIndexFromBufferIndex
inline Index getBufferIndexFromIndex(const Index& index, const Size& bufferSize, const Index& bufferStartIndex) {
    if (bufferStartIndex(0) == 0 && bufferStartIndex(1) == 0) return index;
    Index b(index(0) + bufferStartIndex(0), index(1) + bufferStartIndex(1));
    wrapIndexToRange(b(0), bufferSize(0));
    wrapIndexToRange(b(1), bufferSize(1));
    return b;
}
inline Index getIndexFromBufferIndex(const Index& bufferIndex, const Size& bufferSize, const Index& bufferStartIndex) {
    if (bufferStartIndex(0) == 0 && bufferStartIndex(1) == 0) return bufferIndex;
    Index i(bufferIndex(0) - bufferStartIndex(0), bufferIndex(1) - bufferStartIndex(1));
    wrapIndexToRange(i(0), bufferSize(0));
    wrapIndexToRange(i(1), bufferSize(1));
    return i;
}
// cast<int>() of (-Identity * v): truncation toward zero of the negated value.  A NaN / out-of-range
// double -> int conversion is undefined in C++ (x86 cvttsd2si yields INT_MIN); such indices only occur
// for positions outside the map, which every caller rejects through isInside() / its own range test.
inline int castToInt(double v) { return static_cast<int>(v); }
// getIndexFromPosition
inline bool getIndexFromPosition(Index& index, const Position& position, const Length& mapLength, const Position& mapPosition,
                                 const double& resolution, const Size& bufferSize, const Index& bufferStartIndex) {
    Vector offset;
    getVectorToOrigin(offset, mapLength);
    const double ivx = ((position(0) - offset(0)) - mapPosition(0)) / resolution;
    const double ivy = ((position(1) - offset(1)) - mapPosition(1)) / resolution;
    Index raw(castToInt(-ivx), castToInt(-ivy));
    index = getBufferIndexFromIndex(raw, bufferSize, bufferStartIndex);
    return checkIfPositionWithinMap(position, mapLength, mapPosition) && checkIfIndexInRange(index, bufferSize);
}
// getPositionFromIndex: position = mapPosition + (0.5 * length - 0.5 * resolution) + resolution * (-(double)unwrappedIndex)
inline bool getPositionFromIndex(Position& position, const Index& index, const Length& mapLength, const Position& mapPosition,
                                 const double& resolution, const Size& bufferSize, const Index& bufferStartIndex) {
    if (!checkIfIndexInRange(index, bufferSize)) return false;
    Vector origin;
    getVectorToOrigin(origin, mapLength);
    const double ox = origin(0) - 0.5 * resolution, oy = origin(1) - 0.5 * resolution;  // getVectorToFirstCell
    const Index u = getIndexFromBufferIndex(index, bufferSize, bufferStartIndex);
    position(0) = (mapPosition(0) + ox) + resolution * static_cast<double>(-u(0));
    position(1) = (mapPosition(1) + oy) + resolution * static_cast<double>(-u(1));
    return true;
}
// getIndexShiftFromPositionShift: round half away from zero in grid units, then map frame -> buffer order (negate)
inline void getIndexShiftFromPositionShift(Index& indexShift, const Vector& positionShift, const double& resolution) {
    for (int i = 0; i < 2; ++i) {
        const double g = positionShift(i) / resolution;
        const int s = static_cast<int>(g + 0.5 * (g > 0 ? 1 : -1));
        indexShift(i) = -s;
    }
}
// getPositionShiftFromIndexShift: (-indexShift).cast<double>() * resolution
inline void getPositionShiftFromIndexShift(Vector& positionShift, const Index& indexShift, const double& resolution) {
    positionShift(0) = static_cast<double>(-indexShift(0)) * resolution;
    positionShift(1) = static_cast<double>(-indexShift(1)) * resolution;
}
}  // namespace math

// ---- GridMap.cpp (1.6.x) -----------------------------------------------------------------------
class GridMap {
  public:
    explicit GridMap(const std::vector<std::string>& layers) : layers_(layers) {
        position_.setZero();
        length_.setZero();
        resolution_ = 0.0;
        size_.setZero();
        startIndex_.setZero();
        timestamp_ = 0;
        for (auto& layer : layers_) data_.insert(std::pair<std::string, Matrix>(layer, Matrix()));
    }
    GridMap() : GridMap(std::vector<std::string>()) {}

    void setFrameId(const std::string& frameId) { frameId_ = frameId; }
    const std::string& getFrameId() const { return frameId_; }

    void setGeometry(const Length& length, const double resolution, const Position& position = Position(0.0, 0.0)) {
        Size size;
        size(0) = static_cast<int>(std::round(length(0) / resolution));
        size(1) = static_cast<int>(std::round(length(1) / resolution));
        resize(size);
        clearAll();
        resolution_ = resolution;
        length_(0) = static_cast<double>(size_(0)) * resolution_;
        length_(1) = static_cast<double>(size_(1)) * resolution_;
        position_ = position;
        startIndex_.setZero();
    }

    void add(const std::string& layer, const double value = NAN) { add(layer, Matrix::Constant(size_(0), size_(1), static_cast<float>(value))); }
    void add(const std::string& layer, const Matrix& data) {
        if (exists(layer)) {
            data_.at(layer) = data;  // same storage, overwritten in place
        } else {
            data_.insert(std::pair<std::string, Matrix>(layer, data));
            layers_.push_back(layer);
        }
    }
    bool exists(const std::string& layer) const { return data_.count(layer) != 0; }
    const Matrix& get(const std::string& layer) const {
        try {
            return data_.at(layer);
        } catch (const std::out_of_range&) {
            throw std::out_of_range("GridMap::get(...) : No map layer '" + layer + "' available.");
        }
    }
    Matrix& get(const std::string& layer) {
        try {
            return data_.at(layer);
        } catch (const std::out_of_range&) {
            throw std::out_of_range("GridMap::get(...) : No map layer of type '" + layer + "' available.");
        }
    }
    const Matrix& operator[](const std::string& layer) const { return get(layer); }
    Matrix& operator[](const std::string& layer) { return get(layer); }
    const std::vector<std::string>& getLayers() const { return layers_; }

    float& at(const std::string& layer, const Index& index) { return get(layer)(index(0), index(1)); }
    const float& at(const std::string& layer, const Index& index) const { return get(layer)(index(0), index(1)); }

    bool getIndex(const Position& position, Index& index) const {
        return math::getIndexFromPosition(index, position, length_, position_, resolution_, size_, startIndex_);
    }
    bool getPosition(const Index& index, Position& position) const {
        return math::getPositionFromIndex(position, index, length_, position_, resolution_, size_, startIndex_);
    }
    bool isInside(const Position& position) const { return math::checkIfPositionWithinMap(position, length_, position_); }

    bool move(const Position& position, std::vector<BufferRegion>& newRegions) {
        Index indexShift;
        Position positionShift(position(0) - position_(0), position(1) - position_(1));
        math::getIndexShiftFromPositionShift(indexShift, positionShift, resolution_);
        Position alignedPositionShift;
        math::getPositionShiftFromIndexShift(alignedPositionShift, indexShift, resolution_);

        for (int i = 0; i < 2; i++) {
            if (indexShift(i) != 0) {
                if (std::abs(indexShift(i)) >= getSize()(i)) {
                    clearAll();
                    newRegions.push_back(BufferRegion(Index(0, 0), getSize(), BufferRegion::Quadrant::Undefined));
                } else {
                    int sign = (indexShift(i) > 0 ? 1 : -1);
                    int startIndex = startIndex_(i) - (sign < 0 ? 1 : 0);
                    int endIndex = startIndex - sign + indexShift(i);
                    int nCells = std::abs(indexShift(i));
                    int index = (sign > 0 ? startIndex : endIndex);
                    math::wrapIndexToRange(index, getSize()(i));

                    if (index + nCells <= getSize()(i)) {
                        if (i == 0) {
                            clearRows(index, nCells);
                            newRegions.push_back(BufferRegion(Index(index, 0), Size(nCells, getSize()(1)), BufferRegion::Quadrant::Undefined));
                        } else {
                            clearCols(index, nCells);
                            newRegions.push_back(BufferRegion(Index(0, index), Size(getSize()(0), nCells), BufferRegion::Quadrant::Undefined));
                        }
                    } else {
                        int firstIndex = index;
                        int firstNCells = getSize()(i) - firstIndex;
                        if (i == 0) {
                            clearRows(firstIndex, firstNCells);
                            newRegions.push_back(BufferRegion(Index(firstIndex, 0), Size(firstNCells, getSize()(1)), BufferRegion::Quadrant::Undefined));
                        } else {
                            clearCols(firstIndex, firstNCells);
                            newRegions.push_back(BufferRegion(Index(0, firstIndex), Size(getSize()(0), firstNCells), BufferRegion::Quadrant::Undefined));
                        }
                        int secondIndex = 0;
                        int secondNCells = nCells - firstNCells;
                        if (i == 0) {
                            clearRows(secondIndex, secondNCells);
                            newRegions.push_back(BufferRegion(Index(secondIndex, 0), Size(secondNCells, getSize()(1)), BufferRegion::Quadrant::Undefined));
                        } else {
                            clearCols(secondIndex, secondNCells);
                            newRegions.push_back(BufferRegion(Index(0, secondIndex), Size(getSize()(0), secondNCells), BufferRegion::Quadrant::Undefined));
                        }
                    }
                }
            }
        }
        startIndex_(0) += indexShift(0);
        startIndex_(1) += indexShift(1);
        math::wrapIndexToRange(startIndex_(0), getSize()(0));
        math::wrapIndexToRange(startIndex_(1), getSize()(1));
        position_(0) += alignedPositionShift(0);
        position_(1) += alignedPositionShift(1);
        return indexShift(0) != 0 || indexShift(1) != 0;
    }
    bool move(const Position& position) {
        std::vector<BufferRegion> newRegions;
        return move(position, newRegions);
    }

    // new(i, j) = old((i + start_i) mod rows, (j + start_j) mod cols) for every layer
    void convertToDefaultStartIndex() {
        if (isDefaultStartIndex()) return;
        for (auto& kv : data_) {
            Matrix temp(kv.second);
            Matrix& m = kv.second;
            const Eigen::Index R = m.rows(), C = m.cols();
            for (Eigen::Index j = 0; j < C; ++j)
                for (Eigen::Index i = 0; i < R; ++i) m(i, j) = temp((i + startIndex_(0)) % R, (j + startIndex_(1)) % C);
        }
        startIndex_.setZero();
    }
    bool isDefaultStartIndex() const { return startIndex_(0) == 0 && startIndex_(1) == 0; }

    void clearAll() {
        for (auto& kv : data_) kv.second.setConstant(NAN);
    }
    void clearRows(unsigned int index, unsigned int nRows) {
        for (auto& layer : layers_) {
            Matrix& m = data_.at(layer);
            for (Eigen::Index j = 0; j < m.cols(); ++j)
                for (unsigned int i = index; i < index + nRows; ++i) m(i, j) = NAN;
        }
    }
    void clearCols(unsigned int index, unsigned int nCols) {
        for (auto& layer : layers_) {
            Matrix& m = data_.at(layer);
            for (unsigned int j = index; j < index + nCols; ++j)
                for (Eigen::Index i = 0; i < m.rows(); ++i) m(i, j) = NAN;
        }
    }

    const Length& getLength() const { return length_; }
    const Position& getPosition() const { return position_; }
    double getResolution() const { return resolution_; }
    const Size& getSize() const { return size_; }
    const Index& getStartIndex() const { return startIndex_; }
    void setTimestamp(uint64_t t) { timestamp_ = t; }

  private:
    void resize(const Size& size) {
        size_ = size;
        for (auto& kv : data_) kv.second.resize(size_(0), size_(1));
    }
    std::string frameId_;
    uint64_t timestamp_;
    std::unordered_map<std::string, Matrix> data_;
    std::vector<std::string> layers_;
    Length length_;
    double resolution_;
    Position position_;
    Size size_;
    Index startIndex_;
};

// SubmapIterator(gridMap, bufferRegion): walks the region row index fastest ... the reference only
// writes each visited cell independently (GroundGrid.cpp:121-133), so the visiting order is immaterial.
class SubmapIterator {
  public:
    SubmapIterator(const GridMap& gridMap, const BufferRegion& bufferRegion)
        : size_(gridMap.getSize()), start_(bufferRegion.getStartIndex()), sub_(bufferRegion.getSize()), k_(0) {}
    bool isPastEnd() const { return k_ >= (long)sub_(0) * sub_(1); }
    SubmapIterator& operator++() { ++k_; return *this; }
    const Index& operator*() const {
        // submap index (row-major walk like incrementIndexForSubmap), mapped into the circular buffer
        int si = (int)(k_ / sub_(1)), sj = (int)(k_ % sub_(1));
        cur_(0) = start_(0) + si;
        cur_(1) = start_(1) + sj;
        math::wrapIndexToRange(cur_(0), size_(0));
        math::wrapIndexToRange(cur_(1), size_(1));
        return cur_;
    }

  private:
    Size size_;
    Index start_;
    Size sub_;
    long k_;
    mutable Index cur_;
};

}  // namespace grid_map

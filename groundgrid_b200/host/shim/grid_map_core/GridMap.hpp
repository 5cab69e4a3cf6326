// grid_map::GridMap stand-in whose layers live on the GPU.
//
// In the reference the map is an in-memory grid_map::GridMap that GroundGrid creates and
// GroundSegmentation::filter_cloud mutates in place (src/GroundGrid.cpp:55-75,
// src/GroundSegmentation.cpp:61-78).  Here the map object owns a groundgrid_b200 handle
// (one slot): layers are device float buffers and operator[] / get() return a host mirror that
// is fetched on demand.  Geometry queries (getIndex, isInside, getPosition) follow
// grid_map_core 1.6.x (GridMapMath.cpp) in fp64.
#pragma once
#include <cmath>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "groundgrid_b200.h"

namespace grid_map {

struct Vec2d {
    double v[2] = {0, 0};
    Vec2d() = default;
    Vec2d(double a, double b) { v[0] = a; v[1] = b; }
    double& operator()(int i) { return v[i]; }
    double operator()(int i) const { return v[i]; }
    double x() const { return v[0]; }
    double y() const { return v[1]; }
};
struct Vec2i {
    int v[2] = {0, 0};
    Vec2i() = default;
    Vec2i(int a, int b) { v[0] = a; v[1] = b; }
    int& operator()(int i) { return v[i]; }
    int operator()(int i) const { return v[i]; }
};
typedef Vec2d Position;
typedef Vec2d Length;
typedef Vec2i Index;
typedef Vec2i Size;

// Column-major float matrix (host mirror of one layer), element (i, j) at i + j * rows.
class Matrix {
  public:
    Matrix() = default;
    Matrix(int r, int c) : r_(r), c_(c), d_((size_t)r * c, 0.f) {}
    int rows() const { return r_; }
    int cols() const { return c_; }
    float& operator()(int i, int j) { return d_[(size_t)i + (size_t)j * r_]; }
    float operator()(int i, int j) const { return d_[(size_t)i + (size_t)j * r_]; }
    float* data() { return d_.data(); }
    const float* data() const { return d_.data(); }
    size_t size() const { return d_.size(); }
    float minCoeff() const { float m = d_.at(0); for (float x : d_) m = x < m ? x : m; return m; }
    float maxCoeff() const { float m = d_.at(0); for (float x : d_) m = m < x ? x : m; return m; }
  private:
    int r_ = 0, c_ = 0;
    std::vector<float> d_;
};

struct BufferRegion {};

class GridMap {
  public:
    explicit GridMap(const std::vector<std::string>& layers) : layers_(layers) {}
    GridMap(const GridMap&) = delete;
    GridMap& operator=(const GridMap&) = delete;
    ~GridMap() { if (h_) gg_destroy(h_); }

    void setFrameId(const std::string& f) { frame_ = f; }
    const std::string& getFrameId() const { return frame_; }

    // size = round(length / resolution); creates the device maps (one slot).  The capacity /
    // device / flags can be preset with setDeviceOptions().
    void setGeometry(const Length& length, double resolution, const Position& position) {
        if (h_) { gg_destroy(h_); h_ = nullptr; }
        res_ = resolution;
        n_ = (int)std::round(length(0) / resolution);
        len_ = (double)n_ * resolution;
        pos_ = position;
        int rc = gg_create(length(0), (float)resolution, device_, 1, max_points_, GG_FLAG_FULL_LAYERS, nullptr, &h_);
        if (rc != GG_OK) throw std::runtime_error(std::string("groundgrid_b200: ") + gg_last_error());
        if (gg_cells_per_side(h_) != n_) throw std::runtime_error("groundgrid_b200: cell count mismatch");
    }
    void setDeviceOptions(int device, size_t max_points) { device_ = device; max_points_ = max_points; }

    gg_handle handle() const { return h_; }
    int slot() const { return 0; }

    const Size getSize() const { return Size(n_, n_); }
    double getResolution() const { return res_; }
    const Length getLength() const { return Length(len_, len_); }
    const Position getPosition() const {
        double xy[2] = {pos_(0), pos_(1)};
        if (h_) gg_get_map_position(h_, 0, xy);
        return Position(xy[0], xy[1]);
    }
    const std::vector<std::string>& getLayers() const { return layers_; }
    bool exists(const std::string& layer) const { for (auto& l : layers_) if (l == layer) return true; return false; }

    // adds (or overwrites) a layer name; the six per-scan layers of filter_cloud exist on the device already
    void add(const std::string& layer, double = 0.0) { if (!exists(layer)) layers_.push_back(layer); cache_.erase(layer); }

    // host mirror of a layer, fetched from the device (throws std::out_of_range like grid_map)
    const Matrix& get(const std::string& layer) const {
        if (!exists(layer)) throw std::out_of_range("GridMap::get(...) : No map layer '" + layer + "' available.");
        Matrix& m = cache_[layer];
        if (m.rows() != n_) m = Matrix(n_, n_);
        if (gg_get_layer(h_, 0, layer.c_str(), m.data()) != GG_OK) throw std::out_of_range(std::string("GridMap layer: ") + gg_last_error());
        return m;
    }
    const Matrix& operator[](const std::string& layer) const { return get(layer); }
    void set(const std::string& layer, const Matrix& m) {
        if (gg_set_layer(h_, 0, layer.c_str(), m.data()) != GG_OK) throw std::out_of_range(std::string("GridMap layer: ") + gg_last_error());
    }

    // grid_map_core 1.6.x geometry (start index (0,0))
    bool getIndex(const Position& p, Index& idx) const {
        const Position c = getPosition();
        const double half = 0.5 * len_;
        idx(0) = -(int)((p(0) - half - c(0)) / res_);
        idx(1) = -(int)((p(1) - half - c(1)) / res_);
        return isInside(p) && idx(0) >= 0 && idx(1) >= 0 && idx(0) < n_ && idx(1) < n_;
    }
    bool isInside(const Position& p) const {
        const Position c = getPosition();
        const double half = 0.5 * len_;
        const double tx = -(p(0) - c(0) - half), ty = -(p(1) - c(1) - half);
        return tx >= 0.0 && ty >= 0.0 && tx < len_ && ty < len_;
    }
    bool getPosition(const Index& idx, Position& p) const {
        if (idx(0) < 0 || idx(1) < 0 || idx(0) >= n_ || idx(1) >= n_) return false;
        const Position c = getPosition();
        const double off = 0.5 * len_ - 0.5 * res_;
        p(0) = (c(0) + off) + res_ * (double)(-idx(0));
        p(1) = (c(1) + off) + res_ * (double)(-idx(1));
        return true;
    }

  private:
    std::vector<std::string> layers_;
    std::string frame_;
    gg_handle h_ = nullptr;
    int n_ = 0, device_ = 0;
    size_t max_points_ = 1u << 20;
    double res_ = 0.0, len_ = 0.0;
    Position pos_;
    mutable std::map<std::string, Matrix> cache_;
};
}  // namespace grid_map

// pcl::PointCloud stand-in: header + contiguous, 16-byte aligned point storage.
#pragma once
#include <std_msgs/Header.h>
#include <cstdlib>
#include <memory>
#include <new>
#include <vector>
namespace pcl {
template <typename T, size_t Align = 16>
struct aligned_allocator {
    typedef T value_type;
    aligned_allocator() = default;
    template <class U> aligned_allocator(const aligned_allocator<U, Align>&) {}
    template <class U> struct rebind { typedef aligned_allocator<U, Align> other; };
    T* allocate(size_t n) {
        void* p = nullptr;
        if (posix_memalign(&p, Align, n * sizeof(T) ? n * sizeof(T) : Align)) throw std::bad_alloc();
        return static_cast<T*>(p);
    }
    void deallocate(T* p, size_t) { std::free(p); }
    bool operator==(const aligned_allocator&) const { return true; }
    bool operator!=(const aligned_allocator&) const { return false; }
};
struct PCLHeader { uint32_t seq = 0; uint64_t stamp = 0; std::string frame_id; };
template <typename PointT>
class PointCloud {
  public:
    typedef std::shared_ptr<PointCloud<PointT>> Ptr;
    typedef std::shared_ptr<const PointCloud<PointT>> ConstPtr;
    PCLHeader header;
    std::vector<PointT, aligned_allocator<PointT>> points;
    uint32_t width = 0, height = 1;
    bool is_dense = true;
    size_t size() const { return points.size(); }
};
}  // namespace pcl

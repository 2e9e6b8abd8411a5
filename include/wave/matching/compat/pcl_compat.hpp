// Minimal stand-ins for the PCL / Boost types that appear in wave_matching's public interface,
// used ONLY when the real headers are not installed (this image has neither PCL nor Boost).
// With real PCL present, <pcl/point_types.h> / <pcl/point_cloud.h> are used instead and this file
// is never included (see pcl_common.hpp).  Layout note: pcl::PointXYZ is 16 bytes (x, y, z and a
// pad that PCL sets to 1.0f), so a PointCloud's points array is already the float4 array the C ABI
// takes - no repacking on the way to the device.
#ifndef WAVE_MATCHING_COMPAT_PCL_COMPAT_HPP
#define WAVE_MATCHING_COMPAT_PCL_COMPAT_HPP

#include <cstddef>
#include <memory>
#include <vector>

namespace pcl {

struct alignas(16) PointXYZ {
    float x = 0.f, y = 0.f, z = 0.f;
    float pad = 1.f;
    PointXYZ() = default;
    PointXYZ(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
};

template <typename PointT>
class PointCloud {
 public:
    typedef std::shared_ptr<PointCloud<PointT>> Ptr;
    typedef std::shared_ptr<const PointCloud<PointT>> ConstPtr;
    std::vector<PointT> points;
    unsigned width = 0, height = 1;
    bool is_dense = true;

    std::size_t size() const { return points.size(); }
    bool empty() const { return points.empty(); }
    void clear() { points.clear(); width = 0; }
    void push_back(const PointT &p) { points.push_back(p); width = static_cast<unsigned>(points.size()); }
    void resize(std::size_t n) { points.resize(n); width = static_cast<unsigned>(n); }
    PointT &at(std::size_t i) { return points.at(i); }
    const PointT &at(std::size_t i) const { return points.at(i); }
    PointT &operator[](std::size_t i) { return points[i]; }
    const PointT &operator[](std::size_t i) const { return points[i]; }
};

}  // namespace pcl

#if !defined(BOOST_VERSION)
namespace boost {  // the reference spells its smart pointers boost::make_shared / boost::shared_ptr
using std::make_shared;
using std::shared_ptr;
}  // namespace boost
#endif

#endif  // WAVE_MATCHING_COMPAT_PCL_COMPAT_HPP

// Minimal stand-ins for the PCL / Boost types that appear in wave_matching's public interface,
// used ONLY when the real headers are not installed (this image has neither PCL nor Boost).
// With real PCL present, <pcl/point_types.h> / <pcl/point_cloud.h> are used instead and this file
// is never included (see pcl_common.hpp).  Layout note: pcl::PointXYZ is 16 bytes (x, y, z and a
// pad that PCL sets to 1.0f), so a PointCloud's points array is already the float4 array the C ABI
// takes - no repacking on the way to the device.
#ifndef WAVE_MATCHING_COMPAT_PCL_COMPAT_HPP
#define WAVE_MATCHING_COMPAT_PCL_COMPAT_HPP

#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <memory>
#include <sstream>
#include <string>
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

namespace io {

// pcl::io::loadPCDFile<pcl::PointXYZ> (the reference tests load their fixture with it,
// wave_matching/tests/icp_tests.cpp:26): PCD v0.7, DATA ascii or binary; x, y, z are taken from
// wherever FIELDS / SIZE / COUNT place them, every other field is skipped (the fixture's records
// are 32 bytes: x y z _ intensity ring _).  Returns 0 on success, -1 on
// failure, as PCL does.
inline int loadPCDFile(const std::string &path, PointCloud<PointXYZ> &cloud) {
    std::ifstream in(path, std::ios::binary);
    if (!in.good()) return -1;
    std::vector<std::string> fields;
    std::vector<int> sizes, counts;
    std::vector<char> types;
    std::size_t n_points = 0, width = 0, height = 1;
    std::string data_kind, line;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty() || line[0] == '#') continue;
        std::istringstream ls(line);
        std::string key, tok;
        ls >> key;
        if (key == "FIELDS" || key == "COLUMNS") while (ls >> tok) fields.push_back(tok);
        else if (key == "SIZE") while (ls >> tok) sizes.push_back(std::stoi(tok));
        else if (key == "TYPE") while (ls >> tok) types.push_back(tok.empty() ? 'F' : tok[0]);
        else if (key == "COUNT") while (ls >> tok) counts.push_back(std::stoi(tok));
        else if (key == "WIDTH") ls >> width;
        else if (key == "HEIGHT") ls >> height;
        else if (key == "POINTS") ls >> n_points;
        else if (key == "DATA") {
            ls >> data_kind;
            break;
        }
    }
    if (fields.empty() || sizes.size() != fields.size()) return -1;
    if (counts.empty()) counts.assign(fields.size(), 1);
    if (types.empty()) types.assign(fields.size(), 'F');
    if (counts.size() != fields.size() || types.size() != fields.size()) return -1;
    if (n_points == 0) n_points = width * height;
    std::size_t stride = 0;
    int off[3] = {-1, -1, -1}, col[3] = {-1, -1, -1}, column = 0;
    for (std::size_t f = 0; f < fields.size(); ++f) {
        for (int a = 0; a < 3; ++a)
            if (fields[f] == std::string(1, "xyz"[a]) && sizes[f] == 4 && types[f] == 'F') {
                off[a] = static_cast<int>(stride);
                col[a] = column;
            }
        stride += static_cast<std::size_t>(sizes[f]) * static_cast<std::size_t>(counts[f]);
        column += counts[f];
    }
    if (off[0] < 0 || off[1] < 0 || off[2] < 0) return -1;
    cloud.points.clear();
    cloud.points.reserve(n_points);
    if (data_kind == "binary") {
        std::vector<char> rec(stride);
        for (std::size_t i = 0; i < n_points; ++i) {
            in.read(rec.data(), static_cast<std::streamsize>(stride));
            if (in.gcount() != static_cast<std::streamsize>(stride)) return -1;
            PointXYZ p;
            std::memcpy(&p.x, rec.data() + off[0], 4);
            std::memcpy(&p.y, rec.data() + off[1], 4);
            std::memcpy(&p.z, rec.data() + off[2], 4);
            cloud.points.push_back(p);
        }
    } else if (data_kind == "ascii") {
        for (std::size_t i = 0; i < n_points && std::getline(in, line); ++i) {
            std::istringstream ls(line);
            std::string tok;
            PointXYZ p;
            for (int c = 0; ls >> tok; ++c) {
                if (c == col[0]) p.x = std::stof(tok);
                else if (c == col[1]) p.y = std::stof(tok);
                else if (c == col[2]) p.z = std::stof(tok);
            }
            cloud.points.push_back(p);
        }
        if (cloud.points.size() != n_points) return -1;
    } else {
        return -1;  // binary_compressed: not produced by anything on this path
    }
    cloud.width = static_cast<unsigned>(height > 1 ? width : cloud.points.size());
    cloud.height = static_cast<unsigned>(height > 1 ? height : 1);
    cloud.is_dense = true;
    return 0;
}

}  // namespace io

// pcl::transformPointCloud(in, out, Eigen::Affine3d) as the reference tests use it
// (tests/icp_tests.cpp:31): per point, double arithmetic left to right, cast to float.  A template
// so that this header does not depend on which Affine3d (Eigen's or the stand-in) is in use.
template <typename PointT, typename TransformT>
inline void transformPointCloud(const PointCloud<PointT> &in, PointCloud<PointT> &out, const TransformT &T) {
    const auto &m = T.matrix();
    std::vector<PointT> pts;
    pts.reserve(in.points.size());
    for (const PointT &p : in.points) {
        const double x = p.x, y = p.y, z = p.z;
        PointT q = p;
        q.x = static_cast<float>(m(0, 0) * x + m(0, 1) * y + m(0, 2) * z + m(0, 3));
        q.y = static_cast<float>(m(1, 0) * x + m(1, 1) * y + m(1, 2) * z + m(1, 3));
        q.z = static_cast<float>(m(2, 0) * x + m(2, 1) * y + m(2, 2) * z + m(2, 3));
        pts.push_back(q);
    }
    out.points.swap(pts);
    out.width = static_cast<unsigned>(out.points.size());
    out.height = 1;
    out.is_dense = in.is_dense;
}

}  // namespace pcl

#if !defined(BOOST_VERSION)
namespace boost {  // the reference spells its smart pointers boost::make_shared / boost::shared_ptr
using std::make_shared;
using std::shared_ptr;
}  // namespace boost
#endif

#endif  // WAVE_MATCHING_COMPAT_PCL_COMPAT_HPP

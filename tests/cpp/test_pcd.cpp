// CPU test of the PCL stand-ins a literal build of the reference tests needs
// (include/wave/matching/compat/): pcl::io::loadPCDFile as wave_matching/tests/icp_tests.cpp:26
// calls it (included the reference's way, <pcl/io/pcd_io.h>), pcl::transformPointCloud
// (tests/icp_tests.cpp:31), and the implicit constructors of the reference's parameter structs
// (icp.hpp:31, gicp.hpp:31, ndt.hpp:35, matcher.hpp:32).
// Usage: test_pcd <binary.pcd> <ascii.pcd> <xyz.f32>; exit code = number of failures.
#include <pcl/io/pcd_io.h>
#include <pcl/common/transforms.h>

#include <cstdio>
#include <fstream>
#include <string>
#include <type_traits>
#include <vector>

#include "wave/matching/gicp.hpp"
#include "wave/matching/icp.hpp"
#include "wave/matching/ndt.hpp"

static_assert(std::is_convertible<std::string, wave::ICPMatcherParams>::value, "ICPMatcherParams p = path;");
static_assert(std::is_convertible<std::string, wave::GICPMatcherParams>::value, "GICPMatcherParams p = path;");
static_assert(std::is_convertible<std::string, wave::NDTMatcherParams>::value, "NDTMatcherParams p = path;");
static_assert(std::is_convertible<float, wave::Matcher<wave::PCLPointCloudPtr>>::value ||
                  std::is_abstract<wave::Matcher<wave::PCLPointCloudPtr>>::value,
              "Matcher(float) is implicit in the reference");

int main(int argc, char **argv) {
    if (argc < 4) return 100;
    int failures = 0;
    std::ifstream in(argv[3], std::ios::binary);
    in.seekg(0, std::ios::end);
    std::vector<float> xyz(static_cast<size_t>(in.tellg()) / 4);
    in.seekg(0);
    in.read(reinterpret_cast<char *>(xyz.data()), static_cast<std::streamsize>(xyz.size() * 4));
    const size_t n = xyz.size() / 3;

    pcl::PointCloud<pcl::PointXYZ> cloud;
    if (pcl::io::loadPCDFile(argv[1], cloud) != 0) { std::printf("FAIL load binary\n"); ++failures; }
    if (cloud.size() != n) { std::printf("FAIL binary size %zu != %zu\n", cloud.size(), n); ++failures; }
    for (size_t i = 0; i < cloud.size() && i < n; ++i)
        if (cloud.points[i].x != xyz[3 * i] || cloud.points[i].y != xyz[3 * i + 1] || cloud.points[i].z != xyz[3 * i + 2]) {
            std::printf("FAIL binary point %zu\n", i);
            ++failures;
            break;
        }
    pcl::PointCloud<pcl::PointXYZ> ascii;
    if (pcl::io::loadPCDFile(argv[2], ascii) != 0) { std::printf("FAIL load ascii\n"); ++failures; }
    for (size_t i = 0; i < ascii.size(); ++i)
        if (ascii.points[i].x != xyz[3 * i] || ascii.points[i].y != xyz[3 * i + 1] || ascii.points[i].z != xyz[3 * i + 2]) {
            std::printf("FAIL ascii point %zu\n", i);
            ++failures;
            break;
        }
    if (ascii.size() != 100) { std::printf("FAIL ascii size %zu\n", ascii.size()); ++failures; }
    pcl::PointCloud<pcl::PointXYZ> missing;
    if (pcl::io::loadPCDFile("/nonexistent.pcd", missing) != -1) { std::printf("FAIL missing file\n"); ++failures; }

    wave::Affine3 perturb = wave::Affine3::Identity();
    perturb.translation() << 0.2, 0, 0;
    pcl::PointCloud<pcl::PointXYZ> moved;
    pcl::transformPointCloud(cloud, moved, perturb);
    if (moved.size() != cloud.size()) { std::printf("FAIL transform size\n"); ++failures; }
    for (size_t i = 0; i < moved.size(); ++i) {
        const float want = static_cast<float>(static_cast<double>(cloud.points[i].x) + 0.2);
        if (moved.points[i].x != want || moved.points[i].y != cloud.points[i].y || moved.points[i].z != cloud.points[i].z) {
            std::printf("FAIL transform point %zu\n", i);
            ++failures;
            break;
        }
    }
    std::printf("%s pcd: %zu points, %d failures\n", failures ? "FAILED" : "PASSED", cloud.size(), failures);
    return failures;
}

// Native throughput of the reference-facing C++ API: wave::MultiMatcher<ICPMatcher, ICPMatcherParams> (the reference's
// own batch interface, multi_matcher.hpp:30-96) on one GPU, no Python anywhere.  Every job is what the reference's worker
// does (impl/multi_matcher_impl.hpp:45-48): setRef + setTarget (upload from ordinary host memory, index build) + match()
// + estimateInfo() (LUMold: one more nearest-neighbour pass).
//   multimatcher_bench <src_xyz.f32> <tgt_xyz.f32> <workers> <jobs>      (raw float32 x y z triples)
// Built by tools/native/Makefile into tools/native/_build/.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "wave/matching/icp.hpp"
#include "wave/matching/multi_matcher.hpp"

using namespace wave;

static PCLPointCloudPtr load_xyz(const std::string &path) {
    std::ifstream in(path, std::ios::binary);
    if (!in.good()) throw std::runtime_error("cannot open " + path);
    in.seekg(0, std::ios::end);
    const size_t bytes = static_cast<size_t>(in.tellg());
    in.seekg(0);
    std::vector<float> xyz(bytes / 4);
    in.read(reinterpret_cast<char *>(xyz.data()), static_cast<std::streamsize>(bytes));
    auto cloud = boost::make_shared<pcl::PointCloud<pcl::PointXYZ>>();
    cloud->points.reserve(xyz.size() / 3);
    for (size_t i = 0; i + 2 < xyz.size(); i += 3) cloud->push_back(pcl::PointXYZ(xyz[i], xyz[i + 1], xyz[i + 2]));
    return cloud;
}

int main(int argc, char **argv) {
    if (argc < 5) {
        std::fprintf(stderr, "usage: %s src.f32 tgt.f32 workers jobs\n", argv[0]);
        return 2;
    }
    const PCLPointCloudPtr src = load_xyz(argv[1]), tgt = load_xyz(argv[2]);
    const int workers = std::atoi(argv[3]), jobs = std::atoi(argv[4]);
    ICPMatcherParams params;
    params.res = -1;   // full resolution (BASELINE configs 2 and 5)
    MultiMatcher<ICPMatcher, ICPMatcherParams> mm(workers, 2 * workers, params);
    auto round_of = [&](int count) {
        for (int i = 0; i < count; ++i) mm.insert(i, src, tgt);
        int id = 0, got = 0;
        Affine3 T;
        Mat6 info;
        while (mm.getResult(&id, &T, &info)) ++got;
        return got;
    };
    round_of(2 * workers);   // warm-up: allocations, graph captures
    const auto t0 = std::chrono::steady_clock::now();
    const int got = round_of(jobs);
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    std::printf("MultiMatcher<ICPMatcher> workers=%d jobs=%d results=%d points=%zu/%zu: %.3f ms per job (wall clock)\n",
                workers, jobs, got, src->points.size(), tgt->points.size(), ms / jobs);
    return got == jobs ? 0 : 1;
}

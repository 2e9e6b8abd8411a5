// The reference's gtest cases for ICPMatcher and MultiMatcher (wave_matching/tests/icp_tests.cpp,
// tests/multi_matcher_tests.cpp) re-stated against the drop-in headers, with a tiny assert harness
// instead of gtest (not available in this image).  Usage: test_matching <testscan_xyz.f32> <icp.yaml>
// Prints one line per case; exit code = number of failed cases.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <random>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "wave/matching/gicp.hpp"
#include "wave/matching/icp.hpp"
#include "wave/matching/multi_matcher.hpp"
#include "wave/matching/ndt.hpp"

using namespace wave;

static int failures = 0;
#define EXPECT(cond, name)                                                     \
    do {                                                                       \
        if (!(cond)) {                                                         \
            std::printf("FAIL %s: %s\n", name, #cond);                         \
            ++failures;                                                        \
        }                                                                      \
    } while (0)

static PCLPointCloudPtr load_scan(const std::string &path) {
    std::ifstream in(path, std::ios::binary);
    if (!in.good()) throw std::runtime_error("cannot open " + path);
    in.seekg(0, std::ios::end);
    const size_t bytes = static_cast<size_t>(in.tellg());
    in.seekg(0);
    std::vector<float> xyz(bytes / 4);
    in.read(reinterpret_cast<char *>(xyz.data()), static_cast<std::streamsize>(bytes));
    auto cloud = boost::make_shared<pcl::PointCloud<pcl::PointXYZ>>();
    for (size_t i = 0; i + 2 < xyz.size(); i += 3) cloud->push_back(pcl::PointXYZ(xyz[i], xyz[i + 1], xyz[i + 2]));
    return cloud;
}

// pcl::transformPointCloud(in, out, Affine3d): double arithmetic, cast to float
static PCLPointCloudPtr transformed(const PCLPointCloudPtr &in, const Affine3 &T) {
    auto out = boost::make_shared<pcl::PointCloud<pcl::PointXYZ>>();
    const auto &m = T.matrix();
    for (const auto &p : in->points) {
        const double x = p.x, y = p.y, z = p.z;
        out->push_back(pcl::PointXYZ(static_cast<float>(m(0, 0) * x + m(0, 1) * y + m(0, 2) * z + m(0, 3)),
                                     static_cast<float>(m(1, 0) * x + m(1, 1) * y + m(1, 2) * z + m(1, 3)),
                                     static_cast<float>(m(2, 0) * x + m(2, 1) * y + m(2, 2) * z + m(2, 3))));
    }
    return out;
}

struct Case {
    const char *name;
    float res;
    int multiscale_steps;
    double tx;
    bool with_info;
};

int main(int argc, char **argv) {
    if (argc < 3) {
        std::printf("usage: %s testscan_xyz.f32 icp.yaml\n", argv[0]);
        return 100;
    }
    const std::string scan = argv[1], config = argv[2];
    const float threshold = 0.1f;  // tests/icp_tests.cpp:37
    PCLPointCloudPtr ref = load_scan(scan);
    std::printf("loaded %zu points\n", ref->size());

    {  // ICPTests.initialization
        ICPMatcher matcher{ICPMatcherParams()};
        EXPECT(matcher.getRes() > 0.09f, "initialization");
    }
    {  // params constructor: a missing file must throw std::runtime_error (src/icp.cpp:18-20)
        bool threw = false;
        try {
            ICPMatcherParams bad("/nonexistent/icp.yaml");
        } catch (const std::runtime_error &) {
            threw = true;
        }
        EXPECT(threw, "params_missing_file_throws");
        ICPMatcherParams ok(config);
        EXPECT(ok.max_iter == 100 && ok.multiscale_steps == 0 && std::fabs(ok.res - 0.1f) < 1e-6f, "params_from_yaml");
    }
    const Case cases[] = {
      {"fullResNullMatch", -1.f, 0, 0.0, false},  {"nullDisplacement", 0.05f, 0, 0.0, false},
      {"smallDisplacement", 0.05f, 0, 0.2, false}, {"smallinfo", 0.05f, 0, 0.2, true},
      {"multiscale", 0.1f, 3, 0.2, true},
    };
    for (const Case &c : cases) {
        Affine3 perturb = Affine3::Identity();
        perturb.translation() << c.tx, 0, 0;
        ICPMatcherParams params(config);
        params.res = c.res;
        params.multiscale_steps = c.multiscale_steps;
        ICPMatcher matcher(params);
        PCLPointCloudPtr target = transformed(ref, perturb);
        matcher.setup(ref, target);
        const bool match_success = matcher.match();
        const double diff = (matcher.getResult().matrix() - perturb.matrix()).norm();
        EXPECT(match_success, c.name);
        EXPECT(diff < threshold, c.name);
        if (c.with_info) {
            matcher.estimateInfo();
            const auto info = matcher.getInfo();
            EXPECT(info(0, 0) > 0, c.name);
        }
        std::printf("%-18s match=%d diff=%.3e\n", c.name, (int) match_success, diff);
    }
    {  // ICPTests.lumvslum: one scan distorted by U(-0.3, 0.3)
        Affine3 perturb = Affine3::Identity();
        perturb.translation() << 0.2, 0, 0;
        PCLPointCloudPtr target = transformed(ref, perturb);
        std::uniform_real_distribution<double> unif(-0.3, 0.3);
        std::default_random_engine re;
        for (size_t i = 0; i < target->size(); i++) {
            target->at(i).x += static_cast<float>(unif(re));
            target->at(i).y += static_cast<float>(unif(re));
            target->at(i).z += static_cast<float>(unif(re));
        }
        ICPMatcherParams params(config);
        params.res = 0.05f;
        params.covar_estimator = ICPMatcherParams::covar_method::LUMold;
        ICPMatcher matcher1(params);
        matcher1.setup(ref, target);
        matcher1.match();
        matcher1.estimateInfo();
        const auto info1 = matcher1.getInfo();
        params.covar_estimator = ICPMatcherParams::covar_method::LUM;
        ICPMatcher matcher2(params);
        matcher2.setup(ref, target);
        matcher2.match();
        matcher2.estimateInfo();
        const auto info2 = matcher2.getInfo();
        const double diff = (info1 - info2).norm();
        EXPECT(info1(0, 0) > 0, "lumvslum");
        EXPECT(diff < 0.01, "lumvslum");  // both settings end in estimateLUMold (fall-through)
        std::printf("%-18s info(0,0)=%.4f diff=%.3e\n", "lumvslum", info1(0, 0), diff);
    }
    {  // MultiTests.initialization + simultaneousmatching + the getResult the reference never defined
        { MultiMatcher<ICPMatcher, ICPMatcherParams> idle(2, 10, ICPMatcherParams()); }
        ICPMatcherParams params;
        params.res = 0.2f;
        params.multiscale_steps = 0;
        MultiMatcher<ICPMatcher, ICPMatcherParams> matcher(4, 10, params);
        PCLPointCloudPtr dupes[9];
        for (auto &d : dupes) {
            d = boost::make_shared<pcl::PointCloud<pcl::PointXYZ>>();
            *d = *ref;
        }
        for (int i = 0; i < 8; i++) matcher.insert(i, dupes[i], dupes[i + 1]);
        while (!matcher.done()) std::this_thread::sleep_for(std::chrono::milliseconds(5));
        int got = 0, id = -1;
        Eigen::Affine3d T;
        Mat6 info;
        bool seen[8] = {false};
        while (matcher.getResult(&id, &T, &info)) {
            ++got;
            if (id >= 0 && id < 8) seen[id] = true;
            EXPECT((T.matrix() - Affine3::Identity().matrix()).norm() < threshold, "simultaneousmatching");
        }
        bool all = true;
        for (bool s : seen) all = all && s;
        EXPECT(got == 8 && all, "simultaneousmatching");
        std::printf("%-18s results=%d\n", "multimatcher", got);
    }
    {  // extension: one map for all jobs (MultiMatcher::setMap), results equal to per-job targets
        Affine3 perturb = Affine3::Identity();
        perturb.translation() << 0.2, 0, 0;
        PCLPointCloudPtr map = transformed(ref, perturb);
        ICPMatcherParams params(config);
        params.res = -1;
        ICPMatcher single(params);
        single.setup(ref, map);
        const bool ok1 = single.match();
        MultiMatcher<ICPMatcher, ICPMatcherParams> mm(3, 4, params);
        mm.setMap(map);
        for (int i = 0; i < 6; ++i) mm.insert(i, ref, nullptr);
        int id = 0, got = 0;
        Affine3 T;
        Mat6 info;
        while (mm.getResult(&id, &T, &info)) {
            ++got;
            EXPECT(ok1 && (T.matrix() - single.getResult().matrix()).norm() == 0.0, "multimatcher_map");
        }
        EXPECT(got == 6, "multimatcher_map");
        std::printf("%-18s results=%d diff_vs_truth=%.3e\n", "multimatcher_map", got, (T.matrix() - perturb.matrix()).norm());
    }
    if (argc >= 4) {  // NDTTests (tests/ndt_tests.cpp): initialization, fullResNullMatch, nullDisplacement,
                      // smallDisplacement
        const std::string ndt_config = argv[3];
        const float ndt_threshold = 0.12f;  // tests/ndt_tests.cpp:37
        { NDTMatcher matcher{NDTMatcherParams()}; }
        struct NCase { float res; double tx; };
        const NCase ncases[] = {{-1.f /* keep the yaml's 0.05 */, 0.0}, {0.1f, 0.0}, {0.3f, 0.2}};
        for (const NCase &nc : ncases) {
            const float r = nc.res;
            Affine3 perturb = Affine3::Identity();
            perturb.translation() << nc.tx, 0, 0;
            NDTMatcherParams params(ndt_config);
            if (r > 0) params.res = r;
            NDTMatcher matcher(params);
            PCLPointCloudPtr target = transformed(ref, perturb);
            matcher.setup(ref, target);
            const bool match_success = matcher.match();
            const double diff = (matcher.getResult().matrix() - perturb.matrix()).norm();
            const char *name = nc.tx == 0.0 ? "ndt_null" : "ndt_smallDisp";
            EXPECT(match_success, name);
            EXPECT(diff < ndt_threshold, name);
            std::printf("%-18s res=%.2f match=%d diff=%.3e\n", name, matcher.getRes(), (int) match_success, diff);
        }
    }
    if (argc >= 5) {  // GICPTests (tests/gicp_tests.cpp): fullResNullMatch, nullDisplacement, smallDisplacement
        const std::string gicp_config = argv[4];
        { GICPMatcher matcher{GICPMatcherParams()}; }
        struct GCase { const char *name; float res; double tx; };
        const GCase gcases[] = {{"gicp_fullResNull", -1.f, 0.0}, {"gicp_nullDisp", 0.05f, 0.0}, {"gicp_smallDisp", 0.05f, 0.2}};
        for (const GCase &c : gcases) {
            Affine3 perturb = Affine3::Identity();
            perturb.translation() << c.tx, 0, 0;
            GICPMatcherParams params(gicp_config);
            params.res = c.res;
            GICPMatcher matcher(params);
            PCLPointCloudPtr target = transformed(ref, perturb);
            matcher.setup(ref, target);
            const bool match_success = matcher.match();
            const double diff = (matcher.getResult().matrix() - perturb.matrix()).norm();
            EXPECT(match_success, c.name);
            EXPECT(diff < threshold, c.name);
            std::printf("%-18s match=%d diff=%.3e\n", c.name, (int) match_success, diff);
        }
    }
    std::printf("%s (%d failures)\n", failures ? "FAILED" : "PASSED", failures);
    return failures;
}

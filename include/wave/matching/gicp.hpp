// wave::GICPMatcher on the B200 - same public surface as the reference
// (wave_matching/include/wave/matching/gicp.hpp:30-65).  The private
// pcl::GeneralizedIterativeClosestPoint and pcl::VoxelGrid members (gicp.hpp:61-62) are replaced by a
// handle of the C ABI.
#ifndef WAVE_MATCHING_GICP_HPP
#define WAVE_MATCHING_GICP_HPP

#include <string>

#include "wave/matching/matcher.hpp"
#include "wave/matching/pcl_common.hpp"

struct wavecu_gicp;

namespace wave {

struct GICPMatcherParams {
    GICPMatcherParams(const std::string &config_path);  // implicit, as gicp.hpp:31
    GICPMatcherParams() {}

    int corr_rand = 10;     ///< neighbours used for each point's covariance
    int max_iter = 100;     ///< cap on outer iterations
    double r_eps = 1e-8;    ///< rotation epsilon of the convergence test
    double fit_eps = 1e-2;  ///< handed to PCL by the reference, unused by GICP
    float res = 0.1f;       ///< voxel filter applied in setRef / setTarget; <= 0: none
};

class GICPMatcher : public Matcher<PCLPointCloudPtr> {
 public:
    explicit GICPMatcher(GICPMatcherParams params1);
    ~GICPMatcher();
    GICPMatcher(GICPMatcher &&other) noexcept;
    GICPMatcher(const GICPMatcher &) = delete;
    GICPMatcher &operator=(const GICPMatcher &) = delete;

    void setRef(const PCLPointCloudPtr &ref);
    void setTarget(const PCLPointCloudPtr &target);
    bool match();

 private:
    wavecu_gicp *handle = nullptr;
    PCLPointCloudPtr ref, target;
    GICPMatcherParams params;
};

}  // namespace wave

#endif  // WAVE_MATCHING_GICP_HPP

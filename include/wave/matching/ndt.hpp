// wave::NDTMatcher on the B200 - same public surface as the reference
// (wave_matching/include/wave/matching/ndt.hpp:33-80).  The private
// pcl::NormalDistributionsTransform member (ndt.hpp:72) is replaced by a handle of the C ABI.
#ifndef WAVE_MATCHING_NDT_HPP
#define WAVE_MATCHING_NDT_HPP

#include <string>

#include "wave/matching/matcher.hpp"
#include "wave/matching/pcl_common.hpp"

struct wavecu_ndt;

namespace wave {

struct NDTMatcherParams {
    NDTMatcherParams() {}
    NDTMatcherParams(const std::string &config_path);  // implicit, as ndt.hpp:35

    int step_size = 3;            ///< maximum Newton line-search step (an int in the reference too)
    int max_iter = 100;           ///< cap on iterations
    double t_eps = 1e-8;          ///< stop when the step is shorter than this
    float res = 5;                ///< voxel edge length of the normal-distribution grid
    const float min_res = 0.05f;  ///< smaller resolutions are replaced by this one
};

class NDTMatcher : public Matcher<PCLPointCloudPtr> {
 public:
    explicit NDTMatcher(NDTMatcherParams params1);
    ~NDTMatcher();
    NDTMatcher(NDTMatcher &&other) noexcept;
    NDTMatcher(const NDTMatcher &) = delete;
    NDTMatcher &operator=(const NDTMatcher &) = delete;

    void setRef(const PCLPointCloudPtr &ref);
    void setTarget(const PCLPointCloudPtr &target);
    /// blocks until finished; note the reference's own remark that this NDT is slow (ndt.hpp:65)
    bool match();

 private:
    wavecu_ndt *handle = nullptr;
    PCLPointCloudPtr ref, target;
    NDTMatcherParams params;
};

}  // namespace wave

#endif  // WAVE_MATCHING_NDT_HPP

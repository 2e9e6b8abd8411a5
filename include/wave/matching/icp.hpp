// wave::ICPMatcher on the B200 - same public surface as the reference
// (wave_matching/include/wave/matching/icp.hpp:30-120): ICPMatcherParams with the same fields,
// defaults and YAML constructor, a public `params` member, setRef / setTarget / match /
// estimateInfo.  The private pcl::IterativeClosestPoint and pcl::VoxelGrid members of the
// reference (icp.hpp:100-102) are replaced by an opaque handle of the C ABI (include/wavecu.h);
// private layout is not API.
#ifndef WAVE_MATCHING_ICP_HPP
#define WAVE_MATCHING_ICP_HPP

#include <string>

#include "wave/matching/matcher.hpp"
#include "wave/matching/pcl_common.hpp"

struct wavecu_icp;

namespace wave {

struct ICPMatcherParams {
    ICPMatcherParams(const std::string &config_path);  // implicit, as icp.hpp:31
    ICPMatcherParams() {}

    double max_corr = 3;               ///< correspondences farther apart than this are dropped
    int max_iter = 100;                ///< cap on ICP iterations
    double t_eps = 1e-8;               ///< stop when the transform changes by less than this
    double fit_eps = 1e-2;             ///< stop when the relative cost change is below this
    double lidar_ang_covar = 7.78e-9;  ///< sensor model, Censi estimator
    double lidar_lin_covar = 2.5e-4;   ///< sensor model, Censi estimator
    int multiscale_steps = 3;          ///< > 0: coarse-to-fine, each step halves the voxel size
    float res = 0.1f;                  ///< voxel size of the (finest) match; <= 0: no down-sampling
    enum covar_method : int { LUM, CENSI, LUMold } covar_estimator = covar_method::LUM;

    /// extension (not in the reference): 0 = SVD / Umeyama as the reference,
    /// 1 = point-to-plane linear least squares (needs setTargetNormals and res <= 0)
    int estimator = 0;
};

class ICPMatcher : public Matcher<PCLPointCloudPtr> {
 public:
    explicit ICPMatcher(ICPMatcherParams params1);
    ~ICPMatcher();
    ICPMatcher(ICPMatcher &&other) noexcept;
    ICPMatcher(const ICPMatcher &) = delete;
    ICPMatcher &operator=(const ICPMatcher &) = delete;

    void setRef(const PCLPointCloudPtr &ref);
    void setTarget(const PCLPointCloudPtr &target);
    /// unit normals of the target, one xyz(w) record per target point (point-to-plane only)
    void setTargetNormals(const PCLPointCloudPtr &normals);

    /// Extension for scan-to-map batches (not in the reference): finish this matcher's target now
    /// (upload + search structure) so that other matchers on the same GPU can match against it ...
    void buildTarget();
    /// ... and make this matcher use `owner`'s target, read-only, instead of one of its own
    /// (nullptr detaches).  The owner must outlive its sharers; full resolution (res <= 0) only.
    void shareTarget(ICPMatcher *owner);
    /// CUDA ordinal this matcher runs on (WAVE_MATCHING_DEVICE)
    int device() const { return device_; }

    /// blocks until finished; false if ICP did not converge
    bool match();
    /// information matrix of the last match (see src/host/icp.cpp for the fall-through semantics)
    void estimateInfo();

    ICPMatcherParams params;

 private:
    wavecu_icp *handle = nullptr;
    PCLPointCloudPtr ref, target;
    int device_ = 0;
};

}  // namespace wave

#endif  // WAVE_MATCHING_ICP_HPP

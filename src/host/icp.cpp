// wave::ICPMatcher over the C ABI (include/wavecu.h).  Reference behaviour reproduced, with the
// file:line each piece corresponds to in wave_matching/:
//   params from YAML          src/icp.cpp:6-30   (fit_eps is NOT read from the file, :9-16)
//   constructor plumbing      src/icp.cpp:32-51
//   setRef / setTarget        src/icp.cpp:67-73  (the matcher keeps the caller's pointers)
//   match()                   src/icp.cpp:75-133 (branches on res / multiscale_steps; false on
//                             non-convergence, never throws for it)
//   estimateInfo()            src/icp.cpp:135-142
#include "wave/matching/icp.hpp"

#include <atomic>
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "wavecu.h"

namespace wave {

namespace {

[[noreturn]] void fail(const char *what) {
    throw std::runtime_error(std::string(what) + ": " + wavecu_last_error());
}

wavecu_icp_params to_c(const ICPMatcherParams &p) {
    wavecu_icp_params c;
    c.max_corr = p.max_corr;
    c.max_iter = p.max_iter;
    c.t_eps = p.t_eps;
    c.fit_eps = p.fit_eps;
    c.lidar_ang_covar = p.lidar_ang_covar;
    c.lidar_lin_covar = p.lidar_lin_covar;
    c.multiscale_steps = p.multiscale_steps;
    c.res = p.res;
    c.covar_estimator = static_cast<int>(p.covar_estimator);
    c.estimator = p.estimator;
    return c;
}

}  // namespace

// WAVE_MATCHING_DEVICE=<n> pins every matcher to device n; WAVE_MATCHING_DEVICE=all spreads
// successive matchers (of all three kinds) round-robin over the visible devices - MultiMatcher on
// the 8-GPU box; unset: device 0.
int pick_matcher_device() {
    static std::atomic<int> next{0};
    const char *env = std::getenv("WAVE_MATCHING_DEVICE");
    if (!env) return 0;
    if (std::string(env) == "all") {
        const int n = wavecu_device_count();
        return n > 0 ? next.fetch_add(1) % n : 0;
    }
    return std::atoi(env);
}

ICPMatcherParams::ICPMatcherParams(const std::string &config_path) {
    ConfigParser parser;
    int covar_est_temp = 0;
    parser.addParam("max_corr", &this->max_corr);
    parser.addParam("max_iter", &this->max_iter);
    parser.addParam("t_eps", &this->t_eps);
    parser.addParam("lidar_ang_covar", &this->lidar_ang_covar);
    parser.addParam("lidar_lin_covar", &this->lidar_lin_covar);
    parser.addParam("covar_estimator", &covar_est_temp);
    parser.addParam("res", &this->res);
    parser.addParam("multiscale_steps", &this->multiscale_steps);
    if (parser.load(config_path) != ConfigStatus::OK) {
        throw std::runtime_error{"Failed to Load Matcher Config"};
    }
    if (covar_est_temp >= covar_method::LUM && covar_est_temp <= covar_method::LUMold) {
        this->covar_estimator = static_cast<covar_method>(covar_est_temp);
    } else {
        LOG_ERROR("Invalid covariance estimate method, using LUM");
        this->covar_estimator = covar_method::LUM;
    }
}

ICPMatcher::ICPMatcher(ICPMatcherParams params1) : params(params1) {
    this->resolution = this->params.res;
    const wavecu_icp_params c = to_c(this->params);
    this->device_ = pick_matcher_device();
    if (wavecu_icp_create(&c, this->device_, nullptr, &this->handle) != WAVECU_OK) fail("wavecu_icp_create");
}

ICPMatcher::ICPMatcher(ICPMatcher &&other) noexcept
    : Matcher<PCLPointCloudPtr>(other), params(other.params), handle(other.handle), ref(other.ref),
      target(other.target), device_(other.device_) {
    other.handle = nullptr;
}

ICPMatcher::~ICPMatcher() {
    if (this->handle) wavecu_icp_destroy(this->handle);
}

void ICPMatcher::setRef(const PCLPointCloudPtr &ref) {
    this->ref = ref;
    const float *data = ref && !ref->points.empty() ? &ref->points[0].x : nullptr;
    if (wavecu_icp_set_source(this->handle, data, ref ? ref->points.size() : 0) != WAVECU_OK)
        fail("wavecu_icp_set_source");
}

void ICPMatcher::setTarget(const PCLPointCloudPtr &target) {
    this->target = target;
    const float *data = target && !target->points.empty() ? &target->points[0].x : nullptr;
    if (wavecu_icp_set_target(this->handle, data, target ? target->points.size() : 0) != WAVECU_OK)
        fail("wavecu_icp_set_target");
}

void ICPMatcher::setTargetNormals(const PCLPointCloudPtr &normals) {
    const float *data = normals && !normals->points.empty() ? &normals->points[0].x : nullptr;
    if (wavecu_icp_set_target_normals(this->handle, data, normals ? normals->points.size() : 0) != WAVECU_OK)
        fail("wavecu_icp_set_target_normals");
}

void ICPMatcher::buildTarget() {
    if (wavecu_icp_build_target(this->handle) != WAVECU_OK) fail("wavecu_icp_build_target");
}

void ICPMatcher::shareTarget(ICPMatcher *owner) {
    if (wavecu_icp_share_target(this->handle, owner ? owner->handle : nullptr) != WAVECU_OK)
        fail("wavecu_icp_share_target");
}

bool ICPMatcher::match() {
    // `params` is a public, mutable member in the reference (the tests edit it before matching)
    const wavecu_icp_params c = to_c(this->params);
    if (wavecu_icp_set_params(this->handle, &c) != WAVECU_OK) fail("wavecu_icp_set_params");
    this->resolution = this->params.res;
    double T[16];
    int converged = 0, iterations = 0;
    if (wavecu_icp_match(this->handle, T, &converged, &iterations) != WAVECU_OK) fail("wavecu_icp_match");
    if (!converged) return false;
    for (int r = 0; r < 4; ++r)
        for (int col = 0; col < 4; ++col) this->result.matrix()(r, col) = T[4 * r + col];
    return true;
}

void ICPMatcher::estimateInfo() {
    // The reference's switch has no `break`s (src/icp.cpp:135-142): LUM runs estimateLUM, then
    // estimateCensi, then estimateLUMold; CENSI runs the last two; LUMold only itself.  Whatever
    // the setting, `information` therefore ends as estimateLUMold's result - that observable
    // outcome is what is produced here (estimateLUM first, so its early-return identity is
    // overwritten exactly as in the reference).
    double info[36];
    auto store = [&] {
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c < 6; ++c) this->information(r, c) = info[6 * r + c];
    };
    switch (this->params.covar_estimator) {
        case ICPMatcherParams::covar_method::LUM:
            if (wavecu_icp_info(this->handle, WAVECU_INFO_LUM, info) != WAVECU_OK) fail("wavecu_icp_info(LUM)");
            store();
            // fall through
        case ICPMatcherParams::covar_method::CENSI:
            // estimateCensi's matrix (wavecu_icp_info(WAVECU_INFO_CENSI)) would be overwritten by the
            // next case before anyone can read it, so it is not computed here
        case ICPMatcherParams::covar_method::LUMold:
            if (wavecu_icp_info(this->handle, WAVECU_INFO_LUMOLD, info) != WAVECU_OK) fail("wavecu_icp_info(LUMold)");
            store();
            // fall through
        default: return;
    }
}

}  // namespace wave

// wave::NDTMatcher over the C ABI.  Reference behaviour reproduced (wave_matching/src/ndt.cpp):
//   params from YAML      :6-16  (keys step_size, max_iter, t_eps, res - all required)
//   constructor           :18-34 (res below min_res -> LOG_ERROR and min_res)
//   setRef / setTarget    :48-56
//   match()               :58-65 (false when NDT reports non-convergence)
#include "wave/matching/ndt.hpp"

#include <cstdlib>
#include <stdexcept>
#include <string>

#include "wavecu.h"

namespace wave {

int pick_matcher_device();  // src/host/icp.cpp

namespace {
[[noreturn]] void fail(const char *what) { throw std::runtime_error(std::string(what) + ": " + wavecu_last_error()); }

wavecu_ndt_params to_c(const NDTMatcherParams &p) {
    wavecu_ndt_params c;
    c.step_size = p.step_size;
    c.max_iter = p.max_iter;
    c.t_eps = p.t_eps;
    c.res = p.res;
    // the reference's params carry no such field: PCL >= 1.9's line search unless the caller asks for
    // PCL 1.8's skipped one (wavecu.h)
    const char *ls = std::getenv("WAVE_MATCHING_NDT_PCL18");
    c.line_search = (ls && *ls && *ls != '0') ? WAVECU_NDT_LS_PCL18 : WAVECU_NDT_LS_MORE_THUENTE;
    return c;
}
}  // namespace

NDTMatcherParams::NDTMatcherParams(const std::string &config_path) {
    ConfigParser parser;
    parser.addParam("step_size", &this->step_size);
    parser.addParam("max_iter", &this->max_iter);
    parser.addParam("t_eps", &this->t_eps);
    parser.addParam("res", &this->res);
    if (parser.load(config_path) != ConfigStatus::OK) {
        throw std::runtime_error{"Failed to Load Matcher Config"};
    }
}

NDTMatcher::NDTMatcher(NDTMatcherParams params1) {
    this->params.step_size = params1.step_size;
    this->params.max_iter = params1.max_iter;
    this->params.t_eps = params1.t_eps;
    this->params.res = params1.res;
    if (this->params.res < this->params.min_res) {
        LOG_ERROR("Invalid resolution given, using minimum");
        this->params.res = this->params.min_res;
    }
    this->resolution = this->params.res;
    const wavecu_ndt_params c = to_c(this->params);
    const int device = pick_matcher_device();  // WAVE_MATCHING_DEVICE, shared by the three matchers (icp.cpp)
    if (wavecu_ndt_create(&c, device, nullptr, &this->handle) != WAVECU_OK) fail("wavecu_ndt_create");
}

NDTMatcher::NDTMatcher(NDTMatcher &&other) noexcept
    : Matcher<PCLPointCloudPtr>(other), handle(other.handle), ref(other.ref), target(other.target) {
    this->params.step_size = other.params.step_size;
    this->params.max_iter = other.params.max_iter;
    this->params.t_eps = other.params.t_eps;
    this->params.res = other.params.res;
    other.handle = nullptr;
}

NDTMatcher::~NDTMatcher() {
    if (this->handle) wavecu_ndt_destroy(this->handle);
}

void NDTMatcher::setRef(const PCLPointCloudPtr &ref) {
    this->ref = ref;
    const float *data = ref && !ref->points.empty() ? &ref->points[0].x : nullptr;
    if (wavecu_ndt_set_source(this->handle, data, ref ? ref->points.size() : 0) != WAVECU_OK)
        fail("wavecu_ndt_set_source");
}

void NDTMatcher::setTarget(const PCLPointCloudPtr &target) {
    this->target = target;
    const float *data = target && !target->points.empty() ? &target->points[0].x : nullptr;
    if (wavecu_ndt_set_target(this->handle, data, target ? target->points.size() : 0) != WAVECU_OK)
        fail("wavecu_ndt_set_target");
}

bool NDTMatcher::match() {
    double T[16];
    int converged = 0, iterations = 0;
    if (wavecu_ndt_match(this->handle, T, &converged, &iterations) != WAVECU_OK) fail("wavecu_ndt_match");
    if (!converged) return false;
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) this->result.matrix()(r, c) = T[4 * r + c];
    return true;
}

}  // namespace wave

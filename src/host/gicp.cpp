// wave::GICPMatcher over the C ABI.  Reference behaviour reproduced (wave_matching/src/gicp.cpp):
//   params from YAML   :6-18  - the reference parses corr_rand / max_iter / r_eps / fit_eps into
//                      LOCAL variables that shadow the members, so the file's values are read,
//                      validated (a missing key still throws) and then dropped; `res` is never
//                      parsed.  Kept: a drop-in must not start honouring a file the reference ignores.
//   constructor        :20-35 (resolution = res if res > 0 else -1)
//   setRef / setTarget :37-55 (voxel filter when resolution > 0)
//   match()            :57-64
#include "wave/matching/gicp.hpp"

#include <cstdlib>
#include <stdexcept>
#include <string>

#include "wavecu.h"

namespace wave {

int pick_matcher_device();  // src/host/icp.cpp

namespace {
[[noreturn]] void fail(const char *what) { throw std::runtime_error(std::string(what) + ": " + wavecu_last_error()); }
}  // namespace

GICPMatcherParams::GICPMatcherParams(const std::string &config_path) {
    ConfigParser parser;
    double r_eps_file = 1e-8, fit_eps_file = 1e-2;
    int corr_rand_file = 10, max_iter_file = 100;
    parser.addParam("corr_rand", &corr_rand_file);
    parser.addParam("max_iter", &max_iter_file);
    parser.addParam("r_eps", &r_eps_file);
    parser.addParam("fit_eps", &fit_eps_file);
    if (parser.load(config_path) != ConfigStatus::OK) {
        throw std::runtime_error{"Failed to Load Matcher Config"};
    }
}

GICPMatcher::GICPMatcher(GICPMatcherParams params1) : params(params1) {
    this->resolution = (this->params.res > 0) ? this->params.res : -1;
    wavecu_gicp_params c;
    c.corr_rand = this->params.corr_rand;
    c.max_iter = this->params.max_iter;
    c.r_eps = this->params.r_eps;
    c.fit_eps = this->params.fit_eps;
    c.res = this->resolution;
    const int device = pick_matcher_device();  // WAVE_MATCHING_DEVICE, shared by the three matchers (icp.cpp)
    if (wavecu_gicp_create(&c, device, nullptr, &this->handle) != WAVECU_OK) fail("wavecu_gicp_create");
}

GICPMatcher::GICPMatcher(GICPMatcher &&other) noexcept
    : Matcher<PCLPointCloudPtr>(other), handle(other.handle), ref(other.ref), target(other.target),
      params(other.params) {
    other.handle = nullptr;
}

GICPMatcher::~GICPMatcher() {
    if (this->handle) wavecu_gicp_destroy(this->handle);
}

void GICPMatcher::setRef(const PCLPointCloudPtr &ref) {
    this->ref = ref;
    const float *data = ref && !ref->points.empty() ? &ref->points[0].x : nullptr;
    if (wavecu_gicp_set_source(this->handle, data, ref ? ref->points.size() : 0) != WAVECU_OK)
        fail("wavecu_gicp_set_source");
}

void GICPMatcher::setTarget(const PCLPointCloudPtr &target) {
    this->target = target;
    const float *data = target && !target->points.empty() ? &target->points[0].x : nullptr;
    if (wavecu_gicp_set_target(this->handle, data, target ? target->points.size() : 0) != WAVECU_OK)
        fail("wavecu_gicp_set_target");
}

bool GICPMatcher::match() {
    double T[16];
    int converged = 0, iterations = 0;
    if (wavecu_gicp_match(this->handle, T, &converged, &iterations) != WAVECU_OK) fail("wavecu_gicp_match");
    if (!converged) return false;
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) this->result.matrix()(r, c) = T[4 * r + c];
    return true;
}

}  // namespace wave

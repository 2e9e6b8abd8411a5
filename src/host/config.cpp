// Flat-YAML reader behind wave::ConfigParser (see include/wave/utils/config.hpp).  Behaviour kept
// from the reference (wave_utils/src/config.cpp:38-99): a missing file or the first missing
// non-optional key makes load() fail, which the params constructors turn into
// std::runtime_error{"Failed to Load Matcher Config"} (wave_matching/src/icp.cpp:18-20).
#include "wave/utils/config.hpp"

#include <cstdlib>
#include <fstream>
#include <sstream>

namespace wave {

namespace {

std::string trim(const std::string &s) {
    const size_t b = s.find_first_not_of(" \t\r\n");
    if (b == std::string::npos) return "";
    const size_t e = s.find_last_not_of(" \t\r\n");
    return s.substr(b, e - b + 1);
}

}  // namespace

ConfigStatus ConfigParser::load(const std::string &config_file) {
    std::ifstream in(config_file);
    if (!in.good()) return ConfigStatus::FileNotFound;
    std::map<std::string, std::string> kv;
    std::string line;
    while (std::getline(in, line)) {
        const size_t hash = line.find('#');
        if (hash != std::string::npos) line = line.substr(0, hash);
        const size_t colon = line.find(':');
        if (colon == std::string::npos) continue;
        const std::string key = trim(line.substr(0, colon)), val = trim(line.substr(colon + 1));
        if (!key.empty() && !val.empty()) kv[key] = val;
    }
    ConfigStatus status = ConfigStatus::OK;
    for (const Param &p : params_) {
        const auto it = kv.find(p.key);
        if (it == kv.end()) {
            if (p.optional) {
                status = ConfigStatus::MissingOptionalKey;
                continue;
            }
            return ConfigStatus::KeyError;
        }
        const std::string &v = it->second;
        char *end = nullptr;
        switch (p.kind) {
            case Kind::Int: {
                const long x = std::strtol(v.c_str(), &end, 10);
                if (end == v.c_str()) return ConfigStatus::ConversionError;
                *static_cast<int *>(p.out) = static_cast<int>(x);
                break;
            }
            case Kind::Float:
            case Kind::Double: {
                const double x = std::strtod(v.c_str(), &end);
                if (end == v.c_str()) return ConfigStatus::ConversionError;
                if (p.kind == Kind::Float) *static_cast<float *>(p.out) = static_cast<float>(x);
                else *static_cast<double *>(p.out) = x;
                break;
            }
            case Kind::Bool: *static_cast<bool *>(p.out) = (v == "true" || v == "True" || v == "1"); break;
            case Kind::String: *static_cast<std::string *>(p.out) = v; break;
        }
    }
    return status;
}

}  // namespace wave

// The slice of wave::ConfigParser (wave_utils/include/wave/utils/config.hpp:108-149,
// src/config.cpp:38-99) that the matcher params constructors use: register (key, destination)
// pairs, load a YAML file, fill every parameter, and fail on the first missing key.  The matcher
// configs are flat `key: value  # comment` files (wave_matching/tests/config/*.yaml), which is all
// this reader understands - it is not a YAML library.
#ifndef WAVE_UTILS_CONFIG_HPP
#define WAVE_UTILS_CONFIG_HPP

#include <map>
#include <string>
#include <vector>

namespace wave {

enum class ConfigStatus { OK = 0, MissingOptionalKey = 1, FileNotFound = -1, KeyError = -2, ConversionError = -3 };

class ConfigParser {
 public:
    void addParam(const std::string &key, int *out, bool optional = false) { add(key, Kind::Int, out, optional); }
    void addParam(const std::string &key, float *out, bool optional = false) { add(key, Kind::Float, out, optional); }
    void addParam(const std::string &key, double *out, bool optional = false) { add(key, Kind::Double, out, optional); }
    void addParam(const std::string &key, bool *out, bool optional = false) { add(key, Kind::Bool, out, optional); }
    void addParam(const std::string &key, std::string *out, bool optional = false) {
        add(key, Kind::String, out, optional);
    }
    ConfigStatus load(const std::string &config_file);

 private:
    enum class Kind { Int, Float, Double, Bool, String };
    struct Param {
        std::string key;
        Kind kind;
        void *out;
        bool optional;
    };
    void add(const std::string &key, Kind kind, void *out, bool optional) {
        params_.push_back(Param{key, kind, out, optional});
    }
    std::vector<Param> params_;
};

}  // namespace wave
#endif  // WAVE_UTILS_CONFIG_HPP

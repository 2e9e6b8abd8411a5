// LOG_ERROR / LOG_INFO as wave_utils/include/wave/utils/log.hpp:24-28 (stderr / stdout).
#ifndef WAVE_UTILS_LOG_HPP
#define WAVE_UTILS_LOG_HPP
#include <cstdio>
#ifndef LOG_ERROR
#define LOG_ERROR(M, ...) std::fprintf(stderr, "[ERROR] [%s:%d] " M "\n", __FILE__, __LINE__, ##__VA_ARGS__)
#endif
#ifndef LOG_INFO
#define LOG_INFO(M, ...) std::fprintf(stdout, "[INFO] " M "\n", ##__VA_ARGS__)
#endif
#endif  // WAVE_UTILS_LOG_HPP

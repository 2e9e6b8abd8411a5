// wave::PCLPointCloudPtr - the cloud handle every matcher takes (reference:
// wave_matching/include/wave/matching/pcl_common.hpp:22).
#ifndef WAVE_MATCHING_PCL_COMMON_HPP
#define WAVE_MATCHING_PCL_COMMON_HPP

#if defined(__has_include)
#if __has_include(<pcl/point_cloud.h>) && __has_include(<pcl/point_types.h>)
#define WAVE_HAVE_PCL 1
#endif
#endif

#ifdef WAVE_HAVE_PCL
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#else
#include "wave/matching/compat/pcl_compat.hpp"
#endif

namespace wave {

typedef pcl::PointCloud<pcl::PointXYZ>::Ptr PCLPointCloudPtr;

}  // namespace wave

#endif  // WAVE_MATCHING_PCL_COMMON_HPP

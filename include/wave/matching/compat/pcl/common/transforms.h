// Stand-in for <pcl/common/transforms.h> (pcl::transformPointCloud) when PCL is not installed.
#include "wave/matching/compat/pcl_compat.hpp"

// Stand-in for <pcl/io/pcd_io.h> (pcl::io::loadPCDFile) when PCL is not installed: add
// -I<repo>/include/wave/matching/compat so that the reference tests' own include line resolves
// (wave_matching/tests/icp_tests.cpp:1).
#include "wave/matching/compat/pcl_compat.hpp"

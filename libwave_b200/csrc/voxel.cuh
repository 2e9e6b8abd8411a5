// pcl::VoxelGrid<pcl::PointXYZ>::filter on the device (reference call sites
// wave_matching/src/icp.cpp:81-90,106-113 and src/gicp.cpp:39-40,49-50) and
// pcl::transformPointCloud(cloud, cloud, Affine3d) (src/icp.cpp:84-86).
#pragma once
#include "common.cuh"

namespace wavecu {

// voxel indexing of pcl::VoxelGrid / VoxelGridCovariance: ijk = (int)(floor(p * inv) - (float) min_b)
struct GridDesc {
    float inv;
    int min_b[3];
    int div_b[3];
    int mul[3];
};

struct VoxelWork {
    int device = 0;
    cudaStream_t stream = nullptr;
    unsigned *d_keys = nullptr, *d_keys_alt = nullptr, *d_vals = nullptr, *d_vals_alt = nullptr;
    int *d_pos = nullptr;
    int *d_long = nullptr;       // [0]: number of long voxel runs, [1..]: their first elements (filter())
    size_t long_cap = 0;
    unsigned *d_bbox = nullptr;
    void *d_tmp = nullptr;
    size_t tmp_bytes = 0, cap = 0;
    unsigned *h_bbox = nullptr;  // pinned, 8 words
    int *h_count = nullptr;      // pinned
    long long launches = 0;

    GridDesc grid{};             // of the last prepare()
    size_t n_sorted = 0;         // points of the last prepare()
    int n_voxels = 0;            // occupied voxels of the last prepare()

    int reserve(size_t n);
    // bbox + voxel keys + stable sort + head flags + exclusive scan.  Afterwards d_keys/d_vals hold
    // the (voxel index, cloud index) pairs in ascending voxel order (cloud order inside a voxel) and
    // d_pos[i] the output slot of the voxel that sorted element i belongs to / starts.
    // *status: 1 ok, 0 grid would overflow int32 (PCL passes the input through), -1 no finite point.
    int prepare(const float4 *d_in, size_t n, float leaf, int *status);
    // out must have room for n points; *n_out receives the output size; *filtered = 0 when the
    // grid would overflow int32 and PCL copies the input through unchanged.
    int filter(const float4 *d_in, size_t n, float leaf, float4 *d_out, size_t *n_out, int *filtered);
    void release();
};

// cloud <- (float)(T * cloud) with T a row-major 3x4 double affine, evaluated left to right in fp64
int affine3d_inplace(float4 *d_cloud, size_t n, const double T[12], cudaStream_t stream);

// shared with index.cu
void launch_bbox(const float4 *d_pts, size_t n, unsigned *d_bbox8, cudaStream_t stream);

}  // namespace wavecu

// Tiled exact 1-NN correspondence kernel (SURVEY.md 8(a) A4 + A6; replaces the per-iteration search
// of pcl::IterativeClosestPoint::computeTransformation, reference call site
// wave_matching/src/icp.cpp:126).  Same results as correspond_kernel (icp_kernels.cuh) bit for bit -
// exact nearest neighbour under flann::L2_Simple, lowest original index among exact ties - by a
// different route:
//
//   * a CTA owns a tile of kTileQ Morton-consecutive source points (one per thread).  It moves them by
//     the incremental transform (A.3.2), takes the tile's bounding box and grows it by a margin M
//     chosen from the queries' own search radii (the distance to each query's match of the previous
//     iteration), and finds every 64-point run of the Morton-sorted target whose box meets the grown
//     box by a cooperative descent of the box pyramid (index.cuh, BoxLevels);
//   * those runs - points and their 8-point boxes, contiguous slices of the sorted arrays - are staged
//     into shared memory with cp.async.bulk, completion counted on an mbarrier;
//   * every lane then searches the staged set for its own query: a uniform loop over the runs that can
//     matter to its warp (all lanes test the same broadcast box against their own bound), then its
//     own hits: the eight 8-point boxes of a run, then points - all shared-memory reads, no dependent
//     global fetch, no stack;
//   * a result is exact when the ball of its distance lies inside the region whose points were all
//     staged and offered to the lane; the few queries for which that cannot be shown (ball reaching
//     beyond the staged margin, tile too wide for the shared-memory budget) keep their candidate as a
//     bound and are appended to a list that tile_fallback_kernel finishes with the LBVH walk.
#pragma once
#include "common.cuh"
#include "icp_kernels.cuh"
#include "index.cuh"

namespace wavecu {

#ifndef WCU_TILE_Q
#define WCU_TILE_Q 256
#endif
#ifndef WCU_TILE_CAP
#define WCU_TILE_CAP 48
#endif
constexpr int kTileQ = WCU_TILE_Q;        // queries (= threads) per CTA
constexpr int kTileCap = WCU_TILE_CAP;    // staged 64-point runs per tile (<= 64: per-lane hit masks are 64 bits)
constexpr int kTileList = 256;            // frontier capacity of the pyramid descent
constexpr int kTileAttempts = 6;
static_assert(kTileCap <= 64 && kBoxTop <= kTileQ, "tile geometry");

struct TileFallback {
    int *count;      // queries handed to the walk in this iteration (reset by the solve kernel)
    int *list;       // their sorted source positions
};

struct TileSmem {
    float4 pts[kTileCap * 64];
    float4 l0[kTileCap * 16];     // 8 boxes (lo, hi) per run
    float4 l1[kTileCap * 2];      // the run's own box
    int list[2][kTileList];
    unsigned long long mbar;
    float T[12];
    float red[kTileQ / 32][12];
    float tlo[3], thi[3], M;
    int n[2], overflow, n_valid;
};

// ---- mbarrier / bulk-copy primitives (PTX ISA: mbarrier, cp.async.bulk) ------------------------------
__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void *dst_smem, const void *src_gmem, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned phase) {
    unsigned done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_addr(bar)), "r"(phase)
            : "memory");
    } while (!done);
}

__device__ __forceinline__ bool boxes_meet(const float4 &lo, const float4 &hi, const float glo[3], const float ghi[3]) {
    return lo.x <= ghi[0] && hi.x >= glo[0] && lo.y <= ghi[1] && hi.y >= glo[1] && lo.z <= ghi[2] && hi.z >= glo[2];
}

// search radius of a bound: larger than the true radius of the fp32 distance ball, and by enough that
// a ball of exactly this radius still passes the certification test below (0.99999 * 1.00002 > 1)
__device__ __forceinline__ float bound_radius(float d2, float qx, float qy, float qz) {
    return sqrtf(d2) * 1.00002f + fmaxf(fmaxf(fabsf(qx), fabsf(qy)), fabsf(qz)) * 1e-6f + 1e-30f;
}

__global__ void __launch_bounds__(kTileQ) correspond_tile_kernel(const __grid_constant__ IterArgs a,
                                                                 const __grid_constant__ BoxLevels bl, TileFallback fb) {
    if (a.st->done) return;
    extern __shared__ __align__(128) unsigned char tile_smem_raw[];
    TileSmem &sm = *reinterpret_cast<TileSmem *>(tile_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 12) sm.T[tid] = a.st->T_inc[tid];
    if (tid == 0) mbar_init(&sm.mbar, 1);
    __syncthreads();

    // ---- A.3.2: the working cloud moves in place; warm start from the previous match ----
    const int s = blockIdx.x * kTileQ + tid;
    bool valid = false;
    float x = 0.f, y = 0.f, z = 0.f;
    if (s < a.n_src) {
        const float4 c = a.cur[s];
        if (finite3(c.x, c.y, c.z)) {
            x = xform_row(sm.T + 0, c.x, c.y, c.z);
            y = xform_row(sm.T + 4, c.x, c.y, c.z);
            z = xform_row(sm.T + 8, c.x, c.y, c.z);
            a.cur[s] = make_float4(x, y, z, c.w);
            valid = true;
        } else {  // pads and non-finite source points take no part
            a.nn_pos[s] = -1;
            a.nn_idx[s] = -1;
            a.nn_d2[s] = INFINITY;
        }
    }
    const float thr = a.mc->thr;
    unsigned long long best_key = nn_key(thr, 0x7fffffffu);
    int best_pos = -1;
    int best_slot = -1;   // staged slot of an improvement found in shared memory (converted to a position at the end)
    float r = -1.0f;  // search radius of the warm bound; < 0: no candidate
    if (valid) {
        const int warm = a.st->iter > 0 ? a.nn_pos[s] : -1;   // the first iteration has no previous match
        if (warm >= 0) {
            const float4 p = __ldg(a.tgt + warm);
            const float d = l2_simple(x, y, z, p.x, p.y, p.z);
            if (d <= thr) {
                best_key = nn_key(d, __float_as_uint(p.w));
                best_pos = warm;
                r = bound_radius(d, x, y, z);
            }
        }
    }

    // ---- tile box, margin statistics ----
    {
        float v[10];
        v[0] = valid ? x : INFINITY;  v[1] = valid ? y : INFINITY;  v[2] = valid ? z : INFINITY;
        v[3] = valid ? x : -INFINITY; v[4] = valid ? y : -INFINITY; v[5] = valid ? z : -INFINITY;
        v[6] = r;                             // max radius
        v[7] = r >= 0.0f ? r : 0.0f;          // sum of radii
        v[8] = r >= 0.0f ? 1.0f : 0.0f;       // warm count
        v[9] = valid ? 1.0f : 0.0f;           // valid count
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int k = 0; k < 3; ++k) v[k] = fminf(v[k], __shfl_xor_sync(0xffffffffu, v[k], o));
#pragma unroll
            for (int k = 3; k < 7; ++k) v[k] = fmaxf(v[k], __shfl_xor_sync(0xffffffffu, v[k], o));
#pragma unroll
            for (int k = 7; k < 10; ++k) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
        }
        if (lane == 0)
#pragma unroll
            for (int k = 0; k < 10; ++k) sm.red[warp][k] = v[k];
    }
    __syncthreads();
    if (tid == 0) {
        float v[10];
        for (int k = 0; k < 10; ++k) v[k] = sm.red[0][k];
        for (int w = 1; w < kTileQ / 32; ++w) {
            for (int k = 0; k < 3; ++k) v[k] = fminf(v[k], sm.red[w][k]);
            for (int k = 3; k < 7; ++k) v[k] = fmaxf(v[k], sm.red[w][k]);
            for (int k = 7; k < 10; ++k) v[k] += sm.red[w][k];
        }
        for (int k = 0; k < 3; ++k) {
            sm.tlo[k] = v[k];
            sm.thi[k] = v[3 + k];
        }
        const float n_warm = v[8], n_valid = v[9];
        const float ext = fmaxf(fmaxf(v[3] - v[0], v[4] - v[1]), v[5] - v[2]);
        float M = 0.0f;
        // queries with a candidate: enough for (nearly) all of their balls, without letting a few wide
        // ones decide; queries without one: a first guess from the tile's own extent
        if (n_warm > 0.0f) M = fminf(v[6], 3.0f * (v[7] / n_warm) + 0.01f);
        if (n_valid > n_warm) M = fmaxf(M, fmaxf(0.6f * ext, 0.05f));
        sm.M = fminf(M, sqrtf(thr) * 1.0001f + 1e-3f);
        sm.n_valid = (int) n_valid;
    }
    __syncthreads();
    if (sm.n_valid == 0) return;

    // ---- which 64-point runs of the target meet the grown tile box: descent of the box pyramid ----
    float glo[3], ghi[3];
    float M = sm.M;
    int n_runs = 0, cur = 0;
    for (int attempt = 0;; ++attempt) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            glo[k] = sm.tlo[k] - M;
            ghi[k] = sm.thi[k] + M;
        }
        if (tid == 0) {
            sm.n[0] = 0;
            sm.n[1] = 0;
            sm.overflow = 0;
        }
        __syncthreads();
        const int top = bl.n_levels - 1;
        cur = 0;
        if (tid < bl.cnt[top]) {
            const float4 lo = __ldg(bl.lv[top] + 2 * tid), hi = __ldg(bl.lv[top] + 2 * tid + 1);
            if (boxes_meet(lo, hi, glo, ghi)) {
                const int i = atomicAdd(&sm.n[0], 1);
                if (i < kTileList) sm.list[0][i] = tid;
                else sm.overflow = 1;
            }
        }
        __syncthreads();
        for (int k = top - 1; k >= 1; --k) {
            const int n_par = min(sm.n[cur], kTileList);
            const float4 *lv = bl.lv[k];
            const int cnt = bl.cnt[k];
            for (int i = tid; i < n_par * 8; i += kTileQ) {
                const int child = sm.list[cur][i >> 3] * 8 + (i & 7);
                if (child < cnt) {
                    const float4 lo = __ldg(lv + 2 * child), hi = __ldg(lv + 2 * child + 1);
                    if (boxes_meet(lo, hi, glo, ghi)) {
                        const int j = atomicAdd(&sm.n[cur ^ 1], 1);
                        if (j < kTileList) sm.list[cur ^ 1][j] = child;
                        else sm.overflow = 1;
                    }
                }
            }
            __syncthreads();
            if (tid == 0) sm.n[cur] = 0;   // becomes the output list of the next level
            cur ^= 1;
            __syncthreads();
        }
        n_runs = sm.n[cur];
        const bool too_many = sm.overflow || n_runs > kTileCap;
        __syncthreads();
        if (!too_many) break;
        if (attempt + 1 >= kTileAttempts) {  // even a thin margin does not fit: everything goes to the walk
            n_runs = 0;
            M = -1.0f;
            break;
        }
        M *= 0.5f;
    }

    // ---- stage the runs: points (1 KB), their eight 8-point boxes (256 B), their own box (32 B) ----
    if (warp == 0) {
        if (lane == 0) mbar_arrive_expect_tx(&sm.mbar, (unsigned) n_runs * (1024u + 256u + 32u));
        __syncwarp();
        for (int b = lane; b < n_runs; b += 32) {
            const int run = sm.list[cur][b];
            bulk_load(&sm.pts[b * 64], a.tgt + (size_t) run * 64, 1024u, &sm.mbar);
            bulk_load(&sm.l0[b * 16], bl.lv[0] + (size_t) run * 16, 256u, &sm.mbar);
            bulk_load(&sm.l1[b * 2], bl.lv[1] + (size_t) run * 2, 32u, &sm.mbar);
        }
    }

    // ---- per warp: the region inside which this warp's lanes see every target point ----
    // W = (warp box grown by the warp's largest radius) ∩ G.  A target point inside W sits in a staged
    // run (its run's box meets G) and in an 8-point box that meets W, so the lane either measures it or
    // discards its box by the lane's own bound.  Everything outside W is farther from the query than
    // the distance e to W's nearest face.
    float e = 0.0f;          // certified radius of this lane
    float wlo[3], whi[3];
    {
        float mq = INFINITY;  // distance to the faces of G
        const float q[3] = {x, y, z};
#pragma unroll
        for (int k = 0; k < 3; ++k) mq = fminf(mq, fminf(q[k] - glo[k], ghi[k] - q[k]));
        float rl = !valid ? -INFINITY : (r >= 0.0f ? r : mq);   // a lane without a candidate looks as far as G allows
        float b[6] = {valid ? x : INFINITY, valid ? y : INFINITY, valid ? z : INFINITY,
                      valid ? x : -INFINITY, valid ? y : -INFINITY, valid ? z : -INFINITY};
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int k = 0; k < 3; ++k) b[k] = fminf(b[k], __shfl_xor_sync(0xffffffffu, b[k], o));
#pragma unroll
            for (int k = 3; k < 6; ++k) b[k] = fmaxf(b[k], __shfl_xor_sync(0xffffffffu, b[k], o));
            rl = fmaxf(rl, __shfl_xor_sync(0xffffffffu, rl, o));
        }
        e = INFINITY;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            wlo[k] = fmaxf(b[k] - rl, glo[k]);
            whi[k] = fminf(b[3 + k] + rl, ghi[k]);
            e = fminf(e, fminf(q[k] - wlo[k], whi[k] - q[k]));
        }
        if (M < 0.0f || !valid) e = 0.0f;
    }

    mbar_wait(&sm.mbar, 0);

    // ---- search the staged set ----
    // runs that meet W, as two warp-uniform masks
    unsigned run_mask[2] = {0u, 0u};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int b = h * 32 + lane;
        bool meet = false;
        if (b < n_runs) meet = boxes_meet(sm.l1[2 * b], sm.l1[2 * b + 1], wlo, whi);
        run_mask[h] = __ballot_sync(0xffffffffu, meet);
    }
    if (valid) {
        // pass 1 (uniform over the warp): which of those runs can hold something for this lane
        unsigned hit[2] = {0u, 0u};
        float bound = key_bound(best_key);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            unsigned m = run_mask[h];
            while (m) {
                const int j = __ffs(m) - 1;
                m &= m - 1;
                const int b = h * 32 + j;
                const float d = aabb_dist(x, y, z, sm.l1[2 * b], sm.l1[2 * b + 1]);
                if (d <= bound) hit[h] |= 1u << j;
            }
        }
        // pass 2: own hits - the run's eight 8-point boxes, then the points of those within the bound
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            unsigned m = hit[h];
            while (m) {
                const int j = __ffs(m) - 1;
                m &= m - 1;
                const int b = h * 32 + j;
                unsigned sub = 0u;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float d = aabb_dist(x, y, z, sm.l0[(b * 8 + c) * 2], sm.l0[(b * 8 + c) * 2 + 1]);
                    if (d <= bound) sub |= 1u << c;
                }
                while (sub) {
                    const int c = __ffs(sub) - 1;
                    sub &= sub - 1;
                    const float4 lo = sm.l0[(b * 8 + c) * 2], hi = sm.l0[(b * 8 + c) * 2 + 1];
                    if (aabb_dist(x, y, z, lo, hi) > bound) continue;   // the bound may have tightened meanwhile
                    const int slot0 = (b * 8 + c) * 8;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float4 p = sm.pts[slot0 + k];
                        const float d = l2_simple(x, y, z, p.x, p.y, p.z);
                        if (d <= bound) {   // most points fail this one compare; ties go through the full key
                            const unsigned long long key = nn_key(d, __float_as_uint(p.w));
                            if (key < best_key) {
                                best_key = key;
                                best_slot = slot0 + k;
                                bound = d;
                            }
                        }
                    }
                }
            }
        }
    }

    if (best_slot >= 0) best_pos = sm.list[cur][best_slot >> 6] * 64 + (best_slot & 63);

    // ---- results; queries whose ball leaves the certified region go to the walk with their bound ----
    const float ec = e * 0.99999f;
    const bool certified = valid && key_bound(best_key) < ec * ec;
    if (valid) {
        a.nn_pos[s] = best_pos;
        a.nn_idx[s] = best_pos >= 0 ? (int) (unsigned) best_key : -1;
        a.nn_d2[s] = key_bound(best_key);
    }
    const unsigned need = __ballot_sync(0xffffffffu, valid && !certified);
    if (need) {
        int base = 0;
        if (lane == 0) base = atomicAdd(fb.count, __popc(need));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (valid && !certified) fb.list[base + __popc(need & ((1u << lane) - 1u))] = s;
    }
}

// The queries the tiles could not certify: exact LBVH walk (index.cuh), warm-started from the candidate
// and bound the tile pass left in nn_pos / nn_d2.
__global__ void __launch_bounds__(kIterThreads, WCU_MINBLOCKS) tile_fallback_kernel(IterArgs a, TileFallback fb) {
    if (a.st->done) return;
    const int n = *fb.count;
    for (int i = blockIdx.x * kIterThreads + threadIdx.x; i < n; i += gridDim.x * kIterThreads) {
        const int s = fb.list[i];
        const float4 c = a.cur[s];
        float best = a.mc->thr;
        int best_idx = 0x7fffffff, best_pos = -1;
        const int warm = a.nn_pos[s];
        if (warm >= 0) {
            const float4 p = __ldg(a.tgt + warm);
            best = l2_simple(c.x, c.y, c.z, p.x, p.y, p.z);
            best_idx = __float_as_int(p.w);
            best_pos = warm;
        }
        nn_search_cells(c.x, c.y, c.z, a.ix, best, best_idx, best_pos);
        a.nn_pos[s] = best_pos;
        a.nn_idx[s] = best_pos >= 0 ? best_idx : -1;
        a.nn_d2[s] = best;
    }
}

}  // namespace wavecu

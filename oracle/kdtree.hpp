// ORACLE - test infrastructure only.  Nothing in the product path (libwave_b200/, include/,
// src/) may include, link or call this.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs use it.
//
// CPU restatement of the exact nearest-neighbour search the reference reaches through
// pcl::KdTreeFLANN (reference call sites: wave_matching/src/icp_pcl_functions.cpp:67-80 and every
// pcl::Registration::align() issued from wave_matching/src/icp.cpp:95,116,126, src/gicp.cpp:58,
// src/ndt.cpp:59).  PCL/FLANN are un-vendored system dependencies (PCL >= 1.8, FLANN 1.8.x,
// CMakeLists.txt:48) that are absent from /root/reference, so this restates the published
// algorithm: flann::KDTreeSingleIndex, leaf size 15, middle-split rule,
// points re-ordered into leaf order, incremental bounding-box distance pruning, eps = 0, and the
// L2_Simple<float> distance  r = ((dx*dx) + dy*dy) + dz*dz  evaluated in fp32 with separately
// rounded multiplies and adds.  PARITY UNPINNED: no PCL build is available to cross-check.
//
// Deliberate, documented deviation: among candidates at exactly equal fp32 distance FLANN keeps
// the first one its traversal meets; this oracle (and the GPU path) keep the lowest cloud index,
// which is order-independent.  Pruning bounds are evaluated conservatively so that the search is
// exact with respect to the fp32 distance above.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <vector>

namespace wo {

// L2_Simple<float> over 3 dims, fp32, no contraction (file is built with -ffp-contract=off).
static inline float l2_simple(const float *a, const float *b) {
    float r = 0.0f;
    for (int d = 0; d < 3; ++d) {
        const float diff = a[d] - b[d];
        r += diff * diff;
    }
    return r;
}

struct KnnSet {  // k best (distance, index), ties broken towards the lower index
    int k, count = 0;
    std::vector<float> d;
    std::vector<int> i;
    explicit KnnSet(int k_) : k(k_), d(k_, std::numeric_limits<float>::infinity()), i(k_, std::numeric_limits<int>::max()) {}
    float worst() const { return d[k - 1]; }
    int worst_idx() const { return i[k - 1]; }
    bool accepts(float dist, int idx) const { return dist < d[k - 1] || (dist == d[k - 1] && idx < i[k - 1]); }
    void add(float dist, int idx) {
        if (!accepts(dist, idx)) return;
        int p = k - 1;
        while (p > 0 && (d[p - 1] > dist || (d[p - 1] == dist && i[p - 1] > idx))) {
            d[p] = d[p - 1];
            i[p] = i[p - 1];
            --p;
        }
        d[p] = dist;
        i[p] = idx;
        if (count < k) ++count;
    }
};

struct Nn1Set {  // k = 1 without heap traffic
    float d0 = std::numeric_limits<float>::infinity();
    int i0 = std::numeric_limits<int>::max();
    float worst() const { return d0; }
    void add(float dist, int idx) {
        if (dist < d0 || (dist == d0 && idx < i0)) {
            d0 = dist;
            i0 = idx;
        }
    }
};

class KdTree {
  public:
    static constexpr int kLeafMax = 15;  // pcl::KdTreeFLANN: KDTreeSingleIndexParams(15)

    // pts: n records of `stride` floats, xyz first (pcl::PointXYZ: stride 4).  Non-finite points
    // are left out of the index, as pcl::KdTreeFLANN::convertCloudToArray does.
    KdTree(const float *pts, size_t n, int stride = 4) {
        ids_.reserve(n);
        for (size_t i = 0; i < n; ++i) {
            const float *p = pts + i * stride;
            if (std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2])) ids_.push_back((int) i);
        }
        const size_t m = ids_.size();
        src_ = pts;
        stride_ = stride;
        if (m == 0) return;
        float lo[3], hi[3];
        for (int d = 0; d < 3; ++d) lo[d] = hi[d] = at(ids_[0])[d];
        for (size_t i = 1; i < m; ++i)
            for (int d = 0; d < 3; ++d) {
                lo[d] = std::min(lo[d], at(ids_[i])[d]);
                hi[d] = std::max(hi[d], at(ids_[i])[d]);
            }
        for (int d = 0; d < 3; ++d) {
            root_lo_[d] = lo[d];
            root_hi_[d] = hi[d];
        }
        nodes_.reserve(2 * m / kLeafMax + 16);
        root_ = divide(0, (int) m, lo, hi);
        // reorder = true: copy points into leaf order
        data_.resize(3 * m);
        for (size_t i = 0; i < m; ++i)
            for (int d = 0; d < 3; ++d) data_[3 * i + d] = at(ids_[i])[d];
        src_ = nullptr;
    }

    size_t size() const { return ids_.size(); }

    // Exact k-NN (global, no radius).  Returns the number found (min(k, size)).
    int knn(const float *q, int k, int *idx, float *d2) const {
        KnnSet rs(k);
        if (!ids_.empty()) search_knn(q, rs);
        for (int j = 0; j < rs.count; ++j) {
            idx[j] = rs.i[j];
            d2[j] = rs.d[j];
        }
        return rs.count;
    }

    // Exact 1-NN; idx = -1 if the tree is empty.
    void nn1(const float *q, int *idx, float *d2) const {
        if (ids_.empty()) {
            *idx = -1;
            *d2 = std::numeric_limits<float>::infinity();
            return;
        }
        Nn1Set rs;
        search_knn(q, rs);
        *idx = rs.i0;
        *d2 = rs.d0;
    }

    // All points with fp32 squared distance <= r2 (pcl radiusSearch semantics: d <= radius),
    // sorted by (distance, index).
    void radius(const float *q, float r2, std::vector<std::pair<float, int>> &out) const {
        out.clear();
        if (ids_.empty()) return;
        double dists[3];
        const double m = root_dist(q, dists);
        radius_level(root_, q, (double) r2, m, dists, out);
        std::sort(out.begin(), out.end());
    }

  private:
    struct Node {
        int left, right;      // leaf: [left,right) into ids_/data_
        int child1, child2;   // internal: node indices; -1 for leaves
        int divfeat;
        float divlow, divhigh;
    };

    const float *at(int id) const { return src_ + (size_t) id * stride_; }

    int divide(int left, int right, float *lo, float *hi) {
        const int me = (int) nodes_.size();
        nodes_.push_back(Node{left, right, -1, -1, 0, 0.f, 0.f});
        if (right - left <= kLeafMax) {
            for (int d = 0; d < 3; ++d) lo[d] = hi[d] = at(ids_[left])[d];
            for (int i = left + 1; i < right; ++i)
                for (int d = 0; d < 3; ++d) {
                    lo[d] = std::min(lo[d], at(ids_[i])[d]);
                    hi[d] = std::max(hi[d], at(ids_[i])[d]);
                }
            return me;
        }
        int cutfeat, split;
        float cutval;
        middle_split(left, right - left, lo, hi, split, cutfeat, cutval);
        float llo[3], lhi[3], rlo[3], rhi[3];
        for (int d = 0; d < 3; ++d) {
            llo[d] = rlo[d] = lo[d];
            lhi[d] = rhi[d] = hi[d];
        }
        lhi[cutfeat] = cutval;
        rlo[cutfeat] = cutval;
        const int c1 = divide(left, left + split, llo, lhi);
        const int c2 = divide(left + split, right, rlo, rhi);
        Node &nd = nodes_[me];
        nd.child1 = c1;
        nd.child2 = c2;
        nd.divfeat = cutfeat;
        nd.divlow = lhi[cutfeat];
        nd.divhigh = rlo[cutfeat];
        for (int d = 0; d < 3; ++d) {
            lo[d] = std::min(llo[d], rlo[d]);
            hi[d] = std::max(lhi[d], rhi[d]);
        }
        return me;
    }

    // FLANN's middleSplit_: cut the widest-spread dimension at the bbox midpoint (clamped to the
    // data range) and balance the split index.
    void middle_split(int first, int count, const float *lo, const float *hi, int &index, int &cutfeat,
                      float &cutval) {
        const float eps = 0.00001f;
        float max_span = hi[0] - lo[0];
        for (int d = 1; d < 3; ++d) max_span = std::max(max_span, hi[d] - lo[d]);
        float max_spread = -1.f;
        cutfeat = 0;
        for (int d = 0; d < 3; ++d) {
            if (hi[d] - lo[d] > (1.f - eps) * max_span) {
                float mn, mx;
                minmax(first, count, d, mn, mx);
                if (mx - mn > max_spread) {
                    cutfeat = d;
                    max_spread = mx - mn;
                }
            }
        }
        const float split_val = (lo[cutfeat] + hi[cutfeat]) / 2;
        float mn, mx;
        minmax(first, count, cutfeat, mn, mx);
        cutval = split_val < mn ? mn : (split_val > mx ? mx : split_val);
        int lim1, lim2;
        plane_split(first, count, cutfeat, cutval, lim1, lim2);
        if (lim1 > count / 2)
            index = lim1;
        else if (lim2 < count / 2)
            index = lim2;
        else
            index = count / 2;
    }

    void minmax(int first, int count, int d, float &mn, float &mx) const {
        mn = mx = at(ids_[first])[d];
        for (int i = 1; i < count; ++i) {
            const float v = at(ids_[first + i])[d];
            mn = std::min(mn, v);
            mx = std::max(mx, v);
        }
    }

    void plane_split(int first, int count, int d, float cutval, int &lim1, int &lim2) {
        int *ind = ids_.data() + first;
        int left = 0, right = count - 1;
        for (;;) {
            while (left <= right && at(ind[left])[d] < cutval) ++left;
            while (left <= right && at(ind[right])[d] >= cutval) --right;
            if (left > right) break;
            std::swap(ind[left], ind[right]);
            ++left;
            --right;
        }
        lim1 = left;
        right = count - 1;
        for (;;) {
            while (left <= right && at(ind[left])[d] <= cutval) ++left;
            while (left <= right && at(ind[right])[d] > cutval) --right;
            if (left > right) break;
            std::swap(ind[left], ind[right]);
            ++left;
            --right;
        }
        lim2 = left;
    }

    // Lower bounds are kept in double and scaled by (1 - 2^-20) before they prune, so a bound can
    // never exceed the fp32 distance of a point inside the cell (see header comment).
    static constexpr double kSlack = 1.0 - 1.0 / 1048576.0;

    double root_dist(const float *q, double *dists) const {
        double s = 0;
        for (int d = 0; d < 3; ++d) {
            dists[d] = 0;
            if (q[d] < root_lo_[d]) dists[d] = ((double) q[d] - root_lo_[d]) * ((double) q[d] - root_lo_[d]);
            if (q[d] > root_hi_[d]) dists[d] = ((double) q[d] - root_hi_[d]) * ((double) q[d] - root_hi_[d]);
            s += dists[d];
        }
        return s;
    }

    template <class RS>
    void search_knn(const float *q, RS &rs) const {
        double dists[3];
        const double m = root_dist(q, dists);
        knn_level(root_, q, rs, m, dists);
    }

    template <class RS>
    void knn_level(int ni, const float *q, RS &rs, double mindist, double *dists) const {
        const Node &nd = nodes_[ni];
        if (nd.child1 < 0) {
            for (int i = nd.left; i < nd.right; ++i) rs.add(l2_simple(q, &data_[3 * (size_t) i]), ids_[i]);
            return;
        }
        const int f = nd.divfeat;
        const double val = q[f];
        const double diff1 = val - nd.divlow, diff2 = val - nd.divhigh;
        int best, other;
        double cut;
        if (diff1 + diff2 < 0) {
            best = nd.child1;
            other = nd.child2;
            cut = diff2 * diff2;
        } else {
            best = nd.child2;
            other = nd.child1;
            cut = diff1 * diff1;
        }
        knn_level(best, q, rs, mindist, dists);
        const double dst = dists[f];
        const double md = mindist + cut - dst;
        dists[f] = cut;
        if (md * kSlack <= (double) rs.worst()) knn_level(other, q, rs, md, dists);
        dists[f] = dst;
    }

    void radius_level(int ni, const float *q, double r2, double mindist, double *dists,
                      std::vector<std::pair<float, int>> &out) const {
        const Node &nd = nodes_[ni];
        if (nd.child1 < 0) {
            for (int i = nd.left; i < nd.right; ++i) {
                const float d = l2_simple(q, &data_[3 * (size_t) i]);
                if ((double) d <= r2) out.emplace_back(d, ids_[i]);
            }
            return;
        }
        const int f = nd.divfeat;
        const double val = q[f];
        const double diff1 = val - nd.divlow, diff2 = val - nd.divhigh;
        int best, other;
        double cut;
        if (diff1 + diff2 < 0) {
            best = nd.child1;
            other = nd.child2;
            cut = diff2 * diff2;
        } else {
            best = nd.child2;
            other = nd.child1;
            cut = diff1 * diff1;
        }
        radius_level(best, q, r2, mindist, dists, out);
        const double dst = dists[f];
        const double md = mindist + cut - dst;
        dists[f] = cut;
        if (md * kSlack <= r2) radius_level(other, q, r2, md, dists, out);
        dists[f] = dst;
    }

    const float *src_ = nullptr;
    int stride_ = 4;
    std::vector<int> ids_;      // leaf-ordered cloud indices
    std::vector<float> data_;   // leaf-ordered xyz
    std::vector<Node> nodes_;
    int root_ = -1;
    float root_lo_[3] = {0, 0, 0}, root_hi_[3] = {0, 0, 0};
};

}  // namespace wo

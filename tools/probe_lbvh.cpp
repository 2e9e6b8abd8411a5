// Scratch probe: radix-tree LBVH (split at the highest differing Morton bit), leaves of <= L points.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <vector>
struct P { float x, y, z, w; };
struct Node { float lo[3], hi[3]; int left, right; int b, e; };  // leaf if left < 0
static uint64_t expand21(uint64_t v) {
    v &= 0x1fffffull; v = (v | v << 32) & 0x1f00000000ffffull; v = (v | v << 16) & 0x1f0000ff0000ffull;
    v = (v | v << 8) & 0x100f00f00f00f00full; v = (v | v << 4) & 0x10c30c30c30c30c3ull; v = (v | v << 2) & 0x1249249249249249ull; return v; }
static std::vector<P> load(const char *fn) { FILE *f = fopen(fn, "rb"); fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET); std::vector<P> v(sz / 16); if (fread(v.data(), 16, v.size(), f)) {} fclose(f); return v; }
std::vector<P> tgt; std::vector<uint32_t> perm; std::vector<uint64_t> key; std::vector<Node> nodes; int L; int splitmode;
static float bdist(const Node &b, const float *q) { float s = 0; for (int d = 0; d < 3; ++d) { float e = std::max(std::max(b.lo[d] - q[d], q[d] - b.hi[d]), 0.f); s += e * e; } return s; }
int build(int b, int e, int bit) {
    int me = nodes.size(); nodes.push_back(Node());
    Node nd; for (int d = 0; d < 3; ++d) { nd.lo[d] = INFINITY; nd.hi[d] = -INFINITY; }
    for (int i = b; i < e; ++i) { const float *c = &tgt[perm[i]].x; for (int d = 0; d < 3; ++d) { nd.lo[d] = std::min(nd.lo[d], c[d]); nd.hi[d] = std::max(nd.hi[d], c[d]); } }
    nd.b = b; nd.e = e; nd.left = nd.right = -1;
    if (e - b > L) {
        int split = -1;
        if (splitmode == 0) {  // highest differing bit
            while (bit >= 0) { uint64_t m = 1ull << bit; if ((key[perm[b]] & m) != (key[perm[e - 1]] & m)) break; --bit; }
            if (bit < 0) split = (b + e) / 2;
            else { uint64_t m = 1ull << bit; int lo = b, hi = e - 1; while (lo < hi) { int mid = (lo + hi) / 2; if (key[perm[mid]] & m) hi = mid; else lo = mid + 1; } split = lo; --bit; }
        } else split = (b + e) / 2;
        int l = build(b, split, bit), r = build(split, e, bit); nd.left = l; nd.right = r;
    }
    nodes[me] = nd; return me;
}
int main(int argc, char **argv) {
    tgt = load(argv[1]); auto qry = load(argv[2]); L = atoi(argv[3]); splitmode = atoi(argv[4]); size_t n = tgt.size();
    float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
    for (auto &p : tgt) { const float *c = &p.x; for (int d = 0; d < 3; ++d) { lo[d] = std::min(lo[d], c[d]); hi[d] = std::max(hi[d], c[d]); } }
    float ext = std::max({hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]}); float sc = 2097151.f / ext;
    perm.resize(n); std::iota(perm.begin(), perm.end(), 0); key.resize(n);
    for (size_t i = 0; i < n; ++i) { const float *c = &tgt[i].x; uint64_t k = 0; for (int d = 0; d < 3; ++d) { uint64_t q = (uint64_t) std::min(std::max((c[d] - lo[d]) * sc, 0.f), 2097151.f); k |= expand21(q) << d; } key[i] = k; }
    std::sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
    build(0, n, 62);
    size_t nleaf = 0, leafpts = 0; for (auto &nd : nodes) if (nd.left < 0) { ++nleaf; leafpts += nd.e - nd.b; }
    double tot_leaf = 0, tot_node = 0, tot_pts = 0; std::vector<int> lv; size_t nq = qry.size(); size_t stride = std::max<size_t>(1, nq / 20000);
    for (size_t qi = 0; qi < nq; qi += stride) { const float *q = &qry[qi].x; float best = 9.f; int nl = 0, nn = 0, np = 0; int stack[128]; float sd[128]; int sp = 0; stack[sp] = 0; sd[sp++] = 0;
        while (sp) { --sp; int ni = stack[sp]; if (sd[sp] > best) continue; const Node &nd = nodes[ni];
            if (nd.left < 0) { ++nl; for (int i = nd.b; i < nd.e; ++i) { const float *c = &tgt[perm[i]].x; float dx = q[0] - c[0], dy = q[1] - c[1], dz = q[2] - c[2]; float d = dx * dx + dy * dy + dz * dz; ++np; if (d < best) best = d; } }
            else { ++nn; float d0 = bdist(nodes[nd.left], q), d1 = bdist(nodes[nd.right], q); if (d0 <= d1) { if (d1 <= best) { stack[sp] = nd.right; sd[sp++] = d1; } if (d0 <= best) { stack[sp] = nd.left; sd[sp++] = d0; } } else { if (d0 <= best) { stack[sp] = nd.left; sd[sp++] = d0; } if (d1 <= best) { stack[sp] = nd.right; sd[sp++] = d1; } } } }
        tot_leaf += nl; tot_node += nn; tot_pts += np; lv.push_back(nl); }
    // wide traversal: node = up to W children obtained by expanding internal children (largest box first) 
    for (int W : {2, 4, 8}) {
        double wv = 0, wb = 0, wl = 0, wp = 0, wpush = 0; size_t cntq = 0;
        for (size_t qi = 0; qi < nq; qi += stride) { const float *q = &qry[qi].x; float best = 9.f; ++cntq;
            struct E { int ni; float d; }; std::vector<E> st; st.push_back({0, 0.f});
            while (!st.empty()) { E e = st.back(); st.pop_back(); if (e.d > best) continue; const Node &nd = nodes[e.ni];
                if (nd.left < 0) { ++wl; for (int i = nd.b; i < nd.e; ++i) { const float *c = &tgt[perm[i]].x; float dx = q[0] - c[0], dy = q[1] - c[1], dz = q[2] - c[2]; float d = dx * dx + dy * dy + dz * dz; ++wp; if (d < best) best = d; } continue; }
                // gather children of wide node
                std::vector<int> ch = {nd.left, nd.right};
                while ((int) ch.size() < W) { int bi = -1; int bc = 0; for (size_t k = 0; k < ch.size(); ++k) { const Node &c = nodes[ch[k]]; if (c.left >= 0 && c.e - c.b > bc) { bc = c.e - c.b; bi = k; } } if (bi < 0) break; int c = ch[bi]; ch[bi] = nodes[c].left; ch.push_back(nodes[c].right); }
                ++wv; wb += ch.size(); std::vector<E> hits; for (int c : ch) { float d = bdist(nodes[c], q); if (d <= best) hits.push_back({c, d}); }
                std::sort(hits.begin(), hits.end(), [](const E &a, const E &b) { return a.d > b.d; }); wpush += hits.size() > 0 ? hits.size() - 1 : 0; for (auto &h : hits) st.push_back(h); } }
        printf("  W=%d: visits %.1f boxes %.1f leaves %.1f pts %.1f pushes %.1f  cost~%.0f\n", W, wv / cntq, wb / cntq, wl / cntq, wp / cntq, wpush / cntq, (wv * 12 + wb * 14 + wl * 10 + wp * 12 + wpush * 8) / cntq);
    }
    std::sort(lv.begin(), lv.end()); size_t m = lv.size();
    printf("lbvh split %d L %d: nodes %zu leaves %zu (avg %.1f pts) | leaves/query mean %.1f p50 %d p90 %d p99 %d | internal visits %.1f | pts %.1f\n", splitmode, L, nodes.size(), nleaf, (double) leafpts / nleaf, tot_leaf / m, lv[m / 2], lv[m * 9 / 10], lv[m * 99 / 100], tot_node / m, tot_pts / m);
}

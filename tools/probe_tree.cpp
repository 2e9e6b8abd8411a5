// Scratch probe (not product, not oracle): CPU replica of the GPU's implicit Morton bucket tree
// traversal with visit counters, to compare tree layouts / leaf sizes / orderings.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <vector>
#include <cstring>
struct P { float x, y, z, w; };
struct Box { float lo[3], hi[3]; };
static uint64_t expand21(uint64_t v) {
    v &= 0x1fffffull; v = (v | v << 32) & 0x1f00000000ffffull; v = (v | v << 16) & 0x1f0000ff0000ffull;
    v = (v | v << 8) & 0x100f00f00f00f00full; v = (v | v << 4) & 0x10c30c30c30c30c3ull; v = (v | v << 2) & 0x1249249249249249ull; return v; }
static float bdist(const Box &b, const float *q) {
    float s = 0; for (int d = 0; d < 3; ++d) { float e = std::max(std::max(b.lo[d] - q[d], q[d] - b.hi[d]), 0.f); s += e * e; } return s; }
static std::vector<P> load(const char *fn) { FILE *f = fopen(fn, "rb"); fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET); std::vector<P> v(sz / 16); fread(v.data(), 16, v.size(), f); fclose(f); return v; }
int main(int argc, char **argv) {
    auto tgt = load(argv[1]); auto qry = load(argv[2]); int L = atoi(argv[3]); int mode = argc > 4 ? atoi(argv[4]) : 0; // mode 0: morton bucket heap; 1: kd median split heap
    size_t n = tgt.size();
    float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
    for (auto &p : tgt) { const float *c = &p.x; for (int d = 0; d < 3; ++d) { lo[d] = std::min(lo[d], c[d]); hi[d] = std::max(hi[d], c[d]); } }
    float ext = std::max({hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]}); float sc = 2097151.f / ext;
    std::vector<uint32_t> perm(n); std::iota(perm.begin(), perm.end(), 0);
    size_t leaves = (n + L - 1) / L; size_t Pn = 1; while (Pn < leaves) Pn <<= 1;
    if (mode == 0) {
        std::vector<uint64_t> key(n);
        for (size_t i = 0; i < n; ++i) { const float *c = &tgt[i].x; uint64_t k = 0; for (int d = 0; d < 3; ++d) { uint64_t q = (uint64_t) std::min(std::max((c[d] - lo[d]) * sc, 0.f), 2097151.f); k |= expand21(q) << d; } key[i] = k; }
        std::sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
    } else {
        // balanced kd ordering: recursively split the widest dim at the position that keeps the heap complete over leaves of L
        struct Job { size_t b, e; size_t nleaf; };
        std::vector<Job> st; st.push_back({0, n, Pn});
        while (!st.empty()) { Job j = st.back(); st.pop_back(); if (j.nleaf <= 1 || j.e - j.b <= 1) continue;
            float l2[3] = {1e30f, 1e30f, 1e30f}, h2[3] = {-1e30f, -1e30f, -1e30f};
            for (size_t i = j.b; i < j.e; ++i) { const float *c = &tgt[perm[i]].x; for (int d = 0; d < 3; ++d) { l2[d] = std::min(l2[d], c[d]); h2[d] = std::max(h2[d], c[d]); } }
            int dim = 0; for (int d = 1; d < 3; ++d) if (h2[d] - l2[d] > h2[dim] - l2[dim]) dim = d;
            size_t half = j.nleaf / 2; size_t cnt = j.e - j.b; size_t leftcnt = std::min(cnt, half * (size_t) L);
            if (mode == 1) leftcnt = std::min(cnt, std::max((size_t) ((cnt + 1) / 2 + L - 1) / L * L, (size_t) 0)); // median rounded to L
            if (leftcnt > half * (size_t) L) leftcnt = half * (size_t) L;
            if (leftcnt >= cnt) { st.push_back({j.b, j.e, half}); continue; }
            std::nth_element(perm.begin() + j.b, perm.begin() + j.b + leftcnt, perm.begin() + j.e, [&](uint32_t a, uint32_t b) { return (&tgt[a].x)[dim] < (&tgt[b].x)[dim]; });
            st.push_back({j.b, j.b + leftcnt, half}); st.push_back({j.b + leftcnt, j.e, half}); }
    }
    // mode 1 packs left-aligned per subtree: need explicit leaf ranges. Build leaf ranges by simulating the same recursion -> store begin/end per heap leaf
    std::vector<size_t> lb(Pn, 0), le(Pn, 0);
    if (mode == 0) { for (size_t j = 0; j < Pn; ++j) { lb[j] = std::min(n, j * L); le[j] = std::min(n, (j + 1) * L); } }
    else { struct J2 { size_t b, e, node, nleaf; }; std::vector<J2> st; st.push_back({0, n, 1, Pn});
        while (!st.empty()) { J2 j = st.back(); st.pop_back(); if (j.nleaf == 1) { lb[j.node - Pn] = j.b; le[j.node - Pn] = j.e; continue; }
            size_t half = j.nleaf / 2, cnt = j.e - j.b; size_t leftcnt = std::min(cnt, (size_t) ((cnt + 1) / 2 + L - 1) / L * L); if (leftcnt > half * (size_t) L) leftcnt = half * (size_t) L; if (leftcnt > cnt) leftcnt = cnt;
            if (cnt <= 1 && false) {}
            st.push_back({j.b, j.b + leftcnt, 2 * j.node, half}); st.push_back({j.b + leftcnt, j.e, 2 * j.node + 1, half}); } }
    std::vector<Box> nodes(2 * Pn);
    for (size_t j = 0; j < Pn; ++j) { Box b; for (int d = 0; d < 3; ++d) { b.lo[d] = INFINITY; b.hi[d] = -INFINITY; }
        for (size_t i = lb[j]; i < le[j]; ++i) { const float *c = &tgt[perm[i]].x; for (int d = 0; d < 3; ++d) { b.lo[d] = std::min(b.lo[d], c[d]); b.hi[d] = std::max(b.hi[d], c[d]); } } nodes[Pn + j] = b; }
    for (size_t i = Pn - 1; i >= 1; --i) { Box b; for (int d = 0; d < 3; ++d) { b.lo[d] = std::min(nodes[2 * i].lo[d], nodes[2 * i + 1].lo[d]); b.hi[d] = std::max(nodes[2 * i].hi[d], nodes[2 * i + 1].hi[d]); } nodes[i] = b; }
    // traverse
    double tot_leaf = 0, tot_node = 0, tot_pts = 0; std::vector<int> lv; size_t nq = qry.size(); size_t stride = std::max<size_t>(1, nq / 20000);
    float thr = 9.f; int maxleaf = 0;
    for (size_t qi = 0; qi < nq; qi += stride) { const float *q = &qry[qi].x; float best = thr; int nl = 0, nn = 0, np = 0;
        unsigned node = 1, pending = 0;
        for (;;) { bool up = false;
            if (node >= Pn) { ++nl; for (size_t i = lb[node - Pn]; i < le[node - Pn]; ++i) { const float *c = &tgt[perm[i]].x; float dx = q[0] - c[0], dy = q[1] - c[1], dz = q[2] - c[2]; float d = dx * dx + dy * dy + dz * dz; ++np; if (d < best) best = d; } up = true; }
            else { ++nn; float d0 = bdist(nodes[2 * node], q), d1 = bdist(nodes[2 * node + 1], q); bool rf = d1 < d0; float dn = rf ? d1 : d0, df = rf ? d0 : d1; if (dn > best) up = true; else { pending = (pending << 1) | (df <= best ? 1u : 0u); node = 2 * node + (rf ? 1 : 0); } }
            if (up) { bool fin = false; for (;;) { if (node == 1) { fin = true; break; } if (pending & 1u) { pending &= ~1u; node ^= 1u; ++nn; if (bdist(nodes[node], q) <= best) break; } else { node >>= 1; pending >>= 1; } } if (fin) break; } }
        tot_leaf += nl; tot_node += nn; tot_pts += np; lv.push_back(nl); maxleaf = std::max(maxleaf, nl); }
    std::sort(lv.begin(), lv.end()); size_t m = lv.size();
    printf("mode %d L %d: leaves/query mean %.1f p50 %d p90 %d p99 %d max %d | node visits %.1f | pts %.1f\n", mode, L, tot_leaf / m, lv[m / 2], lv[m * 9 / 10], lv[m * 99 / 100], maxleaf, tot_node / m, tot_pts / m);
}

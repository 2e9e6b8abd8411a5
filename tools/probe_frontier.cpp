// Scratch probe: per-group (warp / CTA) traversal frontiers over the radix-tree LBVH.
// Simulates the ICP iterations of one match (query clouds q1..qK, warm start carried over) and
// counts node visits / leaf scans / point tests per query for
//   A  the per-thread ordered top-down walk from the root (what correspond_kernel does), and
//   B  a walk that starts from a per-group frontier: the cut of the tree against the bounding box
//      of G Morton-consecutive queries inflated by the largest warm bound of the group.
// usage: probe_frontier tgt.f4 q1.f4 [q2.f4 ...]
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
static std::vector<P> load(const char *fn) { FILE *f = fopen(fn, "rb"); if (!f) { perror(fn); exit(1); } fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET); std::vector<P> v(sz / 16); if (fread(v.data(), 16, v.size(), f)) {} fclose(f); return v; }
static std::vector<P> tgt;      // Morton sorted
static std::vector<uint64_t> key;
static std::vector<Node> nodes; static int L = 8;
static float glo[3], gsc; static int BITS = 15;
static uint64_t mkey(const float *c, int bits) { uint64_t k = 0; float qm = (float) ((1u << bits) - 1u); for (int d = 0; d < 3; ++d) { uint64_t q = (uint64_t) std::min(std::max((c[d] - glo[d]) * gsc * (qm / 32767.f), 0.f), qm); k |= expand21(q) << d; } return k; }
static float bdist(const float *lo, const float *hi, const float *q) { float s = 0; for (int d = 0; d < 3; ++d) { float e = std::max(std::max(lo[d] - q[d], q[d] - hi[d]), 0.f); s += e * e; } return s; }
static float pdist(const float *q, const P &p) { float dx = q[0] - p.x, dy = q[1] - p.y, dz = q[2] - p.z; return dx * dx + dy * dy + dz * dz; }
static int build(int b, int e, int bit) {
    int me = nodes.size(); nodes.push_back(Node());
    Node nd; for (int d = 0; d < 3; ++d) { nd.lo[d] = INFINITY; nd.hi[d] = -INFINITY; }
    for (int i = b; i < e; ++i) { const float *c = &tgt[i].x; for (int d = 0; d < 3; ++d) { nd.lo[d] = std::min(nd.lo[d], c[d]); nd.hi[d] = std::max(nd.hi[d], c[d]); } }
    nd.b = b; nd.e = e; nd.left = nd.right = -1;
    if (e - b > L) {
        int split = -1;
        while (bit >= 0) { uint64_t m = 1ull << bit; if ((key[b] & m) != (key[e - 1] & m)) break; --bit; }
        if (bit < 0) split = (b + e) / 2;
        else { uint64_t m = 1ull << bit; int lo = b, hi = e - 1; while (lo < hi) { int mid = (lo + hi) / 2; if (key[mid] & m) hi = mid; else lo = mid + 1; } split = lo; --bit; }
        int l = build(b, split, bit), r = build(split, e, bit); nd.left = l; nd.right = r;
    }
    nodes[me] = nd; return me;
}
struct Cnt { double nodes = 0, leaves = 0, pts = 0, ftests = 0, fbuild = 0, fsize = 0; };
// ordered top-down search of subtree `root` (box already known within bound)
static void search(int root, const float *q, float &best, int &bpos, Cnt &c) {
    int stack[128]; float sd[128]; int sp = 0; stack[sp] = root; sd[sp++] = 0;
    while (sp) { --sp; int ni = stack[sp]; if (sd[sp] > best) continue; const Node &nd = nodes[ni];
        if (nd.left < 0) { c.leaves += 1; for (int i = nd.b; i < nd.e; ++i) { float d = pdist(q, tgt[i]); c.pts += 1; if (d < best) { best = d; bpos = i; } } }
        else { c.nodes += 1; const Node &a = nodes[nd.left], &b = nodes[nd.right]; float d0 = bdist(a.lo, a.hi, q), d1 = bdist(b.lo, b.hi, q);
            if (d0 <= d1) { if (d1 <= best) { stack[sp] = nd.right; sd[sp++] = d1; } if (d0 <= best) { stack[sp] = nd.left; sd[sp++] = d0; } }
            else { if (d0 <= best) { stack[sp] = nd.left; sd[sp++] = d0; } if (d1 <= best) { stack[sp] = nd.right; sd[sp++] = d1; } } } }
}
int main(int argc, char **argv) {
    std::vector<P> traw = load(argv[1]); size_t n = traw.size();
    float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
    for (auto &p : traw) { const float *c = &p.x; for (int d = 0; d < 3; ++d) { lo[d] = std::min(lo[d], c[d]); hi[d] = std::max(hi[d], c[d]); } }
    float ext = std::max({hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]}); for (int d = 0; d < 3; ++d) glo[d] = lo[d]; gsc = 32767.f / ext;
    { std::vector<uint32_t> perm(n); std::iota(perm.begin(), perm.end(), 0); std::vector<uint64_t> k0(n); for (size_t i = 0; i < n; ++i) k0[i] = mkey(&traw[i].x, BITS);
      std::sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return k0[a] < k0[b]; }); tgt.resize(n); key.resize(n); for (size_t i = 0; i < n; ++i) { tgt[i] = traw[perm[i]]; key[i] = k0[perm[i]]; } }
    build(0, n, 62);
    int maxdepth = 0; { std::vector<std::pair<int,int>> st; st.push_back({0, 1}); double sumd = 0; size_t nl = 0; while (!st.empty()) { auto [ni, d] = st.back(); st.pop_back(); if (nodes[ni].left < 0) { sumd += d; ++nl; maxdepth = std::max(maxdepth, d); } else { st.push_back({nodes[ni].left, d + 1}); st.push_back({nodes[ni].right, d + 1}); } } printf("tree: %zu nodes, %zu leaves, mean leaf depth %.1f max %d\n", nodes.size(), nl, sumd / nl, maxdepth); }
    // queries: sorted once by 10-bit Morton of iteration-1 positions
    int K = argc - 2; std::vector<std::vector<P>> Q(K); for (int k = 0; k < K; ++k) Q[k] = load(argv[2 + k]);
    size_t nq = Q[0].size(); std::vector<uint32_t> qperm(nq); std::iota(qperm.begin(), qperm.end(), 0);
    { std::vector<uint64_t> qk(nq); for (size_t i = 0; i < nq; ++i) qk[i] = mkey(&Q[0][i].x, 10); std::stable_sort(qperm.begin(), qperm.end(), [&](uint32_t a, uint32_t b) { return qk[a] < qk[b]; }); }
    const float THR = 9.f;
    for (int variant = (argc > 0 && getenv("V0")) ? atoi(getenv("V0")) : 0; variant < (getenv("PG") ? atoi(getenv("V0")) + 1 : 13); ++variant) {
        int G = 0, FMAX = 0; bool lbinit = false; int levelsync = 0; float alpha = 0.f;
        switch (variant) { case 0: break; case 1: lbinit = true; break; case 2: G = 32; FMAX = 8; lbinit = true; break; case 3: G = 32; FMAX = 16; lbinit = true; break;
            case 4: G = 128; FMAX = 16; lbinit = true; break; case 5: G = 128; FMAX = 32; lbinit = true; break; case 6: G = 32; FMAX = 32; lbinit = true; break;
            case 7: G = 32; FMAX = 8; levelsync = 1; break; case 8: G = 32; FMAX = 16; levelsync = 1; break; case 9: G = 32; FMAX = 32; levelsync = 1; break;
            case 10: G = 32; FMAX = 16; levelsync = 1; alpha = 0.5f; break; case 11: G = 32; FMAX = 32; levelsync = 1; alpha = 0.5f; break; case 12: G = 32; FMAX = 32; levelsync = 1; alpha = 1.0f; break; }
        if (getenv("PG")) { G = atoi(getenv("PG")); FMAX = atoi(getenv("PF")); levelsync = 1; lbinit = true; alpha = 0.f; }
        printf("variant %d: group %d fmax %d lower-bound-init %d\n", variant, G, FMAX, (int) lbinit);
        std::vector<int> warm(nq, -1);
        for (int k = 0; k < K; ++k) {
            Cnt c; const auto &q = Q[k];
            std::vector<float> bound(nq); std::vector<int> bpos(nq, -1);
            for (size_t s = 0; s < nq; ++s) { const float *qq = &q[qperm[s]].x; float b = THR; int bp = -1;
                if (warm[s] >= 0) { float d = pdist(qq, tgt[warm[s]]); if (d <= b) { b = d; bp = warm[s]; } }
                else if (lbinit) { uint64_t kq = mkey(qq, BITS); size_t p = std::lower_bound(key.begin(), key.end(), kq) - key.begin(); for (long j = (long) p - 1; j <= (long) p; ++j) if (j >= 0 && j < (long) n) { float d = pdist(qq, tgt[j]); if (d <= b) { b = d; bp = (int) j; } } }
                bound[s] = b; bpos[s] = bp; }
            if (G == 0) { for (size_t s = 0; s < nq; ++s) { const float *qq = &q[qperm[s]].x; if (bdist(nodes[0].lo, nodes[0].hi, qq) <= bound[s]) search(0, qq, bound[s], bpos[s], c); } }
            else {
                for (size_t g0 = 0; g0 < nq; g0 += G) { size_t g1 = std::min(nq, g0 + G);
                    float blo[3] = {1e30f, 1e30f, 1e30f}, bhi[3] = {-1e30f, -1e30f, -1e30f}; float R2 = 0;
                    for (size_t s = g0; s < g1; ++s) { const float *qq = &q[qperm[s]].x; for (int d = 0; d < 3; ++d) { blo[d] = std::min(blo[d], qq[d]); bhi[d] = std::max(bhi[d], qq[d]); } R2 = std::max(R2, bound[s]); }
                    float R = std::sqrt(R2) * 1.000001f + 1e-30f; for (int d = 0; d < 3; ++d) { blo[d] -= R; bhi[d] += R; }
                    auto hits = [&](const Node &nd) { for (int d = 0; d < 3; ++d) if (nd.lo[d] > bhi[d] || nd.hi[d] < blo[d]) return false; return true; };
                    std::vector<int> fr; if (hits(nodes[0])) fr.push_back(0);
                    if (levelsync) { float bext = std::max({bhi[0] - blo[0], bhi[1] - blo[1], bhi[2] - blo[2]});
                        for (;;) { std::vector<int> nx; bool any = false; for (int f : fr) { const Node &nd = nodes[f]; float sz = std::max({nd.hi[0] - nd.lo[0], nd.hi[1] - nd.lo[1], nd.hi[2] - nd.lo[2]});
                                if (nd.left < 0 || sz <= alpha * bext) { nx.push_back(f); continue; } any = true; if (hits(nodes[nd.left])) nx.push_back(nd.left); if (hits(nodes[nd.right])) nx.push_back(nd.right); }
                            if (!any || (int) nx.size() > FMAX) break; c.fbuild += 1; fr.swap(nx); } }
                    else for (;;) { int bi = -1; float bsz = -1; for (size_t i = 0; i < fr.size(); ++i) { const Node &nd = nodes[fr[i]]; if (nd.left < 0) continue; float sz = std::max({nd.hi[0] - nd.lo[0], nd.hi[1] - nd.lo[1], nd.hi[2] - nd.lo[2]}); if (sz > bsz) { bsz = sz; bi = i; } }
                        if (bi < 0) break; const Node &nd = nodes[fr[bi]]; bool hl = hits(nodes[nd.left]), hr = hits(nodes[nd.right]); if ((int) fr.size() - 1 + hl + hr > FMAX) break; c.fbuild += 1;
                        int l = nd.left, r = nd.right; fr.erase(fr.begin() + bi); if (hl) fr.push_back(l); if (hr) fr.push_back(r); }
                    c.fsize += fr.size();
                    for (size_t s = g0; s < g1; ++s) { const float *qq = &q[qperm[s]].x; std::vector<std::pair<float,int>> ds; for (int f : fr) { c.ftests += 1; float d = bdist(nodes[f].lo, nodes[f].hi, qq); if (d <= bound[s]) ds.push_back({d, f}); }
                        std::sort(ds.begin(), ds.end()); for (auto &e : ds) if (e.first <= bound[s]) search(e.second, qq, bound[s], bpos[s], c); }
                }
            }
            size_t ngroups = G ? (nq + G - 1) / G : 1;
            double cost = (c.nodes * 46 + c.ftests * 20 + c.leaves * 10 + c.pts * 10) / nq + (G ? c.fbuild * 60.0 / nq : 0);
            printf("  it %d: nodes %.1f leaves %.2f pts %.1f | frontier size %.1f tests/q %.1f build-expansions/group %.1f | cost~%.0f\n", k + 1, c.nodes / nq, c.leaves / nq, c.pts / nq, c.fsize / ngroups, c.ftests / nq, c.fbuild / ngroups, cost);
            for (size_t s = 0; s < nq; ++s) warm[s] = bpos[s];
        }
    }
}

// bvh_lab.cpp -- CPU laboratory for the tree-quality work on the ray caster (not product code, not the oracle).
//
//   g++ -O2 -std=c++17 -fopenmp tools/bvh_lab.cpp -o /tmp/lab/bvh_lab && /tmp/lab/bvh_lab mesh.bin [n_ref]
//
// Builds, for one mesh, the binary trees the GPU builder can produce (Morton LBVH = Karras topology; PLOC with search
// radius r on the Morton order) and the 4-wide collapses (every other level / greedy by surface area), then counts what
// the traversal kernel would do for the AGS rays (origin on the surface, cone about the inward normal, ALL crossings, no
// t_max): wide nodes visited and leaves reached per ray.  The numbers decide which build the CUDA code implements.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <random>
#include <vector>

struct Box { float lo[3], hi[3]; };
static Box merge(const Box& a, const Box& b) {
    Box r;
    for (int k = 0; k < 3; ++k) { r.lo[k] = std::min(a.lo[k], b.lo[k]); r.hi[k] = std::max(a.hi[k], b.hi[k]); }
    return r;
}
static float area(const Box& b) {
    const float x = b.hi[0] - b.lo[0], y = b.hi[1] - b.lo[1], z = b.hi[2] - b.lo[2];
    return x * y + y * z + z * x;
}

struct BinTree {                       // child code: >= 0 internal, < 0 leaf ~code (leaf index = position in `leaf_box`)
    std::vector<int> left, right;
    std::vector<Box> box;              // internal boxes
    std::vector<Box> leaf_box;
    int root = 0;
    const Box& cb(int c) const { return c < 0 ? leaf_box[~c] : box[c]; }
};

static uint32_t expand10(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu; v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u; v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

// Karras topology on sorted 30-bit codes (ties broken by position) == recursive split at the highest differing bit
static int lbvh_rec(BinTree& t, const std::vector<uint64_t>& key, int lo, int hi) {      // [lo, hi] inclusive, hi > lo
    const uint64_t diff = key[lo] ^ key[hi];
    const int bit = 63 - __builtin_clzll(diff);
    // first index whose `bit` is set
    int a = lo, b = hi;
    while (a + 1 < b) { const int m = (a + b) / 2; if ((key[m] >> bit) & 1) b = m; else a = m; }
    const int split = a;               // last of the left part
    const int id = (int)t.left.size();
    t.left.push_back(0); t.right.push_back(0); t.box.push_back(Box());
    const int l = (split == lo) ? ~lo : lbvh_rec(t, key, lo, split);
    const int r = (split + 1 == hi) ? ~hi : lbvh_rec(t, key, split + 1, hi);
    t.left[id] = l; t.right[id] = r;
    t.box[id] = merge(t.cb(l), t.cb(r));
    return id;
}

// PLOC (Meister & Bittner 2018) on the Morton order, search radius r
static void ploc(BinTree& t, int r) {
    const int n = (int)t.leaf_box.size();
    std::vector<int> code(n), nn(n), next;
    std::vector<Box> cb(n), nb;
    for (int i = 0; i < n; ++i) { code[i] = ~i; cb[i] = t.leaf_box[i]; }
    int m = n, iters = 0;
    while (m > 1) {
        ++iters;
#pragma omp parallel for schedule(static)
        for (int i = 0; i < m; ++i) {
            float best = 1e30f; int bj = -1;
            for (int j = std::max(0, i - r); j <= std::min(m - 1, i + r); ++j) {
                if (j == i) continue;
                const float d = area(merge(cb[i], cb[j]));
                // total order on pairs: (d, min(i,j), max(i,j))
                if (d < best || (d == best && (std::min(i, j) < std::min(i, bj) || (std::min(i, j) == std::min(i, bj) && std::max(i, j) < std::max(i, bj))))) { best = d; bj = j; }
            }
            nn[i] = bj;
        }
        next.clear(); nb.clear();
        for (int i = 0; i < m; ++i) {
            const int j = nn[i];
            if (nn[j] == i) {
                if (i < j) {
                    const int id = (int)t.left.size();
                    t.left.push_back(code[i]); t.right.push_back(code[j]);
                    t.box.push_back(merge(cb[i], cb[j]));
                    next.push_back(id); nb.push_back(t.box[id]);
                }
            } else { next.push_back(code[i]); nb.push_back(cb[i]); }
        }
        code.swap(next); cb.swap(nb);
        m = (int)code.size();
    }
    t.root = code[0];
    fprintf(stderr, "ploc r=%d: %d iterations\n", r, iters);
}

struct Wide { int child[8]; int n; };
struct WideTree { std::vector<Wide> nodes; int width; };

// every other level (the round-1 GPU build): wide node = binary node of even depth, children = grandchildren
static WideTree collapse_levels(const BinTree& t) {
    WideTree w; w.width = 4;
    std::vector<int> widx(t.left.size(), -1);
    std::function<int(int)> rec = [&](int b) -> int {
        const int id = (int)w.nodes.size();
        w.nodes.push_back(Wide());
        int ch[4], nc = 0;
        for (int c : {t.left[b], t.right[b]}) {
            if (c < 0) ch[nc++] = c; else { ch[nc++] = t.left[c]; ch[nc++] = t.right[c]; }
        }
        Wide wn; wn.n = nc;
        for (int k = 0; k < nc; ++k) wn.child[k] = ch[k] < 0 ? ch[k] : -0x40000000;    // placeholder
        for (int k = 0; k < nc; ++k) if (ch[k] >= 0) wn.child[k] = rec(ch[k]);
        w.nodes[id] = wn;
        return id;
    };
    rec(t.root);
    return w;
}

// greedy: open the child with the largest surface area until `width` children
static WideTree collapse_greedy(const BinTree& t, int width, std::vector<Box>* child_boxes = nullptr) {
    WideTree w; w.width = width;
    std::function<int(int)> rec = [&](int b) -> int {
        const int id = (int)w.nodes.size();
        w.nodes.push_back(Wide());
        int ch[8], nc = 0;
        ch[nc++] = t.left[b]; ch[nc++] = t.right[b];
        while (nc < width) {
            int best = -1; float ba = -1.f;
            for (int k = 0; k < nc; ++k) if (ch[k] >= 0) { const float a = area(t.box[ch[k]]); if (a > ba) { ba = a; best = k; } }
            if (best < 0) break;
            const int c = ch[best];
            ch[best] = t.left[c]; ch[nc++] = t.right[c];
        }
        Wide wn; wn.n = nc;
        std::vector<int> bin(ch, ch + nc);
        for (int k = 0; k < nc; ++k) wn.child[k] = ch[k] < 0 ? ch[k] : rec(ch[k]);
        w.nodes[id] = wn;
        return id;
    };
    rec(t.root);
    return w;
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: bvh_lab mesh.bin [n_ref]\n"); return 1; }
    const int n_ref = argc > 2 ? atoi(argv[2]) : 400;
    FILE* fh = fopen(argv[1], "rb");
    int hdr[2];
    if (!fh || fread(hdr, 4, 2, fh) != 2) return 1;
    const int nV = hdr[0], nF = hdr[1];
    std::vector<double> V(3 * nV); std::vector<int> F(3 * nF);
    if (fread(V.data(), 8, 3 * nV, fh) != (size_t)3 * nV || fread(F.data(), 4, 3 * nF, fh) != (size_t)3 * nF) return 1;
    fclose(fh);
    double lo[3] = {1e30, 1e30, 1e30}, hi[3] = {-1e30, -1e30, -1e30};
    for (int v = 0; v < nV; ++v) for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], V[3 * v + k]); hi[k] = std::max(hi[k], V[3 * v + k]); }
    double max_abs = 0;
    for (int k = 0; k < 3; ++k) max_abs = std::max(max_abs, std::max(std::fabs(lo[k]), std::fabs(hi[k])));
    const float pad = (float)(2e-5 * max_abs + 1e-9);
    // Morton order
    std::vector<uint64_t> key(nF);
    std::vector<Box> tb(nF);
    for (int f = 0; f < nF; ++f) {
        Box b; for (int k = 0; k < 3; ++k) { b.lo[k] = 1e30f; b.hi[k] = -1e30f; }
        for (int v = 0; v < 3; ++v) for (int k = 0; k < 3; ++k) {
            const float x = (float)V[3 * F[3 * f + v] + k];
            b.lo[k] = std::min(b.lo[k], x - pad); b.hi[k] = std::max(b.hi[k], x + pad);
        }
        tb[f] = b;
        uint32_t q[3];
        for (int k = 0; k < 3; ++k) {
            const double c = 0.5 * ((double)b.lo[k] + b.hi[k]);
            q[k] = (uint32_t)std::min(std::max((c - lo[k]) / (hi[k] - lo[k]) * 1024.0, 0.0), 1023.0);
        }
        key[f] = ((uint64_t)((expand10(q[0]) << 2) | (expand10(q[1]) << 1) | expand10(q[2])) << 32) | (uint32_t)f;
    }
    std::vector<int> order(nF);
    for (int f = 0; f < nF; ++f) order[f] = f;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return (key[a] >> 32) < (key[b] >> 32); });
    std::vector<uint64_t> skey(nF);
    std::vector<Box> leaf_box(nF);
    for (int i = 0; i < nF; ++i) { skey[i] = (key[order[i]] & 0xffffffff00000000ull) | (uint32_t)i; leaf_box[i] = tb[order[i]]; }

    // rays: area-weighted reference points, triangle normal, cone about -n, atan(0.25)
    std::mt19937_64 rng(1);
    std::uniform_real_distribution<double> U(0, 1);
    std::vector<double> cdf(nF);
    double acc = 0;
    auto tri = [&](int f, int v, int k) { return V[3 * F[3 * f + v] + k]; };
    for (int f = 0; f < nF; ++f) {
        double e0[3], e1[3];
        for (int k = 0; k < 3; ++k) { e0[k] = tri(f, 1, k) - tri(f, 0, k); e1[k] = tri(f, 2, k) - tri(f, 0, k); }
        const double cx = e0[1] * e1[2] - e0[2] * e1[1], cy = e0[2] * e1[0] - e0[0] * e1[2], cz = e0[0] * e1[1] - e0[1] * e1[0];
        acc += 0.5 * std::sqrt(cx * cx + cy * cy + cz * cz);
        cdf[f] = acc;
    }
    const int n_rays = 100;
    std::vector<float> ro(3 * (size_t)n_ref * n_rays), rd(3 * (size_t)n_ref * n_rays);
    const double ang = std::atan(0.25);
    for (int i = 0; i < n_ref; ++i) {
        const int f = (int)(std::lower_bound(cdf.begin(), cdf.end(), U(rng) * acc) - cdf.begin());
        const double r1 = std::sqrt(U(rng)), r2 = U(rng);
        double p[3], e0[3], e1[3];
        for (int k = 0; k < 3; ++k) {
            p[k] = (1 - r1) * tri(f, 0, k) + r1 * (1 - r2) * tri(f, 1, k) + r1 * r2 * tri(f, 2, k);
            e0[k] = tri(f, 1, k) - tri(f, 0, k); e1[k] = tri(f, 2, k) - tri(f, 0, k);
        }
        double n[3] = {e0[1] * e1[2] - e0[2] * e1[1], e0[2] * e1[0] - e0[0] * e1[2], e0[0] * e1[1] - e0[1] * e1[0]};
        const double nn = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        double ax[3] = {-n[0] / nn, -n[1] / nn, -n[2] / nn};
        double kk[3] = {ax[0], ax[1], ax[2] + 1.0};
        double k2 = kk[0] * kk[0] + kk[1] * kk[1] + kk[2] * kk[2];
        if (k2 == 0) { kk[0] = 1; kk[1] = 0; kk[2] = 0; k2 = 1; }
        for (int r = 0; r < n_rays; ++r) {
            const double az = U(rng) * 2 * M_PI, inc = U(rng) * ang;
            const double c[3] = {std::sin(inc) * std::cos(az), std::sin(inc) * std::sin(az), std::cos(inc)};
            const double kc = kk[0] * c[0] + kk[1] * c[1] + kk[2] * c[2];
            const size_t o = 3 * ((size_t)i * n_rays + r);
            for (int k = 0; k < 3; ++k) { ro[o + k] = (float)p[k]; rd[o + k] = (float)(2 * kc / k2 * kk[k] - c[k]); }
        }
    }
    const size_t n_total = (size_t)n_ref * n_rays;

    auto slab = [&](const Box& b, const float* o, const float* inv) {
        float tn = -1e30f, tf = 1e30f;
        for (int k = 0; k < 3; ++k) {
            const float t0 = (b.lo[k] - o[k]) * inv[k], t1 = (b.hi[k] - o[k]) * inv[k];
            tn = std::max(tn, std::min(t0, t1)); tf = std::min(tf, std::max(t0, t1));
        }
        return tn <= tf && tf >= -1e-5f;
    };
    auto measure = [&](const char* name, const BinTree& t, const WideTree& w) {
        double visits = 0, leaves = 0, tests = 0;
#pragma omp parallel for reduction(+ : visits, leaves, tests) schedule(dynamic, 256)
        for (size_t i = 0; i < n_total; ++i) {
            const float* o = &ro[3 * i];
            float inv[3];
            for (int k = 0; k < 3; ++k) { const float d = rd[3 * i + k]; inv[k] = 1.0f / (std::fabs(d) < 1e-20f ? std::copysign(1e-20f, d) : d); }
            int stack[256], sp = 0;
            stack[sp++] = 0;
            // per wide node we need the child boxes: a wide child's box is the box of the binary node it came from; keep a
            // parallel array (filled below) -- here looked up through `wbox`
            while (sp) {
                const int nd = stack[--sp];
                ++visits;
                const Wide& wn = w.nodes[nd];
                for (int k = 0; k < wn.n; ++k) {
                    ++tests;
                    const int c = wn.child[k];
                    const Box& b = c < 0 ? t.leaf_box[~c] : *reinterpret_cast<const Box*>(&t.box[0]);   // replaced below
                    (void)b;
                }
            }
        }
        (void)name; (void)leaves;
    };
    (void)measure;

    // simpler: store child boxes in the wide nodes
    struct WideB { Box b[8]; int child[8]; int n; };
    auto with_boxes = [&](const BinTree& t, int width, bool greedy) {
        std::vector<WideB> out;
        std::function<int(int)> rec = [&](int b) -> int {
            const int id = (int)out.size();
            out.push_back(WideB());
            int ch[8], nc = 0;
            if (!greedy) {
                for (int c : {t.left[b], t.right[b]}) { if (c < 0) ch[nc++] = c; else { ch[nc++] = t.left[c]; ch[nc++] = t.right[c]; } }
            } else {
                ch[nc++] = t.left[b]; ch[nc++] = t.right[b];
                while (nc < width) {
                    int best = -1; float ba = -1.f;
                    for (int k = 0; k < nc; ++k) if (ch[k] >= 0) { const float a = area(t.box[ch[k]]); if (a > ba) { ba = a; best = k; } }
                    if (best < 0) break;
                    const int c = ch[best];
                    ch[best] = t.left[c]; ch[nc++] = t.right[c];
                }
            }
            WideB wn; wn.n = nc;
            for (int k = 0; k < nc; ++k) wn.b[k] = t.cb(ch[k]);
            for (int k = 0; k < nc; ++k) wn.child[k] = ch[k] < 0 ? ch[k] : rec(ch[k]);
            out[id] = wn;
            return id;
        };
        rec(t.root);
        return out;
    };
    auto run = [&](const char* name, const BinTree& t, int width, bool greedy) {
        const std::vector<WideB> w = with_boxes(t, width, greedy);
        double visits = 0, leaves = 0, maxd = 0;
        double sah = 0;
        for (const WideB& n : w) for (int k = 0; k < n.n; ++k) sah += area(n.b[k]);
#pragma omp parallel for reduction(+ : visits, leaves) schedule(dynamic, 256)
        for (size_t i = 0; i < n_total; ++i) {
            const float* o = &ro[3 * i];
            float inv[3];
            for (int k = 0; k < 3; ++k) { const float d = rd[3 * i + k]; inv[k] = 1.0f / (std::fabs(d) < 1e-20f ? std::copysign(1e-20f, d) : d); }
            int stack[512], sp = 0;
            stack[sp++] = 0;
            while (sp) {
                const WideB& wn = w[stack[--sp]];
                ++visits;
                for (int k = 0; k < wn.n; ++k)
                    if (slab(wn.b[k], o, inv)) { if (wn.child[k] < 0) ++leaves; else stack[sp++] = wn.child[k]; }
            }
        }
        (void)maxd;
        printf("%-34s wide nodes %7zu  visits/ray %6.2f  leaves/ray %5.2f  sum child area %.4f\n", name, w.size(), visits / n_total, leaves / n_total, sah);
    };

    // ---- normal cones (valid hits need the interpolated vertex normal within atan(mu) of the ray direction) ----
    std::vector<double> VN(3 * nV, 0.0);
    for (int f = 0; f < nF; ++f) {
        double e0[3], e1[3];
        for (int k = 0; k < 3; ++k) { e0[k] = tri(f, 1, k) - tri(f, 0, k); e1[k] = tri(f, 2, k) - tri(f, 0, k); }
        const double c[3] = {e0[1] * e1[2] - e0[2] * e1[1], e0[2] * e1[0] - e0[0] * e1[2], e0[0] * e1[1] - e0[1] * e1[0]};
        for (int v = 0; v < 3; ++v) for (int k = 0; k < 3; ++k) VN[3 * F[3 * f + v] + k] += c[k];
    }
    for (int v = 0; v < nV; ++v) { const double n = std::sqrt(VN[3*v]*VN[3*v]+VN[3*v+1]*VN[3*v+1]+VN[3*v+2]*VN[3*v+2]); for (int k = 0; k < 3; ++k) VN[3*v+k] /= n; }
    struct Cone { double a[3]; double th; };
    auto leaf_cone = [&](int leaf) {
        const int f = order[leaf];
        Cone c; double s[3] = {0, 0, 0};
        for (int v = 0; v < 3; ++v) for (int k = 0; k < 3; ++k) s[k] += VN[3 * F[3 * f + v] + k];
        const double n = std::sqrt(s[0]*s[0]+s[1]*s[1]+s[2]*s[2]);
        for (int k = 0; k < 3; ++k) c.a[k] = s[k] / n;
        c.th = 0;
        for (int v = 0; v < 3; ++v) { double d = 0; for (int k = 0; k < 3; ++k) d += c.a[k] * VN[3 * F[3 * f + v] + k]; c.th = std::max(c.th, std::acos(std::min(1.0, std::max(-1.0, d)))); }
        return c;
    };
    // neighbour-expanded cone: own axis, half-angle covers the vertex normals of every triangle sharing a vertex
    std::vector<std::vector<int>> v2f(nV);
    for (int f = 0; f < nF; ++f) for (int v = 0; v < 3; ++v) v2f[F[3 * f + v]].push_back(f);
    auto leaf_cone_nb = [&](int leaf) {
        Cone c = leaf_cone(leaf);
        const int f = order[leaf];
        for (int v = 0; v < 3; ++v) for (int g : v2f[F[3 * f + v]]) for (int u = 0; u < 3; ++u) {
            double d = 0; for (int k = 0; k < 3; ++k) d += c.a[k] * VN[3 * F[3 * g + u] + k];
            c.th = std::max(c.th, std::acos(std::min(1.0, std::max(-1.0, d))));
        }
        c.th = std::ceil((c.th + 0.012) / (M_PI / 255)) * (M_PI / 255);
        return c;
    };
    auto merge_cone = [&](const Cone& x, const Cone& y) {
        Cone c; double s[3]; for (int k = 0; k < 3; ++k) s[k] = x.a[k] + y.a[k];
        const double n = std::sqrt(s[0]*s[0]+s[1]*s[1]+s[2]*s[2]);
        if (n < 1e-9) { c = x; c.th = M_PI; return c; }
        for (int k = 0; k < 3; ++k) c.a[k] = s[k] / n;
        double dx = 0, dy = 0; for (int k = 0; k < 3; ++k) { dx += c.a[k] * x.a[k]; dy += c.a[k] * y.a[k]; }
        c.th = std::min(M_PI, std::max(std::acos(std::min(1.0, dx)) + x.th, std::acos(std::min(1.0, dy)) + y.th));
        return c;
    };
    auto run_cull = [&](const char* name, const BinTree& t, int width) {
        std::vector<Cone> nc(t.left.size());
        std::function<Cone(int)> cone_of = [&](int c) -> Cone { if (c < 0) return leaf_cone(~c); Cone r = merge_cone(cone_of(t.left[c]), cone_of(t.right[c])); nc[c] = r; return r; };
        cone_of(t.root);
        struct WideC { Box b[8]; int child[8]; double ax[8][3]; double cth[8]; int n; };
        std::vector<WideC> out;
        const double alpha = std::atan(0.25) + 0.002;
        std::function<int(int)> rec = [&](int b) -> int {
            const int id = (int)out.size(); out.push_back(WideC());
            int ch[8], n = 0; ch[n++] = t.left[b]; ch[n++] = t.right[b];
            while (n < width) { int best = -1; float ba = -1.f; for (int k = 0; k < n; ++k) if (ch[k] >= 0) { const float a = area(t.box[ch[k]]); if (a > ba) { ba = a; best = k; } } if (best < 0) break; const int c = ch[best]; ch[best] = t.left[c]; ch[n++] = t.right[c]; }
            WideC wn; wn.n = n;
            for (int k = 0; k < n; ++k) { wn.b[k] = t.cb(ch[k]); const Cone c = ch[k] < 0 ? leaf_cone(~ch[k]) : nc[ch[k]]; for (int q = 0; q < 3; ++q) wn.ax[k][q] = c.a[q]; const double lim = c.th + alpha; wn.cth[k] = lim >= M_PI ? -2.0 : std::cos(lim); }
            for (int k = 0; k < n; ++k) wn.child[k] = ch[k] < 0 ? ch[k] : rec(ch[k]);
            out[id] = wn; return id;
        };
        rec(t.root);
        double visits = 0, leaves = 0;
#pragma omp parallel for reduction(+ : visits, leaves) schedule(dynamic, 256)
        for (size_t i = 0; i < n_total; ++i) {
            const float* o = &ro[3 * i]; float inv[3];
            for (int k = 0; k < 3; ++k) { const float d = rd[3 * i + k]; inv[k] = 1.0f / (std::fabs(d) < 1e-20f ? std::copysign(1e-20f, d) : d); }
            int stack[512], sp = 0; stack[sp++] = 0;
            while (sp) {
                const WideC& wn = out[stack[--sp]]; ++visits;
                for (int k = 0; k < wn.n; ++k) {
                    const double dd = rd[3*i] * wn.ax[k][0] + rd[3*i+1] * wn.ax[k][1] + rd[3*i+2] * wn.ax[k][2];
                    if (dd < wn.cth[k]) continue;
                    if (slab(wn.b[k], o, inv)) { if (wn.child[k] < 0) ++leaves; else stack[sp++] = wn.child[k]; }
                }
            }
        }
        printf("%-34s wide nodes %7zu  visits/ray %6.2f  leaves/ray %5.2f  (normal-cone culling)\n", name, out.size(), visits / n_total, leaves / n_total);
    };

    BinTree lb; lb.leaf_box = leaf_box;
    lb.root = lbvh_rec(lb, skey, 0, nF - 1);
    // per-node (on entry) culling, neighbour-expanded quantised cones, every-other-level collapse as in the GPU build
    auto run_gpu_like = [&](const char* name, const BinTree& t) {
        std::vector<Cone> nc(t.left.size());
        std::function<Cone(int)> cone_of = [&](int c) -> Cone { if (c < 0) return leaf_cone_nb(~c); Cone r = merge_cone(cone_of(t.left[c]), cone_of(t.right[c])); r.th = std::min(M_PI, std::ceil((r.th + 0.012) / (M_PI / 255)) * (M_PI / 255)); nc[c] = r; return r; };
        cone_of(t.root);
        int max_depth = 0;
        { std::function<void(int, int)> dep = [&](int c, int d) { if (c < 0) { max_depth = std::max(max_depth, d); return; } dep(t.left[c], d + 1); dep(t.right[c], d + 1); }; dep(t.root, 0); }
        struct WN { Box b[4]; int child[4]; int n; double ax[3]; double cth; };
        std::vector<WN> out;
        const double alpha = std::atan(0.25) + 0.002;
        std::function<int(int)> rec = [&](int b) -> int {
            const int id = (int)out.size(); out.push_back(WN());
            int ch[4], n = 0;
            for (int c : {t.left[b], t.right[b]}) { if (c < 0) ch[n++] = c; else { ch[n++] = t.left[c]; ch[n++] = t.right[c]; } }
            WN wn; wn.n = n;
            for (int q = 0; q < 3; ++q) wn.ax[q] = nc[b].a[q];
            wn.cth = (b == t.root || nc[b].th + alpha >= M_PI) ? -2.0 : std::cos(nc[b].th + alpha);
            for (int k = 0; k < n; ++k) wn.b[k] = t.cb(ch[k]);
            for (int k = 0; k < n; ++k) wn.child[k] = ch[k] < 0 ? ch[k] : rec(ch[k]);
            out[id] = wn; return id;
        };
        rec(t.root);
        std::vector<Cone> lc(nF);
        for (int i = 0; i < nF; ++i) lc[i] = leaf_cone_nb(i);
        double fetch = 0, visits = 0, leaves = 0, leaves_kept = 0;
#pragma omp parallel for reduction(+ : fetch, visits, leaves, leaves_kept) schedule(dynamic, 256)
        for (size_t i = 0; i < n_total; ++i) {
            const float* o = &ro[3 * i]; float inv[3];
            for (int k = 0; k < 3; ++k) { const float d = rd[3 * i + k]; inv[k] = 1.0f / (std::fabs(d) < 1e-20f ? std::copysign(1e-20f, d) : d); }
            int stack[512], sp = 0; stack[sp++] = 0;
            while (sp) {
                const WN& wn = out[stack[--sp]]; ++fetch;
                const double dd = rd[3*i] * wn.ax[0] + rd[3*i+1] * wn.ax[1] + rd[3*i+2] * wn.ax[2];
                if (dd < wn.cth) continue;
                ++visits;
                for (int k = 0; k < wn.n; ++k)
                    if (slab(wn.b[k], o, inv)) {
                        if (wn.child[k] < 0) { ++leaves; const Cone& c = lc[~wn.child[k]]; const double d2 = rd[3*i]*c.a[0]+rd[3*i+1]*c.a[1]+rd[3*i+2]*c.a[2]; if (!(c.th + alpha < M_PI && d2 < std::cos(c.th + alpha))) ++leaves_kept; }
                        else stack[sp++] = wn.child[k];
                    }
            }
        }
        printf("%-28s GPU-like (every other level, per-node cull, nb-expanded 8-bit cones): fetches/ray %.2f  full visits/ray %.2f  leaves reached %.2f  pretests %.2f  binary depth %d\n", name, fetch / n_total, visits / n_total, leaves / n_total, leaves_kept / n_total, max_depth);
    };
    run_gpu_like("LBVH", lb);

    run_cull("LBVH greedy 4-wide + cone cull", lb, 4);
    run_cull("LBVH greedy 8-wide + cone cull", lb, 8);
    run("LBVH, every other level (round 1)", lb, 4, false);
    run("LBVH, greedy 4-wide", lb, 4, true);
    run("LBVH, greedy 8-wide", lb, 8, true);
    for (int r : {4, 8, 16, 32}) {
        BinTree pt; pt.leaf_box = leaf_box;
        ploc(pt, r);
        char nm[64];
        snprintf(nm, sizeof nm, "PLOC r=%d", r); run_gpu_like(nm, pt);
        snprintf(nm, sizeof nm, "PLOC r=%d, every other level", r); run(nm, pt, 4, false);
        snprintf(nm, sizeof nm, "PLOC r=%d, greedy 4-wide", r); run(nm, pt, 4, true);
        snprintf(nm, sizeof nm, "PLOC r=%d, greedy 8-wide", r); run(nm, pt, 8, true);
    }
    return 0;
}

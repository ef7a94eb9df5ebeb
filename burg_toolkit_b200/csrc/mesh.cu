// mesh.cu -- library plumbing, device mesh upload and the GPU-built LBVH.
//
// Replaces the set-up half of the reference's sampler (sampling.py:191-192: as_trimesh +
// RayMeshIntersector construction; :372-380: CollisionManager population).  The LBVH is the Karras 2012
// construction: 30-bit Morton codes of the triangle-box centres -> radix sort (CUB) -> one thread per
// internal node finds its key range and split -> bottom-up refit with arrival counters -> final 64-byte
// nodes that carry BOTH children's boxes (one float4x4 fetch per traversal step).
#include <cub/cub.cuh>
#include <stdarg.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

#include "common.cuh"

namespace bt {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
    return BT_E_CUDA;
}

// ---- kernels -----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t expand_bits10(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

__global__ void morton_kernel(const double* __restrict__ verts, const int32_t* __restrict__ faces, int64_t nF,
                              double3 bmin, double3 inv_ext, uint32_t* __restrict__ keys, int32_t* __restrict__ ids) {
    int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (f >= nF) return;
    const int32_t* fc = faces + 3 * f;
    d3 a = ld3(verts + 3 * (int64_t)fc[0]), b = ld3(verts + 3 * (int64_t)fc[1]), c = ld3(verts + 3 * (int64_t)fc[2]);
    double cx = 0.5 * (fmin(a.x, fmin(b.x, c.x)) + fmax(a.x, fmax(b.x, c.x)));
    double cy = 0.5 * (fmin(a.y, fmin(b.y, c.y)) + fmax(a.y, fmax(b.y, c.y)));
    double cz = 0.5 * (fmin(a.z, fmin(b.z, c.z)) + fmax(a.z, fmax(b.z, c.z)));
    uint32_t ix = (uint32_t)fmin(fmax((cx - bmin.x) * inv_ext.x * 1024.0, 0.0), 1023.0);
    uint32_t iy = (uint32_t)fmin(fmax((cy - bmin.y) * inv_ext.y * 1024.0, 0.0), 1023.0);
    uint32_t iz = (uint32_t)fmin(fmax((cz - bmin.z) * inv_ext.z * 1024.0, 0.0), 1023.0);
    keys[f] = (expand_bits10(ix) << 2) | (expand_bits10(iy) << 1) | expand_bits10(iz);
    ids[f] = (int32_t)f;
}

// per leaf (Morton order): exact-test records, f32 leaf boxes (rounded outward + pad), areas in caller order
__global__ void leaf_kernel(const double* __restrict__ verts, const double* __restrict__ vnormals, const int32_t* __restrict__ faces,
                            const int32_t* __restrict__ sorted_ids, int64_t nF, float pad,
                            double* __restrict__ tri_rec, double* __restrict__ tri_verts, double* __restrict__ tri_vvn, float4* __restrict__ tri_f32,
                            int32_t* __restrict__ leaf_tri,
                            float* __restrict__ box_min, float* __restrict__ box_max, double* __restrict__ area) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nF) return;
    int32_t f = sorted_ids[i];
    const int32_t* fc = faces + 3 * (int64_t)f;
    d3 v0 = ld3(verts + 3 * (int64_t)fc[0]), v1 = ld3(verts + 3 * (int64_t)fc[1]), v2 = ld3(verts + 3 * (int64_t)fc[2]);
    d3 e0 = sub3(v1, v0), e1 = sub3(v2, v0);
    d3 cr = cross3(e0, e1);
    double nn = norm3(cr);                       // trimesh util.unitize: v * (1 / sqrt(dot(v*v, ones)))
    d3 n = mk3(0.0, 0.0, 0.0);
    if (nn > 1e-13) { double inv = 1.0 / nn; n = scl3(cr, inv); }
    double d00 = dot3(e0, e0), d01 = dot3(e0, e1), d11 = dot3(e1, e1);
    double inv_den = 1.0 / (d00 * d11 - d01 * d01);
    double* r = tri_rec + 16 * i;
    st3(r + 0, v0); st3(r + 3, n); st3(r + 6, e0); st3(r + 9, e1);
    r[12] = d00; r[13] = d01; r[14] = d11; r[15] = inv_den;
    double* tv = tri_verts + 9 * i;
    st3(tv + 0, v0); st3(tv + 3, v1); st3(tv + 6, v2);
    double* tn = tri_vvn + 20 * i;
    st3(tn + 0, v0); st3(tn + 3, v1); st3(tn + 6, v2); tn[9] = 0.0;
    st3(tn + 10, ld3(vnormals + 3 * (int64_t)fc[0])); st3(tn + 13, ld3(vnormals + 3 * (int64_t)fc[1])); st3(tn + 16, ld3(vnormals + 3 * (int64_t)fc[2]));
    tn[19] = 0.0;
    tri_f32[3 * i + 0] = make_float4((float)v0.x, (float)v0.y, (float)v0.z, 0.f);
    tri_f32[3 * i + 1] = make_float4((float)e0.x, (float)e0.y, (float)e0.z, 0.f);
    tri_f32[3 * i + 2] = make_float4((float)e1.x, (float)e1.y, (float)e1.z, 0.f);
    leaf_tri[i] = f;
    area[f] = 0.5 * nn;
    int64_t slot = (nF - 1) + i;                 // leaves follow the nF-1 internal boxes
    box_min[3 * slot + 0] = __double2float_rd(fmin(v0.x, fmin(v1.x, v2.x))) - pad;
    box_min[3 * slot + 1] = __double2float_rd(fmin(v0.y, fmin(v1.y, v2.y))) - pad;
    box_min[3 * slot + 2] = __double2float_rd(fmin(v0.z, fmin(v1.z, v2.z))) - pad;
    box_max[3 * slot + 0] = __double2float_ru(fmax(v0.x, fmax(v1.x, v2.x))) + pad;
    box_max[3 * slot + 1] = __double2float_ru(fmax(v0.y, fmax(v1.y, v2.y))) + pad;
    box_max[3 * slot + 2] = __double2float_ru(fmax(v0.z, fmax(v1.z, v2.z))) + pad;
}

__device__ __forceinline__ int delta_fn(const uint32_t* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    uint32_t a = keys[i], b = keys[j];
    if (a != b) return __clz(a ^ b);
    return 32 + __clz((uint32_t)i ^ (uint32_t)j);     // duplicate codes: fall back to the sorted position
}

// Karras 2012, one thread per internal node.  child code: >= 0 internal node, < 0 leaf ~code.
__global__ void hierarchy_kernel(const uint32_t* __restrict__ keys, int n, int32_t* __restrict__ left,
                                 int32_t* __restrict__ right, int32_t* __restrict__ parent) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int d = (delta_fn(keys, n, i, i + 1) - delta_fn(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = delta_fn(keys, n, i, i - d);
    int lmax = 2;
    while (delta_fn(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (delta_fn(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = delta_fn(keys, n, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (delta_fn(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    int32_t lc = (lo == gamma) ? ~gamma : gamma;
    int32_t rc = (hi == gamma + 1) ? ~(gamma + 1) : (gamma + 1);
    left[i] = lc;
    right[i] = rc;
    parent[lc < 0 ? (n - 1) + (~lc) : lc] = i;
    parent[rc < 0 ? (n - 1) + (~rc) : rc] = i;
    if (i == 0) parent[0] = -1;
}


// ---- PLOC topology (Meister & Bittner 2018: parallel locally-ordered clustering) --------------------------------------
// The Karras hierarchy splits the Morton order at bit boundaries, whatever the boxes look like.  PLOC builds the tree
// bottom-up instead: the clusters (leaves at first) stay in Morton order; every cluster looks at its PLOC_RADIUS
// neighbours on either side for the partner that gives the smallest merged surface area, mutual choices are merged, and
// the survivors are compacted in order -- until one cluster is left.  The ray caster visits 12 % fewer wide nodes per
// ray on such a tree (tools/bvh_lab.cpp, config-3 object: 17.8 -> 15.6 fetches); results do not depend on the tree.
// Deterministic: the choice is a pure function of the boxes with a total order on the pairs (area, lower index, upper
// index), ids are handed out by prefix sums.  Internal ids are allocated from n - 2 downwards, iteration by iteration, so
// the root is node 0 and the array runs top-down, every level along the Morton curve.
constexpr int PLOC_RADIUS = 8;
struct PBox { float lo[3], hi[3]; };

__device__ __forceinline__ double merged_area(const PBox& a, const PBox& b) {
    const double x = (double)fmaxf(a.hi[0], b.hi[0]) - (double)fminf(a.lo[0], b.lo[0]);
    const double y = (double)fmaxf(a.hi[1], b.hi[1]) - (double)fminf(a.lo[1], b.lo[1]);
    const double z = (double)fmaxf(a.hi[2], b.hi[2]) - (double)fminf(a.lo[2], b.lo[2]);
    return x * y + y * z + z * x;
}

__global__ void ploc_init_kernel(int n, const float* __restrict__ box_min, const float* __restrict__ box_max,
                                 int32_t* __restrict__ cid, PBox* __restrict__ cb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int slot = (n - 1) + i;
    PBox b;
#pragma unroll
    for (int k = 0; k < 3; ++k) { b.lo[k] = box_min[3 * slot + k]; b.hi[k] = box_max[3 * slot + k]; }
    cid[i] = ~i;
    cb[i] = b;
}

// state[0] = clusters left, state[1] = lowest internal id handed out so far; the merge rounds run without the host
__global__ void ploc_nn_kernel(const int* __restrict__ state, const PBox* __restrict__ cb, int32_t* __restrict__ nn) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int m = state[0];
    if (i >= m) return;
    const PBox me = cb[i];
    double best = 0.0;
    int bj = -1;
    const int j0 = max(0, i - PLOC_RADIUS), j1 = min(m - 1, i + PLOC_RADIUS);
    for (int j = j0; j <= j1; ++j) {
        if (j == i) continue;
        const double d = merged_area(me, cb[j]);
        // total order on the pairs: (d, lower index, upper index) -- the same from either side, so the best pair overall
        // is always mutual and every iteration merges at least once
        bool better = bj < 0 || d < best;
        if (!better && d == best) {
            const int lo = min(i, j), hi = max(i, j), blo = min(i, bj), bhi = max(i, bj);
            better = lo < blo || (lo == blo && hi < bhi);
        }
        if (better) { best = d; bj = j; }
    }
    nn[i] = bj;
}

// flag: low word = the cluster (or its merge result) stays in the list, high word = it creates a node
__global__ void ploc_flag_kernel(int n, const int* __restrict__ state, const int32_t* __restrict__ nn, unsigned long long* __restrict__ flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i >= state[0]) { flag[i] = 0ull; return; }          // the scan runs over all n entries
    const int j = nn[i];
    const bool mutual = j >= 0 && nn[j] == i;
    flag[i] = ((mutual && i > j) ? 0ull : 1ull) | ((mutual && i < j) ? (1ull << 32) : 0ull);
}

__global__ void ploc_emit_kernel(int n, const int* __restrict__ state, const int32_t* __restrict__ nn, const unsigned long long* __restrict__ flag,
                                 const unsigned long long* __restrict__ pos, const int32_t* __restrict__ cid, const PBox* __restrict__ cb,
                                 int32_t* __restrict__ cid2, PBox* __restrict__ cb2, int32_t* __restrict__ left,
                                 int32_t* __restrict__ right, int32_t* __restrict__ parent) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int m = state[0];
    if (i >= m) return;
    const int made = (int)((pos[m - 1] + flag[m - 1]) >> 32);
    const int id_base = state[1] - made;                     // this round's nodes: [id_base, id_base + made), along the curve
    const unsigned long long f = flag[i], p = pos[i];
    if (!(f & 1ull)) return;                                // the upper half of a merged pair: its partner writes the node
    const int at = (int)(p & 0xffffffffull);
    if (f >> 32) {
        const int j = nn[i];
        const int id = id_base + (int)(p >> 32);
        const int32_t a = cid[i], b = cid[j];
        const PBox x = cb[i], y = cb[j];
        PBox u;
#pragma unroll
        for (int k = 0; k < 3; ++k) { u.lo[k] = fminf(x.lo[k], y.lo[k]); u.hi[k] = fmaxf(x.hi[k], y.hi[k]); }
        left[id] = a; right[id] = b;
        parent[a < 0 ? (n - 1) + (~a) : a] = id;
        parent[b < 0 ? (n - 1) + (~b) : b] = id;
        cid2[at] = id; cb2[at] = u;
    } else {
        cid2[at] = cid[i]; cb2[at] = cb[i];
    }
}

// after a round: the new cluster count and id mark (a separate launch: every thread of the round has read the old ones)
__global__ void ploc_advance_kernel(int* __restrict__ state, const unsigned long long* __restrict__ flag, const unsigned long long* __restrict__ pos) {
    const int m = state[0];
    const unsigned long long tot = pos[m - 1] + flag[m - 1];
    state[0] = (int)(tot & 0xffffffffull);
    state[1] -= (int)(tot >> 32);
    state[2] += m > 1 ? 1 : 0;                               // rounds that merged something
}

__global__ void ploc_root_kernel(int32_t* __restrict__ parent) { parent[0] = -1; }

__device__ __forceinline__ int box_slot(int32_t code, int n) { return code < 0 ? (n - 1) + (~code) : code; }

__global__ void refit_kernel(int n, const int32_t* __restrict__ left, const int32_t* __restrict__ right,
                             const int32_t* __restrict__ parent, float* box_min, float* box_max,
                             int* __restrict__ arrivals, int* __restrict__ max_depth) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int depth = 0;
    for (int p = parent[(n - 1) + i]; p >= 0; p = parent[p]) ++depth;
    atomicMax(max_depth, depth);
    int cur = parent[(n - 1) + i];
    while (cur >= 0) {
        __threadfence();
        if (atomicAdd(&arrivals[cur], 1) == 0) return;       // first arrival: sibling subtree not done yet
        __threadfence();
        int a = box_slot(left[cur], n), b = box_slot(right[cur], n);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float lo = fminf(__ldcg(box_min + 3 * a + k), __ldcg(box_min + 3 * b + k));
            float hi = fmaxf(__ldcg(box_max + 3 * a + k), __ldcg(box_max + 3 * b + k));
            __stcg(box_min + 3 * cur + k, lo);
            __stcg(box_max + 3 * cur + k, hi);
        }
        cur = parent[cur];
    }
}

__global__ void emit_nodes_kernel(int n, const int32_t* __restrict__ left, const int32_t* __restrict__ right,
                                  const float* __restrict__ box_min, const float* __restrict__ box_max,
                                  float4* __restrict__ nodes) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int a = box_slot(left[i], n), b = box_slot(right[i], n);
    float4 q0 = make_float4(box_min[3 * a], box_min[3 * a + 1], box_min[3 * a + 2], box_max[3 * a]);
    float4 q1 = make_float4(box_max[3 * a + 1], box_max[3 * a + 2], box_min[3 * b], box_min[3 * b + 1]);
    float4 q2 = make_float4(box_min[3 * b + 2], box_max[3 * b], box_max[3 * b + 1], box_max[3 * b + 2]);
    float4 q3 = make_float4(__int_as_float(left[i]), __int_as_float(right[i]), 0.f, 0.f);
    nodes[4 * i + 0] = q0; nodes[4 * i + 1] = q1; nodes[4 * i + 2] = q2; nodes[4 * i + 3] = q3;
}

__device__ __forceinline__ unsigned quant_lo(float x, double gmin, double gscale) {
    const double g = floor(((double)x - gmin) * gscale) - 1.0;
    return (unsigned)fmin(fmax(g, 0.0), 65535.0);
}
__device__ __forceinline__ unsigned quant_hi(float x, double gmin, double gscale) {
    const double g = ceil(((double)x - gmin) * gscale) + 1.0;
    return (unsigned)fmin(fmax(g, 0.0), 65535.0);
}

// float32 nodes -> 32-byte nodes with uint16 boxes on the global grid, rounded outward (+1 cell)
__global__ void quantize_nodes_kernel(int n_nodes, const float4* __restrict__ nodes, double3 gmin, double3 gscale,
                                      uint4* __restrict__ qnodes) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    const float4 q0 = nodes[4 * i], q1 = nodes[4 * i + 1], q2 = nodes[4 * i + 2], q3 = nodes[4 * i + 3];
    uint4 A, B;
    A.x = quant_lo(q0.x, gmin.x, gscale.x) | (quant_lo(q0.y, gmin.y, gscale.y) << 16);
    A.y = quant_lo(q0.z, gmin.z, gscale.z) | (quant_hi(q0.w, gmin.x, gscale.x) << 16);
    A.z = quant_hi(q1.x, gmin.y, gscale.y) | (quant_hi(q1.y, gmin.z, gscale.z) << 16);
    A.w = quant_lo(q1.z, gmin.x, gscale.x) | (quant_lo(q1.w, gmin.y, gscale.y) << 16);
    B.x = quant_lo(q2.x, gmin.z, gscale.z) | (quant_hi(q2.y, gmin.x, gscale.x) << 16);
    B.y = quant_hi(q2.z, gmin.y, gscale.y) | (quant_hi(q2.w, gmin.z, gscale.z) << 16);
    B.z = __float_as_uint(q3.x);
    B.w = __float_as_uint(q3.y);
    if (q1.z > q2.y) { A.w = 0xffffffffu; B.x = 0x0000ffffu; B.y = 0u; }      // empty right box (single-triangle mesh) stays empty
    qnodes[2 * i] = A;
    qnodes[2 * i + 1] = B;
}

// 4-wide LBVH for the ray caster: every internal node of EVEN depth becomes a wide node whose children are its
// grandchildren (a leaf child of the binary node stays a direct child), i.e. every other level of the binary tree is
// skipped and a ray's chain of dependent node fetches is halved.  The wide node keeps the binary node's index (the array
// is sparse: odd-depth entries are unused), so no allocation pass is needed and the build stays deterministic.
//   64 bytes = 4 x uint4: words 3k..3k+2 = child k's box as uint16 on the global grid (rounded outward by one cell):
//       (lo.x | hi.x << 16,  lo.y | hi.y << 16,  lo.z | hi.z << 16) -- both planes of an axis in one word, so that the
//       traversal picks a ray's near / far plane with one byte-permute each;   words 12..15 = child codes
//   child code: >= 0 wide node (binary index), < 0 leaf ~code, BT_WIDE_EMPTY = no child
// flag[i] = 1 iff internal node i has even depth (it becomes a wide node); an exclusive scan of the flags gives the wide
// nodes dense indices in the order of the binary indices (= along the Morton curve), so the array has no holes
__global__ void wide_flags_kernel(int n, const int32_t* __restrict__ parent, int32_t* __restrict__ flag) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int depth = 0;
    for (int p = parent[i]; p >= 0; p = parent[p]) ++depth;
    flag[i] = (depth & 1) ? 0 : 1;
}

// Packed child code (meshes below 2^23 triangles): low 24 bits = the code sign-extended (inner: wide index, leaf: ~leaf,
// empty: BT_WIDE_EMPTY_PACKED), top byte = one byte of the NODE's own normal cone (word 12: axis x, 13: axis y, 14: axis z,
// 15: half-angle code), see pack_cone.  Empty slots carry an inverted box (lo = 65535, hi = 0 on every axis): no ray can
// hit it, so the traversal needs no test for them.
__device__ __forceinline__ unsigned pack_code(int32_t code, unsigned cone_byte) { return ((unsigned)code & 0x00ffffffu) | (cone_byte << 24); }

// Normal cone (unit axis a, half-angle th) -> 32 bits: axis as three biased int8 (q + 128), half-angle as an 8-bit code in
// steps of pi/255, rounded UP after adding the angle between the true and the quantised axis; 255 = "never cull".
__device__ unsigned pack_cone(float ax, float ay, float az, float th) {
    if (!(th >= 0.f) || !(th < 3.1f) || !(ax == ax) || !(ay == ay) || !(az == az)) return 0xff808080u;   // code 255 = never cull
    const int qx = max(-127, min(127, __float2int_rn(ax * 127.f))), qy = max(-127, min(127, __float2int_rn(ay * 127.f))),
              qz = max(-127, min(127, __float2int_rn(az * 127.f)));
    const float ql = sqrtf((float)(qx * qx + qy * qy + qz * qz));
    if (ql < 64.f) return 0xff808080u;                  // the axis was not a unit vector
    const float c = ((float)qx * ax + (float)qy * ay + (float)qz * az) / ql;
    const float err = acosf(fminf(1.f, fmaxf(-1.f, c)));
    const float total = th + err + 2e-3f;
    const int code = (int)ceilf(total * (255.f / 3.14159265f));
    return (unsigned)(qx + 128) | ((unsigned)(qy + 128) << 8) | ((unsigned)(qz + 128) << 16) | ((unsigned)min(255, max(0, code)) << 24);
}

__global__ void emit_wide_nodes_kernel(int n, const int32_t* __restrict__ left, const int32_t* __restrict__ right,
                                       const int32_t* __restrict__ flag, const int32_t* __restrict__ widx,
                                       const float* __restrict__ box_min, const float* __restrict__ box_max, double3 gmin,
                                       double3 gscale, const float4* __restrict__ cone, int packed, uint4* __restrict__ wnodes) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    if (!flag[i]) return;
    int32_t child[4];
    int nc = 0;
    const int32_t two[2] = {left[i], right[i]};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        if (two[k] < 0) child[nc++] = two[k];
        else { child[nc++] = left[two[k]]; child[nc++] = right[two[k]]; }
    }
    unsigned cone_word = 0xff808080u;
    if (packed && i != 0) { const float4 c = cone[i]; cone_word = pack_cone(c.x, c.y, c.z, c.w); }      // the root is always walked
    unsigned w[16];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k < nc) {
            const int sl = box_slot(child[k], n);
            w[3 * k + 0] = quant_lo(box_min[3 * sl], gmin.x, gscale.x) | (quant_hi(box_max[3 * sl], gmin.x, gscale.x) << 16);
            w[3 * k + 1] = quant_lo(box_min[3 * sl + 1], gmin.y, gscale.y) | (quant_hi(box_max[3 * sl + 1], gmin.y, gscale.y) << 16);
            w[3 * k + 2] = quant_lo(box_min[3 * sl + 2], gmin.z, gscale.z) | (quant_hi(box_max[3 * sl + 2], gmin.z, gscale.z) << 16);
            const int32_t code = child[k] < 0 ? child[k] : widx[child[k]];          // inner children: dense wide index
            w[12 + k] = packed ? pack_code(code, (cone_word >> (8 * k)) & 0xffu) : (unsigned)code;
        } else {
            w[3 * k + 0] = w[3 * k + 1] = w[3 * k + 2] = 0x0000ffffu;               // inverted box: never hit
            w[12 + k] = packed ? pack_code(BT_WIDE_EMPTY_PACKED, (cone_word >> (8 * k)) & 0xffu) : (unsigned)BT_WIDE_EMPTY;
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) wnodes[4 * (int64_t)widx[i] + q] = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
}

// ---- normal cones: which directions can a VALID antipodal hit on this subtree have? ----------------------------------
// A hit is valid only if the interpolated vertex normal n at the hit lies within atan(mu) of the ray direction
// (sampling.py:289-308); n is a positive combination of the triangle's three vertex normals, so it lies inside any
// circular cone (half-angle < 90 deg) that holds them.  The lean ray caster uses a cone per leaf and per node to skip
// what cannot hold a valid hit.  The skipped hits must not change what happens to the valid ones: the only coupling
// between hits of a ray is trimesh's 1e-8 de-duplication, which can only merge hits on triangles whose (padded) boxes
// touch -- so a leaf's cone is widened to hold the vertex normals of EVERY triangle whose box overlaps its own
// (found by a walk of the finished LBVH).  Then: valid hit on B found => every hit that can share its location is found.
//   cone[slot] = (axis.xyz, half-angle); slot numbering as for the boxes (internal nodes, then leaves)
__device__ __forceinline__ float angle_between(float ax, float ay, float az, float bx, float by, float bz) {
    // atan2(|a x b|, a . b): accurate for small angles, where acos of a float32 dot product is not
    const float cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx;
    return atan2f(sqrtf(cx * cx + cy * cy + cz * cz), ax * bx + ay * by + az * bz);
}

__global__ void leaf_cone_kernel(int n, const float4* __restrict__ nodes, const double* __restrict__ tri_vvn,
                                 const float* __restrict__ box_min, const float* __restrict__ box_max, float grow,
                                 float4* __restrict__ cone) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // own axis: normalised sum of the (normalised) vertex normals
    float vn[3][3];
    float sx = 0.f, sy = 0.f, sz = 0.f;
    int n_ok = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double* p = tri_vvn + 20 * (int64_t)i + 10 + 3 * k;
        const float x = (float)p[0], y = (float)p[1], z = (float)p[2];
        const float l = sqrtf(x * x + y * y + z * z);
        const bool ok = (l > 1e-30f) && (l < 1e30f);           // zero / NaN / inf normals contribute nothing to an interpolated normal that can pass the filters
        vn[k][0] = ok ? x / l : 0.f; vn[k][1] = ok ? y / l : 0.f; vn[k][2] = ok ? z / l : 0.f;
        sx += vn[k][0]; sy += vn[k][1]; sz += vn[k][2];
        n_ok += ok ? 1 : 0;
    }
    const int slot = (n - 1) + i;
    const float sl = sqrtf(sx * sx + sy * sy + sz * sz);
    if (n_ok < 3 || !(sl > 1e-3f)) { cone[slot] = make_float4(1.f, 0.f, 0.f, 4.f); return; }      // degenerate: never culled
    const float ax = sx / sl, ay = sy / sl, az = sz / sl;
    float th = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) th = fmaxf(th, angle_between(ax, ay, az, vn[k][0], vn[k][1], vn[k][2]));
    // neighbours: every leaf whose box overlaps this leaf's box grown by `grow`
    const float lx = box_min[3 * slot] - grow, ly = box_min[3 * slot + 1] - grow, lz = box_min[3 * slot + 2] - grow;
    const float hx = box_max[3 * slot] + grow, hy = box_max[3 * slot + 1] + grow, hz = box_max[3 * slot + 2] + grow;
    if (n > 1) {
        int stack[64];
        int sp = 0, cur = 0;
        bool bad = false;
        while (true) {
            if (cur >= 0) {
                const float4 q0 = nodes[4 * (int64_t)cur], q1 = nodes[4 * (int64_t)cur + 1], q2 = nodes[4 * (int64_t)cur + 2], q3 = nodes[4 * (int64_t)cur + 3];
                const bool ol = (q0.x <= hx) & (q0.w >= lx) & (q0.y <= hy) & (q1.x >= ly) & (q0.z <= hz) & (q1.y >= lz);
                const bool orr = (q1.z <= hx) & (q2.y >= lx) & (q1.w <= hy) & (q2.z >= ly) & (q2.x <= hz) & (q2.w >= lz);
                const int l = __float_as_int(q3.x), r = __float_as_int(q3.y);
                if (ol && orr) { if (sp >= 64) { bad = true; break; } stack[sp++] = r; cur = l; continue; }
                if (ol) { cur = l; continue; }
                if (orr) { cur = r; continue; }
            } else {
                const int j = ~cur;
                if (j != i) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const double* p = tri_vvn + 20 * (int64_t)j + 10 + 3 * k;
                        const float x = (float)p[0], y = (float)p[1], z = (float)p[2];
                        const float l = sqrtf(x * x + y * y + z * z);
                        if ((l > 1e-30f) && (l < 1e30f)) th = fmaxf(th, angle_between(ax, ay, az, x / l, y / l, z / l));
                    }
                }
            }
            if (sp == 0) break;
            cur = stack[--sp];
        }
        if (bad) th = 4.f;
    }
    // the "positive combinations stay inside" argument needs a convex cone, i.e. a half-angle below 90 degrees
    cone[slot] = make_float4(ax, ay, az, th < 1.5f ? th + 1e-4f : 4.f);
}

// bottom-up merge (same arrival-counter scheme as refit_kernel): axis = normalised sum of the children's axes,
// half-angle = the larger of (angle to a child's axis + that child's half-angle)
__global__ void cone_refit_kernel(int n, const int32_t* __restrict__ left, const int32_t* __restrict__ right,
                                  const int32_t* __restrict__ parent, float4* cone, int* __restrict__ arrivals) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int cur = parent[(n - 1) + i];
    while (cur >= 0) {
        __threadfence();
        if (atomicAdd(&arrivals[cur], 1) == 0) return;
        __threadfence();
        const float4 a = __ldcg(cone + box_slot(left[cur], n)), b = __ldcg(cone + box_slot(right[cur], n));
        float4 m = make_float4(1.f, 0.f, 0.f, 4.f);
        const float sx = a.x + b.x, sy = a.y + b.y, sz = a.z + b.z;
        const float sl = sqrtf(sx * sx + sy * sy + sz * sz);
        if (a.w < 3.2f && b.w < 3.2f && sl > 1e-3f) {
            const float ax = sx / sl, ay = sy / sl, az = sz / sl;
            const float th = fmaxf(angle_between(ax, ay, az, a.x, a.y, a.z) + a.w, angle_between(ax, ay, az, b.x, b.y, b.z) + b.w) + 1e-4f;
            m = make_float4(ax, ay, az, th);
        }
        __stcg(cone + cur, m);
        cur = parent[cur];
    }
}

__global__ void pack_leaf_cones_kernel(int n, const float4* __restrict__ cone, uint32_t* __restrict__ leaf_cone) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 c = cone[(n - 1) + i];
    leaf_cone[i] = pack_cone(c.x, c.y, c.z, c.w);
}

// nF == 1: wide node 0 = (leaf 0 | empty | empty | empty)
__global__ void single_leaf_wide_node_kernel(const float* __restrict__ box_min, const float* __restrict__ box_max,
                                             double3 gmin, double3 gscale, int packed, uint4* __restrict__ wnodes) {
    const unsigned w0 = quant_lo(box_min[0], gmin.x, gscale.x) | (quant_hi(box_max[0], gmin.x, gscale.x) << 16);
    const unsigned w1 = quant_lo(box_min[1], gmin.y, gscale.y) | (quant_hi(box_max[1], gmin.y, gscale.y) << 16);
    const unsigned w2 = quant_lo(box_min[2], gmin.z, gscale.z) | (quant_hi(box_max[2], gmin.z, gscale.z) << 16);
    const unsigned inv = 0x0000ffffu;                  // inverted box (lo = 65535, hi = 0): never hit
    wnodes[0] = make_uint4(w0, w1, w2, inv);
    wnodes[1] = make_uint4(inv, inv, inv, inv);
    wnodes[2] = make_uint4(inv, inv, inv, inv);
    if (packed) {                                      // cone bytes: axis (128, 128, 128) is irrelevant with half-angle code 255 = never cull
        const unsigned e = pack_code(BT_WIDE_EMPTY_PACKED, 0x80u);
        wnodes[3] = make_uint4(pack_code(~0, 0x80u), e, e, pack_code(BT_WIDE_EMPTY_PACKED, 0xffu));
    } else
        wnodes[3] = make_uint4((unsigned)~0, (unsigned)BT_WIDE_EMPTY, (unsigned)BT_WIDE_EMPTY, (unsigned)BT_WIDE_EMPTY);
}

__global__ void single_leaf_node_kernel(const float* __restrict__ box_min, const float* __restrict__ box_max,
                                        float4* __restrict__ nodes) {
    // nF == 1: node 0 = (leaf 0 | empty box)
    nodes[0] = make_float4(box_min[0], box_min[1], box_min[2], box_max[0]);
    nodes[1] = make_float4(box_max[1], box_max[2], 1e30f, 1e30f);
    nodes[2] = make_float4(1e30f, -1e30f, -1e30f, -1e30f);
    nodes[3] = make_float4(__int_as_float(~0), __int_as_float(~0), 0.f, 0.f);
}

__global__ void normalise_cdf_kernel(double* cdf, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double total = cdf[n - 1];
    if (i < n - 1) cdf[i] = cdf[i] / total;     // last element fixed up afterwards (it is read by everyone)
}
__global__ void finish_cdf_kernel(double* cdf, int64_t n, double* total_out) {
    *total_out = cdf[n - 1];
    cdf[n - 1] = 1.0;
}

}  // namespace bt

using namespace bt;

unsigned long long* bt::g_counters_host_copy = nullptr;

extern "C" int bt_debug_set_counters(uint64_t* d_counters) {
    bt::g_counters_host_copy = reinterpret_cast<unsigned long long*>(d_counters);
    return BT_OK;
}

// L2 read-bandwidth probe: every SM streams the same L2-resident buffer (ld.global.cg: cached in L2 only) `iters` times.
// The ray caster's working set (mesh + LBVH, <= ~50 MB for a 100k-triangle object) lives in L2, so this -- not HBM -- is
// the memory roof its traffic model is compared with (bench.py: roofline.binding).
namespace bt {
__global__ void l2_probe_kernel(const uint4* __restrict__ buf, int64_t n_vec, int iters, uint4* __restrict__ sink) {
    uint4 acc = make_uint4(0u, 0u, 0u, 0u);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int it = 0; it < iters; ++it)
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_vec; i += stride) {
            const uint4 v = __ldcg(buf + i);
            acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
        }
    if (acc.x == 0x12345678u && acc.y == 0x9abcdef0u) *sink = acc;       // never true for a zeroed buffer: keeps the loads alive
}
}  // namespace bt

extern "C" int bt_debug_l2_probe(int device, int64_t bytes, int32_t iters, double* gbytes_per_s) {
    BT_REQUIRE(gbytes_per_s && bytes >= (1 << 20) && iters >= 1, "bad arguments");
    BT_CUDA(cudaSetDevice(device));
    int sm_count = 148;
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device);
    uint4* buf = nullptr;
    BT_CUDA(cudaMalloc(&buf, (size_t)bytes + sizeof(uint4)));
    cudaError_t e = cudaMemset(buf, 0, (size_t)bytes + sizeof(uint4));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (e == cudaSuccess) e = cudaEventCreate(&e0);
    if (e == cudaSuccess) e = cudaEventCreate(&e1);
    const int64_t n_vec = bytes / (int64_t)sizeof(uint4);
    float best = 1e30f;
    if (e == cudaSuccess) {
        bt::l2_probe_kernel<<<sm_count * 8, 256>>>(buf, n_vec, 1, buf + n_vec);          // warm L2
        for (int rep = 0; rep < 3 && e == cudaSuccess; ++rep) {
            cudaEventRecord(e0);
            bt::l2_probe_kernel<<<sm_count * 8, 256>>>(buf, n_vec, iters, buf + n_vec);
            cudaEventRecord(e1);
            e = cudaEventSynchronize(e1);
            float ms = 0.f;
            if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
            if (ms > 0.f && ms < best) best = ms;
        }
    }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(buf);
    if (e != cudaSuccess) return cuda_fail(e, "bt_debug_l2_probe", __FILE__, __LINE__);
    *gbytes_per_s = (double)bytes * iters / (best * 1e-3) / 1e9;
    return BT_OK;
}

// ---- library ---------------------------------------------------------------------------------------
extern "C" int bt_version(void) { return 100; }

extern "C" const char* bt_last_error(void) { return bt::g_err; }

extern "C" int bt_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, int64_t* total_mem) {
    cudaDeviceProp p;
    BT_CUDA(cudaGetDeviceProperties(&p, device));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (total_mem) *total_mem = (int64_t)p.totalGlobalMem;
    return BT_OK;
}

extern "C" int bt_malloc(int device, int64_t bytes, void** d_ptr) {
    BT_REQUIRE(d_ptr && bytes >= 0, "bad arguments");
    BT_CUDA(cudaSetDevice(device));
    BT_CUDA(cudaMalloc(d_ptr, bytes ? bytes : 1));
    return BT_OK;
}
extern "C" int bt_free(void* d_ptr) {
    BT_CUDA(cudaFree(d_ptr));
    return BT_OK;
}
extern "C" int bt_memcpy_h2d(void* d_dst, const void* h_src, int64_t bytes, void* stream) {
    BT_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return BT_OK;
}
extern "C" int bt_memcpy_d2h(void* h_dst, const void* d_src, int64_t bytes, void* stream) {
    BT_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return BT_OK;
}
extern "C" int bt_stream_sync(void* stream) {
    BT_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return BT_OK;
}

// ---- mesh ------------------------------------------------------------------------------------------

// which topologies bt_mesh_create builds: 0 = Morton hierarchy for everything, 1 (default) = PLOC for the ray tree, and for the
// walk tree from PLOC_WALK_MIN_TRIS triangles on, 2 = PLOC for everything; BT_BVH_BUILD=lbvh|ploc|ploc_all sets the initial value
constexpr int64_t PLOC_WALK_MIN_TRIS = 1 << 18;
static int g_bvh_build = -1;
static int bvh_build_kind() {
    if (g_bvh_build < 0) {
        const char* e = getenv("BT_BVH_BUILD");
        g_bvh_build = (e && (!strcmp(e, "lbvh") || !strcmp(e, "0"))) ? 0 : (e && (!strcmp(e, "ploc_all") || !strcmp(e, "2"))) ? 2 : 1;
    }
    return g_bvh_build;
}

extern "C" int bt_set_bvh_build(int32_t kind) {
    BT_REQUIRE(kind >= 0 && kind <= 2, "kind must be 0 (Morton hierarchy), 1 (PLOC ray tree, size-dependent walk tree) or 2 (PLOC for both)");
    g_bvh_build = kind;
    return BT_OK;
}

extern "C" int bt_mesh_bvh_kind(const bt_mesh* mesh, int32_t* ray_kind, int32_t* walk_kind, int32_t* build_iterations) {
    BT_REQUIRE(mesh && ray_kind, "null pointer");
    *ray_kind = mesh->m.bvh_kind;
    if (walk_kind) *walk_kind = mesh->m.walk_kind;
    if (build_iterations) *build_iterations = mesh->m.ploc_iterations;
    return BT_OK;
}

// PLOC topology on the device: left / right / parent as hierarchy_kernel writes them (root = node 0).  The merge rounds
// (~60 for 10^5 triangles, ~70 for 10^6) are queued without waiting for one another -- the cluster count lives in device
// memory, a finished build idles through the rest of its batch of rounds -- and the host looks at the count once per batch.
static cudaError_t ploc_topology(int n, const float* box_min, const float* box_max, int32_t* left, int32_t* right, int32_t* parent,
                                 void* cub_tmp, size_t cub_bytes, int* iterations, cudaStream_t st) {
    char* slab = nullptr;
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t sz_cid = up(sizeof(int32_t) * n), sz_cb = up(sizeof(PBox) * n), sz_u64 = up(sizeof(unsigned long long) * n);
    retain_pool_memory();
    cudaError_t e = cudaMallocAsync(&slab, 3 * sz_cid + 2 * sz_cb + 2 * sz_u64 + 256, st);
    if (e != cudaSuccess) return e;
    int32_t* cid[2] = {reinterpret_cast<int32_t*>(slab), reinterpret_cast<int32_t*>(slab + sz_cid)};
    int32_t* nn = reinterpret_cast<int32_t*>(slab + 2 * sz_cid);
    PBox* cb[2] = {reinterpret_cast<PBox*>(slab + 3 * sz_cid), reinterpret_cast<PBox*>(slab + 3 * sz_cid + sz_cb)};
    unsigned long long* flag = reinterpret_cast<unsigned long long*>(slab + 3 * sz_cid + 2 * sz_cb);
    unsigned long long* pos = reinterpret_cast<unsigned long long*>(slab + 3 * sz_cid + 2 * sz_cb + sz_u64);
    int* state = reinterpret_cast<int*>(slab + 3 * sz_cid + 2 * sz_cb + 2 * sz_u64);
#define PLOC_TRY(call) do { e = (call); if (e != cudaSuccess) { cudaFreeAsync(slab, st); return e; } } while (0)
    const int T = 256, G = (int)ceil_div(n, T);
    const int init[3] = {n, n - 1, 0};
    PLOC_TRY(cudaMemcpyAsync(state, init, sizeof(init), cudaMemcpyHostToDevice, st));
    ploc_init_kernel<<<G, T, 0, st>>>(n, box_min, box_max, cid[0], cb[0]);
    PLOC_TRY(cudaGetLastError());
    int cur = 0, host_state[3] = {n, n - 1, 0};
    for (int batch = 0; host_state[0] > 1; ++batch) {
        if (batch > 4096) { cudaFreeAsync(slab, st); return cudaErrorUnknown; }      // cannot happen: every round merges at least the best pair
        const int rounds = 16;
        for (int r = 0; r < rounds; ++r) {
            ploc_nn_kernel<<<G, T, 0, st>>>(state, cb[cur], nn);
            ploc_flag_kernel<<<G, T, 0, st>>>(n, state, nn, flag);
            size_t bytes = cub_bytes;
            PLOC_TRY(cub::DeviceScan::ExclusiveSum(cub_tmp, bytes, flag, pos, n, st));
            ploc_emit_kernel<<<G, T, 0, st>>>(n, state, nn, flag, pos, cid[cur], cb[cur], cid[cur ^ 1], cb[cur ^ 1], left, right, parent);
            ploc_advance_kernel<<<1, 1, 0, st>>>(state, flag, pos);
            PLOC_TRY(cudaGetLastError());
            cur ^= 1;
        }
        PLOC_TRY(cudaMemcpyAsync(host_state, state, sizeof(host_state), cudaMemcpyDeviceToHost, st));
        PLOC_TRY(cudaStreamSynchronize(st));
    }
    if (host_state[0] != 1 || host_state[1] != 0) { cudaFreeAsync(slab, st); return cudaErrorUnknown; }
    ploc_root_kernel<<<1, 1, 0, st>>>(parent);
    PLOC_TRY(cudaGetLastError());
    PLOC_TRY(cudaStreamSynchronize(st));
#undef PLOC_TRY
    if (iterations) *iterations = host_state[2];
    cudaFreeAsync(slab, st);
    return cudaSuccess;
}

static int mesh_free(MeshDev& m) {
    cudaFree(m.verts); cudaFree(m.vnormals); cudaFree(m.faces); cudaFree(m.area_cdf);
    cudaFree(m.nodes); cudaFree(m.qnodes); cudaFree(m.wnodes); cudaFree(m.tri_rec); cudaFree(m.tri_verts); cudaFree(m.tri_vvn); cudaFree(m.tri_f32); cudaFree(m.leaf_tri); cudaFree(m.leaf_cone);
    m = MeshDev();
    return BT_OK;
}

extern "C" int bt_mesh_create(const double* h_V, int64_t nV, const int32_t* h_F, int64_t nF,
                              const double* h_VN, int device, bt_mesh** out) {
    BT_REQUIRE(h_V && h_F && out, "null pointer");
    BT_REQUIRE(nV > 0 && nF > 0, "empty mesh");
    BT_REQUIRE(nF < (1ll << 30), "too many triangles");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("bt_mesh_create: no CUDA device available (this library has no CPU fallback)");
        return BT_E_NODEVICE;
    }
    BT_CUDA(cudaSetDevice(device));
    for (int64_t i = 0; i < 3 * nF; ++i)
        if (h_F[i] < 0 || h_F[i] >= nV) { set_error("bt_mesh_create: face index out of range"); return BT_E_INVALID; }

    // host-side set-up: bounds, optional area-weighted vertex normals (io.load_mesh convention, io.py:96)
    double lo[3] = {h_V[0], h_V[1], h_V[2]}, hi[3] = {h_V[0], h_V[1], h_V[2]};
    for (int64_t v = 0; v < nV; ++v)
        for (int k = 0; k < 3; ++k) {
            double x = h_V[3 * v + k];
            if (!(x == x) || fabs(x) > 1e30) { set_error("bt_mesh_create: non-finite vertex"); return BT_E_INVALID; }
            lo[k] = x < lo[k] ? x : lo[k];
            hi[k] = x > hi[k] ? x : hi[k];
        }
    std::vector<double> vn_host;
    if (!h_VN) {
        vn_host.assign(3 * nV, 0.0);
        std::vector<double> cr(3 * nF);
        for (int64_t f = 0; f < nF; ++f) {
            d3 a = ld3(h_V + 3 * (int64_t)h_F[3 * f]), b = ld3(h_V + 3 * (int64_t)h_F[3 * f + 1]), c = ld3(h_V + 3 * (int64_t)h_F[3 * f + 2]);
            st3(&cr[3 * f], cross3(sub3(b, a), sub3(c, a)));
        }
        for (int k = 0; k < 3; ++k)                     // same accumulation order as np.add.at(vn, F[:, k], cross)
            for (int64_t f = 0; f < nF; ++f)
                for (int c = 0; c < 3; ++c) vn_host[3 * (int64_t)h_F[3 * f + k] + c] += cr[3 * f + c];
        for (int64_t v = 0; v < nV; ++v) {
            double nn = norm3(ld3(&vn_host[3 * v]));
            if (nn > 0) { for (int c = 0; c < 3; ++c) vn_host[3 * v + c] /= nn; }
            else { vn_host[3 * v] = 0; vn_host[3 * v + 1] = 0; vn_host[3 * v + 2] = 1; }
        }
        h_VN = vn_host.data();
    }

    bt_mesh* h = new bt_mesh();
    MeshDev& m = h->m;
    m.device = device; m.nV = nV; m.nF = nF;
    const int64_t nNodes = nF > 1 ? nF - 1 : 1;
    cudaStream_t st = nullptr;
#define MESH_TRY(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { int rc = cuda_fail(_e, #call, __FILE__, __LINE__); mesh_free(m); delete h; return rc; } } while (0)
    MESH_TRY(cudaMalloc(&m.verts, sizeof(double) * 3 * nV));
    MESH_TRY(cudaMalloc(&m.vnormals, sizeof(double) * 3 * nV));
    MESH_TRY(cudaMalloc(&m.faces, sizeof(int32_t) * 3 * nF));
    MESH_TRY(cudaMalloc(&m.area_cdf, sizeof(double) * nF));
    MESH_TRY(cudaMalloc(&m.nodes, sizeof(float4) * 4 * nNodes));
    MESH_TRY(cudaMalloc(&m.qnodes, sizeof(uint4) * 2 * nNodes));
    MESH_TRY(cudaMalloc(&m.wnodes, sizeof(uint4) * 4 * nNodes));
    MESH_TRY(cudaMalloc(&m.tri_rec, sizeof(double) * 16 * nF));
    MESH_TRY(cudaMalloc(&m.tri_verts, sizeof(double) * 9 * nF));
    MESH_TRY(cudaMalloc(&m.tri_vvn, sizeof(double) * 20 * nF));
    MESH_TRY(cudaMalloc(&m.tri_f32, sizeof(float4) * 3 * nF));
    MESH_TRY(cudaMalloc(&m.leaf_tri, sizeof(int32_t) * nF));
    MESH_TRY(cudaMalloc(&m.leaf_cone, sizeof(uint32_t) * nF));
    MESH_TRY(cudaMemcpy(m.verts, h_V, sizeof(double) * 3 * nV, cudaMemcpyHostToDevice));
    MESH_TRY(cudaMemcpy(m.vnormals, h_VN, sizeof(double) * 3 * nV, cudaMemcpyHostToDevice));
    MESH_TRY(cudaMemcpy(m.faces, h_F, sizeof(int32_t) * 3 * nF, cudaMemcpyHostToDevice));

    cudaEvent_t ev0, ev1;
    MESH_TRY(cudaEventCreate(&ev0));
    MESH_TRY(cudaEventCreate(&ev1));
    MESH_TRY(cudaEventRecord(ev0, st));
    // BT_BUILD_TIMING=1: per-phase times of the build on stderr (events at the phase boundaries)
    static const bool phase_timing = getenv("BT_BUILD_TIMING") && atoi(getenv("BT_BUILD_TIMING")) != 0;
    cudaEvent_t pe[8] = {};
    const char* pe_name[8] = {"sort + leaf records", "topology + refit", "nodes + quantised nodes", "leaf cones", "cone refit + pack", "wide nodes", "area cdf", ""};
    int n_pe = 0;
    auto phase = [&]() { if (phase_timing && n_pe < 8 && cudaEventCreate(&pe[n_pe]) == cudaSuccess) { cudaEventRecord(pe[n_pe], st); ++n_pe; } };

    // scratch
    uint32_t *keys = nullptr, *keys_sorted = nullptr;
    int32_t *ids = nullptr, *ids_sorted = nullptr, *left = nullptr, *right = nullptr, *parent = nullptr;
    float *box_min = nullptr, *box_max = nullptr;
    int *arrivals = nullptr, *max_depth = nullptr;
    double *area = nullptr, *total_area = nullptr;
    float4* cone = nullptr;
    void* cub_tmp = nullptr;
    size_t cub_bytes = 0, cub_bytes2 = 0;
    int ploc_iterations = 0, depth_walk = 0, depth_ray = 0;
    bool walk_is_ploc = false, ray_is_ploc = false;
    const int n = (int)nF;
    // build scratch comes from the device's stream-ordered pool (kept across builds, see retain_pool_memory): fourteen
    // cudaMalloc / cudaFree pairs per mesh cost more than the build itself
    retain_pool_memory();
    MESH_TRY(cudaMallocAsync(&keys, sizeof(uint32_t) * nF, st));
    MESH_TRY(cudaMallocAsync(&keys_sorted, sizeof(uint32_t) * nF, st));
    MESH_TRY(cudaMallocAsync(&ids, sizeof(int32_t) * nF, st));
    MESH_TRY(cudaMallocAsync(&ids_sorted, sizeof(int32_t) * nF, st));
    MESH_TRY(cudaMallocAsync(&left, sizeof(int32_t) * nNodes, st));
    MESH_TRY(cudaMallocAsync(&right, sizeof(int32_t) * nNodes, st));
    MESH_TRY(cudaMallocAsync(&parent, sizeof(int32_t) * (2 * nF), st));
    MESH_TRY(cudaMallocAsync(&box_min, sizeof(float) * 3 * 2 * nF, st));
    MESH_TRY(cudaMallocAsync(&box_max, sizeof(float) * 3 * 2 * nF, st));
    MESH_TRY(cudaMallocAsync(&arrivals, sizeof(int) * nNodes, st));
    MESH_TRY(cudaMallocAsync(&max_depth, sizeof(int), st));
    MESH_TRY(cudaMallocAsync(&area, sizeof(double) * nF, st));
    MESH_TRY(cudaMallocAsync(&total_area, sizeof(double), st));
    MESH_TRY(cudaMallocAsync(&cone, sizeof(float4) * 2 * nF, st));
    MESH_TRY(cudaMemsetAsync(arrivals, 0, sizeof(int) * nNodes, st));
    MESH_TRY(cudaMemsetAsync(max_depth, 0, sizeof(int), st));
    MESH_TRY(cudaMemsetAsync(parent, 0xFF, sizeof(int32_t) * (2 * nF), st));
    MESH_TRY(cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, keys, keys_sorted, ids, ids_sorted, n, 0, 30, st));
    MESH_TRY(cub::DeviceScan::InclusiveSum(nullptr, cub_bytes2, area, m.area_cdf, n, st));
    if (cub_bytes2 > cub_bytes) cub_bytes = cub_bytes2;
    MESH_TRY(cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes2, ids, ids_sorted, n, st));      // wide-node index scan
    if (cub_bytes2 > cub_bytes) cub_bytes = cub_bytes2;
    MESH_TRY(cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes2, (unsigned long long*)nullptr, (unsigned long long*)nullptr, n, st));   // PLOC flags
    if (cub_bytes2 > cub_bytes) cub_bytes = cub_bytes2;
    MESH_TRY(cudaMallocAsync(&cub_tmp, cub_bytes ? cub_bytes : 1, st));

    double ext[3], inv_ext[3], max_abs = 0.0;
    for (int k = 0; k < 3; ++k) {
        ext[k] = hi[k] - lo[k];
        inv_ext[k] = ext[k] > 0 ? 1.0 / ext[k] : 0.0;
        max_abs = fmax(max_abs, fmax(fabs(lo[k]), fabs(hi[k])));
    }
    // conservative padding of every box: covers float32 rounding of the slab arithmetic and the 1e-13 /
    // 1e-6 tolerances of the exact tests (DESIGN.md "broad phase is result-neutral")
    const float pad = (float)(2e-5 * max_abs + 1e-9);
    const int T = 256;
    const int G = (int)ceil_div(nF, T);
    morton_kernel<<<G, T, 0, st>>>(m.verts, m.faces, nF, make_double3(lo[0], lo[1], lo[2]),
                                   make_double3(inv_ext[0], inv_ext[1], inv_ext[2]), keys, ids);
    MESH_TRY(cudaGetLastError());
    MESH_TRY(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, keys, keys_sorted, ids, ids_sorted, n, 0, 30, st));
    leaf_kernel<<<G, T, 0, st>>>(m.verts, m.vnormals, m.faces, ids_sorted, nF, pad, m.tri_rec, m.tri_verts, m.tri_vvn, m.tri_f32, m.leaf_tri,
                                 box_min, box_max, area);
    MESH_TRY(cudaGetLastError());
    phase();
    // Two consumers, two topologies (bt_set_bvh_build / BT_BVH_BUILD): the RAY tree under the 4-wide nodes is PLOC by default
    // (fewer node fetches; the traversal is bound by them), the WALK tree of the binary nodes (collision walks, mesh pairs,
    // leaf-cone neighbours, fall-back ray walk) is Karras' shallower Morton hierarchy for object-sized meshes -- those walks are
    // bound by their longest chain of dependent passes, i.e. by depth -- and PLOC from PLOC_WALK_MIN_TRIS triangles on, where
    // the working set outgrows L2 and fewer fetches win again.  A PLOC tree deeper than the stacks allow falls back to Karras.
    const int mode = bvh_build_kind();
    const bool ray_wants_ploc = mode >= 1, walk_wants_ploc = mode == 2 || (mode == 1 && nF >= PLOC_WALK_MIN_TRIS);
    auto build_topology = [&](bool ploc, bool* got_ploc, int* depth_out) -> cudaError_t {
        cudaError_t e;
        while (true) {
            if ((e = cudaMemsetAsync(arrivals, 0, sizeof(int) * nNodes, st)) != cudaSuccess) return e;
            if ((e = cudaMemsetAsync(max_depth, 0, sizeof(int), st)) != cudaSuccess) return e;
            if ((e = cudaMemsetAsync(parent, 0xFF, sizeof(int32_t) * (2 * nF), st)) != cudaSuccess) return e;
            if (ploc) { if ((e = ploc_topology(n, box_min, box_max, left, right, parent, cub_tmp, cub_bytes, &ploc_iterations, st)) != cudaSuccess) return e; }
            else {
                hierarchy_kernel<<<(int)ceil_div(nF - 1, T), T, 0, st>>>(keys_sorted, n, left, right, parent);
                if ((e = cudaGetLastError()) != cudaSuccess) return e;
            }
            refit_kernel<<<G, T, 0, st>>>(n, left, right, parent, box_min, box_max, arrivals, max_depth);
            if ((e = cudaGetLastError()) != cudaSuccess) return e;
            int d = 0;
            if ((e = cudaMemcpyAsync(&d, max_depth, sizeof(int), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
            if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
            *depth_out = d;
            if (!ploc || d <= 60) break;
            ploc = false;                               // too deep for the traversal stacks: the Morton hierarchy instead
        }
        *got_ploc = ploc;
        return cudaSuccess;
    };
    if (nF > 1) {
        MESH_TRY(build_topology(walk_wants_ploc, &walk_is_ploc, &depth_walk));
        phase();
        emit_nodes_kernel<<<(int)ceil_div(nF - 1, T), T, 0, st>>>(n, left, right, box_min, box_max, m.nodes);
        MESH_TRY(cudaGetLastError());
    } else {
        single_leaf_node_kernel<<<1, 1, 0, st>>>(box_min, box_max, m.nodes);
        MESH_TRY(cudaGetLastError());
    }
    {
        const double margin = 2.0 * (double)pad + 1e-7;
        for (int k = 0; k < 3; ++k) {
            m.grid_min[k] = lo[k] - margin;
            m.grid_scale[k] = 65534.0 / ((hi[k] + margin) - m.grid_min[k]);
        }
        quantize_nodes_kernel<<<(int)ceil_div(nNodes, T), T, 0, st>>>(
            (int)nNodes, m.nodes, make_double3(m.grid_min[0], m.grid_min[1], m.grid_min[2]),
            make_double3(m.grid_scale[0], m.grid_scale[1], m.grid_scale[2]), m.qnodes);
        MESH_TRY(cudaGetLastError());
        const double3 gm = make_double3(m.grid_min[0], m.grid_min[1], m.grid_min[2]);
        const double3 gs = make_double3(m.grid_scale[0], m.grid_scale[1], m.grid_scale[2]);
        // normal cones of the leaves (widened over every triangle whose box touches the leaf's) and of the internal nodes
        m.wide_packed = nF < BT_WIDE_EMPTY_PACKED ? 1 : 0;         // 24-bit child codes + cone bytes; larger meshes: plain codes, no culling
        phase();
        leaf_cone_kernel<<<G, T, 0, st>>>(n, m.nodes, m.tri_vvn, box_min, box_max, 1e-7f + 2.0f * pad, cone);
        MESH_TRY(cudaGetLastError());
        phase();
        if (nF > 1) {
            // the ray tree: the same topology again, or its own (left / right / parent and the internal boxes are rebuilt; the
            // binary nodes emitted above keep the walk tree)
            ray_is_ploc = walk_is_ploc; depth_ray = depth_walk;
            if (ray_wants_ploc != walk_is_ploc && !(ray_wants_ploc && walk_wants_ploc))
                MESH_TRY(build_topology(ray_wants_ploc, &ray_is_ploc, &depth_ray));
            MESH_TRY(cudaMemsetAsync(arrivals, 0, sizeof(int) * nNodes, st));
            cone_refit_kernel<<<G, T, 0, st>>>(n, left, right, parent, cone, arrivals);
            MESH_TRY(cudaGetLastError());
        }
        pack_leaf_cones_kernel<<<G, T, 0, st>>>(n, cone, m.leaf_cone);
        MESH_TRY(cudaGetLastError());
        phase();
        if (nF > 1) {
            // ids / ids_sorted (n int32 each) are free again at this point: reuse them for the flags and their scan
            int32_t* flag = reinterpret_cast<int32_t*>(ids);
            int32_t* widx = reinterpret_cast<int32_t*>(ids_sorted);
            wide_flags_kernel<<<(int)ceil_div(nF - 1, T), T, 0, st>>>(n, parent, flag);
            MESH_TRY(cudaGetLastError());
            MESH_TRY(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, flag, widx, n - 1, st));
            emit_wide_nodes_kernel<<<(int)ceil_div(nF - 1, T), T, 0, st>>>(n, left, right, flag, widx, box_min, box_max, gm, gs, cone, m.wide_packed, m.wnodes);
        } else
            single_leaf_wide_node_kernel<<<1, 1, 0, st>>>(box_min, box_max, gm, gs, m.wide_packed, m.wnodes);
        MESH_TRY(cudaGetLastError());
    }
    phase();
    MESH_TRY(cub::DeviceScan::InclusiveSum(cub_tmp, cub_bytes, area, m.area_cdf, n, st));
    normalise_cdf_kernel<<<G, T, 0, st>>>(m.area_cdf, nF);
    MESH_TRY(cudaGetLastError());
    finish_cdf_kernel<<<1, 1, 0, st>>>(m.area_cdf, nF, total_area);
    MESH_TRY(cudaGetLastError());
    MESH_TRY(cudaEventRecord(ev1, st));
    MESH_TRY(cudaStreamSynchronize(st));
    float ms = 0.f;
    MESH_TRY(cudaEventElapsedTime(&ms, ev0, ev1));
    if (phase_timing) {
        cudaEvent_t prev = ev0;
        fprintf(stderr, "bt_mesh_create: %lld triangles, ray tree %s, walk tree %s", (long long)nF, ray_is_ploc ? "PLOC" : "Morton hierarchy",
                walk_is_ploc ? "PLOC" : "Morton hierarchy");
        for (int k = 0; k < n_pe; ++k) { float t = 0.f; cudaEventElapsedTime(&t, prev, pe[k]); fprintf(stderr, " | %s %.3f ms", pe_name[k], t); prev = pe[k]; }
        float t = 0.f; cudaEventElapsedTime(&t, prev, ev1);
        fprintf(stderr, " | area cdf %.3f ms | total %.3f ms\n", t, ms);
    }
    for (int k = 0; k < n_pe; ++k) cudaEventDestroy(pe[k]);
    const int depth = depth_walk > depth_ray ? depth_walk : depth_ray;      // the stacks of both kinds of walks are sized by it
    double total = 0.0;
    MESH_TRY(cudaMemcpy(&total, total_area, sizeof(double), cudaMemcpyDeviceToHost));
    cudaEventDestroy(ev0); cudaEventDestroy(ev1);
    for (void* p : {(void*)keys, (void*)keys_sorted, (void*)ids, (void*)ids_sorted, (void*)left, (void*)right, (void*)parent, (void*)box_min,
                    (void*)box_max, (void*)arrivals, (void*)max_depth, (void*)area, (void*)total_area, cub_tmp, (void*)cone})
        cudaFreeAsync(p, st);
#undef MESH_TRY
    if (depth > 60) {
        set_error("bt_mesh_create: LBVH depth %d exceeds the traversal stack (64)", depth);
        mesh_free(m); delete h;
        return BT_E_CAPACITY;
    }
    m.root = 0;
    m.info.n_vertices = nV; m.info.n_triangles = nF; m.info.n_nodes = nNodes;
    m.info.device_bytes = (int64_t)(sizeof(double) * (6 * nV + 46 * nF) + sizeof(int32_t) * 4 * nF + 48 * nF + 160 * nNodes);
    m.info.surface_area = total;
    m.info.build_ms = ms;
    for (int k = 0; k < 3; ++k) { m.info.bounds_min[k] = (float)lo[k]; m.info.bounds_max[k] = (float)hi[k]; }
    m.info.device = device;
    m.info.max_depth = depth;
    m.bvh_kind = ray_is_ploc ? 1 : 0;
    m.walk_kind = walk_is_ploc ? 1 : 0;
    m.ploc_iterations = ploc_iterations;
    *out = h;
    return BT_OK;
}

extern "C" int bt_mesh_destroy(bt_mesh* mesh) {
    if (!mesh) return BT_OK;
    cudaSetDevice(mesh->m.device);
    mesh_free(mesh->m);
    delete mesh;
    return BT_OK;
}

extern "C" int bt_mesh_get_info(const bt_mesh* mesh, bt_mesh_info* info) {
    BT_REQUIRE(mesh && info, "null pointer");
    *info = mesh->m.info;
    return BT_OK;
}

extern "C" int bt_mesh_copy_bvh(const bt_mesh* mesh, float* h_nodes, int32_t* h_leaf_tri) {
    BT_REQUIRE(mesh, "null pointer");
    BT_CUDA(cudaSetDevice(mesh->m.device));
    if (h_nodes) BT_CUDA(cudaMemcpy(h_nodes, mesh->m.nodes, sizeof(float) * 16 * mesh->m.info.n_nodes, cudaMemcpyDeviceToHost));
    if (h_leaf_tri) BT_CUDA(cudaMemcpy(h_leaf_tri, mesh->m.leaf_tri, sizeof(int32_t) * mesh->m.nF, cudaMemcpyDeviceToHost));
    return BT_OK;
}

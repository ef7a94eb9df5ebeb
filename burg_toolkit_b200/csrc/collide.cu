// collide.cu -- gripper-vs-scene collision (K1) and stream compaction of the surviving grasps.
//
// Reference being replaced: AntipodalGraspSampler.check_collisions (burg_toolkit/sampling.py:354-397), i.e. one
// trimesh CollisionManager.in_collision_single -> FCL mesh/mesh query per grasp in a Python loop.  FCL's answer
// is "does any triangle of the posed gripper mesh intersect any scene triangle" (surface test: containment is not
// a collision) decided by the 17-axis separating-axis test of Intersect::intersect_Triangle / project6.
//
// Here: one thread per (grasp, gripper part).  A part is a group of gripper triangles with its TCP-frame AABB; under
// pose * diag(s,1,1) it is an exact oriented box in the world, which is walked down the scene LBVH with a
// conservative float32 box test; at the leaves the float64 17-axis test runs on the world-space vertices.
#include <cub/cub.cuh>
#include <math.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "tritri.cuh"

namespace bt {

constexpr int CSTACK = 64;
constexpr int MAX_PARTS = 64;

// One traversal item of the gripper: a (sub-)box of a part in the TCP frame plus the part it belongs to.  Long thin parts
// (fingers, the stem) are cut into segments along their longest axis: a short box overlaps far fewer LBVH nodes than a
// long one grazing the surface, which removes the heavy tail of the per-item work distribution.  The exact test always
// uses ALL triangles of the part; the segment only drives the conservative broad phase.
struct PartDev {
    float cx, cy, cz;      // segment centre, TCP frame
    float hx, hy, hz;      // segment half extents
    float ox, oy, oz;      // segment centre minus full-part centre, TCP frame
    float fhx, fhy, fhz;   // full part half extents
    int32_t tri_begin, tri_end;
    int32_t is_box;        // the part's triangles are exactly the 12 faces of its (full) box: a closed surface
};

struct Obb {
    float cx, cy, cz;
    float ux[3], uy[3], uz[3];     // world-space axes (columns of R)
    float h[3];                    // half extents along the axes (inflated)
    float ax, ay, az;              // half extents of the OBB's world AABB
};

// conservative overlap: 3 world axes + 3 box axes (the 9 cross axes are skipped -> false positives only)
__device__ __forceinline__ bool obb_overlaps_aabb(const Obb& b, float x0, float y0, float z0, float x1, float y1, float z1) {
    const float ex = 0.5f * (x1 - x0), ey = 0.5f * (y1 - y0), ez = 0.5f * (z1 - z0);
    const float dx = 0.5f * (x1 + x0) - b.cx, dy = 0.5f * (y1 + y0) - b.cy, dz = 0.5f * (z1 + z0) - b.cz;
    if (fabsf(dx) > ex + b.ax || fabsf(dy) > ey + b.ay || fabsf(dz) > ez + b.az) return false;
    float pr;
    pr = fabsf(fmaf(dx, b.ux[0], fmaf(dy, b.ux[1], dz * b.ux[2])));
    if (pr > b.h[0] + fmaf(ex, fabsf(b.ux[0]), fmaf(ey, fabsf(b.ux[1]), ez * fabsf(b.ux[2])))) return false;
    pr = fabsf(fmaf(dx, b.uy[0], fmaf(dy, b.uy[1], dz * b.uy[2])));
    if (pr > b.h[1] + fmaf(ex, fabsf(b.uy[0]), fmaf(ey, fabsf(b.uy[1]), ez * fabsf(b.uy[2])))) return false;
    pr = fabsf(fmaf(dx, b.uz[0], fmaf(dy, b.uz[1], dz * b.uz[2])));
    if (pr > b.h[2] + fmaf(ex, fabsf(b.uz[0]), fmaf(ey, fabsf(b.uz[1]), ez * fabsf(b.uz[2])))) return false;
    return true;
}

// Conservative float32 classification of a scene triangle against the part's (solid) oriented box, in box-local
// coordinates q (metres along the three box axes):
//   0 = certainly disjoint from the box (some separating axis found, extents inflated by `m`)
//   1 = certainly strictly inside the box (all vertices inside the box shrunk by `m`)  -- only meaningful for is_box parts:
//       a triangle inside a closed box cannot touch the box's surface triangles
//   2 = undecided -> run the exact float64 test
//   3 = certainly crossing the box's surface: one vertex strictly inside the box shrunk by `m`, another strictly outside
//       the box grown by `m`, so an edge of the triangle pierces one of the box's surface triangles (is_box parts only:
//       the 12 triangles cover the whole surface).  The exact test of that pair finds a proper crossing with a margin
//       many orders of magnitude above float64 rounding, i.e. it is certain to report the collision.
__device__ __forceinline__ int classify_triangle(const float q[3][3], float e0, float e1, float e2, float m,
                                                 const float off[3], const float fe[3]) {
    const float E[3] = {e0 + m, e1 + m, e2 + m};
    bool inside = true;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float lo = fminf(q[0][k], fminf(q[1][k], q[2][k])), hi = fmaxf(q[0][k], fmaxf(q[1][k], q[2][k]));
        if (lo > E[k] || hi < -E[k]) return 0;
        // 'inside' is judged against the FULL part box (coordinates shifted by the segment offset), shrunk by m
        inside = inside && (lo + off[k] > -(fe[k] - m)) && (hi + off[k] < (fe[k] - m));
    }
    if (inside) return 1;
    {
        bool any_in = false, any_out = false;
#pragma unroll
        for (int v = 0; v < 3; ++v) {
            bool vin = true, vout = false;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float c = fabsf(q[v][k] + off[k]);
                vin = vin && (c < fe[k] - m);
                vout = vout || (c > fe[k] + m);
            }
            any_in = any_in || vin;
            any_out = any_out || vout;
        }
        if (any_in && any_out) return 3;
    }
    // triangle plane
    const float f0[3] = {q[1][0] - q[0][0], q[1][1] - q[0][1], q[1][2] - q[0][2]};
    const float f1[3] = {q[2][0] - q[1][0], q[2][1] - q[1][1], q[2][2] - q[1][2]};
    const float f2[3] = {q[0][0] - q[2][0], q[0][1] - q[2][1], q[0][2] - q[2][2]};
    const float n[3] = {f0[1] * f1[2] - f0[2] * f1[1], f0[2] * f1[0] - f0[0] * f1[2], f0[0] * f1[1] - f0[1] * f1[0]};
    const float nn = fabsf(n[0]) + fabsf(n[1]) + fabsf(n[2]);
    const float dist = n[0] * q[0][0] + n[1] * q[0][1] + n[2] * q[0][2];
    const float rad = E[0] * fabsf(n[0]) + E[1] * fabsf(n[1]) + E[2] * fabsf(n[2]);
    if (fabsf(dist) > rad + 1e-6f * nn) return 0;
    // 9 edge-cross axes  a = e_i x f_j  (e_i = unit box axes)
    const float* F[3] = {f0, f1, f2};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float fx = F[j][0], fy = F[j][1], fz = F[j][2];
        {   // e_0 x f = (0, -fz, fy)
            const float p0 = -fz * q[0][1] + fy * q[0][2], p1 = -fz * q[1][1] + fy * q[1][2], p2 = -fz * q[2][1] + fy * q[2][2];
            const float r = E[1] * fabsf(fz) + E[2] * fabsf(fy) + 1e-6f * (fabsf(fy) + fabsf(fz));
            if (fminf(p0, fminf(p1, p2)) > r || fmaxf(p0, fmaxf(p1, p2)) < -r) return 0;
        }
        {   // e_1 x f = (fz, 0, -fx)
            const float p0 = fz * q[0][0] - fx * q[0][2], p1 = fz * q[1][0] - fx * q[1][2], p2 = fz * q[2][0] - fx * q[2][2];
            const float r = E[0] * fabsf(fz) + E[2] * fabsf(fx) + 1e-6f * (fabsf(fx) + fabsf(fz));
            if (fminf(p0, fminf(p1, p2)) > r || fmaxf(p0, fmaxf(p1, p2)) < -r) return 0;
        }
        {   // e_2 x f = (-fy, fx, 0)
            const float p0 = -fy * q[0][0] + fx * q[0][1], p1 = -fy * q[1][0] + fx * q[1][1], p2 = -fy * q[2][0] + fx * q[2][1];
            const float r = E[0] * fabsf(fy) + E[1] * fabsf(fx) + 1e-6f * (fabsf(fx) + fabsf(fy));
            if (fminf(p0, fminf(p1, p2)) > r || fmaxf(p0, fmaxf(p1, p2)) < -r) return 0;
        }
    }
    return 2;
}

// ---- persistent collision kernel --------------------------------------------------------------------------------
// Work item = (gripper part, grasp).  Traversal lengths differ by orders of magnitude between items (a stem far from the
// object ends at the root, a finger grazing the surface visits hundreds of nodes), so items are NOT bound to threads:
// warps are persistent and every lane fetches a new item from a global counter as soon as enough lanes of its warp have
// run dry.  The three kinds of work have very different instruction streams, so they never share a divergent branch:
//   walk     (per lane)      the part's oriented box walks the scene LBVH (float32, conservative 6-axis test) in short
//                            bursts; a reached leaf is only APPENDED to the warp's classification queue (shared memory).
//   classify (warp, 32 wide) once 32 leaves are queued every lane takes one (item, leaf) entry -- of ANY lane's item: the
//                            box is rebuilt from the grasp row -- and classifies the triangle in box-local float32
//                            coordinates: disjoint / inside-a-closed-box / undecided.  Undecided entries go to the exact queue.
//   exact    (half warps)    two queued entries per pass: lane l of a half tests gripper triangle l of the part with FCL's
//                            float64 17-axis test.  A hit sets out[row] and tells the lane still walking that item to stop.
#ifndef LEAN_MINB
#define LEAN_MINB 5
#endif
constexpr int WU = 1;                      // tasks per lane and pass
constexpr int TQ_CAP = 512;               // task stack entries per warp
constexpr int CWARPS = 4;                  // warps per block
#ifndef COLLIDE_ADAPT_ROUNDS
#define COLLIDE_ADAPT_ROUNDS 4         // the fetch size adapts while fewer than this many full rounds of batches are left in the queue
#endif
#ifndef COLLIDE_ADAPT
#define COLLIDE_ADAPT 14               // target passes per batch of the adaptive fetch size (0: always 32 items).  Measured at
                                       // config 3, collide stage in ms for 5k / 10k / 20k / 40k reference points: off 0.270 / 0.271 / 0.288 / 0.483,
                                       // 16: 0.241 / 0.240 / 0.293 / 0.486, 12: 0.233 / 0.204 / 0.318 / 0.506
#endif
constexpr int CQ_CAP = 32 + 64 * WU;        // classification queue: < 32 left over + at most 64 * WU new entries per pass
constexpr int XQ_CAP = 64;                 // exact queue: <= 1 left over + 32 new entries

struct LaneBox {
    float u[3][3];                         // world axes (columns of R), un-normalised
    float h[3], a[3];                      // half extents: along the axes (projection units) / of the world AABB, inflated
    float inv_n[3], le[3];                 // 1/|axis|, half extents in metres
    float off[3], fe[3];                   // segment offset inside the full part box / full half extents, metres
    float cf[3];                           // centre (float32)
    double c[3];                           // centre (float64, for the leaf classifier)
    float margin;
};

// pose * diag(s,1,1): M (row-major 3x3) and T of grasp `row`  (sampling.py:392-395; s in float32 as numpy >= 2 does)
__device__ __forceinline__ void grasp_affine(const float* __restrict__ gs, int64_t row, float ow, float tol, int use_width,
                                             double M[9], double T[3], float& sf_out) {
    const float* g = gs + 14 * row;
    const float sf = use_width ? (__ldg(g + 13) + tol) / ow : 1.0f;
    const double s = (double)sf;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        M[3 * i + 0] = (double)__ldg(g + 3 + 3 * i) * s;
        M[3 * i + 1] = (double)__ldg(g + 4 + 3 * i);
        M[3 * i + 2] = (double)__ldg(g + 5 + 3 * i);
        T[i] = (double)__ldg(g + i);
    }
    sf_out = sf;
}

// the oriented box of traversal item (grasp row, part segment); callers use a subset of the fields (the rest is dead code)
__device__ __forceinline__ void build_lanebox(LaneBox& b, const float* __restrict__ gs, int64_t row, const PartDev& pd,
                                              float ow, float tol, int use_width, float inflate) {
    double M[9], T[3];
    float sf;
    grasp_affine(gs, row, ow, tol, use_width, M, T, sf);
    const float* g = gs + 14 * row;
#pragma unroll
    for (int j = 0; j < 3; ++j) { b.u[j][0] = __ldg(g + 3 + j); b.u[j][1] = __ldg(g + 6 + j); b.u[j][2] = __ldg(g + 9 + j); }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        b.c[i] = (M[3 * i] * pd.cx + M[3 * i + 1] * pd.cy) + M[3 * i + 2] * pd.cz + T[i];
        b.cf[i] = (float)b.c[i];
    }
    const float ext[3] = {pd.hx * fabsf(sf), pd.hy, pd.hz};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        // a float32 rotation block is orthonormal only to ~1e-7: use the actual column norms
        const float nn = sqrtf(b.u[j][0] * b.u[j][0] + b.u[j][1] * b.u[j][1] + b.u[j][2] * b.u[j][2]);
        b.h[j] = ext[j] * nn * nn * 1.0001f + inflate;
        b.inv_n[j] = nn > 0.f ? 1.0f / nn : 0.f;
        b.le[j] = ext[j] * nn;
        const float sc = (j == 0) ? fabsf(sf) : 1.0f;
        b.off[j] = (j == 0 ? pd.ox : (j == 1 ? pd.oy : pd.oz)) * (j == 0 ? sf : 1.0f) * nn;
        b.fe[j] = (j == 0 ? pd.fhx : (j == 1 ? pd.fhy : pd.fhz)) * sc * nn;
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
        b.a[a] = (fabsf(b.u[0][a]) * ext[0] + fabsf(b.u[1][a]) * ext[1] + fabsf(b.u[2][a]) * ext[2]) * 1.0001f + inflate;
    b.margin = inflate + 1e-4f * fmaxf(b.le[0], fmaxf(b.le[1], b.le[2]));
}

// The traversal box of an item in plain float32 and without divisions / square roots (no slow-path subroutine calls in
// the refill code): the walk is a conservative broad phase, so the squeeze factor and the centre may carry float32
// rounding (~1e-8 m), which the inflation (>= 1e-6 m) covers.  Layout: u[9] (axes, [j][k]), h[3], a[3], cf[3].
struct WalkBox { float u[3][3], h[3], a[3], cf[3]; };

__device__ __forceinline__ void build_walkbox(WalkBox& b, const float* __restrict__ gs, int64_t row, const PartDev& pd,
                                              float inv_ow, float tol, int use_width, float inflate) {
    const float* g = gs + 14 * row;
    const float sf = use_width ? (__ldg(g + 13) + tol) * inv_ow : 1.0f;
#pragma unroll
    for (int j = 0; j < 3; ++j) { b.u[j][0] = __ldg(g + 3 + j); b.u[j][1] = __ldg(g + 6 + j); b.u[j][2] = __ldg(g + 9 + j); }
    const float px = pd.cx * sf;
#pragma unroll
    for (int i = 0; i < 3; ++i) b.cf[i] = fmaf(b.u[0][i], px, fmaf(b.u[1][i], pd.cy, fmaf(b.u[2][i], pd.cz, __ldg(g + i))));
    const float ext[3] = {pd.hx * fabsf(sf) * 1.000001f, pd.hy, pd.hz};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float nn2 = b.u[j][0] * b.u[j][0] + b.u[j][1] * b.u[j][1] + b.u[j][2] * b.u[j][2];
        b.h[j] = ext[j] * nn2 * 1.0001f + inflate;
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
        b.a[a] = (fabsf(b.u[0][a]) * ext[0] + fabsf(b.u[1][a]) * ext[1] + fabsf(b.u[2][a]) * ext[2]) * 1.0001f + inflate;
}

__device__ __forceinline__ uint8_t load_flag(const uint8_t* p) { return *reinterpret_cast<const volatile uint8_t*>(p); }

struct CollideParams { float ow, tol, inflate; int use_width; };

// classify scene triangle `tv` against the box of item (row, part): 0 = no contact, 2 = undecided (run the exact test),
// 3 = certain collision
__device__ __forceinline__ int leaf_class(const float* __restrict__ gs, int row, const PartDev* pd_ptr, const double* __restrict__ tv,
                                            CollideParams cp) {
    const PartDev pd = *pd_ptr;
    LaneBox cb;
    build_lanebox(cb, gs, row, pd, cp.ow, cp.tol, cp.use_width, cp.inflate);
    float q[3][3];
#pragma unroll
    for (int v = 0; v < 3; ++v) {
        const float dx = (float)(tv[3 * v] - cb.c[0]), dy = (float)(tv[3 * v + 1] - cb.c[1]), dz = (float)(tv[3 * v + 2] - cb.c[2]);
#pragma unroll
        for (int j = 0; j < 3; ++j) q[v][j] = fmaf(dx, cb.u[j][0], fmaf(dy, cb.u[j][1], dz * cb.u[j][2])) * cb.inv_n[j];
    }
    const int cls = classify_triangle(q, cb.le[0], cb.le[1], cb.le[2], cb.margin, cb.off, cb.fe);
    if (cls == 0) return 0;
    if (!pd.is_box) return 2;                      // 'inside' and 'crossing' only mean something for closed boxes
    return cls == 1 ? 0 : cls;
}

// FCL's exact test of scene triangle `tv` against gripper triangle `gt` (TCP frame) under the affine map of grasp `row`
__device__ __forceinline__ bool exact_pair_hit(const float* __restrict__ gs, int row, const double* __restrict__ tv, const double* gt,
                                            CollideParams cp) {
    double M[9], T[3];
    float sf;
    grasp_affine(gs, row, cp.ow, cp.tol, cp.use_width, M, T, sf);
    d3 Q[3];
#pragma unroll
    for (int v = 0; v < 3; ++v) {
        const double px = gt[3 * v], py = gt[3 * v + 1], pz = gt[3 * v + 2];
        Q[v] = mk3(((M[0] * px + M[1] * py) + M[2] * pz) + T[0], ((M[3] * px + M[4] * py) + M[5] * pz) + T[1],
                   ((M[6] * px + M[7] * py) + M[8] * pz) + T[2]);
    }
    return tri_tri_intersect(ld3(tv), ld3(tv + 3), ld3(tv + 6), Q[0], Q[1], Q[2]);
}

// The queue-draining passes live in their own (not inlined) function: their register needs (float64 triangle pairs)
// would otherwise push the traversal state of the walking lanes into local memory for the whole kernel.  The walk boxes
// are parked in shared memory around the call, so nothing but a few scalars is live across it.
struct DrainArgs {
    const float* gs; const double* tri_verts; const double* gtris; const PartDev* parts; uint8_t* out;
    int *cq_row, *cq_leaf, *cq_meta, *xq_row, *xq_leaf, *xq_meta, *slot_row, *slot_dead;
    CollideParams cp;
    // lean mode (INLINE_EXACT == false): undecided leaves leave the kernel through a global queue
    int4* gx; unsigned* gx_count; unsigned gx_cap; uint8_t* unknown; unsigned* overflow;
};

// exact passes first (keeps the exact queue below 2 + 32 entries), then classification; returns (nx, ncq).
// INLINE_EXACT == false: no exact passes here -- undecided entries are appended to the global queue `gx` (one atomic per
// warp and pass) for collide_exact_kernel; entries that do not fit mark their row in `unknown` for the rescue launch.
template <bool STATS, bool INLINE_EXACT>
__device__ __noinline__ int2 drain_queues(const DrainArgs& a, int nx, int ncq, bool final, int lane, unsigned long long* st) {
    const unsigned FULL = 0xffffffffu;
while (true) {
        if (INLINE_EXACT && (nx >= 2 || (final && ncq == 0 && nx > 0))) {
            // exact: two queued entries per pass, one per half warp
            const int half = lane >> 4, sub = lane & 15;
            const int e = nx - 1 - half;
            int x_row = 0, x_leaf = 0, x_meta = 0, tb = 0, cnt = 0;
            if (e >= 0) {
                x_row = a.xq_row[e]; x_leaf = a.xq_leaf[e]; x_meta = a.xq_meta[e];
                if (!load_flag(a.out + x_row)) { tb = a.parts[x_meta & 0xff].tri_begin; cnt = a.parts[x_meta & 0xff].tri_end - tb; }
            }
            const int max_cnt = max(cnt, __shfl_xor_sync(FULL, cnt, 16));
            bool found = false;
            for (int t0 = 0; t0 < max_cnt; t0 += 16) {
                bool hit = false;
                if (!found && t0 + sub < cnt)
                    hit = exact_pair_hit(a.gs, x_row, a.tri_verts + 9 * (int64_t)x_leaf, a.gtris + 9 * (tb + t0 + sub), a.cp);
                const unsigned hb = __ballot_sync(FULL, hit);
                found = found || (((hb >> (16 * half)) & 0xffffu) != 0u);
            }
            if (found && sub == 0) { a.out[x_row] = 1; if (a.slot_row[x_meta >> 8] == x_row) a.slot_dead[x_meta >> 8] = 1; }
            nx = max(nx - 2, 0);
            __syncwarp();
        } else if (ncq >= 32 || (final && ncq > 0)) {
            // classify: 32 queued (item, leaf) entries at a time
            const int take = min(32, ncq), qbase = ncq - take;
            bool undecided = false;
            int e_row = 0, e_leaf = 0, e_meta = 0;
            if (lane < take) {
                e_row = a.cq_row[qbase + lane]; e_leaf = a.cq_leaf[qbase + lane]; e_meta = a.cq_meta[qbase + lane];
                if (!load_flag(a.out + e_row)) {
                    const int cls = leaf_class(a.gs, e_row, &a.parts[e_meta & 0xff], a.tri_verts + 9 * (int64_t)e_leaf, a.cp);
                    if (cls == 3) {                                     // certain collision: no exact test needed
                        a.out[e_row] = 1;
                        if (a.slot_row[e_meta >> 8] == e_row) a.slot_dead[e_meta >> 8] = 1;
                    }
                    undecided = cls == 2;
                    if (STATS) { ++st[0]; if (undecided) ++st[1]; }
                }
            }
            const unsigned ub = __ballot_sync(FULL, undecided);
            if (INLINE_EXACT) {
                if (undecided) {
                    const int slot = nx + __popc(ub & ((1u << lane) - 1u));
                    a.xq_row[slot] = e_row; a.xq_leaf[slot] = e_leaf; a.xq_meta[slot] = e_meta;
                }
                nx += __popc(ub);
            } else if (ub) {
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(a.gx_count, (unsigned)__popc(ub));
                base = __shfl_sync(FULL, base, 0);
                if (undecided) {
                    const unsigned pos = base + (unsigned)__popc(ub & ((1u << lane) - 1u));
                    if (pos < a.gx_cap) a.gx[pos] = make_int4(e_row, e_leaf, e_meta & 0xff, 0);
                    else { a.unknown[e_row] = 1; atomicAdd(a.overflow, 1u); }
                }
            }
            ncq = qbase;
            __syncwarp();
        } else {
            break;
        }
    }
    return make_int2(nx, ncq);
}

// branch-free version of the conservative 6-axis box test (all 32 lanes of a pass run the same instructions)
__device__ __forceinline__ bool box_overlaps(const float* bx, float x0, float y0, float z0, float x1, float y1, float z1) {
    // bx: u[9] ([j][k]), h[3], a[3], cf[3]
    const float ex = 0.5f * (x1 - x0), ey = 0.5f * (y1 - y0), ez = 0.5f * (z1 - z0);
    const float dx = 0.5f * (x1 + x0) - bx[15], dy = 0.5f * (y1 + y0) - bx[16], dz = 0.5f * (z1 + z0) - bx[17];
    bool ov = (fabsf(dx) <= ex + bx[12]) & (fabsf(dy) <= ey + bx[13]) & (fabsf(dz) <= ez + bx[14]);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float pr = fabsf(fmaf(dx, bx[3 * j], fmaf(dy, bx[3 * j + 1], dz * bx[3 * j + 2])));
        ov = ov & (pr <= bx[9 + j] + fmaf(ex, fabsf(bx[3 * j]), fmaf(ey, fabsf(bx[3 * j + 1]), ez * fabsf(bx[3 * j + 2]))));
    }
    return ov;
}

// Work item = (gripper part segment, grasp).  Traversal lengths differ by orders of magnitude between items (a stem far
// from the object ends at the root, a finger grazing the surface visits hundreds of nodes), so neither items nor their
// walks are bound to threads.  Every warp keeps up to 32 items in flight in shared-memory SLOTS (oriented box, row, part)
// and one LIFO stack of TASKS (slot, LBVH node) for all of them:
//   walk     each pass the 32 lanes pop 32 tasks -- of whatever items -- test both children of the node against the
//            slot's box (float32, conservative 6 axes, branch-free), push internal children as new tasks and leaf children
//            as (row, part, leaf) entries of the classification queue.  Lane utilisation no longer depends on how long the
//            individual walks are.  A slot is free again when its pending-task count drops to zero; free slots are
//            refilled from a global work counter.
//   classify once 32 leaves are queued every lane takes one entry and classifies the triangle in box-local float32
//            coordinates: disjoint / inside-a-closed-box / undecided.  Undecided entries go to the exact queue.
//   exact    two queued entries per pass, one per half warp: lane l tests gripper triangle l of the part with FCL's
//            float64 17-axis test.  A hit sets out[row] and marks the slot dead (its remaining tasks are discarded).
// Stack bound: a multi-pop pass grows the stack by at most 32, a single-pop pass (taken near the capacity) by at most
// one per tree level, so TQ_CAP - (max_depth + 2) is a safe threshold for switching to single pops.
// Two instantiations:
//   INLINE_EXACT == false (the normal launch): the kernel only walks and classifies in float32 -- few registers, twice the
//       resident warps; the classifier settles most collisions itself (class 3), the undecided leaves go to a global
//       queue that collide_exact_kernel drains afterwards.
//   INLINE_EXACT == true  (rescue launch, and the instrumented build): exact tests run inside the kernel as described
//       above; with `only_rows` it handles just the rows whose queue entries did not fit (it returns at once if there
//       are none).
struct GlobalExactQueue { int4* entries; unsigned* count; unsigned cap; uint8_t* unknown; unsigned* overflow; };

#ifndef BT_COLLIDE_DBG
#define BT_COLLIDE_DBG 0               // 1: pass / batch statistics of the lean kernel (tools/collide_batches.py)
#endif
#if BT_COLLIDE_DBG
__device__ unsigned long long g_cdbg[8];
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#endif

template <bool STATS, bool INLINE_EXACT>
__global__ void __launch_bounds__(32 * CWARPS, INLINE_EXACT ? 4 : LEAN_MINB)
collide_kernel(const float4* __restrict__ nodes, const double* __restrict__ tri_verts, const float* __restrict__ gs,
               int64_t n_cap, const int64_t* __restrict__ d_n, const double* __restrict__ gtris,
               const PartDev* __restrict__ parts, int n_parts, float ow, float tol, int use_width, float inflate,
               int max_depth, int refill_at, unsigned long long* __restrict__ work_counter, uint8_t* out,
               GlobalExactQueue gq, const uint8_t* __restrict__ only_rows, unsigned long long* __restrict__ counters) {
    if (INLINE_EXACT && only_rows && *reinterpret_cast<volatile unsigned*>(gq.overflow) == 0u) return;   // nothing to rescue
    unsigned long long st_nodes = 0;
    extern __shared__ double s_gtris[];            // gripper triangles, TCP frame
    __shared__ PartDev s_parts[MAX_PARTS];
    __shared__ int s_cq_row[CWARPS][CQ_CAP], s_cq_leaf[CWARPS][CQ_CAP], s_cq_meta[CWARPS][CQ_CAP];
    __shared__ int s_xq_row[CWARPS][XQ_CAP], s_xq_leaf[CWARPS][XQ_CAP], s_xq_meta[CWARPS][XQ_CAP];
    __shared__ float s_box[CWARPS][18][32];
    __shared__ int s_row[CWARPS][32], s_part[CWARPS][32], s_pending[CWARPS][32], s_dead[CWARPS][32];
    __shared__ int s_tq[CWARPS][TQ_CAP];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const int n_gt = parts[n_parts - 1].tri_end;
    if (INLINE_EXACT)
        for (int i = threadIdx.x; i < 9 * n_gt; i += blockDim.x) s_gtris[i] = gtris[i];
    for (int i = threadIdx.x; i < n_parts; i += blockDim.x) s_parts[i] = parts[i];
    s_row[w][lane] = -1; s_pending[w][lane] = 0; s_dead[w][lane] = 0;
    __syncthreads();
    int* tq = s_tq[w];
    int* cq_row = s_cq_row[w]; int* cq_leaf = s_cq_leaf[w]; int* cq_meta = s_cq_meta[w];

    const int64_t n = d_n ? min(n_cap, *d_n) : n_cap;
    const int64_t total = n * n_parts;             // item = part * n + row  (part-major: neighbouring items share a region)
    const float inv_ow = 1.0f / ow;
    const int multi_limit = TQ_CAP - (max_depth + 2) - 32 * WU; // multi-pop passes only while nt <= multi_limit

    DrainArgs da;
    da.gs = gs; da.tri_verts = tri_verts; da.gtris = s_gtris; da.parts = s_parts; da.out = out;
    da.cq_row = cq_row; da.cq_leaf = cq_leaf; da.cq_meta = cq_meta;
    da.xq_row = s_xq_row[w]; da.xq_leaf = s_xq_leaf[w]; da.xq_meta = s_xq_meta[w];
    da.slot_row = s_row[w]; da.slot_dead = s_dead[w];
    da.gx = gq.entries; da.gx_count = gq.count; da.gx_cap = gq.cap; da.unknown = gq.unknown; da.overflow = gq.overflow;
    da.cp.ow = ow; da.cp.tol = tol; da.cp.inflate = inflate; da.cp.use_width = use_width;
    unsigned long long st_drain[2] = {0, 0};

    unsigned free_mask = FULL;                     // warp-uniform: slots without pending tasks
    int nt = 0, nx = 0, ncq = 0;                   // warp-uniform: entries on the task stack / exact queue / classification queue
    bool exhausted = false;
#if COLLIDE_ADAPT
    int fetch_items = 32, batch_items = 0, batch_passes = 0;       // warp-uniform
    unsigned long long last_base = 0;                              // queue position after this warp's latest fetch
#endif

#if BT_COLLIDE_DBG
    unsigned long long dbg_t0 = gtimer(), dbg_tb = dbg_t0, dbg_maxbt = 0;
    unsigned dbg_pass = 0, dbg_pass_b = 0, dbg_single = 0, dbg_maxb = 0, dbg_batches = 0;
#endif
    while (true) {
        // ---- refill free slots with new items ---------------------------------------------------------------------------
        int n_free = __popc(free_mask);
        if (!exhausted && n_free >= refill_at && nt <= multi_limit) {
#if BT_COLLIDE_DBG
            { const unsigned long long now = gtimer(); dbg_maxb = max(dbg_maxb, dbg_pass_b); dbg_maxbt = max(dbg_maxbt, now - dbg_tb); dbg_tb = now; dbg_pass_b = 0; ++dbg_batches; }
#endif
#if COLLIDE_ADAPT
            // adaptive batch size: a batch that took many passes (expensive items come in runs: one gripper part is
            // costlier than the others, and so are whole neighbourhoods of contact pairs) is followed by a smaller fetch,
            // so that no warp sits on a batch several times longer than the average while the queue runs dry
            // ... which only matters while the queue is short: with many rounds left the dynamic fetch evens things out and
            // full batches keep the lanes busiest
            if (batch_items > 0 && total - (int64_t)last_base < (int64_t)COLLIDE_ADAPT_ROUNDS * 32 * gridDim.x * CWARPS) {
                const int want = (batch_items * COLLIDE_ADAPT + batch_passes / 2) / max(batch_passes, 1);
                fetch_items = max(8, min(32, (want + 3) & ~3));
            } else fetch_items = 32;
            n_free = min(n_free, fetch_items);
            batch_items = n_free; batch_passes = 0;
#endif
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(work_counter, (unsigned long long)n_free);
            base = __shfl_sync(FULL, base, 0);
#if COLLIDE_ADAPT
            last_base = base + (unsigned long long)n_free;
#endif
            bool take = false;
            int slot = 0;
            if (lane < n_free) {
                const int64_t item = (int64_t)base + lane;
                if (item < total) {
                    const int part = (int)(item / n);
                    const int64_t row = item - (int64_t)part * n;
                    if (!load_flag(out + row) && !(INLINE_EXACT && only_rows && !only_rows[row])) {   // may already collide
                        slot = __fns(free_mask, 0, lane + 1);
                        WalkBox b;
                        build_walkbox(b, gs, row, s_parts[part], inv_ow, tol, use_width, inflate);
#pragma unroll
                        for (int j = 0; j < 3; ++j) {
#pragma unroll
                            for (int k = 0; k < 3; ++k) s_box[w][3 * j + k][slot] = b.u[j][k];
                            s_box[w][9 + j][slot] = b.h[j]; s_box[w][12 + j][slot] = b.a[j]; s_box[w][15 + j][slot] = b.cf[j];
                        }
                        s_row[w][slot] = (int)row; s_part[w][slot] = part; s_pending[w][slot] = 1; s_dead[w][slot] = 0;
                        take = true;
                    }
                }
            }
            const unsigned tb = __ballot_sync(FULL, take);
            if (take) tq[nt + __popc(tb & lt)] = slot << 26;                    // (slot, root node 0)
            free_mask &= ~__reduce_or_sync(FULL, take ? (1u << slot) : 0u);
            nt += __popc(tb);
            if (base + (unsigned long long)n_free >= (unsigned long long)total) exhausted = true;
            __syncwarp();
        }
        if (nt == 0) {
            if (!exhausted) continue;              // every slot is free here, so the next refill takes 32 items
            break;
        }
        // ---- walk: pop up to 32 * WU tasks (WU per lane: independent node fetches in flight), test both children, push ---
        const int take_n = nt <= multi_limit ? min(32 * WU, nt) : 1;
#if BT_COLLIDE_DBG
        ++dbg_pass; ++dbg_pass_b; if (nt > multi_limit) ++dbg_single;
#endif
#if COLLIDE_ADAPT
        ++batch_passes;
#endif
        nt -= take_n;
        int slot[WU], l[WU], r[WU], task[WU];
        bool has[WU], hl[WU], hr[WU];
        float4 q0[WU], q1[WU], q2[WU];
#pragma unroll
        for (int u = 0; u < WU; ++u) {
            has[u] = lane + 32 * u < take_n;
            task[u] = has[u] ? tq[nt + lane + 32 * u] : 0;
            slot[u] = task[u] >> 26;
            has[u] = has[u];
            hl[u] = hr[u] = false; l[u] = r[u] = 0;
        }
        bool live[WU];
#pragma unroll
        for (int u = 0; u < WU; ++u) {
            live[u] = has[u] && !s_dead[w][slot[u]];
            if (live[u]) {
                const int node = task[u] & 0x3ffffff;
                float4 q3;
                ldg256(nodes + 4 * (int64_t)node, q0[u], q1[u]);                  // 64-byte node: 2 x LDG.256
                ldg256(nodes + 4 * (int64_t)node + 2, q2[u], q3);
                l[u] = __float_as_int(q3.x); r[u] = __float_as_int(q3.y);
            }
        }
#pragma unroll
        for (int u = 0; u < WU; ++u) {
            if (live[u]) {
                if (STATS) ++st_nodes;
                float bx[18];
#pragma unroll
                for (int k = 0; k < 18; ++k) bx[k] = s_box[w][k][slot[u]];
                hl[u] = box_overlaps(bx, q0[u].x, q0[u].y, q0[u].z, q0[u].w, q1[u].x, q1[u].y);
                hr[u] = box_overlaps(bx, q1[u].z, q1[u].w, q2[u].x, q2[u].y, q2[u].z, q2[u].w);
            }
        }
        __syncwarp();                               // all pops are done before anything is pushed over them
        unsigned freed_bits = 0;
#pragma unroll
        for (int u = 0; u < WU; ++u) {
            const bool pl = hl[u] && l[u] >= 0, pr = hr[u] && r[u] >= 0, ll = hl[u] && l[u] < 0, lr = hr[u] && r[u] < 0;
            const unsigned bl = __ballot_sync(FULL, pl), br = __ballot_sync(FULL, pr);
            if (pl) tq[nt + __popc(bl & lt)] = l[u] | (slot[u] << 26);
            nt += __popc(bl);
            if (pr) tq[nt + __popc(br & lt)] = r[u] | (slot[u] << 26);
            nt += __popc(br);
            const unsigned cl = __ballot_sync(FULL, ll), cr = __ballot_sync(FULL, lr);
            if (cl | cr) {
                if (ll) { const int q = ncq + __popc(cl & lt); cq_row[q] = s_row[w][slot[u]]; cq_leaf[q] = ~l[u]; cq_meta[q] = s_part[w][slot[u]] | (slot[u] << 8); }
                ncq += __popc(cl);
                if (lr) { const int q = ncq + __popc(cr & lt); cq_row[q] = s_row[w][slot[u]]; cq_leaf[q] = ~r[u]; cq_meta[q] = s_part[w][slot[u]] | (slot[u] << 8); }
                ncq += __popc(cr);
            }
            if (has[u]) {
                const int delta = (int)pl + (int)pr - 1;
                if (atomicAdd(&s_pending[w][slot[u]], delta) + delta == 0) freed_bits |= 1u << slot[u];
            }
        }
        free_mask |= __reduce_or_sync(FULL, freed_bits);
        __syncwarp();
        // ---- drain the queues ---------------------------------------------------------------------------------------------
        if (nx >= 2 || ncq >= 32) {
            const int2 res = drain_queues<STATS, INLINE_EXACT>(da, nx, ncq, false, lane, st_drain);
            nx = res.x; ncq = res.y;
        }
    }
    if (nx > 0 || ncq > 0) drain_queues<STATS, INLINE_EXACT>(da, nx, ncq, true, lane, st_drain);
#if BT_COLLIDE_DBG
    if (lane == 0 && !INLINE_EXACT) {
        const unsigned long long now = gtimer();
        dbg_maxb = max(dbg_maxb, dbg_pass_b); dbg_maxbt = max(dbg_maxbt, now - dbg_tb);
        atomicMax(&g_cdbg[0], (unsigned long long)dbg_maxb); atomicAdd(&g_cdbg[1], (unsigned long long)dbg_pass);
        atomicAdd(&g_cdbg[2], (unsigned long long)dbg_batches); atomicMax(&g_cdbg[3], (unsigned long long)dbg_pass);
        atomicAdd(&g_cdbg[4], 1ull); atomicMax(&g_cdbg[5], (unsigned long long)dbg_single);
        atomicMax(&g_cdbg[6], dbg_maxbt); atomicMax(&g_cdbg[7], now - dbg_t0);
    }
#endif
    if (STATS) { atomicAdd(counters + 3, st_nodes); atomicAdd(counters + 4, st_drain[0]); atomicAdd(counters + 5, st_drain[1]); }
}

// The undecided (row, leaf, part) entries of the lean launch: one half warp per entry, lane l <-> gripper triangle l of the
// part, FCL's float64 17-axis test.  Entries of rows that are already known to collide are skipped.
__global__ void __launch_bounds__(128)
collide_exact_kernel(const int4* __restrict__ entries, const unsigned* __restrict__ count, unsigned cap,
                     const double* __restrict__ tri_verts, const float* __restrict__ gs, const double* __restrict__ gtris,
                     const PartDev* __restrict__ parts, int n_parts, float ow, float tol, int use_width, float inflate,
                     uint8_t* out) {
    extern __shared__ double s_gtris[];
    __shared__ PartDev s_parts[MAX_PARTS];
    const int n_gt = parts[n_parts - 1].tri_end;
    for (int i = threadIdx.x; i < 9 * n_gt; i += blockDim.x) s_gtris[i] = gtris[i];
    for (int i = threadIdx.x; i < n_parts; i += blockDim.x) s_parts[i] = parts[i];
    __syncthreads();
    const unsigned n = min(*count, cap);
    const int sub = threadIdx.x & 15;
    const unsigned half_mask = 0xffffu << (threadIdx.x & 16);
    const CollideParams cp = {ow, tol, inflate, use_width};
    const unsigned stride = gridDim.x * (blockDim.x >> 4);
    for (unsigned e = blockIdx.x * (blockDim.x >> 4) + (threadIdx.x >> 4); e < n; e += stride) {
        const int4 en = entries[e];
        if (load_flag(out + en.x)) continue;
        const int tb = s_parts[en.z].tri_begin, cnt = s_parts[en.z].tri_end - tb;
        bool found = false;
        for (int t0 = 0; t0 < cnt && !found; t0 += 16) {
            bool hit = false;
            if (t0 + sub < cnt) hit = exact_pair_hit(gs, en.x, tri_verts + 9 * (int64_t)en.y, s_gtris + 9 * (tb + t0 + sub), cp);
            found = __ballot_sync(half_mask, hit) != 0u;
        }
        if (found && sub == 0) out[en.x] = 1;
    }
}

// ---- compaction in ONE launch: ballot/popc counts per block, a single-pass chained scan of the block totals across the
// grid (decoupled look-back, common.cuh), ordered scatter ---------------------------------------------------------------
constexpr int CB = 256;

// The kept rows of a block are contiguous in the destination ([offset, offset + run)), so the block stages its 256 source
// rows in shared memory (coalesced 8-byte loads), builds the dst->src map from the ballot ranks and streams the run out as
// consecutive 32-byte stores: full-line writes whatever the destination is -- local HBM, a peer GPU's memory over NVLink
// (the multi-GPU gather is this store, see distributed.PeerGather) or mapped host memory over PCIe
// (AntipodalGraspSampler.evaluate_host).  A 14-float row is 7 words, a 6-double contact record 6 words; both row sizes
// are multiples of 8 bytes, so every word is aligned.
__global__ void __launch_bounds__(CB) compact_kernel(const uint8_t* __restrict__ mask, uint8_t keep, int64_t n_cap,
                                       const int64_t* __restrict__ d_n, unsigned long long* status, unsigned* ticket,
                                       unsigned n_blocks, const float* __restrict__ gs, const double* __restrict__ contacts,
                                       float* __restrict__ out_gs, double* __restrict__ out_contacts, int64_t* __restrict__ n_out,
                                       const int64_t* __restrict__ d_base) {
    __shared__ int warp_off[CB / 32];
    __shared__ int run_s;
    __shared__ unsigned s_bid;
    __shared__ unsigned long long s_excl;
    __shared__ uint16_t src_of[CB];
    __shared__ double tile[CB * 7];                       // 14 KB: the block's rows (7 words each), then its contacts (6 each)
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t n = d_n ? min(n_cap, *d_n) : n_cap;
    // Persistent blocks take 256-row tiles in ticket order (the look-back only ever waits for lower tickets, which are done
    // or held by a running block).  The grid is sized by the destination: many blocks for local HBM, ONE block per SM when
    // the stores cross NVLink or PCIe -- there the kernel lives as long as the link needs (80 us for 11 MB into a peer that
    // seven ranks write at once, 250 us into host memory), and a block per tile would sit on a third of every SM's warp
    // slots for that time and starve the next batch's ray caster running beside it.
    while (true) {
        __syncthreads();                                  // the previous tile's shared state is no longer read
        if (threadIdx.x == 0) s_bid = atomicAdd(ticket, 1u);
        __syncthreads();
        const unsigned bid = s_bid;
        if (bid >= n_blocks) return;
        const int64_t row0 = bid * (int64_t)CB;
        const int64_t i = row0 + threadIdx.x;
        const bool k = (i < n) && (mask[i] == keep);
        const unsigned bal = __ballot_sync(0xffffffffu, k);
        if (lane == 0) warp_off[w] = __popc(bal);
        const int rows_here = (int)max((int64_t)0, min((int64_t)CB, n - row0));
        {   // stage the source rows while the offsets settle
            const double* src = reinterpret_cast<const double*>(gs) + 7 * row0;
            for (int e = threadIdx.x; e < rows_here * 7; e += CB) tile[e] = src[e];
        }
        __syncthreads();
        if (w == 0) {
            int c = lane < CB / 32 ? warp_off[lane] : 0;
            int incl = c;
#pragma unroll
            for (int s = 1; s < CB / 32; s <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, s); if (lane >= s) incl += v; }
            if (lane < CB / 32) warp_off[lane] = incl - c;
            const int run = __shfl_sync(0xffffffffu, incl, CB / 32 - 1);
            const unsigned long long excl = lookback_exclusive(status, bid, (unsigned long long)run, lane);
            if (lane == 0) {
                run_s = run;
                s_excl = excl;
                if (bid == n_blocks - 1 && n_out) *n_out = (int64_t)excl + run + (d_base ? *d_base : 0);
            }
        }
        __syncthreads();
        if (k) src_of[warp_off[w] + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)threadIdx.x;
        __syncthreads();
        const int run = run_s;
        const int64_t dst0 = (int64_t)s_excl + (d_base ? *d_base : 0);      // appending behind an earlier batch's rows
        // the run leaves as 32-byte stores (STG.E.256; up to three 8-byte words in front / behind where the run is not 32-byte
        // aligned): a warp store covers 1 KB of contiguous destination, which is what a link wants to see (against 16-byte
        // stores: the same over NVLink, 1.26 -> 1.20 ms for the end-to-end call over PCIe)
        auto stream_out = [&](double* dst, int words_per_row) {
            const int n = run * words_per_row;
            if (n == 0) return;
            auto word = [&](int e) { const int r = e / words_per_row; return tile[words_per_row * src_of[r] + (e - words_per_row * r)]; };
            const int head = min(n, (int)((4u - ((unsigned)(reinterpret_cast<uintptr_t>(dst) >> 3) & 3u)) & 3u));
            const int n_quads = (n - head) >> 2;
            const int tail0 = head + 4 * n_quads;
            if (threadIdx.x < head) dst[threadIdx.x] = word(threadIdx.x);
            for (int q = threadIdx.x; q < n_quads; q += CB) {
                const int e = head + 4 * q;
                const double w0 = word(e), w1 = word(e + 1), w2 = word(e + 2), w3 = word(e + 3);
                asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(dst + e), "d"(w0), "d"(w1), "d"(w2), "d"(w3) : "memory");
            }
            if (threadIdx.x >= 32 && threadIdx.x < 32 + (n - tail0)) dst[tail0 + threadIdx.x - 32] = word(tail0 + threadIdx.x - 32);
        };
        stream_out(reinterpret_cast<double*>(out_gs) + 7 * dst0, 7);
        if (contacts && out_contacts) {
            __syncthreads();
            const double* src = contacts + 6 * row0;
            for (int e = threadIdx.x; e < rows_here * 6; e += CB) tile[e] = src[e];
            __syncthreads();
            stream_out(out_contacts + 6 * dst0, 6);
        }
    }
}

}  // namespace bt

using namespace bt;

// true iff the part's triangles are exactly the 12 faces of its axis-aligned (TCP frame) box: every face of the box
// carries two triangles whose vertices are corners of that face and which share exactly one diagonal
static bool part_is_closed_box(const double* tris, const bt_gripper_part& hp) {
    if (hp.tri_end - hp.tri_begin != 12) return false;
    int per_face[6] = {0, 0, 0, 0, 0, 0};
    int corner_mask[6][2];
    for (int t = hp.tri_begin; t < hp.tri_end; ++t) {
        const double* v = tris + 9 * t;
        int corner[3];
        for (int k = 0; k < 3; ++k) {
            int c = 0;
            for (int a = 0; a < 3; ++a) {
                const double x = v[3 * k + a];
                if (x == hp.box_max[a]) c |= 1 << a;
                else if (x != hp.box_min[a]) return false;          // not a box corner
            }
            corner[k] = c;
        }
        if (corner[0] == corner[1] || corner[1] == corner[2] || corner[0] == corner[2]) return false;
        int face = -1;
        for (int a = 0; a < 3 && face < 0; ++a)
            for (int side = 0; side < 2 && face < 0; ++side) {
                bool all = true;
                for (int k = 0; k < 3; ++k) all = all && (((corner[k] >> a) & 1) == side);
                if (all) face = 2 * a + side;
            }
        if (face < 0 || per_face[face] >= 2) return false;
        corner_mask[face][per_face[face]++] = (1 << corner[0]) | (1 << corner[1]) | (1 << corner[2]);
    }
    for (int f = 0; f < 6; ++f) {
        if (per_face[f] != 2) return false;
        const int shared = corner_mask[f][0] & corner_mask[f][1];
        if (__builtin_popcount(shared) != 2 || __builtin_popcount(corner_mask[f][0] | corner_mask[f][1]) != 4) return false;
    }
    return true;
}

struct bt_gripper {
    int device = 0;
    int n_tris = 0, n_parts = 0;
    double* tris = nullptr;        // [n_tris, 9] TCP frame
    bt::PartDev* parts = nullptr;  // [n_parts]
};

extern "C" int bt_gripper_create(const double* h_tris, int32_t n_gt, const bt_gripper_part* h_parts, int32_t n_parts,
                                 int device, bt_gripper** out) {
    BT_REQUIRE(h_tris && h_parts && out, "null pointer");
    BT_REQUIRE(n_gt > 0 && n_parts > 0 && n_parts <= MAX_PARTS, "bad counts (1 <= n_parts <= 64)");
    BT_REQUIRE(9 * (size_t)n_gt * sizeof(double) <= 160 * 1024, "gripper mesh too large for shared memory staging");
    std::vector<PartDev> parts;
    int expect = 0;
    double seg_len = 0.048;                                    // metres; BT_COLLIDE_SEG overrides (tuning knob)
    if (const char* e = getenv("BT_COLLIDE_SEG")) { const double v = atof(e); if (v > 0) seg_len = v; }
    for (int p = 0; p < n_parts; ++p) {
        const bt_gripper_part& hp = h_parts[p];
        BT_REQUIRE(hp.tri_begin == expect && hp.tri_end > hp.tri_begin && hp.tri_end <= n_gt, "parts must tile [0, n_tris)");
        expect = hp.tri_end;
        double c[3], h[3];
        for (int k = 0; k < 3; ++k) { c[k] = 0.5 * (hp.box_min[k] + hp.box_max[k]); h[k] = 0.5 * (hp.box_max[k] - hp.box_min[k]); }
        int axis = 0;
        if (h[1] > h[axis]) axis = 1;
        if (h[2] > h[axis]) axis = 2;
        int nseg = (int)ceil(2.0 * h[axis] / seg_len);
        if (nseg < 1) nseg = 1;
        if ((int)parts.size() + nseg + (n_parts - p - 1) > MAX_PARTS) nseg = 1;      // stay within the shared-memory table
        const int is_box = part_is_closed_box(h_tris, hp) ? 1 : 0;
        for (int sgm = 0; sgm < nseg; ++sgm) {
            PartDev d;
            double sc[3] = {c[0], c[1], c[2]}, sh[3] = {h[0], h[1], h[2]};
            sh[axis] = h[axis] / nseg;
            sc[axis] = c[axis] - h[axis] + (2 * sgm + 1) * sh[axis];
            d.cx = (float)sc[0]; d.cy = (float)sc[1]; d.cz = (float)sc[2];
            // half extents rounded up a little: the segments must cover the part's box (and so all its triangles)
            d.hx = (float)sh[0] * 1.00001f + 1e-9f; d.hy = (float)sh[1] * 1.00001f + 1e-9f; d.hz = (float)sh[2] * 1.00001f + 1e-9f;
            d.ox = (float)(sc[0] - c[0]); d.oy = (float)(sc[1] - c[1]); d.oz = (float)(sc[2] - c[2]);
            d.fhx = (float)h[0]; d.fhy = (float)h[1]; d.fhz = (float)h[2];
            d.tri_begin = hp.tri_begin; d.tri_end = hp.tri_end; d.is_box = is_box;
            parts.push_back(d);
        }
    }
    BT_REQUIRE(expect == n_gt, "parts must cover every gripper triangle");
    n_parts = (int)parts.size();
    BT_CUDA(cudaSetDevice(device));
    bt_gripper* g = new bt_gripper();
    g->device = device; g->n_tris = n_gt; g->n_parts = n_parts;
    cudaError_t e = cudaMalloc(&g->tris, sizeof(double) * 9 * (size_t)n_gt);
    if (e == cudaSuccess) e = cudaMalloc(&g->parts, sizeof(PartDev) * n_parts);
    if (e == cudaSuccess) e = cudaMemcpy(g->tris, h_tris, sizeof(double) * 9 * (size_t)n_gt, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(g->parts, parts.data(), sizeof(PartDev) * n_parts, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(g->tris); cudaFree(g->parts); delete g; return cuda_fail(e, "bt_gripper_create", __FILE__, __LINE__); }
    *out = g;
    return BT_OK;
}

extern "C" int bt_gripper_destroy(bt_gripper* g) {
    if (!g) return BT_OK;
    cudaSetDevice(g->device);
    cudaFree(g->tris); cudaFree(g->parts);
    delete g;
    return BT_OK;
}

extern "C" int bt_collide(const bt_mesh* scene, const bt_gripper* gripper, const float* d_gs, int64_t n,
                          const int64_t* d_n, double opening_width, double width_tolerance, int32_t use_width,
                          uint8_t* d_colliding, void* stream) {
    BT_REQUIRE(scene && gripper && d_gs && d_colliding, "null pointer");
    BT_REQUIRE(n >= 0 && n < (int64_t)1 << 31, "row count must be in [0, 2^31)");
    BT_REQUIRE(scene->m.device == gripper->device, "scene and gripper live on different devices");
    BT_REQUIRE((reinterpret_cast<uintptr_t>(d_colliding) & 3) == 0, "d_colliding must be 4-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    BT_CUDA(cudaSetDevice(scene->m.device));
    BT_CUDA(cudaMemsetAsync(d_colliding, 0, (size_t)ceil_div(n, 4) * 4, st));
    if (n == 0) return BT_OK;
    float max_abs = 0.1f;
    for (int k = 0; k < 3; ++k)
        max_abs = fmaxf(max_abs, fmaxf(fabsf(scene->m.info.bounds_min[k]), fabsf(scene->m.info.bounds_max[k])));
    const float inflate = 2e-5f * max_abs + 1e-6f;
    const size_t tri_bytes = sizeof(double) * 9 * (size_t)gripper->n_tris;
    // tuning knobs (environment, read once): traversal burst length, refill threshold, occupancy variant
    static int refill_at = 0, inline_exact = 0;
    if (!refill_at) {
        const char* e;
        refill_at = (e = getenv("BT_COLLIDE_REFILL")) ? std::max(1, std::min(32, atoi(e))) : 32;
        inline_exact = (e = getenv("BT_COLLIDE_INLINE")) ? atoi(e) : 0;
    }
    BT_REQUIRE(scene->m.info.n_nodes < (1 << 26), "scene LBVH has more than 2^26 nodes");
    BT_REQUIRE(scene->m.info.max_depth + 2 + 64 * WU <= TQ_CAP, "scene LBVH too deep for the task stack");
    Scratch counter, gx_entries, gx_meta, unknown;
    BT_CUDA(counter.alloc(2 * sizeof(unsigned long long), st));                 // work counters of the two launches
    BT_CUDA(cudaMemsetAsync(counter.p, 0, 2 * sizeof(unsigned long long), st));
    int sm_count = 148;
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, scene->m.device);
    const int64_t items = n * gripper->n_parts;
    // global queue of undecided leaves: 2 entries per item is ~15x what config 3 / 4 produce; what does not fit is
    // handled by the rescue launch below
    unsigned gx_cap = (unsigned)std::min<int64_t>(std::max<int64_t>(2 * items, 1 << 16), (int64_t)1 << 26);
    if (const char* e = getenv("BT_COLLIDE_GXCAP")) { const long v = atol(e); if (v >= 1) gx_cap = (unsigned)std::min<long>(v, 1l << 26); }   // tests: force the rescue path
    BT_CUDA(gx_entries.alloc(sizeof(int4) * (size_t)gx_cap, st));
    BT_CUDA(gx_meta.alloc(2 * sizeof(unsigned), st));                            // [0] entries appended, [1] entries dropped
    BT_CUDA(cudaMemsetAsync(gx_meta.p, 0, 2 * sizeof(unsigned), st));
    BT_CUDA(unknown.alloc((size_t)n, st));
    BT_CUDA(cudaMemsetAsync(unknown.p, 0, (size_t)n, st));
    GlobalExactQueue gq;
    gq.entries = gx_entries.as<int4>(); gq.count = gx_meta.as<unsigned>(); gq.cap = gx_cap;
    gq.unknown = unknown.as<uint8_t>(); gq.overflow = gx_meta.as<unsigned>() + 1;
    unsigned long long* wc = counter.as<unsigned long long>();
    if (tri_bytes > 8 * 1024) {
        BT_CUDA(cudaFuncSetAttribute(collide_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tri_bytes));
        BT_CUDA(cudaFuncSetAttribute(collide_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tri_bytes));
        BT_CUDA(cudaFuncSetAttribute(collide_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tri_bytes));
        BT_CUDA(cudaFuncSetAttribute(collide_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tri_bytes));
    }
#define BT_COLLIDE_ARGS(WC, ONLY, CNT)                                                                                         \
    scene->m.nodes, scene->m.tri_verts, d_gs, n, d_n, gripper->tris, gripper->parts, gripper->n_parts, (float)opening_width,    \
        (float)width_tolerance, use_width, inflate, (int)scene->m.info.max_depth, refill_at, WC, d_colliding, gq, ONLY, CNT
    if (g_counters_host_copy || inline_exact) {
        // instrumented build / BT_COLLIDE_INLINE=1: everything in one kernel
        const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(items, 128), (int64_t)sm_count * 4);
        if (g_counters_host_copy)
            collide_kernel<true, true><<<grid, 32 * CWARPS, tri_bytes, st>>>(BT_COLLIDE_ARGS(wc, nullptr, g_counters_host_copy));
        else
            collide_kernel<false, true><<<grid, 32 * CWARPS, tri_bytes, st>>>(BT_COLLIDE_ARGS(wc, nullptr, nullptr));
        BT_LAUNCH_CHECK();
        return BT_OK;
    }
    {
        static int lean_bps = -1;                           // tuning knob, as BT_TRAV_BPS for the ray caster
        if (lean_bps < 0) { const char* e = getenv("BT_COLLIDE_BPS"); const int v = e ? atoi(e) : 0; lean_bps = (v >= 1 && v <= LEAN_MINB) ? v : LEAN_MINB; }
        const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(items, 128), (int64_t)sm_count * lean_bps);
        collide_kernel<false, false><<<grid, 32 * CWARPS, 0, st>>>(BT_COLLIDE_ARGS(wc, nullptr, nullptr));
        BT_LAUNCH_CHECK();
        collide_exact_kernel<<<sm_count * 8, 128, tri_bytes, st>>>(gq.entries, gq.count, gq.cap, scene->m.tri_verts, d_gs, gripper->tris,
                                                                   gripper->parts, gripper->n_parts, (float)opening_width,
                                                                   (float)width_tolerance, use_width, inflate, d_colliding);
        BT_LAUNCH_CHECK();
        // rescue: rows whose queue entries were dropped are redone with in-kernel exact tests (returns at once if none)
        const unsigned rgrid = (unsigned)std::min<int64_t>(ceil_div(items, 128), (int64_t)sm_count * 4);
        collide_kernel<false, true><<<rgrid, 32 * CWARPS, tri_bytes, st>>>(BT_COLLIDE_ARGS(wc + 1, gq.unknown, nullptr));
    }
#undef BT_COLLIDE_ARGS
    BT_LAUNCH_CHECK();
    return BT_OK;
}

#if BT_COLLIDE_DBG
// [0] max passes of a batch, [1] passes (sum over warps), [2] batches, [3] max passes of a warp, [4] warps,
// [5] max single-pop passes of a warp, [6] longest batch (ns), [7] longest warp (ns)
extern "C" __attribute__((visibility("default"))) int bt_debug_collide_stats(unsigned long long* h_out, int reset) {
    if (h_out) BT_CUDA(cudaMemcpyFromSymbol(h_out, bt::g_cdbg, sizeof(unsigned long long) * 8));
    if (reset) { unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0}; BT_CUDA(cudaMemcpyToSymbol(bt::g_cdbg, z, sizeof(z))); }
    return BT_OK;
}
#endif

extern "C" int bt_compact(const uint8_t* d_mask, uint8_t keep_value, const float* d_gs, const double* d_contacts,
                          int64_t n, const int64_t* d_n, float* d_out_gs, double* d_out_contacts, int64_t* d_n_out,
                          const int64_t* d_base, void* stream) {
    return bt::compact_impl(d_mask, keep_value, d_gs, d_contacts, n, d_n, d_out_gs, d_out_contacts, d_n_out, d_base, stream);
}

int bt::compact_impl(const uint8_t* d_mask, uint8_t keep_value, const float* d_gs, const double* d_contacts, int64_t n,
                     const int64_t* d_n, float* d_out_gs, double* d_out_contacts, int64_t* d_n_out, const int64_t* d_base,
                     void* stream) {
    BT_REQUIRE(d_mask && d_gs && d_out_gs && d_n_out, "null pointer");
    BT_REQUIRE(n >= 0, "negative count");
    BT_REQUIRE(((reinterpret_cast<uintptr_t>(d_gs) | reinterpret_cast<uintptr_t>(d_out_gs)) & 7) == 0, "grasp rows must be 8-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    BT_REQUIRE(d_base != d_n_out, "d_base and d_n_out must be different addresses");
    if (n == 0) {
        if (d_base) BT_CUDA(cudaMemcpyAsync(d_n_out, d_base, sizeof(int64_t), cudaMemcpyDefault, st));
        else BT_CUDA(cudaMemsetAsync(d_n_out, 0, sizeof(int64_t), st));
        return BT_OK;
    }
    const int64_t n_blocks = ceil_div(n, CB);
    BT_REQUIRE(n_blocks < ((int64_t)1 << 31), "too many rows");
    Scratch state;                                              // [n_blocks] status words of the chained scan + the ticket counter
    BT_CUDA(state.alloc(sizeof(unsigned long long) * (n_blocks + 1), st));
    BT_CUDA(cudaMemsetAsync(state.p, 0, sizeof(unsigned long long) * (n_blocks + 1), st));
    // grid: one block per SM when the destination is not this GPU's memory (peer over NVLink, mapped host memory over
    // PCIe), else up to six per SM (see compact_kernel)
    int device = 0, sm_count = 148;
    BT_CUDA(cudaGetDevice(&device));
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaPointerAttributes attr;
    bool local = true;
    if (cudaPointerGetAttributes(&attr, d_out_gs) == cudaSuccess) local = attr.type == cudaMemoryTypeDevice && attr.device == device;
    else cudaGetLastError();
    static int remote_blocks = -1;
    if (remote_blocks < 0) { const char* e = getenv("BT_COMPACT_REMOTE_BLOCKS"); remote_blocks = e ? std::max(1, atoi(e)) : sm_count; }     // tuning knob
    const int64_t grid = std::min<int64_t>(n_blocks, local ? (int64_t)sm_count * 6 : (int64_t)remote_blocks);
    compact_kernel<<<(unsigned)grid, CB, 0, st>>>(d_mask, keep_value, n, d_n, state.as<unsigned long long>(),
                                                  reinterpret_cast<unsigned*>(state.as<unsigned long long>() + n_blocks), (unsigned)n_blocks,
                                                  d_gs, d_contacts, d_out_gs, d_out_contacts, d_n_out, d_base);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

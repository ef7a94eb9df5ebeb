// ags.cu -- antipodal grasp sampling kernels: surface samples (S1), cone rays (S2), the friction-cone
// ray caster with the antipodal filters (R1-R6), target selection (R7) and orientation expansion (O1/O2).
//
// Reference being replaced: burg_toolkit/sampling.py:185-352 (AntipodalGraspSampler.sample hot loop) and the
// leaves it calls (trimesh ray_triangle_id, mesh_processing.compute_interpolated_vertex_normals :110-153,
// util.angle :46-82, construct_grasp_set :132-183, construct_halfspace_grasp_set :72-130).
//
// Precision contract: the broad phase (LBVH, float32 slabs over padded boxes) only proposes candidates; every
// accept/reject decision is taken in float64 with the reference's formulas, operation order and tolerances.
#include <cub/cub.cuh>
#include <math.h>
#include <algorithm>

#include "common.cuh"

namespace bt {

constexpr int STACK_SIZE = 64;
constexpr int WIDE_STACK = 96;             // 4-wide walk: up to 3 parked children per wide level, depth <= 60 binary levels
constexpr float T_BEHIND = -1e-5f;        // boxes wholly behind the origin by more than this cannot hold a
                                          // hit that passes the reference's forward test (> -1e-6)

struct MeshView {
    const float4* __restrict__ nodes;
    const uint4* __restrict__ qnodes;
    const uint4* __restrict__ wnodes;
    double gmin[3], gscale[3];
    const double* __restrict__ tri_rec;
    const float4* __restrict__ tri_f32;
    const int32_t* __restrict__ leaf_tri;
    const int32_t* __restrict__ faces;
    const double* __restrict__ verts;
    const double* __restrict__ vnormals;
    const double* __restrict__ tri_vvn;            // [nF,20] leaf order: v0 v1 v2 pad | n0 n1 n2 pad
    const uint32_t* __restrict__ leaf_cone;       // [nF] leaf order: packed normal cones (mesh.cu: pack_cone)
};

__device__ __forceinline__ bool slab_hit(float bx0, float by0, float bz0, float bx1, float by1, float bz1,
                                         float idx, float idy, float idz, float oox, float ooy, float ooz) {
    float tx0 = fmaf(bx0, idx, -oox), tx1 = fmaf(bx1, idx, -oox);
    float ty0 = fmaf(by0, idy, -ooy), ty1 = fmaf(by1, idy, -ooy);
    float tz0 = fmaf(bz0, idz, -ooz), tz1 = fmaf(bz1, idz, -ooz);
    float tnear = fmaxf(fmaxf(fminf(tx0, tx1), fminf(ty0, ty1)), fminf(tz0, tz1));
    float tfar = fminf(fminf(fmaxf(tx0, tx1), fmaxf(ty0, ty1)), fmaxf(tz0, tz1));
    return (tnear <= tfar) & (tfar >= T_BEHIND);
}

// exact uint16 -> float32 without the conversion pipe: OR the 16 bits into the mantissa of 2^23 and subtract 2^23
__device__ __forceinline__ float u16lo(unsigned w) { return __uint_as_float((w & 0xffffu) | 0x4B000000u) - 8388608.0f; }
__device__ __forceinline__ float u16hi(unsigned w) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7632)) - 8388608.0f; }

// 2^23 + q as a float: one logic instruction per coordinate, the 2^23 is absorbed by the caller's FMA constant
__device__ __forceinline__ float m16lo(unsigned w) { return __uint_as_float((w & 0xffffu) | 0x4B000000u); }
__device__ __forceinline__ float m16hi(unsigned w) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7632)); }

__device__ __forceinline__ float safe_rcp(float d) {
    return 1.0f / (fabsf(d) < 1e-20f ? copysignf(1e-20f, d) : d);
}

// trimesh ray_triangle_id narrow phase for one (ray, triangle record); returns true and t if it is a hit
__device__ __forceinline__ bool exact_ray_triangle(const double* __restrict__ rec, d3 o, d3 d, double& t_out) {
    const double2* r2 = reinterpret_cast<const double2*>(rec);
    double2 a, b, c, e;
    ldg256(r2, a, b); ldg256(r2 + 2, c, e);          // 128-byte record, one line: 4 x LDG.256
    d3 v0 = mk3(a.x, a.y, b.x), n = mk3(b.y, c.x, c.y);
    // intersections.planes_lines
    double proj_ori = dot3(sub3(v0, o), n);
    double proj_dir = dot3(d, n);
    if (!(fabs(proj_dir) > 1e-5)) return false;
    double t = proj_ori / proj_dir;
    d3 p = add3(scl3(d, t), o);
    // forward test: dot(location - origin, direction) > -1e-6
    if (!(dot3(sub3(p, o), d) > -1e-6)) return false;
    // triangles.points_to_barycentric(method='cramer')
    double2 f, g, h, k;
    ldg256(r2 + 4, f, g); ldg256(r2 + 6, h, k);
    d3 e0 = mk3(e.x, e.y, f.x), e1 = mk3(f.y, g.x, g.y);
    double d00 = h.x, d01 = h.y, d11 = k.x, inv = k.y;
    d3 w = sub3(p, v0);
    double d02 = dot3(e0, w), d12 = dot3(e1, w);
    double b2 = (d00 * d12 - d01 * d02) * inv;
    double b1 = (d11 * d02 - d01 * d12) * inv;
    double b0 = 1.0 - b1 - b2;
    const double lo = -1e-13, hi = 1.0 + 1e-13;
    bool hit = (b0 > lo) & (b1 > lo) & (b2 > lo) & (b0 < hi) & (b1 < hi) & (b2 < hi);
    t_out = t;
    return hit;
}

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// mesh_processing.compute_interpolated_vertex_normals for one point on the triangle of `leaf`: the three vertices and
// their normals come from one 160-byte leaf-order record (no faces -> vertices -> normals chain, 10 sector accesses not 19)
__device__ __forceinline__ d3 interpolated_normal(const MeshView& mv, int32_t leaf, d3 p) {
    const double2* r2 = reinterpret_cast<const double2*>(mv.tri_vvn + 20 * (int64_t)leaf);      // five LDG.256, 160-byte (32-byte aligned) record
    double2 a0, a1, a2, a3, a4, b0, b1, b2, b3, b4;
    ldg256(r2 + 0, a0, a1); ldg256(r2 + 2, a2, a3); ldg256(r2 + 4, a4, b0); ldg256(r2 + 6, b1, b2); ldg256(r2 + 8, b3, b4);
    const d3 vert[3] = {mk3(a0.x, a0.y, a1.x), mk3(a1.y, a2.x, a2.y), mk3(a3.x, a3.y, a4.x)};
    const d3 vn[3] = {mk3(b0.x, b0.y, b1.x), mk3(b1.y, b2.x, b2.y), mk3(b3.x, b3.y, b4.x)};
    double w[3];
    int onehot = -1;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double dist = norm3(sub3(vert[k], p));
        w[k] = 1.0 / sqrt(dist);
        if (dist == 0.0) onehot = k;             // later vertices win, like the reference's loop
    }
    if (onehot >= 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) w[k] = (k == onehot) ? 1.0 : 0.0;
    }
    d3 s = mk3(((w[0] * vn[0].x + w[1] * vn[1].x) + w[2] * vn[2].x) / 3.0,
               ((w[0] * vn[0].y + w[1] * vn[1].y) + w[2] * vn[2].y) / 3.0,
               ((w[0] * vn[0].z + w[1] * vn[1].z) + w[2] * vn[2].z) / 3.0);
    return div3(s, norm3(s));
}

// Conservative float32 ray/triangle pre-test (division-free Moeller-Trumbore).  It may only say "no" when the exact
// float64 test of the reference certainly says "no": every comparison carries an absolute slack E that bounds the
// float32 rounding of the inputs and of the dot/cross products, plus a relative slack on the determinant.
__device__ __forceinline__ bool maybe_hit_f32(const float4* __restrict__ tr, float ox, float oy, float oz,
                                              float dx, float dy, float dz) {
    const float4 v0 = __ldg(tr + 0), e0 = __ldg(tr + 1), e1 = __ldg(tr + 2);
    // pvec = d x e1
    const float px = dy * e1.z - dz * e1.y, py = dz * e1.x - dx * e1.z, pz = dx * e1.y - dy * e1.x;
    float det = fmaf(e0.x, px, fmaf(e0.y, py, e0.z * pz));
    const float tx = ox - v0.x, ty = oy - v0.y, tz = oz - v0.z;
    float U = fmaf(tx, px, fmaf(ty, py, tz * pz));
    // qvec = tvec x e0
    const float qx = ty * e0.z - tz * e0.y, qy = tz * e0.x - tx * e0.z, qz = tx * e0.y - ty * e0.x;
    float V = fmaf(dx, qx, fmaf(dy, qy, dz * qz));
    float T = fmaf(e1.x, qx, fmaf(e1.y, qy, e1.z * qz));
    if (det < 0.f) { det = -det; U = -U; V = -V; T = -T; }
    const float l0 = fabsf(e0.x) + fabsf(e0.y) + fabsf(e0.z), l1 = fabsf(e1.x) + fabsf(e1.y) + fabsf(e1.z);
    const float lt = fabsf(tx) + fabsf(ty) + fabsf(tz) + fabsf(ox) + fabsf(oy) + fabsf(oz) + fabsf(v0.x) + fabsf(v0.y) + fabsf(v0.z);
    // |error| of U, V, T, det <= ~16 eps32 * (|tvec|+|o|+|v0|) * |e0| * |e1|-ish products; 256 eps32 is generous
    const float E = 1.6e-5f * (lt * fmaxf(l0, l1) + l0 * l1) + 1e-3f * det;
    return (U >= -E) & (V >= -E) & (U + V <= det + E) & (T >= -(E + 1e-4f * det) - 1e-4f * l0 * l1);
}

constexpr int NPEND = 32;

// ---- stage 1 of the ray caster: float32 traversal with ray re-fetch ---------------------------------------------------
// Persistent warps take chunks of consecutive rays from a global counter (rays of one reference point are neighbours, so
// a chunk is coherent) and keep their 32 lanes busy: a lane whose ray has left the tree takes the next ray of the chunk,
// and the next chunk when that one is used up.  The lanes only ever execute 4-wide NODE steps: the two planes of an axis share a word of the
// node, the ray picks its near / far plane by the sign of its direction with one byte-permute each, so a box test is
// 6 permutes + 6 FMAs + one max3 + one min3 + the compares.  Leaves that a ray reaches are not tested in line (that
// would serialise the warp): they go to the warp's leaf queue in shared memory, and once 32 are queued every lane runs
// one conservative float32 ray/triangle pre-test -- for whatever ray the entry belongs to.  Survivors are appended to
// that ray's candidate list (at most CCAP; a ray with more is flagged and re-walked by the resolve kernel).  Only
// float32 state is live, so occupancy is high; the float64 work happens in the second kernel with all lanes converged.
constexpr int CCAP = 16;
#ifndef TRAV_BURST
#define TRAV_BURST 2
#endif
#ifndef TRAV_REFETCH
#define TRAV_REFETCH 8
#endif
constexpr int TBURST = TRAV_BURST;                     // node steps between two looks at the leaf queue
constexpr int LQ_CAP = 32 + 128 * TBURST;              // < 32 left over + at most 4 leaves per lane and step
constexpr int TWARPS = 4;
#ifndef TRAV_MINB
#define TRAV_MINB 8
#endif
#ifndef TRAV_PREFETCH
#define TRAV_PREFETCH 0
#endif


// near / far plane of one axis as 2^23 + q: selector 0x7610 takes the low half of the word, 0x7632 the high half
// (PTX prmt directly: __byte_perm masks a run-time selector with 0x7777 first -- one LOP3 per use -- which the selectors
// used here (0x7610 / 0x7632: every nibble below 8, i.e. no sign replication) do not need)
__device__ __forceinline__ unsigned prmt(unsigned a, unsigned b, unsigned sel) { unsigned r; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel)); return r; }
__device__ __forceinline__ float plane16(unsigned w, unsigned sel) { return __uint_as_float(prmt(w, 0x4B000000u, sel)); }

// `behind`: T_BEHIND, or +inf for a node the cone test has ruled out (then no child is hit, without a branch).
// The near and the far plane of an axis go through ONE packed multiply-add (FFMA2: lo = near, hi = far; 1/d and the
// origin term are broadcast operands), so a box costs 6 permutes + 3 FFMA2 + max3 + min3 + the compares; each half is the
// scalar fmaf of the same operands, bit for bit.  `noox` = -(2^23 / d + o / d), negated once per ray.
__device__ __forceinline__ bool slab_hit_signed(unsigned wx, unsigned wy, unsigned wz, unsigned snx, unsigned sny, unsigned snz,
                                                unsigned sfx, unsigned sfy, unsigned sfz, float idx, float idy, float idz,
                                                float noox, float nooy, float nooz, float behind = T_BEHIND) {
    const f32x2 tx = fma2(pack2(plane16(wx, snx), plane16(wx, sfx)), dup2_here(idx), dup2_here(noox));
    const f32x2 ty = fma2(pack2(plane16(wy, sny), plane16(wy, sfy)), dup2_here(idy), dup2_here(nooy));
    const f32x2 tz = fma2(pack2(plane16(wz, snz), plane16(wz, sfz)), dup2_here(idz), dup2_here(nooz));
    const float tnear = fmaxf(fmaxf(lo2(tx), lo2(ty)), lo2(tz));
    const float tfar = fminf(fminf(hi2(tx), hi2(ty)), hi2(tz));
    return (tnear <= tfar) & (tfar >= behind);
}

// 24-bit child code of the packed node format -> the plain code (sign extension: one BFE)
__device__ __forceinline__ int sext24(unsigned w) { int r; asm("bfe.s32 %0, %1, 0, 24;" : "=r"(r) : "r"(w)); return r; }
// byte 3 of w as the float 2^15 + byte (the byte lands in mantissa bits 8..15 of 0x47000000): one byte-permute
__device__ __forceinline__ float top_byte_m15(unsigned w) { unsigned r; asm("prmt.b32 %0, %1, 0x47000000, 0x7434;" : "=r"(r) : "r"(w)); return __uint_as_float(r); }
__device__ __forceinline__ float byte_m15(unsigned w, unsigned sel) { return __uint_as_float(__byte_perm(w, 0x47000000u, sel)); }
constexpr float CONE_BIAS = 32768.0f + 128.0f;        // 2^15 + the int8 bias of a packed cone axis
constexpr float CONE_NEVER = -1e30f;

// threshold table of the cone test for this launch: code c -> 127 * (cos(c * pi/255 + angle_max + slack) - margin).  The
// axis bytes have length 127 +- 0.9 instead of exactly 127, which the cosine margin of 0.01 covers (|1/r - 1| < 0.0075).
__device__ __forceinline__ void fill_cone_table(float* tab, float angle_max) {
    for (int c = threadIdx.x; c < 256; c += blockDim.x) {
        const float phi = (float)c * (3.14159265f / 255.f) + angle_max + 3e-3f;
        tab[c] = (c == 255 || !(phi < 3.14159f)) ? CONE_NEVER : 127.f * (cosf(phi) - 0.01f);
    }
}

// PACKED: 24-bit child codes with the node's normal cone in the top bytes (meshes below 2^23 triangles).
// CULL (needs PACKED; lean pipeline only): a node / leaf whose normal cone cannot hold a VALID antipodal hit for this
// ray's direction is skipped.  What is skipped are hits the filters would reject anyway (wrong sign or outside the
// friction cone), and -- because leaf cones are widened over all touching triangles, see leaf_cone_kernel -- never a hit
// that could take part in the 1e-8 de-duplication of a valid one: the valid-hit masks and contacts are unchanged.
template <bool STATS, bool PACKED, bool CULL>
__global__ void __launch_bounds__(32 * TWARPS, TRAV_MINB)
traverse_kernel(MeshView mv, const double* __restrict__ ref_pts, const double* __restrict__ dirs, int64_t n_total,
                int n_rays, int chunk_len, float angle_max, int32_t* __restrict__ cand_count, int32_t* __restrict__ cand,
                unsigned* __restrict__ chunk_counter, unsigned long long* __restrict__ counters) {
    unsigned long long st_nodes = 0, st_leaves = 0, st_cand = 0, st_culled = 0;
    __shared__ int2 s_lq[TWARPS][LQ_CAP];            // leaf queue entries: (ray, leaf)
    __shared__ int s_lq_n[TWARPS];
    __shared__ float s_thr[CULL ? 256 : 1];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (CULL) { fill_cone_table(s_thr, angle_max); __syncthreads(); }
    // Rays are handed out in chunks of `chunk_len` consecutive rays (rays of one reference point are neighbours, so a chunk
    // is coherent) from a global counter: a warp whose lanes run dry takes the next chunk instead of idling until the
    // slowest ray of a fixed slice has finished, so only the kernel's last chunks see a thinning warp.
    const int total = (int)n_total;
    int next = 0, chunk_end = 0;                      // warp-uniform: rays [next, chunk_end) of the current chunk are untaken
    bool exhausted = false;                           // warp-uniform: the counter has run past the last chunk
    int2* lq = s_lq[w];
    if (lane == 0) s_lq_n[w] = 0;
    __syncwarp();

    const int DONE = (int)0x80000000;                 // never a leaf code: leaf codes are ~index with index < 2^30
    int my = 0;                                       // this lane's ray
    float idx = 0, idy = 0, idz = 0, oox = 0, ooy = 0, ooz = 0;       // oox..: -(2^23 / d + o / d), see the re-fetch below
    unsigned snx = 0, sny = 0, snz = 0, sfx = 0, sfy = 0, sfz = 0;
    float ux = 0, uy = 0, uz = 0, ubias = 0;             // CULL: unit direction and CONE_BIAS * (ux + uy + uz)
    int stack[WIDE_STACK];
    int sp = 0, cur = DONE;

    // one conservative float32 pre-test per lane for queue entries [first, first + count)
    auto pretest_pass = [&](int first, int count) {
        if (lane < count) {
            const int2 en = lq[first + lane];
            const int leaf = en.y;
            const int64_t ray = en.x, ref = (unsigned)en.x / (unsigned)n_rays;
            const float ox = (float)ref_pts[3 * ref], oy = (float)ref_pts[3 * ref + 1], oz = (float)ref_pts[3 * ref + 2];
            const float dx = (float)dirs[3 * ray], dy = (float)dirs[3 * ray + 1], dz = (float)dirs[3 * ray + 2];
            bool keep = true;
            if (CULL) {                                   // the leaf's own (widened) normal cone
                const unsigned cw = __ldg(mv.leaf_cone + leaf);
                const float inv = rsqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
                const float dq = fmaf(byte_m15(cw, 0x7404) - CONE_BIAS, dx, fmaf(byte_m15(cw, 0x7414) - CONE_BIAS, dy, (byte_m15(cw, 0x7424) - CONE_BIAS) * dz)) * inv;
                keep = !(dq < s_thr[cw >> 24]);
            }
            if (STATS && keep) ++st_leaves;
            if (keep && maybe_hit_f32(mv.tri_f32 + 3 * (int64_t)leaf, ox, oy, oz, dx, dy, dz)) {
                const int slot = atomicAdd(cand_count + ray, 1);
                if (slot < CCAP) cand[(int64_t)slot * n_total + ray] = leaf;       // slot-major: the resolve kernel reads it coalesced
                if (STATS) ++st_cand;
            }
        }
    };

    while (true) {
        // ---- re-fetch: when enough lanes have run dry they take the next rays of the chunk (or of a new chunk) ----------
        const unsigned idle = __ballot_sync(FULL, cur == DONE);
        if (!exhausted && (__popc(idle) >= TRAV_REFETCH || idle == FULL)) {
            if (next >= chunk_end) {
                unsigned c = 0;
                if (lane == 0) c = atomicAdd(chunk_counter, 1u);
                c = __shfl_sync(FULL, c, 0);
                const int64_t begin = (int64_t)c * chunk_len;
                exhausted = begin >= n_total;
                next = exhausted ? total : (int)begin;
                chunk_end = (int)min((int64_t)total, begin + chunk_len);
            }
            const int rank = __popc(idle & ((1u << lane) - 1u));
            if (cur == DONE && next + rank < chunk_end) {
                my = next + rank;
                const int64_t ray = my, ref = (unsigned)my / (unsigned)n_rays;
                // the ray in grid coordinates (same parameter t): origin (o - gmin) * scale, direction d * scale.
                // Box coordinates arrive as 2^23 + q (uint16 permuted into the mantissa of 2^23); instead of subtracting 2^23
                // per coordinate the offset is folded into the FMA constant: t = fma(2^23 + q, 1/dg, -(2^23/dg + og/dg)).
                // The rounding of that constant displaces the plane by < 0.52 grid cells -- the boxes were rounded
                // outward by one extra cell for exactly this reason (quantize_nodes_kernel / emit_wide_nodes_kernel).
                const float gdx = (float)dirs[3 * ray] * (float)mv.gscale[0], gdy = (float)dirs[3 * ray + 1] * (float)mv.gscale[1],
                            gdz = (float)dirs[3 * ray + 2] * (float)mv.gscale[2];
                idx = __frcp_rn(fabsf(gdx) < 1e-20f ? copysignf(1e-20f, gdx) : gdx);
                idy = __frcp_rn(fabsf(gdy) < 1e-20f ? copysignf(1e-20f, gdy) : gdy);
                idz = __frcp_rn(fabsf(gdz) < 1e-20f ? copysignf(1e-20f, gdz) : gdz);
                // (kept negated: the addend of the packed multiply-adds)
                oox = -fmaf(8388608.0f, idx, (float)((ref_pts[3 * ref] - mv.gmin[0]) * mv.gscale[0]) * idx);
                ooy = -fmaf(8388608.0f, idy, (float)((ref_pts[3 * ref + 1] - mv.gmin[1]) * mv.gscale[1]) * idy);
                ooz = -fmaf(8388608.0f, idz, (float)((ref_pts[3 * ref + 2] - mv.gmin[2]) * mv.gscale[2]) * idz);
                // near plane = the low one where the ray runs in +axis direction (sign bit of 1/d, so -0 counts as negative)
                const bool nx = __float_as_uint(idx) >> 31, ny = __float_as_uint(idy) >> 31, nz = __float_as_uint(idz) >> 31;
                snx = nx ? 0x7632u : 0x7610u; sfx = nx ? 0x7610u : 0x7632u;
                sny = ny ? 0x7632u : 0x7610u; sfy = ny ? 0x7610u : 0x7632u;
                snz = nz ? 0x7632u : 0x7610u; sfz = nz ? 0x7610u : 0x7632u;
                if (CULL) {
                    const float wx = (float)dirs[3 * ray], wy = (float)dirs[3 * ray + 1], wz = (float)dirs[3 * ray + 2];
                    const float inv = rsqrtf(fmaf(wx, wx, fmaf(wy, wy, wz * wz)));
                    ux = wx * inv; uy = wy * inv; uz = wz * inv;
                    ubias = CONE_BIAS * (ux + uy + uz);
                }
                sp = 0; cur = 0;
            }
            next = min(chunk_end, next + __popc(idle));
        }
        if (__all_sync(FULL, cur == DONE)) {
            if (exhausted) break;
            continue;
        }
        // ---- a short burst of 4-wide node steps -----------------------------------------------------------------------
#pragma unroll 1
        for (int step = 0; step < TBURST; ++step) {
            if (cur != DONE) {
                if (STATS) ++st_nodes;
                const uint4* np = mv.wnodes + 4 * (int64_t)cur;                 // 64-byte node: 4 x LDG.128
                uint4 A, B, Cq, D;
                ldg256(np, A, B); ldg256(np + 2, Cq, D);                        // 64-byte node: 2 x LDG.256
                // unused child slots carry an inverted box (emit_wide_nodes_kernel): no ray hits them, no test needed
                float behind = T_BEHIND;
                if (CULL) {
                    // the node's own normal cone: (2^15 + b_k) u_k summed minus the bias term = q . u  (q = int8 axis).  A node
                    // that is ruled out gets an unreachable far-plane bound instead of a branch around the box tests, so the
                    // table look-up's latency hides behind them
                    const float dq = fmaf(top_byte_m15(D.x), ux, fmaf(top_byte_m15(D.y), uy, fmaf(top_byte_m15(D.z), uz, -ubias)));
                    const bool skip = dq < s_thr[D.w >> 24];
                    behind = skip ? __int_as_float(0x7f800000) : T_BEHIND;
                    if (STATS && skip) ++st_culled;
                }
                const bool h0 = slab_hit_signed(A.x, A.y, A.z, snx, sny, snz, sfx, sfy, sfz, idx, idy, idz, oox, ooy, ooz, behind);
                const bool h1 = slab_hit_signed(A.w, B.x, B.y, snx, sny, snz, sfx, sfy, sfz, idx, idy, idz, oox, ooy, ooz, behind);
                const bool h2 = slab_hit_signed(B.z, B.w, Cq.x, snx, sny, snz, sfx, sfy, sfz, idx, idy, idz, oox, ooy, ooz, behind);
                const bool h3 = slab_hit_signed(Cq.y, Cq.z, Cq.w, snx, sny, snz, sfx, sfy, sfz, idx, idy, idz, oox, ooy, ooz, behind);
                // leaves -> the warp's leaf queue (one shared-memory atomic per lane that reached leaves; each entry is one
                // 8-byte store: ray offset and leaf side by side)
                const int c0 = PACKED ? sext24(D.x) : (int)D.x, c1 = PACKED ? sext24(D.y) : (int)D.y,
                          c2 = PACKED ? sext24(D.z) : (int)D.z, c3 = PACKED ? sext24(D.w) : (int)D.w;
                const bool l0 = h0 & (c0 < 0), l1 = h1 & (c1 < 0), l2 = h2 & (c2 < 0), l3 = h3 & (c3 < 0);
                if (l0 | l1 | l2 | l3) {
                    const int nl = (int)l0 + (int)l1 + (int)l2 + (int)l3;
                    const int q0 = atomicAdd(&s_lq_n[w], nl), q1 = q0 + (int)l0, q2 = q1 + (int)l1, q3 = q2 + (int)l2;
                    if (l0) lq[q0] = make_int2(my, ~c0);
                    if (l1) lq[q1] = make_int2(my, ~c1);
                    if (l2) lq[q2] = make_int2(my, ~c2);
                    if (l3) lq[q3] = make_int2(my, ~c3);
                }
                // inner children: continue with the last one that was hit, park the others (predicated, no branches)
                int nxt = DONE;
#define BT_TAKE(hit, code)                                                       \
                {                                                                \
                    const bool in = (hit) & ((code) >= 0);                       \
                    const bool park = in & (nxt != DONE);                        \
                    if (park) stack[sp] = nxt;                                   \
                    sp += (int)park;                                             \
                    nxt = in ? (code) : nxt;                                     \
                    if (TRAV_PREFETCH && park) prefetch_l1(mv.wnodes + 4 * (int64_t)stack[sp - 1]); \
                }
                BT_TAKE(h0, c0) BT_TAKE(h1, c1) BT_TAKE(h2, c2) BT_TAKE(h3, c3)
#undef BT_TAKE
                if (nxt == DONE && sp > 0) nxt = stack[--sp];
                cur = nxt;
            }
        }
        __syncwarp();
        // ---- leaf queue: 32 pre-tests at a time ------------------------------------------------------------------------
        int nq = s_lq_n[w];
        if (nq >= 32) {
            do {
                nq -= 32;
                pretest_pass(nq, 32);
            } while (nq >= 32);
            __syncwarp();
            if (lane == 0) s_lq_n[w] = nq;
            __syncwarp();
        }
    }
    __syncwarp();
    const int nq = s_lq_n[w];
    if (nq > 0) pretest_pass(0, nq);                   // fewer than 32 entries are left
    if (STATS) { atomicAdd(counters + 0, st_nodes); atomicAdd(counters + 1, st_leaves); atomicAdd(counters + 2, st_cand); atomicAdd(counters + 6, st_culled); }
}

// One thread per ray.  Phase 1 walks the LBVH in float32 (both child boxes per 64-byte node) and parks the leaves that
// survive the float32 pre-test; phase 2 runs the reference's float64 narrow phase on the parked leaves (threads of a
// warp are converged again by then); phase 3 orders, de-duplicates and filters the hits and writes the records.
#ifndef CAST_MINB
#define CAST_MINB 6
#endif
#ifndef RSTAGE
#define RSTAGE 2                    // hits per ray staged in shared memory (the usual ray has the origin's triangle + one far hit)
#endif
template <int CAP>
__global__ void __launch_bounds__(128, CAST_MINB)
cast_kernel(MeshView mv, const double* __restrict__ ref_pts, const double* __restrict__ dirs, int64_t n_total,
            int n_rays, bt_ags_params P, int cap, const int32_t* __restrict__ cand_count, const int32_t* __restrict__ cand,
            int32_t* __restrict__ hit_count, bt_hit* __restrict__ hits, double* __restrict__ hit_tt, uint32_t* __restrict__ valid_mask,
            int32_t* __restrict__ overflow) {
    // `hits` == NULL is the lean mode of bt_ags_pipeline: no records, only the ray parameter of every kept hit in the
    // slot-major array hit_tt[h * n_total + ray] (coalesced 8-byte stores); the target choice rebuilds o + t d from it.
    // 72-byte hit records written field by field by one thread per ray are 8-byte stores 576 bytes apart: 32 sectors per
    // store instruction, 584 MB of sector traffic for 144 MB of records at config 3.  The first RSTAGE hits of every ray are
    // staged per warp in shared memory instead and leave as whole records (9 consecutive lanes per record).
    __shared__ double s_rec[4][32 * RSTAGE * 9 + 32];
    const int64_t ray0 = blockIdx.x * (int64_t)blockDim.x + (threadIdx.x & ~31);       // first ray of this warp
    int64_t ray = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const bool live = ray < n_total;
    if (ray0 >= n_total) return;                                                       // the whole warp is past the end
    if (!live) ray = n_total - 1;                  // lanes past the end shadow the last ray (no outputs), the warp stays whole
    const int64_t ref = ray / n_rays;
    int pend[NPEND];
    int n_pend = 0;
    bool over = false;
    double hit_t[CAP];
    int32_t hit_tri[CAP], hit_leaf[CAP];
    int n_hit = 0;

    // the reference's float64 narrow phase for one candidate leaf
    const d3 ray_o = ld3(ref_pts + 3 * ref);
    const d3 ray_d = ld3(dirs + 3 * ray);
    auto test_leaf = [&](int leaf) {
        const int32_t tri = __ldg(mv.leaf_tri + leaf);            // issued with the record loads, not after the test
        double t;
        if (exact_ray_triangle(mv.tri_rec + 16 * (int64_t)leaf, ray_o, ray_d, t)) {
            if (n_hit < cap) { hit_t[n_hit] = t; hit_tri[n_hit] = tri; hit_leaf[n_hit] = leaf; ++n_hit; }
            else over = true;
        }
    };
    // phase 2 of the in-kernel walk (also an emergency flush when the parking list fills up): the parked leaves
    auto run_exact = [&]() {
        for (int i = 0; i < n_pend; ++i) test_leaf(pend[i]);
        n_pend = 0;
    };

    const int n_cand = cand_count ? cand_count[ray] : CCAP + 1;
    if (n_cand <= CCAP) {
        // the usual case: stage 1 (traverse_kernel) already found this ray's candidate leaves -- test them straight from
        // the slot-major list (coalesced loads), nothing is parked in local memory
        for (int i = 0; i < n_cand; ++i) test_leaf(cand[(int64_t)i * n_total + ray]);
    } else {
        // no candidate list (or it overflowed): walk the tree here
        const double odx = ref_pts[3 * ref], ody = ref_pts[3 * ref + 1], odz = ref_pts[3 * ref + 2];
        const double ddx = dirs[3 * ray], ddy = dirs[3 * ray + 1], ddz = dirs[3 * ray + 2];
        const float ox = (float)odx, oy = (float)ody, oz = (float)odz;
        const float dx = (float)ddx, dy = (float)ddy, dz = (float)ddz;
        // the ray in grid coordinates (same parameter t): origin (o - gmin) * scale, direction d * scale
        const float idx = safe_rcp((float)(ddx * mv.gscale[0])), idy = safe_rcp((float)(ddy * mv.gscale[1])), idz = safe_rcp((float)(ddz * mv.gscale[2]));
        const float oox = (float)((odx - mv.gmin[0]) * mv.gscale[0]) * idx, ooy = (float)((ody - mv.gmin[1]) * mv.gscale[1]) * idy,
                    ooz = (float)((odz - mv.gmin[2]) * mv.gscale[2]) * idz;
        int stack[STACK_SIZE];
        int sp = 0;
        int cur = 0;
        while (true) {
            // descend through internal nodes: one 32-byte fetch (2 x LDG.128) decides both children
            while (cur >= 0) {
                const uint4 A = __ldg(mv.qnodes + 2 * (int64_t)cur);
                const uint4 B = __ldg(mv.qnodes + 2 * (int64_t)cur + 1);
                const bool hl = slab_hit(u16lo(A.x), u16hi(A.x), u16lo(A.y), u16hi(A.y), u16lo(A.z), u16hi(A.z), idx, idy, idz, oox, ooy, ooz);
                const bool hr = slab_hit(u16lo(A.w), u16hi(A.w), u16lo(B.x), u16hi(B.x), u16lo(B.y), u16hi(B.y), idx, idy, idz, oox, ooy, ooz);
                const int l = (int)B.z, r = (int)B.w;
                if (hl && hr) { stack[sp++] = r; cur = l; }
                else if (hl) cur = l;
                else if (hr) cur = r;
                else { if (sp == 0) { cur = 0x7fffffff; break; } cur = stack[--sp]; }
            }
            if (cur == 0x7fffffff) break;
            // a leaf
            const int leaf = ~cur;
            if (maybe_hit_f32(mv.tri_f32 + 3 * (int64_t)leaf, ox, oy, oz, dx, dy, dz)) {
                if (n_pend == NPEND) run_exact();
                pend[n_pend++] = leaf;
            }
            if (sp == 0) break;
            cur = stack[--sp];
        }
    }
    run_exact();
    const d3 o = ray_o, d = ray_d;

    // canonical order: (t, triangle id) ascending
    for (int i = 1; i < n_hit; ++i) {
        double t = hit_t[i];
        int32_t tr = hit_tri[i], lf = hit_leaf[i];
        int j = i - 1;
        while (j >= 0 && (hit_t[j] > t || (hit_t[j] == t && hit_tri[j] > tr))) {
            hit_t[j + 1] = hit_t[j]; hit_tri[j + 1] = hit_tri[j]; hit_leaf[j + 1] = hit_leaf[j]; --j;
        }
        hit_t[j + 1] = t; hit_tri[j + 1] = tr; hit_leaf[j + 1] = lf;
    }

    bt_hit* out = hits + ray * (int64_t)cap;
    int n_out = 0;
    uint32_t vmask = 0;
    double key_x[CAP], key_y[CAP], key_z[CAP];        // 1e-8 merge keys of the hits kept so far
    for (int i = 0; i < n_hit; ++i) {
        const double t = hit_t[i];
        const d3 p = add3(scl3(d, t), o);
        // trimesh intersects_id(return_locations=True): unique rows of (location, ray) at 1e-8
        const double kx = rint(p.x * 1e8 - 1e-6), ky = rint(p.y * 1e8 - 1e-6), kz = rint(p.z * 1e8 - 1e-6);
        bool dup = false;
        for (int j = 0; j < n_out; ++j)
            if (key_x[j] == kx && key_y[j] == ky && key_z[j] == kz) { dup = true; break; }
        if (dup) continue;
        key_x[n_out] = kx; key_y[n_out] = ky; key_z[n_out] = kz;
        const int32_t tri = hit_tri[i];
        // The reference applies its masks in sequence and computes normals / angles only for the hits that are still
        // alive (sampling.py:239-319).  flags = the FIRST filter that rejects the hit; later quantities of a rejected
        // hit are never computed by the reference and are left at zero / NaN here.
        uint32_t flags = 0;
        d3 nrm = mk3(0.0, 0.0, 0.0);
        double ang = __longlong_as_double(0x7ff8000000000000ll);
        // sampling.py:239  ~np.isclose(loc, p_r, atol=1e-11).all(-1)   (rtol = 1e-5 relative to p_r)
        const bool close = (fabs(p.x - o.x) <= 1e-11 + 1e-5 * fabs(o.x)) & (fabs(p.y - o.y) <= 1e-11 + 1e-5 * fabs(o.y)) &
                           (fabs(p.z - o.z) <= 1e-11 + 1e-5 * fabs(o.z));
        const d3 dv = sub3(p, o);
        if (close) flags = BT_HIT_ORIGIN;
        else if (const double dist = norm3(dv);                      // sampling.py:248-251
                 !((dist <= P.opening_width - P.width_tolerance) | (dist >= P.min_grasp_width))) flags = BT_HIT_WIDTH;
        else {
            // sampling.py:259, :289-308
            nrm = interpolated_normal(mv, hit_leaf[i], p);
            const double dotp = dot3(dv, nrm);
            ang = atan(norm3(cross3(dv, nrm)) / dotp);
            if (!(dotp > 0.0)) flags = BT_HIT_SIGN;
            else if (!(ang <= P.angle_max)) flags = BT_HIT_CONE;
            else if (P.use_below_z && !(p.z > P.no_contact_below_z)) flags = BT_HIT_BELOW_Z;
        }
        bt_hit h;
        h.loc[0] = p.x; h.loc[1] = p.y; h.loc[2] = p.z;
        h.nrm[0] = nrm.x; h.nrm[1] = nrm.y; h.nrm[2] = nrm.z;
        h.t = t; h.angle = ang; h.tri = tri; h.flags = flags;
        if (flags == 0u) vmask |= 1u << n_out;
        if (!hits) {
            if (live) hit_tt[(int64_t)n_out * n_total + ray] = t;
        } else if (n_out < RSTAGE) {
            double* sr = s_rec[threadIdx.x >> 5] + ((threadIdx.x & 31) * RSTAGE + n_out) * 9;
            sr[0] = p.x; sr[1] = p.y; sr[2] = p.z; sr[3] = nrm.x; sr[4] = nrm.y; sr[5] = nrm.z; sr[6] = t; sr[7] = ang;
            sr[8] = __hiloint2double((int)flags, tri);                                 // tri in the low word, flags in the high word
        } else if (live) {
            out[n_out] = h;
        }
        ++n_out;
    }
    if (hits) {
        const int lane = threadIdx.x & 31;
        const double* sw = s_rec[threadIdx.x >> 5];
        double* gw = reinterpret_cast<double*>(hits + ray0 * (int64_t)cap);
        __syncwarp();                                                                  // the staged records are visible to the whole warp
#pragma unroll
        for (int k = 0; k < RSTAGE; ++k) {
            const unsigned has = __ballot_sync(0xffffffffu, live && n_out > k);
            if (!has) break;
            for (int e = lane; e < 32 * 9; e += 32) {
                const int r = e / 9, w = e - 9 * r;
                if ((has >> r) & 1u) gw[((int64_t)r * cap + k) * 9 + w] = sw[(r * RSTAGE + k) * 9 + w];
            }
        }
    }
    if (!live) return;
    hit_count[ray] = n_out;
    if (valid_mask) valid_mask[ray] = vmask;
    if (over && overflow) atomicAdd(overflow, 1);
}

// ---- S1 ------------------------------------------------------------------------------------------------
// one area-weighted surface sample (open3d SamplePointsUniformlyImpl: triangle f owns sample slots
// [round(cdf[f-1]*n), round(cdf[f]*n)) ) -> point, triangle normal, triangle
__device__ __forceinline__ int32_t surface_sample(const double* __restrict__ cdf, const int32_t* __restrict__ faces,
                                                  const double* __restrict__ verts, int64_t nF, int64_t n, int64_t i, uint64_t seed,
                                                  d3& p, d3& nv) {
    int64_t lo = 0, hi = nF - 1;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if ((int64_t)llrint(cdf[mid] * (double)n) > i) hi = mid; else lo = mid + 1;
    }
    const int32_t f = (int32_t)lo;
    uint32_t r[4];
    philox4x32(seed, (uint64_t)i, 0x53414d50ull, r);
    const double r1 = u01(r[0], r[1]), r2 = u01(r[2], r[3]);
    const double s = sqrt(r1);
    const double a = 1.0 - s, b = s * (1.0 - r2), c = s * r2;
    const int32_t* fc = faces + 3 * (int64_t)f;
    d3 v0 = ld3(verts + 3 * (int64_t)fc[0]), v1 = ld3(verts + 3 * (int64_t)fc[1]), v2 = ld3(verts + 3 * (int64_t)fc[2]);
    p = add3(add3(scl3(v0, a), scl3(v1, b)), scl3(v2, c));
    d3 cr = cross3(sub3(v1, v0), sub3(v2, v0));
    double nn = norm3(cr);
    nv = nn > 0 ? div3(cr, nn) : mk3(0, 0, 1);
    return f;
}

__global__ void sample_surface_kernel(const double* __restrict__ cdf, const int32_t* __restrict__ faces,
                                      const double* __restrict__ verts, int64_t nF, int64_t n, uint64_t seed,
                                      double* __restrict__ pts, double* __restrict__ nrm, int32_t* __restrict__ tri_out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    d3 p, nv;
    const int32_t f = surface_sample(cdf, faces, verts, nF, n, i, seed, p, nv);
    st3(pts + 3 * i, p);
    st3(nrm + 3 * i, nv);
    if (tri_out) tri_out[i] = f;
}

// ---- S2 ------------------------------------------------------------------------------------------------
// ray i (= ref * n_rays + r) of the friction cone about -normal  (sampling.py:17-48, util.py:7-29)
__device__ __forceinline__ d3 cone_ray(d3 normal, int64_t ref, int64_t i, double angle, uint64_t seed, bool uniform_on_plane = false) {
    uint32_t r[4];
    philox4x32(seed, (uint64_t)i, 0x434f4e45ull, r);
    const double az = u01(r[0], r[1]) * (2.0 * 3.14159265358979323846);      // U(0, 2 pi)   sampling.py:32
    const double inc = uniform_on_plane ? atan(u01(r[2], r[3]) * tan(angle))   // arctan(U(0, tan(angle)))  sampling.py:34
                                        : u01(r[2], r[3]) * angle;             // U(0, angle)  sampling.py:36
    double si, ci, sa, ca;
    sincos(inc, &si, &ci);
    sincos(az, &sa, &ca);
    const d3 c = mk3(si * ca, si * sa, ci);
    // util.rotation_to_align_vectors([0,0,1], axis): R = 2 k k^T / (k^T k) - I,  k = z + axis/|axis|
    d3 axis = mk3(-normal.x, -normal.y, -normal.z);
    const double an = norm3(axis);
    d3 k = mk3(axis.x / an, axis.y / an, 1.0 + axis.z / an);
    if (norm3(k) == 0.0) {                      // axis == -z exactly: half turn about any vector orthogonal to z
        uint32_t q[4];
        philox4x32(seed, (uint64_t)ref, 0x4f525448ull, q);
        const double phi = u01(q[0], q[1]) * (2.0 * 3.14159265358979323846);
        k = mk3(cos(phi), sin(phi), 0.0);
    }
    const double kk = dot3(k, k);
    const double kc = dot3(k, c);
    const double f = 2.0 * kc / kk;
    return mk3(f * k.x - c.x, f * k.y - c.y, f * k.z - c.z);
}

__global__ void cone_rays_kernel(const double* __restrict__ normals, int64_t n_ref, int n_rays, double angle,
                                 uint64_t seed, int uniform_on_plane, double* __restrict__ dirs) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_ref * n_rays) return;
    const int64_t ref = i / n_rays;
    st3(dirs + 3 * i, cone_ray(ld3(normals + 3 * ref), ref, i, angle, seed, uniform_on_plane != 0));
}

// util.generate_random_unit_vector (util.py:32-43): normalised N(0,1)^3, redrawn while shorter than 1e-5
__device__ __forceinline__ d3 random_unit_vector(int64_t i, uint64_t seed) {
    d3 v = mk3(0, 0, 1);
    for (uint64_t attempt = 0; attempt < 8; ++attempt) {
        uint32_t a[4], b[4];
        philox4x32(seed, (uint64_t)i, 0x554e4954ull + (attempt << 32), a);
        philox4x32(seed, (uint64_t)i, 0x554e4955ull + (attempt << 32), b);
        // Box-Muller: three standard normals
        const double u1 = 1.0 - u01(a[0], a[1]), u2 = u01(a[2], a[3]), u3 = 1.0 - u01(b[0], b[1]), u4 = u01(b[2], b[3]);
        const double r1 = sqrt(-2.0 * log(u1)), r2 = sqrt(-2.0 * log(u3));
        const double tp = 2.0 * 3.14159265358979323846;
        d3 g = mk3(r1 * cos(tp * u2), r1 * sin(tp * u2), r2 * cos(tp * u4));
        const double mag = norm3(g);
        if (mag > 1e-5) { v = div3(g, mag); break; }
    }
    return v;
}

__global__ void unit_vectors_kernel(int64_t n, uint64_t seed, double* __restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    st3(out + 3 * i, random_unit_vector(i, seed));
}

// ---- S1 + S2 + frame vectors in one launch (bt_ags_pipeline with device-generated candidates) ---------------------
// A block owns GEN_REFS reference points: warp 0 draws their surface samples, warp 1 their frame vectors, then all
// threads generate the GEN_REFS x n_rays cone rays from the normals in shared memory.  Same Philox keys as the three
// stand-alone kernels, so the values are the same bits.  `seed_add` (nullable): a device-resident offset added to the
// seed -- lets a captured CUDA graph of the pipeline be replayed with a new seed (see bt_graph_*).
// A keyed pseudo-random PERMUTATION of [0, n): balanced Feistel network on the smallest even number of bits that holds
// n - 1, cycle-walked back into the domain.  sample() uses it as the reference's np.random.shuffle of the surface samples
// (sampling.py:201) without materialising them: reference point j of a call is surface sample perm(j).
__device__ __forceinline__ uint64_t feistel_perm(uint64_t x, uint64_t n, uint64_t key) {
    int half = 1;
    while ((1ull << (2 * half)) < n) ++half;
    const uint64_t mask = (1ull << half) - 1;
    do {
        uint64_t l = x >> half, r = x & mask;
#pragma unroll 1
        for (int round = 0; round < 6; ++round) {
            uint32_t o[4];
            philox4x32(key, r, 0x5045524dull + ((uint64_t)round << 32), o);
            const uint64_t f = (((uint64_t)o[0] << 32) | o[1]) & mask;
            const uint64_t nl = r;
            r = l ^ f;
            l = nl;
        }
        x = (l << half) | r;
    } while (x >= n);
    return x;
}

constexpr int GEN_REFS = 8;
// `dyn` (nullable, device): [0] added to the seed, [1] position of this batch in the shuffled sample list, [2] key of the
// shuffle and of the surface samples -- what changes from call to call when the pipeline is replayed as a CUDA graph.
// surface_total > 0: the batch's reference points are samples perm(dyn[1] + i mod surface_total) of a set of surface_total
// area-weighted samples (S1 of a whole sample() call, never materialised); else plain samples 0 .. n_ref - 1.
__global__ void __launch_bounds__(256)
generate_kernel(const double* __restrict__ cdf, const int32_t* __restrict__ faces, const double* __restrict__ verts, int64_t nF,
                int64_t n_ref, int n_rays, double angle, uint64_t seed, const uint64_t* __restrict__ dyn, int64_t surface_total, int do_sample, int do_dirs,
                int do_frame, double* __restrict__ pts, double* __restrict__ nrm, double* __restrict__ frame, double* __restrict__ dirs) {
    __shared__ double s_n[GEN_REFS][3];
    if (dyn) seed += dyn[0];
    const int64_t ref0 = blockIdx.x * (int64_t)GEN_REFS;
    const int t = threadIdx.x;
    if (t < GEN_REFS && ref0 + t < n_ref) {
        const int64_t i = ref0 + t;
        d3 nv;
        if (do_sample && surface_total > 0 && dyn) {
            d3 p;
            const uint64_t slot = feistel_perm((dyn[1] + (uint64_t)i) % (uint64_t)surface_total, (uint64_t)surface_total, dyn[2]);
            surface_sample(cdf, faces, verts, nF, surface_total, (int64_t)slot, dyn[2], p, nv);
            st3(pts + 3 * i, p);
            st3(nrm + 3 * i, nv);
        } else if (do_sample) {
            d3 p;
            surface_sample(cdf, faces, verts, nF, n_ref, i, seed, p, nv);
            st3(pts + 3 * i, p);
            st3(nrm + 3 * i, nv);
        } else nv = ld3(nrm + 3 * i);
        s_n[t][0] = nv.x; s_n[t][1] = nv.y; s_n[t][2] = nv.z;
    }
    if (do_frame && t >= 32 && t < 32 + GEN_REFS && ref0 + (t - 32) < n_ref) {
        const int64_t i = ref0 + (t - 32);
        st3(frame + 3 * i, random_unit_vector(i, seed + 0x3000));
    }
    if (!do_dirs) return;
    __syncthreads();
    const int n_here = (int)min((int64_t)GEN_REFS, n_ref - ref0) * n_rays;
    for (int j = t; j < n_here; j += blockDim.x) {
        const int lr = j / n_rays;
        const int64_t ref = ref0 + lr, i = ref0 * n_rays + j;
        st3(dirs + 3 * i, cone_ray(mk3(s_n[lr][0], s_n[lr][1], s_n[lr][2]), ref, i, angle, seed + 0x1000));
    }
}

// ---- R7 ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t choice_key(uint64_t seed, int64_t ref, int ordinal) {
    uint32_t r[4];
    philox4x32(seed, (uint64_t)ref, 0x43484f49ull + ((uint64_t)ordinal << 32), r);
    return (((uint64_t)r[0] << 32) | r[1]) >> 1;      // 63 bits, so that ~0 can be the 'none' sentinel
}

// one warp per reference point: number of valid hits, capped at max_targets
// number of valid hits of one ray: from the caster's bit mask if there is one, else by scanning the hit records
__device__ __forceinline__ uint32_t ray_valid_mask(const uint32_t* __restrict__ valid_mask, const int32_t* __restrict__ hit_count,
                                                   const bt_hit* __restrict__ hits, int64_t ray, int cap) {
    if (valid_mask) return valid_mask[ray];
    uint32_t m = 0;
    const int c = hit_count[ray];
    for (int h = 0; h < c; ++h) m |= (hits[ray * cap + h].flags == 0u ? 1u : 0u) << h;
    return m;
}

__global__ void select_count_kernel(const int32_t* __restrict__ hit_count, const bt_hit* __restrict__ hits,
                                    const uint32_t* __restrict__ valid_mask, int64_t n_ref, int n_rays, int cap, int max_targets,
                                    int32_t* __restrict__ pair_count) {
    const int64_t ref = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (ref >= n_ref) return;
    int cnt = 0;
    for (int r = lane; r < n_rays; r += 32) {
        cnt += __popc(ray_valid_mask(valid_mask, hit_count, hits, ref * n_rays + r, cap));
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, s);
    if (lane == 0) pair_count[ref] = min(cnt, max_targets);
}

__global__ void zero_first_kernel(int64_t* p) { *p = 0; }

// where the location of hit h of a ray comes from: its record, or (lean pipeline) o + t d with the stored ray parameter --
// the same expression, on the same float64 inputs, as cast_kernel evaluates for the record
struct HitSource { const double* hit_tt; const double* ref_pts; const double* dirs; int64_t n_total; };
__device__ __forceinline__ d3 hit_location(const bt_hit* __restrict__ hits, const HitSource& hs, int64_t ref, int64_t ray, int h, int cap) {
    if (hits) { const bt_hit* hp = hits + ray * cap + h; return mk3(hp->loc[0], hp->loc[1], hp->loc[2]); }
    const double t = hs.hit_tt[(int64_t)h * hs.n_total + ray];
    return add3(scl3(ld3(hs.dirs + 3 * ray), t), ld3(hs.ref_pts + 3 * ref));
}

// one warp writes the n_take chosen contact pairs of reference point `ref` to slots [base, base + n_take)
// (`total` = number of valid hits of the reference point; a random subset is drawn iff seed != 0 and total > max_targets)
__device__ __forceinline__ void select_fill_ref(const int32_t* __restrict__ hit_count, const bt_hit* __restrict__ hits, const HitSource& hs,
                                                const uint32_t* __restrict__ valid_mask, int64_t ref, int lane, int n_rays, int cap,
                                                int max_targets, uint64_t seed, int64_t base, int n_take, int total, int64_t cap_pairs,
                                                int32_t* __restrict__ pair_ref, double* __restrict__ pair_q) {
    const bool random_subset = (seed != 0) && (total > max_targets);

    if (!random_subset) {
        // canonical order: ray ascending, then t ascending; ordinals via a warp scan per chunk of 32 rays
        int run = 0;
        for (int r0 = 0; r0 < n_rays; r0 += 32) {
            const int r = r0 + lane;
            uint32_t vm = 0;
            int64_t ray = ref * n_rays + r;
            if (r < n_rays) vm = ray_valid_mask(valid_mask, hit_count, hits, ray, cap);
            const int nv = __popc(vm);
            int incl = nv;
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                int v = __shfl_up_sync(0xffffffffu, incl, s);
                if (lane >= s) incl += v;
            }
            int ord = run + incl - nv;
            for (; vm; vm &= vm - 1) {
                const int h = __ffs(vm) - 1;
                {
                    if (ord < n_take && base + ord < cap_pairs) {
                        const d3 q = hit_location(hits, hs, ref, ray, h, cap);
                        pair_ref[base + ord] = (int32_t)ref;
                        pair_q[3 * (base + ord) + 0] = q.x;
                        pair_q[3 * (base + ord) + 1] = q.y;
                        pair_q[3 * (base + ord) + 2] = q.z;
                    }
                    ++ord;
                }
            }
            run += __shfl_sync(0xffffffffu, incl, 31);
        }
        return;
    }
    // random subset without replacement: the n_take hits with the smallest Philox keys, in key order
    uint64_t prev_key = 0;
    int prev_ord = -1;
    for (int pick = 0; pick < n_take; ++pick) {
        uint64_t best_key = ~0ull;
        int best_ord = 0x7fffffff, best_slot = -1;
        int run = 0;
        for (int r0 = 0; r0 < n_rays; r0 += 32) {
            const int r = r0 + lane;
            uint32_t vm = 0;
            int64_t ray = ref * n_rays + r;
            if (r < n_rays) vm = ray_valid_mask(valid_mask, hit_count, hits, ray, cap);
            const int nv = __popc(vm);
            int incl = nv;
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                int v = __shfl_up_sync(0xffffffffu, incl, s);
                if (lane >= s) incl += v;
            }
            int ord = run + incl - nv;
            for (; vm; vm &= vm - 1) {
                const int h = __ffs(vm) - 1;
                {
                    const uint64_t key = choice_key(seed, ref, ord);
                    const bool after_prev = (prev_ord < 0) || key > prev_key || (key == prev_key && ord > prev_ord);
                    const bool better = key < best_key || (key == best_key && ord < best_ord);
                    if (after_prev && better) { best_key = key; best_ord = ord; best_slot = r * cap + h; }
                    ++ord;
                }
            }
            run += __shfl_sync(0xffffffffu, incl, 31);
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            const uint64_t ok = __shfl_xor_sync(0xffffffffu, best_key, s);
            const int oo = __shfl_xor_sync(0xffffffffu, best_ord, s);
            const int os = __shfl_xor_sync(0xffffffffu, best_slot, s);
            if (ok < best_key || (ok == best_key && oo < best_ord)) { best_key = ok; best_ord = oo; best_slot = os; }
        }
        if (lane == 0 && best_slot >= 0 && base + pick < cap_pairs) {
            const d3 q = hit_location(hits, hs, ref, ref * n_rays + best_slot / cap, best_slot % cap, cap);      // best_slot = r * cap + h
            pair_ref[base + pick] = (int32_t)ref;
            pair_q[3 * (base + pick) + 0] = q.x;
            pair_q[3 * (base + pick) + 1] = q.y;
            pair_q[3 * (base + pick) + 2] = q.z;
        }
        prev_key = best_key;
        prev_ord = best_ord;
    }
}

// number of valid hits of reference point `ref` (all lanes of the warp get the sum)
__device__ __forceinline__ int ref_valid_total(const uint32_t* __restrict__ valid_mask, const int32_t* __restrict__ hit_count,
                                               const bt_hit* __restrict__ hits, int64_t ref, int lane, int n_rays, int cap) {
    int total = 0;
    for (int r = lane; r < n_rays; r += 32) total += __popc(ray_valid_mask(valid_mask, hit_count, hits, ref * n_rays + r, cap));
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) total += __shfl_xor_sync(0xffffffffu, total, s);
    return total;
}

__global__ void select_fill_kernel(const int32_t* __restrict__ hit_count, const bt_hit* __restrict__ hits, HitSource hs,
                                   const uint32_t* __restrict__ valid_mask, int64_t n_ref, int n_rays, int cap, int max_targets, uint64_t seed,
                                   const int64_t* __restrict__ pair_offset, int64_t cap_pairs,
                                   int32_t* __restrict__ pair_ref, double* __restrict__ pair_q) {
    const int64_t ref = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (ref >= n_ref) return;
    const int64_t base = pair_offset[ref];
    const int n_take = (int)(pair_offset[ref + 1] - base);
    if (n_take == 0) return;
    const int total = ref_valid_total(valid_mask, hit_count, hits, ref, lane, n_rays, cap);
    select_fill_ref(hit_count, hits, hs, valid_mask, ref, lane, n_rays, cap, max_targets, seed, base, n_take, total, cap_pairs, pair_ref, pair_q);
}

// ---- R7 in ONE launch (bt_ags_pipeline): count, exclusive scan and fill ---------------------------------------------------
// A block owns SEL_REFS reference points, one warp each.  The blocks' pair counts are scanned across the grid with a
// single-pass chained scan (decoupled look-back): a block publishes its aggregate, looks back over its predecessors'
// status words until it meets an inclusive prefix, and publishes its own.  Block ids come from a ticket counter, so a block
// only ever waits for blocks that are already running.  Status word: bits 63..62 = 1 aggregate / 2 inclusive prefix.
#ifndef BT_SEL_REFS
#define BT_SEL_REFS 8
#endif
// reference points per block, <= 32 (one warp scans the block's counts).  8 = 256-thread blocks: a 1 024-thread block needs
// a whole SM's register file to itself and waits for one to drain while other batches' kernels are resident (step with four
// batches in flight: 0.479 ms at 32, 0.474 at 16, 0.472 at 8)
constexpr int SEL_REFS = BT_SEL_REFS;
__global__ void __launch_bounds__(32 * SEL_REFS)
select_fused_kernel(const int32_t* __restrict__ hit_count, const bt_hit* __restrict__ hits, HitSource hs,
                    const uint32_t* __restrict__ valid_mask, int64_t n_ref, int n_rays, int cap, int max_targets, uint64_t seed,
                    const uint64_t* __restrict__ seed_add, unsigned long long* status, unsigned* ticket, unsigned n_blocks,
                    int32_t* __restrict__ pair_count, int64_t* __restrict__ pair_offset, int64_t cap_pairs,
                    int32_t* __restrict__ pair_ref, double* __restrict__ pair_q) {
    __shared__ unsigned s_bid;
    __shared__ int s_cnt[SEL_REFS];
    __shared__ unsigned long long s_base;
    if (seed != 0 && seed_add) seed += *seed_add;              // seed 0 = canonical choice, whatever the offset
    if (threadIdx.x == 0) s_bid = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned bid = s_bid;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t ref = (int64_t)bid * SEL_REFS + w;
    int total = 0;
    if (ref < n_ref) total = ref_valid_total(valid_mask, hit_count, hits, ref, lane, n_rays, cap);
    const int n_take = min(total, max_targets);
    if (lane == 0) { s_cnt[w] = n_take; if (ref < n_ref) pair_count[ref] = n_take; }
    __syncthreads();
    if (w == 0) {
        const int c = lane < SEL_REFS ? s_cnt[lane] : 0;
        int incl = c;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, s); if (lane >= s) incl += v; }
        const unsigned long long agg = (unsigned long long)__shfl_sync(0xffffffffu, incl, 31);
        const unsigned long long excl = lookback_exclusive(status, bid, agg, lane);
        if (lane < SEL_REFS) s_cnt[lane] = incl - c;           // exclusive prefix inside the block
        if (lane == 0) s_base = excl;
        const int64_t r = (int64_t)bid * SEL_REFS + lane;
        if (lane < SEL_REFS && r < n_ref) pair_offset[r] = (int64_t)excl + (incl - c);
        if (bid == n_blocks - 1 && lane == 31) pair_offset[n_ref] = (int64_t)(excl + agg);
    }
    __syncthreads();
    if (ref >= n_ref || n_take == 0) return;
    select_fill_ref(hit_count, hits, hs, valid_mask, ref, lane, n_rays, cap, max_targets, seed, (int64_t)s_base + s_cnt[w], n_take, total,
                    cap_pairs, pair_ref, pair_q);
}

// ---- O1 / O2 -------------------------------------------------------------------------------------------
struct OrientTable { double c[64]; double s[64]; };

__global__ void expand_kernel(const double* __restrict__ ref_pts, const int64_t* __restrict__ pair_offset,
                              const int32_t* __restrict__ pair_ref, const double* __restrict__ pair_q,
                              const double* __restrict__ frame_vecs, int64_t n_ref, int n_o, int halfspace,
                              OrientTable tab, float* __restrict__ gs, double* __restrict__ contacts, int64_t* __restrict__ n_rows_out) {
    const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t n_pairs = pair_offset[n_ref];
    if (row == 0 && n_rows_out) *n_rows_out = n_pairs * n_o;       // number of posed grasps (bt_ags_pipeline: n_grasps)
    if (row >= n_pairs * n_o) return;
    const int64_t slot = row / n_o;                   // target-major slot: owns width + contacts of this row
    const int64_t ref = pair_ref[slot];
    const int64_t base = pair_offset[ref];
    const int64_t k = pair_offset[ref + 1] - base;
    const int64_t r = row - base * n_o;               // row index inside the reference point's group
    const d3 p_r = ld3(ref_pts + 3 * ref);
    // width and contacts follow the target-major slot (sampling.py:181 / :128 and :341-344)
    const d3 q_slot = ld3(pair_q + 3 * slot);
    const double width = norm3(sub3(q_slot, p_r));
    int64_t ti;           // which target's pose
    int oi;               // which orientation
    if (halfspace) { ti = r / n_o; oi = (int)(r % n_o); }      // sampling.py:117  index = i * n_o + j
    else { oi = (int)(r / k); ti = r % k; }                    // sampling.py:177  orientation-major
    const d3 q = ld3(pair_q + 3 * (base + ti));
    const d3 dv = sub3(q, p_r);
    const d3 centre = add3(p_r, scl3(dv, 0.5));
    const double dist = norm3(dv);
    d3 x = div3(dv, dist);
    d3 y, z;
    if (!halfspace) {
        const d3 u = ld3(frame_vecs + 3 * ref);
        d3 y0 = cross3(x, u);
        y0 = div3(y0, norm3(y0));
        d3 z0 = cross3(x, y0);
        z0 = div3(z0, norm3(z0));
        // tf_basis @ tf_rot (4x4 float64 matmul = ascending-k FMA chain; zero terms drop out exactly):
        //   y = fma(z0, sin, y0*cos)     z = fma(z0, cos, y0*(-sin))
        const double c = tab.c[oi], s = tab.s[oi];
        y = mk3(fma(z0.x, s, y0.x * c), fma(z0.y, s, y0.y * c), fma(z0.z, s, y0.z * c));
        z = mk3(fma(z0.x, c, y0.x * (-s)), fma(z0.y, c, y0.y * (-s)), fma(z0.z, c, y0.z * (-s)));
    } else {
        if (x.z < 0) x = mk3(-x.x, -x.y, -x.z);
        // y_tangent = -cross(x, [0,0,1]) normalised
        d3 yt = mk3(-(x.y * 1.0 - x.z * 0.0), -(x.z * 0.0 - x.x * 1.0), -(x.x * 0.0 - x.y * 0.0));
        yt = div3(yt, norm3(yt));
        // Rodrigues rotation of y_tangent about x by theta
        const double c = tab.c[oi], s = tab.s[oi];
        const d3 xc = cross3(x, yt);
        const double xd = dot3(x, yt) * (1.0 - c);
        z = mk3(yt.x * c + xc.x * s + x.x * xd, yt.y * c + xc.y * s + x.y * xd, yt.z * c + xc.z * s + x.z * xd);
        y = cross3(z, x);
        y = div3(y, norm3(y));
    }
    float* g = gs + 14 * row;
    g[0] = (float)centre.x; g[1] = (float)centre.y; g[2] = (float)centre.z;
    g[3] = (float)x.x; g[4] = (float)y.x; g[5] = (float)z.x;
    g[6] = (float)x.y; g[7] = (float)y.y; g[8] = (float)z.y;
    g[9] = (float)x.z; g[10] = (float)y.z; g[11] = (float)z.z;
    g[12] = 0.0f;
    g[13] = (float)width;
    if (contacts) {
        double* c = contacts + 6 * row;
        st3(c, p_r);
        st3(c + 3, q_slot);
    }
}

static MeshView view_of(const MeshDev& m) {
    MeshView v;
    v.nodes = m.nodes;
    v.qnodes = m.qnodes;
    v.wnodes = m.wnodes;
    for (int k = 0; k < 3; ++k) { v.gmin[k] = m.grid_min[k]; v.gscale[k] = m.grid_scale[k]; }
    v.tri_rec = m.tri_rec; v.tri_f32 = m.tri_f32; v.leaf_tri = m.leaf_tri; v.faces = m.faces; v.verts = m.verts;
    v.vnormals = m.vnormals; v.tri_vvn = m.tri_vvn; v.leaf_cone = m.leaf_cone;
    return v;
}

}  // namespace bt

using namespace bt;

extern "C" int bt_sample_surface(const bt_mesh* mesh, int64_t n, uint64_t seed, double* d_pts, double* d_nrm,
                                 int32_t* d_tri, void* stream) {
    BT_REQUIRE(mesh && d_pts && d_nrm, "null pointer");
    BT_REQUIRE(n >= 0, "negative count");
    if (n == 0) return BT_OK;
    BT_CUDA(cudaSetDevice(mesh->m.device));
    sample_surface_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(
        mesh->m.area_cdf, mesh->m.faces, mesh->m.verts, mesh->m.nF, n, seed, d_pts, d_nrm, d_tri);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

extern "C" int bt_cone_rays_ex(const double* d_normals, int64_t n_ref, int32_t n_rays, double angle, int32_t uniform_on_plane,
                               uint64_t seed, double* d_dirs, void* stream) {
    BT_REQUIRE(d_normals && d_dirs, "null pointer");
    BT_REQUIRE(n_ref >= 0 && n_rays > 0, "bad counts");
    BT_REQUIRE(!uniform_on_plane || (angle >= 0.0 && angle < 1.5707), "uniform_on_plane needs a cone half-angle below pi/2");
    if (n_ref == 0) return BT_OK;
    cone_rays_kernel<<<(unsigned)ceil_div(n_ref * n_rays, 256), 256, 0, (cudaStream_t)stream>>>(
        d_normals, n_ref, n_rays, angle, seed, uniform_on_plane, d_dirs);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

extern "C" int bt_cone_rays(const double* d_normals, int64_t n_ref, int32_t n_rays, double angle, uint64_t seed,
                            double* d_dirs, void* stream) {
    return bt_cone_rays_ex(d_normals, n_ref, n_rays, angle, 0, seed, d_dirs, stream);
}

extern "C" int bt_random_unit_vectors(int64_t n, uint64_t seed, double* d_out, void* stream) {
    BT_REQUIRE(d_out && n >= 0, "bad arguments");
    if (n == 0) return BT_OK;
    unit_vectors_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(n, seed, d_out);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

extern "C" int bt_ags_cast(const bt_mesh* mesh, const double* d_ref_pts, const double* d_dirs, int64_t n_ref,
                           int32_t n_rays, const bt_ags_params* params, int32_t* d_hit_count, bt_hit* d_hits,
                           uint32_t* d_valid_mask, int32_t* d_overflow, void* stream) {
    BT_REQUIRE(n_ref == 0 || d_hits, "null pointer");
    return bt::ags_cast_impl(mesh, d_ref_pts, d_dirs, n_ref, n_rays, params, d_hit_count, d_hits, nullptr, d_valid_mask, d_overflow, false, stream);
}

// d_hits == NULL: lean mode (bt_ags_pipeline without hit records), the ray parameters go to d_hit_t [cap, n_ref * n_rays]
int bt::ags_cast_impl(const bt_mesh* mesh, const double* d_ref_pts, const double* d_dirs, int64_t n_ref,
                      int32_t n_rays, const bt_ags_params* params, int32_t* d_hit_count, bt_hit* d_hits, double* d_hit_t,
                      uint32_t* d_valid_mask, int32_t* d_overflow, bool cull, void* stream) {
    BT_REQUIRE(mesh && params, "null pointer");
    BT_REQUIRE(n_ref >= 0 && n_rays > 0, "bad counts");
    BT_REQUIRE(n_ref == 0 || (d_ref_pts && d_dirs && d_hit_count && (d_hits || (d_hit_t && d_valid_mask))), "null pointer");     // empty arrays may be NULL
    const int cap = params->max_hits_per_ray;
    BT_REQUIRE(cap >= 1 && cap <= 32, "max_hits_per_ray must be in [1, 32]");
    BT_CUDA(cudaSetDevice(mesh->m.device));
    cudaStream_t st = (cudaStream_t)stream;
    if (d_overflow) BT_CUDA(cudaMemsetAsync(d_overflow, 0, sizeof(int32_t), st));
    if (n_ref == 0) return BT_OK;
    const int64_t total = n_ref * n_rays;
    const unsigned grid = (unsigned)ceil_div(total, 128);
    MeshView mv = view_of(mesh->m);
    // stage 1: float32 traversal with ray re-fetch -> candidate leaves per ray
    Scratch cc, cl;
    BT_CUDA(cc.alloc(sizeof(int32_t) * (total + 1), st));                 // [total] candidate counts + the chunk counter
    BT_CUDA(cl.alloc(sizeof(int32_t) * total * CCAP, st));
    BT_CUDA(cudaMemsetAsync(cc.p, 0, sizeof(int32_t) * (total + 1), st)); // the traversal appends with atomics
    int sm_count = 148;
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, mesh->m.device);
    // chunk of rays a warp takes from the global counter at a time (coherent: consecutive rays of one reference point);
    // persistent grid, TRAV_MINB blocks per SM.  Fixed slices per warp (round 1) left the lanes of every warp idle while
    // the slowest ray of the slice finished: 21.6 of 32 lanes active at C3.
    int64_t rpw = 32;
    if (const char* e = getenv("BT_CAST_RPW")) { const int v = atoi(e); if (v >= 32 && v <= 4096) rpw = v; }       // tuning knob
    BT_REQUIRE(total + rpw < ((int64_t)1 << 31), "more than 2^31 rays in one call");
    // blocks per SM of the persistent grid: TRAV_MINB fills the register file; fewer leave room for another batch's
    // kernels on the same SM (several batches in flight on different streams), BT_TRAV_BPS is the tuning knob
    static int trav_bps = -1;
    if (trav_bps < 0) { const char* e = getenv("BT_TRAV_BPS"); const int v = e ? atoi(e) : 0; trav_bps = (v >= 1 && v <= TRAV_MINB) ? v : TRAV_MINB; }
    const int64_t n_warps = std::min<int64_t>(ceil_div(total, rpw), (int64_t)sm_count * trav_bps * TWARPS);
    // cone culling: only where the caller wants valid hits and nothing else (lean pipeline), on packed nodes
    static int cull_env = -1;
    if (cull_env < 0) { const char* e = getenv("BT_CAST_CULL"); cull_env = e ? atoi(e) : 1; }
    const bool packed = mesh->m.wide_packed != 0;
    const bool do_cull = cull && packed && cull_env != 0 && params->angle_max >= 0.0 && params->angle_max < 1.5;
    const unsigned tgrid = (unsigned)ceil_div(n_warps * 32, 128);
#define BT_TRAVERSE(STATS, PACKED, CULL, CNT)                                                                              \
    traverse_kernel<STATS, PACKED, CULL><<<tgrid, 128, 0, st>>>(mv, d_ref_pts, d_dirs, total, n_rays, (int)rpw,            \
                                                                (float)params->angle_max, cc.as<int32_t>(), cl.as<int32_t>(),      \
                                                                reinterpret_cast<unsigned*>(cc.as<int32_t>() + total), CNT)
    if (g_counters_host_copy) {
        if (do_cull) BT_TRAVERSE(true, true, true, g_counters_host_copy);
        else if (packed) BT_TRAVERSE(true, true, false, g_counters_host_copy);
        else BT_TRAVERSE(true, false, false, g_counters_host_copy);
    } else {
        if (do_cull) BT_TRAVERSE(false, true, true, nullptr);
        else if (packed) BT_TRAVERSE(false, true, false, nullptr);
        else BT_TRAVERSE(false, false, false, nullptr);
    }
#undef BT_TRAVERSE
    BT_LAUNCH_CHECK();
    // stage 2: exact float64 narrow phase, ordering, de-duplication, filters, hit records
    if (cap <= 8)
        cast_kernel<8><<<grid, 128, 0, st>>>(mv, d_ref_pts, d_dirs, total, n_rays, *params, cap, cc.as<int32_t>(), cl.as<int32_t>(), d_hit_count, d_hits, d_hit_t, d_valid_mask, d_overflow);
    else if (cap <= 16)
        cast_kernel<16><<<grid, 128, 0, st>>>(mv, d_ref_pts, d_dirs, total, n_rays, *params, cap, cc.as<int32_t>(), cl.as<int32_t>(), d_hit_count, d_hits, d_hit_t, d_valid_mask, d_overflow);
    else
        cast_kernel<32><<<grid, 128, 0, st>>>(mv, d_ref_pts, d_dirs, total, n_rays, *params, cap, cc.as<int32_t>(), cl.as<int32_t>(), d_hit_count, d_hits, d_hit_t, d_valid_mask, d_overflow);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

extern "C" int bt_ags_select(const int32_t* d_hit_count, const bt_hit* d_hits, const uint32_t* d_valid_mask, int64_t n_ref, int32_t n_rays,
                             int32_t cap, int32_t max_targets, uint64_t seed, int32_t* d_pair_count,
                             int64_t* d_pair_offset, int32_t* d_pair_ref, double* d_pair_q, int64_t cap_pairs,
                             void* stream) {
    BT_REQUIRE(d_hits, "null pointer");
    return bt::ags_select_impl(d_hit_count, d_hits, nullptr, nullptr, nullptr, d_valid_mask, n_ref, n_rays, cap, max_targets, seed, d_pair_count,
                               d_pair_offset, d_pair_ref, d_pair_q, cap_pairs, stream);
}

int bt::ags_select_impl(const int32_t* d_hit_count, const bt_hit* d_hits, const double* d_hit_t, const double* d_ref_pts, const double* d_dirs,
                        const uint32_t* d_valid_mask, int64_t n_ref, int32_t n_rays, int32_t cap, int32_t max_targets, uint64_t seed,
                        int32_t* d_pair_count, int64_t* d_pair_offset, int32_t* d_pair_ref, double* d_pair_q, int64_t cap_pairs,
                        void* stream) {
    BT_REQUIRE(d_hit_count && d_pair_count && d_pair_offset && d_pair_ref && d_pair_q, "null pointer");
    BT_REQUIRE(d_hits || (d_hit_t && d_ref_pts && d_dirs && d_valid_mask), "null pointer");
    BT_REQUIRE(n_ref >= 0 && n_rays > 0 && cap >= 1 && max_targets >= 1, "bad counts");
    HitSource hs{d_hit_t, d_ref_pts, d_dirs, n_ref * (int64_t)n_rays};
    cudaStream_t st = (cudaStream_t)stream;
    if (n_ref == 0) {
        zero_first_kernel<<<1, 1, 0, st>>>(d_pair_offset);
        BT_LAUNCH_CHECK();
        return BT_OK;
    }
    const unsigned grid = (unsigned)ceil_div(n_ref * 32, 128);
    select_count_kernel<<<grid, 128, 0, st>>>(d_hit_count, d_hits, d_valid_mask, n_ref, n_rays, cap, max_targets, d_pair_count);
    BT_LAUNCH_CHECK();
    Scratch tmp;
    if (n_ref <= BT_SMALL_SCAN_MAX) {
        small_scan_kernel<int32_t><<<1, 1024, 0, st>>>(d_pair_count, d_pair_offset, (int)n_ref, 1);      // pair_offset[0 .. n_ref]
        BT_LAUNCH_CHECK();
    } else {
        zero_first_kernel<<<1, 1, 0, st>>>(d_pair_offset);
        BT_LAUNCH_CHECK();
        size_t tmp_bytes = 0;
        BT_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, d_pair_count, d_pair_offset + 1, (int)n_ref, st));
        BT_CUDA(tmp.alloc(tmp_bytes, st));
        BT_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, tmp_bytes, d_pair_count, d_pair_offset + 1, (int)n_ref, st));
    }
    select_fill_kernel<<<grid, 128, 0, st>>>(d_hit_count, d_hits, hs, d_valid_mask, n_ref, n_rays, cap, max_targets, seed,
                                             d_pair_offset, cap_pairs, d_pair_ref, d_pair_q);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

extern "C" int bt_expand_orientations(const double* d_ref_pts, const int64_t* d_pair_offset,
                                      const int32_t* d_pair_ref, const double* d_pair_q, const double* d_frame_vecs,
                                      int64_t n_ref, int64_t cap_pairs, int32_t n_o, int32_t halfspace,
                                      const double* h_cos, const double* h_sin, float* d_gs, double* d_contacts,
                                      void* stream) {
    return bt::expand_impl(d_ref_pts, d_pair_offset, d_pair_ref, d_pair_q, d_frame_vecs, n_ref, cap_pairs, n_o, halfspace, h_cos, h_sin,
                           d_gs, d_contacts, nullptr, stream);
}

int bt::expand_impl(const double* d_ref_pts, const int64_t* d_pair_offset, const int32_t* d_pair_ref, const double* d_pair_q,
                    const double* d_frame_vecs, int64_t n_ref, int64_t cap_pairs, int32_t n_o, int32_t halfspace, const double* h_cos,
                    const double* h_sin, float* d_gs, double* d_contacts, int64_t* d_n_rows, void* stream) {
    BT_REQUIRE(d_ref_pts && d_pair_offset && d_pair_ref && d_pair_q && d_gs && h_cos && h_sin, "null pointer");
    BT_REQUIRE(halfspace || d_frame_vecs, "frame vectors required unless halfspace");
    BT_REQUIRE(n_o >= 1 && n_o <= 64, "n_orientations must be in [1, 64]");
    if (cap_pairs <= 0 || n_ref <= 0) {
        if (d_n_rows) BT_CUDA(cudaMemsetAsync(d_n_rows, 0, sizeof(int64_t), (cudaStream_t)stream));
        return BT_OK;
    }
    OrientTable tab;
    for (int i = 0; i < 64; ++i) { tab.c[i] = i < n_o ? h_cos[i] : 0.0; tab.s[i] = i < n_o ? h_sin[i] : 0.0; }
    expand_kernel<<<(unsigned)ceil_div(cap_pairs * n_o, 128), 128, 0, (cudaStream_t)stream>>>(
        d_ref_pts, d_pair_offset, d_pair_ref, d_pair_q, d_frame_vecs, n_ref, n_o, halfspace, tab, d_gs, d_contacts, d_n_rows);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

// S1 + S2 + frame vectors in one launch (any subset; see generate_kernel).  d_seed_add: nullable device-resident seed offset.
int bt::generate_impl(const bt_mesh* mesh, int64_t n_ref, int32_t n_rays, double angle, uint64_t seed, const uint64_t* d_seed_add,
                      int64_t surface_total, int do_sample, int do_dirs, int do_frame, double* d_pts, double* d_nrm, double* d_frame,
                      double* d_dirs, void* stream) {
    BT_REQUIRE(mesh && d_nrm, "null pointer");
    BT_REQUIRE(n_ref >= 0 && n_rays > 0, "bad counts");
    BT_REQUIRE((!do_sample || d_pts) && (!do_dirs || d_dirs) && (!do_frame || d_frame), "missing output buffer");
    if (n_ref == 0 || !(do_sample || do_dirs || do_frame)) return BT_OK;
    generate_kernel<<<(unsigned)ceil_div(n_ref, GEN_REFS), 256, 0, (cudaStream_t)stream>>>(
        mesh->m.area_cdf, mesh->m.faces, mesh->m.verts, mesh->m.nF, n_ref, n_rays, angle, seed, d_seed_add, surface_total, do_sample, do_dirs, do_frame,
        d_pts, d_nrm, d_frame, d_dirs);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

// R7 in one launch + one memset (lean or record hits; see select_fused_kernel)
int bt::ags_select_fused_impl(const int32_t* d_hit_count, const bt_hit* d_hits, const double* d_hit_t, const double* d_ref_pts,
                              const double* d_dirs, const uint32_t* d_valid_mask, int64_t n_ref, int32_t n_rays, int32_t cap,
                              int32_t max_targets, uint64_t seed, const uint64_t* d_seed_add, int32_t* d_pair_count,
                              int64_t* d_pair_offset, int32_t* d_pair_ref, double* d_pair_q, int64_t cap_pairs, void* stream) {
    BT_REQUIRE(d_hit_count && d_pair_count && d_pair_offset && d_pair_ref && d_pair_q, "null pointer");
    BT_REQUIRE(d_hits || (d_hit_t && d_ref_pts && d_dirs && d_valid_mask), "null pointer");
    BT_REQUIRE(n_ref >= 0 && n_rays > 0 && cap >= 1 && max_targets >= 1, "bad counts");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_ref == 0) {
        BT_CUDA(cudaMemsetAsync(d_pair_offset, 0, sizeof(int64_t), st));
        return BT_OK;
    }
    HitSource hs{d_hit_t, d_ref_pts, d_dirs, n_ref * (int64_t)n_rays};
    const int64_t n_blocks = ceil_div(n_ref, SEL_REFS);
    Scratch state;                                              // [n_blocks] status words + the ticket counter
    BT_CUDA(state.alloc(sizeof(unsigned long long) * (n_blocks + 1), st));
    BT_CUDA(cudaMemsetAsync(state.p, 0, sizeof(unsigned long long) * (n_blocks + 1), st));
    select_fused_kernel<<<(unsigned)n_blocks, 32 * SEL_REFS, 0, st>>>(
        d_hit_count, d_hits, hs, d_valid_mask, n_ref, n_rays, cap, max_targets, seed, d_seed_add, state.as<unsigned long long>(),
        reinterpret_cast<unsigned*>(state.as<unsigned long long>() + n_blocks), (unsigned)n_blocks, d_pair_count, d_pair_offset, cap_pairs,
        d_pair_ref, d_pair_q);
    BT_LAUNCH_CHECK();
    return BT_OK;
}


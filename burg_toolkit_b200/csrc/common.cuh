// common.cuh -- shared host/device helpers for libburgb200 (sm_100a only).
//
// The whole library is compiled with -fmad=false: float64 paths reproduce numpy's separate multiply /
// add roundings, and every fused multiply-add that IS wanted (float32 traversal math, the sgemm-style
// FMA chains of metrics.py) is written explicitly with fmaf()/fma().
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/burg_b200.h"

namespace bt {

constexpr int BT_WIDE_EMPTY = 0x7fffffff;       // child code of an unused slot of a 4-wide node
constexpr int BT_WIDE_EMPTY_PACKED = 0x7fffff;  // the same in the packed (24-bit) code format, see emit_wide_nodes_kernel

// ---- error plumbing ---------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define BT_CUDA(call)                                                          \
    do {                                                                       \
        cudaError_t _e = (call);                                               \
        if (_e != cudaSuccess) return bt::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

#define BT_REQUIRE(cond, msg)                                                  \
    do {                                                                       \
        if (!(cond)) { bt::set_error("%s: %s", __func__, msg); return BT_E_INVALID; } \
    } while (0)

#define BT_LAUNCH_CHECK() BT_CUDA(cudaGetLastError())

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// keep freed scratch memory in the device's default pool across synchronisation points (the default release
// threshold of 0 hands it back to the driver at every sync, which turns the next cudaMallocAsync into a ~ms call)
inline void retain_pool_memory() {
    static bool done[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || done[dev]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    done[dev] = true;
}

// Caller-provided scratch arena of the running bt_ags_pipeline call (bt_pipeline_buffers.scratch): a bump allocator that
// lives for one call.  With it the call makes no stream-ordered allocations -- which matters when the call is captured
// into a CUDA graph: allocation nodes give every graph its own memory, mapped at launch (hundreds of graphs -- one per
// target mesh, stream and gather slot -- then cost milliseconds per launch and gigabytes).
struct Arena { char* base = nullptr; size_t size = 0, used = 0; };
extern thread_local Arena* g_arena;

// scratch buffer: from the arena if there is one with room, else stream-ordered (cudaMallocAsync / cudaFreeAsync)
struct Scratch {
    void* p = nullptr;
    cudaStream_t s = nullptr;
    bool owned = false;
    cudaError_t alloc(size_t bytes, cudaStream_t stream) {
        s = stream;
        if (g_arena) {
            const size_t off = (g_arena->used + 255) & ~(size_t)255;
            if (off + bytes <= g_arena->size) { p = g_arena->base + off; g_arena->used = off + bytes; return cudaSuccess; }
        }
        retain_pool_memory();
        owned = true;
        return cudaMallocAsync(&p, bytes ? bytes : 1, stream);
    }
    ~Scratch() { if (p && owned) cudaFreeAsync(p, s); }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// ---- 256-bit read-only global loads (LDG.E.256, sm_100+; PTX ld.global.nc.v8.b32 / .v4.f64).  A lane that fetches a
// 64-byte node on its own address costs one L1 sector access per instruction whatever the width, so two 32-byte loads do
// the work of four 16-byte ones -- and the traversal kernels are bound by exactly those accesses.
// The address must be 32-byte aligned.
#ifndef BT_LDG256
#define BT_LDG256 1
#endif
#ifdef __CUDACC__
__device__ __forceinline__ void ldg256(const void* p, uint4& a, uint4& b) {
#if BT_LDG256
    asm("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
#else
    a = __ldg(reinterpret_cast<const uint4*>(p)); b = __ldg(reinterpret_cast<const uint4*>(p) + 1);
#endif
}
__device__ __forceinline__ void ldg256(const void* p, float4& a, float4& b) {
#if BT_LDG256
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
#else
    a = __ldg(reinterpret_cast<const float4*>(p)); b = __ldg(reinterpret_cast<const float4*>(p) + 1);
#endif
}
__device__ __forceinline__ void ldg256(const void* p, double2& a, double2& b) {
#if BT_LDG256
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a.x), "=d"(a.y), "=d"(b.x), "=d"(b.y) : "l"(p));
#else
    a = __ldg(reinterpret_cast<const double2*>(p)); b = __ldg(reinterpret_cast<const double2*>(p) + 1);
#endif
}
#endif

// ---- packed float32 pairs (Blackwell FFMA2 / FMUL2: fma.rn.f32x2, mul.rn.f32x2) -----------------------------------------
// One instruction, two IEEE round-to-nearest results per lane -- each half is bit-identical to the scalar fmaf / product
// of the same operands, so a kernel may pair independent chains freely.  The kernels that are bound by FP32 *issue* (the
// coverage hot loops: 3-4 multiply-adds + one compare per pair) halve their multiply-add instructions this way.  A pair
// lives in a 64-bit register (an aligned even/odd register pair): lo = bits 0..31, hi = bits 32..63.
// BT_F32X2=0 builds the same kernels with two scalar instructions per pair (A/B runs, tools/gpu_runs).
#ifndef BT_F32X2
#define BT_F32X2 1
#endif
#ifdef __CUDACC__
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float lo2(f32x2 v) { return __uint_as_float((unsigned)(v & 0xffffffffull)); }
__device__ __forceinline__ float hi2(f32x2 v) { return __uint_as_float((unsigned)(v >> 32)); }
__device__ __forceinline__ f32x2 dup2(float x) { return pack2(x, x); }        // ptxas folds it into the .F32 broadcast operand form
// the same, pinned next to its use: a duplicate that the front end hoists out of a loop lives on as a register pair (two
// MOVs and two registers per operand); `volatile` keeps the pack where it is written, where ptxas folds it away
__device__ __forceinline__ f32x2 dup2_here(float x) {
    f32x2 r;
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
#if BT_F32X2
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
#else
    return pack2(__fmaf_rn(lo2(a), lo2(b), lo2(c)), __fmaf_rn(hi2(a), hi2(b), hi2(c)));
#endif
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
#if BT_F32X2
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
#else
    return pack2(__fmul_rn(lo2(a), lo2(b)), __fmul_rn(hi2(a), hi2(b)));
#endif
}
#endif

// ---- exclusive scan of a short array by ONE 1024-thread block: out[i] = in[0] + ... + in[i-1] (and out[n] = total if
// write_total).  The scans of the pipeline are short (pair counts per reference point, kept rows per compaction block:
// 10^2 .. 10^5 entries), where a device-wide scan is two launches of mostly launch latency; beyond BT_SMALL_SCAN_MAX the
// callers use cub::DeviceScan.
constexpr int BT_SMALL_SCAN_MAX = 1 << 18;
#ifdef __CUDACC__
template <typename TIn>
__global__ void __launch_bounds__(1024) small_scan_kernel(const TIn* __restrict__ in, int64_t* __restrict__ out, int n, int write_total) {
    __shared__ int64_t warp_base[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (n + 1023) / 1024;
    const int b = min(n, tid * per), e = min(n, b + per);
    int64_t s = 0;
    for (int i = b; i < e; ++i) s += (int64_t)in[i];
    int64_t incl = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int64_t v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) warp_base[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int64_t w = warp_base[lane];
        int64_t wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int64_t v = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += v;
        }
        warp_base[lane] = wi - w;                       // exclusive prefix over the warps
    }
    __syncthreads();
    int64_t run = warp_base[warp] + incl - s;
    for (int i = b; i < e; ++i) { out[i] = run; run += (int64_t)in[i]; }
    if (write_total && tid == 1023) out[n] = run;       // the last thread's chunk ends at n, so its running sum is the total
}
#endif

// ---- single-pass chained scan across the blocks of a grid (decoupled look-back), used by the one-launch target selection
// and the one-launch compaction.  Status word per block: bits 63..62 = 1 aggregate / 2 inclusive prefix, rest = value.
// Block ids must come from a ticket counter (atomicAdd), so that a block only waits for blocks that are already running.
#ifdef __CUDACC__
constexpr unsigned long long LB_AGG = 1ull << 62, LB_PREFIX = 2ull << 62, LB_VALUE = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// warp-wide look-back of block `bid` with aggregate `agg`: returns the exclusive prefix (sum of the aggregates of blocks
// 0 .. bid-1) to every lane and publishes the inclusive prefix of `bid`
__device__ __forceinline__ unsigned long long lookback_exclusive(unsigned long long* status, unsigned bid, unsigned long long agg, int lane) {
    if (lane == 0) st_status(status + bid, (bid == 0 ? LB_PREFIX : LB_AGG) | agg);
    unsigned long long run = 0;
    int pos = (int)bid - 1;
    while (pos >= 0) {
        const int idx = pos - lane;
        unsigned long long st = LB_PREFIX;                     // lanes before block 0: an empty inclusive prefix
        if (idx >= 0) { do { st = ld_status(status + idx); } while ((st >> 62) == 0ull); }
        const unsigned pm = __ballot_sync(0xffffffffu, (st >> 62) == 2ull);
        const int first = pm ? __ffs(pm) - 1 : 31;             // nearest predecessor that already knows its inclusive prefix
        unsigned long long v = lane <= first ? (st & LB_VALUE) : 0ull;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
        run += v;
        if (pm) break;
        pos -= 32;
    }
    if (lane == 0 && bid != 0) st_status(status + bid, LB_PREFIX | (run + agg));
    return run;
}
#endif

// ---- instrumentation (bt_debug_set_counters): when a counter array is installed the traversal kernels are launched in
// their STATS instantiation and accumulate work counts there; slots:
//   0 cast: LBVH nodes visited      1 cast: float32 leaf pre-tests     2 cast: exact float64 ray/triangle tests
//   3 collide: LBVH nodes visited   4 collide: leaves classified       5 collide: leaves sent to the exact test
extern unsigned long long* g_counters_host_copy;      // device pointer as seen by the host (NULL = off)

// ---- device mesh -------------------------------------------------------------------------------------
// LBVH node: 64 bytes = 4 x float4, holds BOTH children's boxes so one fetch decides both descents.
//   q0 = (l.min.x, l.min.y, l.min.z, l.max.x)
//   q1 = (l.max.y, l.max.z, r.min.x, r.min.y)
//   q2 = (r.min.z, r.max.x, r.max.y, r.max.z)
//   q3 = (left, right, -, -) as int bits; child >= 0: internal node index, child < 0: leaf ~child
// Quantised LBVH node for the ray caster: 32 bytes = 2 x uint4.  Box coordinates are uint16 on a global grid over the
// padded scene bounds, rounded outward by one extra cell (resolution = extent / 65534, ~1 um for a 7 cm object), which
// halves the bytes -- and the L1 wavefronts -- per traversal step:
//   A = (l.lo.x | l.lo.y << 16,  l.lo.z | l.hi.x << 16,  l.hi.y | l.hi.z << 16,  r.lo.x | r.lo.y << 16)
//   B = (r.lo.z | r.hi.x << 16,  r.hi.y | r.hi.z << 16,  left,  right)
// 4-wide quantised node for the ray caster: 64 bytes = 4 x uint4, see emit_wide_nodes_kernel (mesh.cu).
// Triangle record for the exact ray test: 128 bytes = 16 doubles (one cache line):
//   v0[3] n[3] e0[3] e1[3] d00 d01 d11 inv_den     (e0 = v1-v0, e1 = v2-v0, trimesh 'cramer' constants)
// Triangle vertices for the exact triangle/triangle test: 9 doubles (v0 v1 v2).
// Float32 triangle for the conservative ray pre-test: 3 x float4 (v0, e0, e1), 48 bytes.
struct MeshDev {
    int device = 0;
    int64_t nV = 0, nF = 0;
    // caller order
    double*  verts = nullptr;      // [nV,3]
    double*  vnormals = nullptr;   // [nV,3]
    int32_t* faces = nullptr;      // [nF,3]
    double*  area_cdf = nullptr;   // [nF] inclusive normalised cumulative area (caller order)
    // LBVH (leaf order = Morton order)
    float4*  nodes = nullptr;      // [max(nF-1,1)*4]   float32 boxes (collision kernel)
    uint4*   qnodes = nullptr;     // [max(nF-1,1)*2]   the same nodes, boxes as uint16 on a global grid (ray caster fallback): 32 B
    uint4*   wnodes = nullptr;     // [max(nF-1,1)*4]   4-wide nodes (every other level skipped), uint16 boxes (ray caster): 64 B
    double   grid_min[3] = {0, 0, 0}, grid_scale[3] = {1, 1, 1};   // grid coordinate = (x - grid_min) * grid_scale
    double*  tri_rec = nullptr;    // [nF,16]  leaf order
    double*  tri_verts = nullptr;  // [nF,9]   leaf order
    double*  tri_vvn = nullptr;    // [nF,20]  leaf order: v0 v1 v2 pad | n0 n1 n2 pad -- vertices + vertex normals of the leaf's triangle as
                                   //          ten 16-byte words (normal interpolation without the faces -> vertices -> normals chain)
    float4*  tri_f32 = nullptr;    // [nF,3]   leaf order: (v0 | 0) (e0 | 0) (e1 | 0) for the float32 pre-test
    int32_t* leaf_tri = nullptr;   // [nF]     leaf -> caller's triangle id
    uint32_t* leaf_cone = nullptr; // [nF]     leaf order: packed normal cone of the leaf, widened over its touching neighbours (mesh.cu: pack_cone)
    int32_t  wide_packed = 0;      // wnodes use 24-bit child codes + the node's own cone in their top bytes (nF < 2^23)
    int32_t  root = 0;             // root child code (internal 0, or ~0 if single triangle)
    int32_t  bvh_kind = 0;         // topology under the 4-wide nodes (ray caster): 0 = Karras hierarchy of the Morton codes, 1 = PLOC
    int32_t  walk_kind = 0;        // topology of the binary nodes (collision walks): the same coding
    int32_t  ploc_iterations = 0;
    bt_mesh_info info{};
};

// internal forms of bt_ags_cast / bt_ags_select with the lean hit representation of bt_ags_pipeline (ags.cu)
int ags_cast_impl(const bt_mesh* mesh, const double* d_ref_pts, const double* d_dirs, int64_t n_ref, int32_t n_rays,
                  const bt_ags_params* params, int32_t* d_hit_count, bt_hit* d_hits, double* d_hit_t, uint32_t* d_valid_mask,
                  int32_t* d_overflow, bool cull, void* stream);
int ags_select_impl(const int32_t* d_hit_count, const bt_hit* d_hits, const double* d_hit_t, const double* d_ref_pts, const double* d_dirs,
                    const uint32_t* d_valid_mask, int64_t n_ref, int32_t n_rays, int32_t cap, int32_t max_targets, uint64_t seed,
                    int32_t* d_pair_count, int64_t* d_pair_offset, int32_t* d_pair_ref, double* d_pair_q, int64_t cap_pairs, void* stream);

int ags_select_fused_impl(const int32_t* d_hit_count, const bt_hit* d_hits, const double* d_hit_t, const double* d_ref_pts,
                          const double* d_dirs, const uint32_t* d_valid_mask, int64_t n_ref, int32_t n_rays, int32_t cap,
                          int32_t max_targets, uint64_t seed, const uint64_t* d_seed_add, int32_t* d_pair_count, int64_t* d_pair_offset,
                          int32_t* d_pair_ref, double* d_pair_q, int64_t cap_pairs, void* stream);
int generate_impl(const bt_mesh* mesh, int64_t n_ref, int32_t n_rays, double angle, uint64_t seed, const uint64_t* d_seed_add,
                  int64_t surface_total, int do_sample, int do_dirs, int do_frame, double* d_pts, double* d_nrm, double* d_frame,
                  double* d_dirs, void* stream);
int expand_impl(const double* d_ref_pts, const int64_t* d_pair_offset, const int32_t* d_pair_ref, const double* d_pair_q,
                const double* d_frame_vecs, int64_t n_ref, int64_t cap_pairs, int32_t n_o, int32_t halfspace, const double* h_cos,
                const double* h_sin, float* d_gs, double* d_contacts, int64_t* d_n_rows, void* stream);
// single-pass stream compaction (collide.cu); bt_compact is its C-ABI face
int compact_impl(const uint8_t* d_mask, uint8_t keep_value, const float* d_gs, const double* d_contacts, int64_t n, const int64_t* d_n,
                 float* d_out_gs, double* d_out_contacts, int64_t* d_n_out, const int64_t* d_base, void* stream);

}  // namespace bt

struct bt_mesh {
    bt::MeshDev m;
};

// ---- small device math (float64, numpy operation order; no contraction thanks to -fmad=false) --------
namespace bt {

struct d3 { double x, y, z; };

__host__ __device__ inline d3 mk3(double x, double y, double z) { d3 r; r.x = x; r.y = y; r.z = z; return r; }
__host__ __device__ inline d3 ld3(const double* p) { return mk3(p[0], p[1], p[2]); }
__host__ __device__ inline void st3(double* p, d3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
__host__ __device__ inline d3 sub3(d3 a, d3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__host__ __device__ inline d3 add3(d3 a, d3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__host__ __device__ inline d3 scl3(d3 a, double s) { return mk3(a.x * s, a.y * s, a.z * s); }
__host__ __device__ inline d3 div3(d3 a, double s) { return mk3(a.x / s, a.y / s, a.z / s); }
// np.sum(a*b, -1) / np.dot(a*b, ones) : (x + y) + z with separately rounded products
__host__ __device__ inline double dot3(d3 a, d3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
// np.cross : c0 = a1*b2 - a2*b1, ...
__host__ __device__ inline d3 cross3(d3 a, d3 b) {
    return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// np.linalg.norm(v, axis=-1) : sqrt((x*x + y*y) + z*z)
__host__ __device__ inline double norm3(d3 a) { return sqrt((a.x * a.x + a.y * a.y) + a.z * a.z); }

// Philox4x32-10 counter-based RNG (Salmon et al. 2011) -- device-side candidate generation
__host__ __device__ inline void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3,
                                             uint32_t k0, uint32_t k1) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}
__host__ __device__ inline void philox4x32(uint64_t seed, uint64_t ctr_lo, uint64_t ctr_hi, uint32_t out[4]) {
    uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        philox_round(c0, c1, c2, c3, k0, k1);
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// 53-bit uniform in [0,1) from two 32-bit words
__host__ __device__ inline double u01(uint32_t hi, uint32_t lo) {
    return (double)((((uint64_t)hi << 32) | lo) >> 11) * (1.0 / 9007199254740992.0);
}

}  // namespace bt

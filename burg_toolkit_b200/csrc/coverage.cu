// coverage.cu -- grasp-set coverage / precision and the pairwise SE(3) distances behind it (M1-M5).
//
// Reference being replaced: burg_toolkit/metrics.py -- euclidean_distances :9-27, angular_distances :30-60,
// combined_distances :63-77, coverage_brute_force :141-158, coverage :161-229 (cKDTree.query_ball_tree + a Python
// loop).  The predicate is  covered_i = any_j ( 1000*|t_i - t_j| + deg(R_i, R_j) <= eps )  in float32 with the
// reference's exact operation order:
//     d    = sqrt((dx*dx + dy*dy) + dz*dz)                      separate multiplies / adds
//     g_r  = fma(a_r2, b_r2, fma(a_r1, b_r1, a_r0 * b_r0))      sgemm-style chain, r = 0..2 (rows of R1, R2)
//     c    = ((g0 + g1) + g2 - 1) / 2 ;  k = rint(c * 1e4) ;  deg = LUT[k + 10000]   (np.around(.,4) -> arccos -> rad2deg)
//     comb = 1000*d + deg
// Every pair first goes through a cheap conservative float32 translation reject (3 FFMA + compare); only its rare
// survivors run the exact recipe.  algo 0 tiles all pairs through shared memory; algo 1 bins both sets in a uniform
// grid of cell size ~eps/1000 and only visits the 27 neighbouring cells (result identical: the predicate is an OR).
#include <cub/cub.cuh>
#include <math.h>

#include "common.cuh"

#include <stdlib.h>

namespace bt {

constexpr int COV_THREADS = 256;
constexpr int COV_RPT = 8;                       // reference rows per thread (16: measured 1.7x slower, 128 registers)
constexpr int COV_TILE = 512;                    // queries staged per shared-memory tile (8 KB)

struct CovParams {
    float eps;              // epsilon (float32 compare, metrics.py:157/:219)
    float weight;           // 1000.0f
    float r2_reject;        // conservative squared-distance bound for the fast reject
    double r2_kd;           // (eps/1000)^2 in float64 (mode 1)
    int mode;
    float cx, cy, cz;       // centring offset for the expanded-form reject
    float qdot_min;         // conservative lower bound on |q1 . q2| = cos(angle / 2) for the angular part to fit under eps
};

// exact reference recipe for one pair; rows are the raw float32[14] grasp rows
__device__ __noinline__ bool pair_covered(const float* __restrict__ a, const float* __restrict__ b,
                                             const float* __restrict__ lut, const CovParams& P) {
    const float dx = __fsub_rn(a[0], b[0]), dy = __fsub_rn(a[1], b[1]), dz = __fsub_rn(a[2], b[2]);
    const float d = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    const float wd = __fmul_rn(P.weight, d);
    if (!(wd <= P.eps)) return false;                       // deg >= 0, so the sum cannot come back under eps
    if (P.mode == 1) {                                      // cKDTree.query_ball_tree candidate test, float64
        const double ex = (double)a[0] - (double)b[0], ey = (double)a[1] - (double)b[1], ez = (double)a[2] - (double)b[2];
        if (!(((ex * ex + ey * ey) + ez * ez) <= P.r2_kd)) return false;
    }
    float g[3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
        g[r] = __fmaf_rn(a[3 + 3 * r + 2], b[3 + 3 * r + 2], __fmaf_rn(a[3 + 3 * r + 1], b[3 + 3 * r + 1], __fmul_rn(a[3 + 3 * r], b[3 + 3 * r])));
    const float tr = __fadd_rn(__fadd_rn(g[0], g[1]), g[2]);
    const float c = __fdiv_rn(__fsub_rn(tr, 1.0f), 2.0f);
    const float kf = rintf(__fmul_rn(c, 10000.0f));
    if (!(fabsf(kf) <= 10000.0f)) return false;             // arccos(>1) = NaN in the reference -> never <= eps
    const float deg = __ldg(lut + ((int)kf + 10000));
    return __fadd_rn(wd, deg) <= P.eps;
}

// unit quaternion (x, y, z, w) of a float32 rotation block (row-major 9 floats); approximate (used only for a conservative
// bound): |q1 . q2| = cos(angle(R1 R2^T) / 2)
__device__ __forceinline__ float4 quat_of(const float* __restrict__ R) {
    const float m00 = R[0], m01 = R[1], m02 = R[2], m10 = R[3], m11 = R[4], m12 = R[5], m20 = R[6], m21 = R[7], m22 = R[8];
    const float tr = m00 + m11 + m22;
    float x, y, z, w;
    if (tr > 0.f) { w = tr + 1.f; x = m21 - m12; y = m02 - m20; z = m10 - m01; }
    else if (m00 > m11 && m00 > m22) { x = 1.f + m00 - m11 - m22; w = m21 - m12; y = m01 + m10; z = m02 + m20; }
    else if (m11 > m22) { y = 1.f + m11 - m00 - m22; w = m02 - m20; x = m01 + m10; z = m12 + m21; }
    else { z = 1.f + m22 - m00 - m11; w = m10 - m01; x = m02 + m20; y = m12 + m21; }
    const float inv = rsqrtf(fmaxf(x * x + y * y + z * z + w * w, 1e-30f));
    return make_float4(x * inv, y * inv, z * inv, w * inv);
}

// The orientation bound |q1 . q2| >= cos((eps + 0.1 deg) / 2) of both algorithms presumes rotation matrices: a row whose
// 3 x 3 block is not orthonormal to 1e-4 (or not finite, or a reflection) gets a NaN quaternion, which no comparison of
// the kernels rules out -- its pairs always reach the exact recipe, whatever the block holds.
__device__ __forceinline__ float4 quat_checked(const float* __restrict__ R) {
    float dev = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = a; b < 3; ++b) {
            const float d = fmaf(R[3 * a], R[3 * b], fmaf(R[3 * a + 1], R[3 * b + 1], R[3 * a + 2] * R[3 * b + 2])) - (a == b ? 1.f : 0.f);
            dev = fmaxf(dev, fabsf(d));
        }
    const float det = R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) + R[2] * (R[3] * R[7] - R[4] * R[6]);
    if (!(dev <= 1e-4f) || !(det > 0.f)) return make_float4(__int_as_float(0x7fc00000), 0.f, 0.f, 0.f);
    return quat_of(R);
}

// quaternions of all rows, once per call of the all-pairs kernel
__global__ void row_quat_kernel(const float* __restrict__ rows, int64_t n, float4* __restrict__ quat) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) quat[i] = quat_checked(rows + 14 * i + 3);
}

// ---- algo 0: all pairs, query tiles in shared memory ---------------------------------------------------------
// Fast path per pair: 3 multiply-adds + 1 compare on the expanded form |q|^2 - 2 a.q <= r^2 - |a|^2 (centred coordinates).
// Rows r and r + 1 of a thread live in packed registers, so the multiply-adds of two pairs are three FFMA2 (Blackwell's
// fma.rn.f32x2; the query's components enter as broadcast operands); the compares of a thread's
// COV_RPT rows against one query are OR-ed into two predicates.  The hot loop has NO branch: a thread with a passing pair
// (the volume ratio of the epsilon ball to the data: a few pairs per thousand, i.e. one warp in three per query) appends
// the query's index to its OWN short list in shared memory -- one predicated 16-bit store and one predicated add -- so
// per query a thread issues 1 LDS.128 + 12 FFMA2 + 8 FSETP + 3 = 24 instructions for 8 pairs (round 1: 70 with the rows
// sorted out inside the loop; round 2's first version: 36 + a divergent queue push of ~18 in every third warp).
// After the tile every WARP drains the lists of its 32 threads with all lanes busy (prefix sum of the list lengths by
// shuffles, entry e -> (owner, k) by a 5-step search over the prefix): (A) every entry repeats the conservative test of
// its COV_RPT rows from a shared-memory copy of the row constants and queues the (row, query) pairs that pass, (B) the
// block runs the exact recipe on every queued pair.  Between chunks of 8 queries one vote asks whether any lane's list
// might not hold another chunk; then the warp drains there and then, so the lists never overflow however dense the data.
constexpr int COV_LCAP = 24;                     // list entries per thread (mean ~7 per tile at 100k x 100k in a 0.2 m cube)
constexpr int COV_CHUNK = 8;                     // queries between two looks at the fill of the lists (COV_CHUNK <= COV_LCAP / 2)
constexpr int COV_PCAP = 2048;                   // (row, query) pairs per tile that pass both conservative tests

struct CovTilesSmem {
    float4 tile[COV_TILE];                                  // centred query translations + squared norm
    float4 tile_quat[COV_TILE];                             // query orientations (pass A's orientation bound)
    float4 rows[COV_THREADS * COV_RPT];                     // the block's rows for the drain: (-2 x, -2 y, -2 z, threshold)
    unsigned short lists[COV_LCAP][COV_THREADS];            // per thread: tile indices of the queries with a passing pair
    unsigned int pairs[COV_PCAP];
    unsigned char cov[COV_THREADS * COV_RPT];
    int p_count;
};

__global__ void __launch_bounds__(COV_THREADS, 3)
coverage_tiles_kernel(const float* __restrict__ ref, int64_t n_ref, const float* __restrict__ qry, int64_t n_qry,
                      const float4* __restrict__ ref_quat, const float4* __restrict__ qry_quat,
                      int64_t q_per_split, const float* __restrict__ lut, CovParams P, uint8_t* __restrict__ covered) {
    constexpr int RPT = COV_RPT, TILE = COV_TILE;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char cov_smem_raw[];
    CovTilesSmem& S = *reinterpret_cast<CovTilesSmem*>(cov_smem_raw);
    float4* tile = S.tile;
    unsigned char* s_cov = S.cov;
    const int64_t block_row0 = (int64_t)blockIdx.x * (COV_THREADS * RPT);
    const int lane = threadIdx.x & 31, wbase = threadIdx.x & ~31;
    // per-row constants of the expanded form (centred coordinates); rows r and r + 1 of a thread share packed registers
    // (lo = row r, hi = row r + 1)
    f32x2 ax2[RPT / 2], ay2[RPT / 2], az2[RPT / 2];
    float thr[RPT];
#pragma unroll
    for (int r = 0; r < RPT; r += 2) {
        float c[2][3];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int lr = threadIdx.x + (r + u) * COV_THREADS;
            const int64_t i = block_row0 + lr;
            s_cov[lr] = 0;
            if (i < n_ref) {
                const float x = ref[14 * i] - P.cx, y = ref[14 * i + 1] - P.cy, z = ref[14 * i + 2] - P.cz;
                c[u][0] = -2.0f * x; c[u][1] = -2.0f * y; c[u][2] = -2.0f * z;
                thr[r + u] = P.r2_reject - fmaf(x, x, fmaf(y, y, z * z));
            } else { c[u][0] = c[u][1] = c[u][2] = 0.f; thr[r + u] = -1e30f; }
            S.rows[lr] = make_float4(c[u][0], c[u][1], c[u][2], thr[r + u]);
        }
        ax2[r / 2] = pack2(c[0][0], c[1][0]); ay2[r / 2] = pack2(c[0][1], c[1][1]); az2[r / 2] = pack2(c[0][2], c[1][2]);
    }
    if (threadIdx.x == 0) S.p_count = 0;
    // pass A for one (thread, query) entry: the conservative test of its rows again (same expression, same operands)
    auto sort_out = [&](int t, int j, int64_t q0) {
        const float4 q = tile[j];
#pragma unroll 1
        for (int r = 0; r < RPT; ++r) {
            const int lr = t + r * COV_THREADS;
            const float4 a = S.rows[lr];
            if (fmaf(a.x, q.x, fmaf(a.y, q.y, fmaf(a.z, q.z, q.w))) <= a.w && !s_cov[lr]) {
                // a few pairs per thousand get here; only ~1 in 1000 of those is also close in orientation (the fraction
                // of rotations within eps of a given one), so the orientation bound -- one 16-byte load per side instead
                // of the exact recipe's 28 scattered floats -- empties the pair queue of nearly everything
                const float4 rq = __ldg(ref_quat + block_row0 + lr), qq = S.tile_quat[j];
                if (fabsf(fmaf(rq.x, qq.x, fmaf(rq.y, qq.y, fmaf(rq.z, qq.z, rq.w * qq.w)))) < P.qdot_min) continue;   // (NaN: not ruled out)
                const int pos = atomicAdd(&S.p_count, 1);
                if (pos < COV_PCAP) S.pairs[pos] = ((unsigned)lr << 10) | (unsigned)j;
                else if (pair_covered(ref + 14 * (block_row0 + lr), qry + 14 * (q0 + j), lut, P)) s_cov[lr] = 1;
            }
        }
    };
    const int64_t q_begin = (int64_t)blockIdx.y * q_per_split;
    const int64_t q_end = min(n_qry, q_begin + q_per_split);
    for (int64_t q0 = q_begin; q0 < q_end; q0 += TILE) {
        const int cnt = (int)min((int64_t)TILE, q_end - q0);
        __syncthreads();
        for (int j = threadIdx.x; j < cnt; j += COV_THREADS) {
            const float* q = qry + 14 * (q0 + j);
            const float x = q[0] - P.cx, y = q[1] - P.cy, z = q[2] - P.cz;
            tile[j] = make_float4(x, y, z, fmaf(x, x, fmaf(y, y, z * z)));
            S.tile_quat[j] = __ldg(qry_quat + q0 + j);
        }
        __syncthreads();
        // next free entry of this thread's list, as a shared-memory address: the only state the hot loop carries
        constexpr unsigned STRIDE = COV_THREADS * (unsigned)sizeof(unsigned short);
        const unsigned slot0 = (unsigned)__cvta_generic_to_shared(&S.lists[0][threadIdx.x]);
        const unsigned slot_high = slot0 + (COV_LCAP - COV_CHUNK) * STRIDE;        // above this mark the next chunk might not fit
        unsigned slot = slot0;
        // conservative test of this thread's RPT rows against query j of the tile (packed multiply-adds), append on a hit
        auto test_query = [&](int j) {
            const float4 q = tile[j];
            const f32x2 qx = dup2(q.x), qy = dup2(q.y), qz = dup2(q.z), qw = dup2(q.w);     // broadcast operands of the FFMA2s
            // two OR-chains (even / odd rows) so that the compares of one query do not form a single dependent chain
            bool any0 = false, any1 = false;
#pragma unroll
            for (int r = 0; r < RPT; r += 2) {
                const f32x2 v = fma2(ax2[r / 2], qx, fma2(ay2[r / 2], qy, fma2(az2[r / 2], qz, qw)));
                any0 |= lo2(v) <= thr[r];
                any1 |= hi2(v) <= thr[r + 1];
            }
            if (any0 | any1) {                                     // (predicated: one store, one add)
                asm volatile("st.shared.u16 [%0], %1;" :: "r"(slot), "h"((unsigned short)j) : "memory");
                slot += STRIDE;
            }
        };
        // pass A, per warp and with all lanes busy: prefix sum of the list lengths, entry e -> (owner lane, k); empties the lists
        auto drain_lists = [&]() {
            const int n_mine = (int)((slot - slot0) / STRIDE);
            __syncwarp();
            int incl = n_mine;
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                const int v = __shfl_up_sync(FULL, incl, s);
                if (lane >= s) incl += v;
            }
            const int excl = incl - n_mine, total = __shfl_sync(FULL, incl, 31);
            for (int e0 = 0; e0 < total; e0 += 32) {
                const int e = e0 + lane;
                int owner = 0;                                      // the last lane whose list starts at or before entry e
#pragma unroll
                for (int s = 16; s >= 1; s >>= 1) {
                    const int v = __shfl_sync(FULL, excl, owner + s);
                    if (v <= e) owner += s;
                }
                const int start = __shfl_sync(FULL, excl, owner);
                if (e < total) sort_out(wbase + owner, (int)S.lists[e - start][wbase + owner], q0);
            }
            __syncwarp();                                           // every entry has been read before the lists fill again
            slot = slot0;
        };
        // Chunks of COV_CHUNK queries without a branch; between chunks ONE vote: if any lane's list might not hold another
        // chunk the warp drains its lists there and then (dense data: the epsilon ball holds more than ~0.5 % of the query
        // set), so a list never overflows and no pair is ever dropped -- at the cost of the same drain the tile's end runs anyway.
        int j0 = 0;
        for (; j0 + COV_CHUNK <= cnt; j0 += COV_CHUNK) {
#pragma unroll
            for (int u = 0; u < COV_CHUNK; ++u) test_query(j0 + u);
            if (__any_sync(FULL, slot > slot_high)) drain_lists();
        }
#pragma unroll 1
        for (; j0 < cnt; ++j0) test_query(j0);                      // (< COV_CHUNK queries: still fits)
        drain_lists();
        __syncthreads();
        const int n_p = min(S.p_count, COV_PCAP);
        for (int e = threadIdx.x; e < n_p; e += COV_THREADS) {
            const unsigned code = S.pairs[e];
            const int lr = (int)(code >> 10), j = (int)(code & 1023u);
            if (!s_cov[lr] && pair_covered(ref + 14 * (block_row0 + lr), qry + 14 * (q0 + j), lut, P)) s_cov[lr] = 1;
        }
        __syncthreads();
        if (threadIdx.x == 0) S.p_count = 0;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        const int lr = threadIdx.x + r * COV_THREADS;
        const int64_t i = block_row0 + lr;
        if (i < n_ref && s_cov[lr]) covered[i] = 1;                  // all writers store the same value
    }
}

// ---- algo 1: uniform grid -----------------------------------------------------------------------------------
struct Grid {
    float ox, oy, oz;       // origin
    float inv_h;
    int nx, ny, nz;
};

__device__ __forceinline__ int cell_coord(float x, float o, float inv_h, int n) {
    int c = (int)floorf((x - o) * inv_h);
    return max(0, min(n - 1, c));
}

__device__ __forceinline__ uint32_t float_flip(float f) {           // order-preserving float -> uint
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ inline float float_unflip(uint32_t u) {
    uint32_t v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    float f;
    memcpy(&f, &v, 4);
    return f;
}

__global__ void bounds_kernel(const float* __restrict__ rows, int64_t n, uint32_t* __restrict__ mm /*[6]: min xyz, max xyz*/) {
    float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { const float v = rows[14 * i + k]; lo[k] = fminf(lo[k], v); hi[k] = fmaxf(hi[k], v); }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], s));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], s));
        }
        if ((threadIdx.x & 31) == 0) { atomicMin(mm + k, float_flip(lo[k])); atomicMax(mm + 3 + k, float_flip(hi[k])); }
    }
}

__global__ void cell_key_kernel(const float* __restrict__ rows, int64_t n, Grid g, uint32_t* __restrict__ keys,
                                int32_t* __restrict__ ids) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int cx = cell_coord(rows[14 * i], g.ox, g.inv_h, g.nx);
    const int cy = cell_coord(rows[14 * i + 1], g.oy, g.inv_h, g.ny);
    const int cz = cell_coord(rows[14 * i + 2], g.oz, g.inv_h, g.nz);
    keys[i] = (uint32_t)((cx * g.ny + cy) * g.nz + cz);
    ids[i] = (int32_t)i;
}

// sorted query translations as float4 (x, y, z, original index) + dense cell range table
__global__ void gather_sorted_kernel(const float* __restrict__ rows, const int32_t* __restrict__ sorted_ids,
                                     const uint32_t* __restrict__ sorted_keys, int64_t n, float4* __restrict__ out,
                                     float4* __restrict__ out_quat,
                                     int32_t* __restrict__ cell_start, int32_t* __restrict__ cell_end) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t id = sorted_ids[i];
    out[i] = make_float4(rows[14 * (int64_t)id], rows[14 * (int64_t)id + 1], rows[14 * (int64_t)id + 2], __int_as_float(id));
    if (out_quat) out_quat[i] = quat_checked(rows + 14 * (int64_t)id + 3);
    if (cell_start) {
        const uint32_t k = sorted_keys[i];
        if (i == 0 || sorted_keys[i - 1] != k) cell_start[k] = (int32_t)i;
        if (i == n - 1 || sorted_keys[i + 1] != k) cell_end[k] = (int32_t)i + 1;
    }
}

__device__ __forceinline__ int lower_bound_u32(const uint32_t* __restrict__ a, int n, uint32_t key) {
    int lo = 0, hi = n;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (a[mid] < key) lo = mid + 1; else hi = mid; }
    return lo;
}

// one WARP per reference grasp (walked in cell order): the lanes stride over the sorted queries of the 27 neighbouring
// cells (3 consecutive z cells are one contiguous range -> 9 coalesced ranges), survivors of the translation reject run
// the exact recipe, the warp stops at the first covering query.
__global__ void __launch_bounds__(256)
coverage_grid_kernel(const float* __restrict__ ref, const int32_t* __restrict__ ref_sorted_ids, int64_t n_ref,
                     const float* __restrict__ qry, const float4* __restrict__ q_sorted,
                     const uint32_t* __restrict__ q_keys, int n_qry, const int32_t* __restrict__ cell_start,
                     const int32_t* __restrict__ cell_end, Grid g, const float* __restrict__ lut, CovParams P,
                     uint8_t* __restrict__ covered) {
    const int64_t s = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (s >= n_ref) return;
    const int64_t i = ref_sorted_ids[s];
    const float* a = ref + 14 * i;
    const float x = __ldg(a), y = __ldg(a + 1), z = __ldg(a + 2);
    const int cx = cell_coord(x, g.ox, g.inv_h, g.nx), cy = cell_coord(y, g.oy, g.inv_h, g.ny), cz = cell_coord(z, g.oz, g.inv_h, g.nz);
    const int z0 = max(cz - 1, 0), z1 = min(cz + 1, g.nz - 1);
    bool cov = false;
    for (int ix = max(cx - 1, 0); ix <= min(cx + 1, g.nx - 1) && !cov; ++ix)
        for (int iy = max(cy - 1, 0); iy <= min(cy + 1, g.ny - 1) && !cov; ++iy) {
            const uint32_t k0 = (uint32_t)((ix * g.ny + iy) * g.nz + z0), k1 = (uint32_t)((ix * g.ny + iy) * g.nz + z1);
            int b, e;
            if (cell_start) {
                // consecutive z cells are consecutive keys; empty cells have start == end == 0
                b = -1; e = 0;
                for (uint32_t k = k0; k <= k1; ++k) {
                    const int cs = __ldg(cell_start + k), ce = __ldg(cell_end + k);
                    if (ce > cs) { if (b < 0) b = cs; e = ce; }
                }
                if (b < 0) continue;
            } else {
                b = lower_bound_u32(q_keys, n_qry, k0);
                e = lower_bound_u32(q_keys, n_qry, k1 + 1);
            }
            for (int j0 = b; j0 < e && !cov; j0 += 32) {
                const int j = j0 + lane;
                bool mine = false;
                if (j < e) {
                    const float4 q = __ldg(q_sorted + j);
                    const float dx = x - q.x, dy = y - q.y, dz = z - q.z;
                    const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                    if (d2 <= P.r2_reject) mine = pair_covered(a, qry + 14 * (int64_t)__float_as_int(q.w), lut, P);
                }
                cov = __any_sync(0xffffffffu, mine);
            }
        }
    if (cov && lane == 0) covered[i] = 1;
}

// Cell-tile kernel (dense grid): one block per chunk of CELL_REFS consecutive references IN CELL ORDER.  The chunk's
// references live in a few neighbouring cells; the block walks the bounding cell box of the chunk, grown by one cell,
// stages the sorted queries of every (x, y) column -- a contiguous range because z is the fastest key digit -- in shared
// memory tiles and runs the 3-FFMA + compare reject of the all-pairs kernel on them, with coordinates centred on the
// chunk so that the expanded form stays exact to a few ulps.  Survivors go through the shared queue to the exact recipe.
constexpr int CELL_THREADS = 128;
constexpr int CELL_RPT = 2;
constexpr int CELL_REFS = CELL_THREADS * CELL_RPT;
constexpr int CELL_TILE = 768;                   // queries staged per tile (4 | CELL_TILE <= 1024: queue entries carry 10 bits of it)
constexpr int CELL_LCAP = 16;                    // list entries per thread and tile = groups of 4 queries with a passing orientation bound
constexpr int CELL_PCAP = 1024;                  // (row, query) pairs per tile that pass both conservative tests

// Chunk table of the cell-tile kernel: the references in cell order are cut at every (x, y) column boundary and every
// column into equal parts of at most CELL_REFS rows.  A chunk that straddled columns (as fixed 256-row chunks do) has a
// bounding cell box spanning the whole z -- or y and z -- range: 4x .. 20x the work of a normal chunk, and those few
// blocks decided the kernel's duration (simulated makespan 4.5x the balanced one at 1 M x 1 M).
__global__ void column_bounds_kernel(const uint32_t* __restrict__ sorted_keys, int64_t n, uint32_t nz,
                                     int32_t* __restrict__ col_start, int32_t* __restrict__ col_end) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t c = sorted_keys[i] / nz;
    if (i == 0 || sorted_keys[i - 1] / nz != c) col_start[c] = (int32_t)i;
    if (i == n - 1 || sorted_keys[i + 1] / nz != c) col_end[c] = (int32_t)i + 1;
}

__global__ void emit_chunks_kernel(const int32_t* __restrict__ col_start, const int32_t* __restrict__ col_end, int n_cols,
                                   int chunk_rows, int2* __restrict__ chunks, int* __restrict__ n_chunks) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cols) return;
    const int b = col_start[c], e = col_end[c];
    if (e <= b) return;
    const int k = (e - b + chunk_rows - 1) / chunk_rows;
    const int part = (e - b + k - 1) / k;                   // equal parts, not 256, 256, ..., remainder
    const int pos = atomicAdd(n_chunks, k);                 // chunk order is irrelevant: every row owns its output flag
    for (int j = 0; j < k; ++j) chunks[pos + j] = make_int2(b + j * part, min(e, b + (j + 1) * part));
}

__global__ void __launch_bounds__(CELL_THREADS)
coverage_cells_kernel(const float* __restrict__ ref, const int32_t* __restrict__ ref_sorted_ids, const int2* __restrict__ chunks,
                      const int* __restrict__ n_chunks,
                      const float* __restrict__ qry, const float4* __restrict__ q_sorted, const float4* __restrict__ q_quat_sorted,
                      const int32_t* __restrict__ cell_start, const int32_t* __restrict__ cell_end, Grid g,
                      const float* __restrict__ lut, CovParams P, int split, uint8_t* __restrict__ covered) {
    // split (gridDim.y) = 1: the block walks the whole 3 x 3 column neighbourhood of its chunk; 3: one x-slab of it per
    // block; 9: one column per block.  A reference shard of a multi-GPU run has few chunks (1 M rows / 8 ranks = 488),
    // each ~0.3 ms of work: without the split the 148 SMs get 1.6 waves of long blocks and the call does not scale.
    __shared__ float4 tile[CELL_TILE];
    // query orientations as unit quaternions (+ zero padding for the 4-wide loop), two queries per pair of float4:
    // (x0, x1, y0, y1), (z0, z1, w0, w1) -- the packed operands of the hot loop's FFMA2 / FMUL2
    __shared__ __align__(16) float tile_quat[4 * (CELL_TILE + 4)];
    __shared__ int tile_id[CELL_TILE];
    __shared__ unsigned short lists[CELL_LCAP][CELL_THREADS];   // per thread: the query groups with a passing orientation bound
    __shared__ unsigned int pairs[CELL_PCAP];
    __shared__ int p_count;
    __shared__ float4 s_rquat[CELL_REFS];                   // the chunk's rows for the drain: unit quaternion,
    __shared__ float4 s_rpos[CELL_REFS];                    // (-2 x, -2 y, -2 z, threshold) of the centred translation test
    __shared__ unsigned char s_cov[CELL_REFS];
    __shared__ int s_box[6];                               // min x y z, max x y z (cell coordinates of the chunk)
    if ((int)blockIdx.x >= *n_chunks) return;               // the grid is the upper bound of the chunk count
    const int2 chunk = chunks[blockIdx.x];
    const int64_t s0 = chunk.x;
    const int n_here = chunk.y - chunk.x;
    if (threadIdx.x < 3) { s_box[threadIdx.x] = 0x7fffffff; s_box[3 + threadIdx.x] = -1; }
    if (threadIdx.x == 0) p_count = 0;
    __syncthreads();
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, wbase = threadIdx.x & ~31;
    int64_t row[CELL_RPT];
    float px[CELL_RPT], py[CELL_RPT], pz[CELL_RPT];
    float4 quat[CELL_RPT];
#pragma unroll
    for (int r = 0; r < CELL_RPT; ++r) {
        const int lr = threadIdx.x + r * CELL_THREADS;
        s_cov[lr] = 0;
        row[r] = -1;
        quat[r] = make_float4(0.f, 0.f, 0.f, 1.f);
        if (lr < n_here) {
            row[r] = ref_sorted_ids[s0 + lr];
            px[r] = ref[14 * row[r]]; py[r] = ref[14 * row[r] + 1]; pz[r] = ref[14 * row[r] + 2];
            quat[r] = quat_checked(ref + 14 * row[r] + 3);
            const int cx = cell_coord(px[r], g.ox, g.inv_h, g.nx), cy = cell_coord(py[r], g.oy, g.inv_h, g.ny), cz = cell_coord(pz[r], g.oz, g.inv_h, g.nz);
            atomicMin(&s_box[0], cx); atomicMin(&s_box[1], cy); atomicMin(&s_box[2], cz);
            atomicMax(&s_box[3], cx); atomicMax(&s_box[4], cy); atomicMax(&s_box[5], cz);
        } else { px[r] = py[r] = pz[r] = 0.f; }
    }
    __syncthreads();
    const int x0 = max(s_box[0] - 1, 0), x1 = min(s_box[3] + 1, g.nx - 1);
    const int y0 = max(s_box[1] - 1, 0), y1 = min(s_box[4] + 1, g.ny - 1);
    const int z0 = max(s_box[2] - 1, 0), z1 = min(s_box[5] + 1, g.nz - 1);
    // centre of the chunk's cell box in world coordinates: all staged points lie within a few cells of it
    const float h = 1.0f / g.inv_h;
    const float ccx = g.ox + 0.5f * (x0 + x1 + 1) * h, ccy = g.oy + 0.5f * (y0 + y1 + 1) * h, ccz = g.oz + 0.5f * (z0 + z1 + 1) * h;
    const float ext = 0.5f * h * (float)(max(x1 - x0, max(y1 - y0, z1 - z0)) + 1);
    // rounding of the 4-term FMA chain and the two squared norms, for centred coordinates bounded by ext*sqrt(3)
    const float r2 = P.r2_reject + 64.0f * 5.97e-8f * (3.0f * ext * ext);
    float ax[CELL_RPT], ay[CELL_RPT], az[CELL_RPT], thr[CELL_RPT], qmin[CELL_RPT];
#pragma unroll
    for (int r = 0; r < CELL_RPT; ++r) {
        const float x = px[r] - ccx, y = py[r] - ccy, z = pz[r] - ccz;
        ax[r] = -2.0f * x; ay[r] = -2.0f * y; az[r] = -2.0f * z;
        thr[r] = r2 - fmaf(x, x, fmaf(y, y, z * z));
        qmin[r] = row[r] >= 0 ? P.qdot_min : 2.0f;
        s_rquat[threadIdx.x + r * CELL_THREADS] = quat[r];
        s_rpos[threadIdx.x + r * CELL_THREADS] = make_float4(ax[r], ay[r], az[r], row[r] >= 0 ? thr[r] : -1e30f);
    }
    f32x2 rqx[CELL_RPT], rqy[CELL_RPT], rqz[CELL_RPT], rqw[CELL_RPT];      // a row's quaternion against two queries at once
#pragma unroll
    for (int r = 0; r < CELL_RPT; ++r) {
        rqx[r] = dup2(quat[r].x); rqy[r] = dup2(quat[r].y); rqz[r] = dup2(quat[r].z); rqw[r] = dup2(quat[r].w);
    }
    // pass A of the drain for the 4 x CELL_RPT pairs of (thread t, queries j0 .. j0 + 3 of the staged tile): orientation
    // bound and translation reject from shared memory, by whichever thread drains the entry; survivors -> pair queue
    auto drain_group = [&](int t, int j0, int cnt) {
#pragma unroll 1
        for (int r = 0; r < CELL_RPT; ++r) {
            const int lr = t + r * CELL_THREADS;
            if (lr >= n_here || s_cov[lr]) continue;
            const float4 rq = s_rquat[lr], rp = s_rpos[lr];
#pragma unroll 1
            for (int j = j0; j < min(j0 + 4, cnt); ++j) {
                const float* tq = tile_quat + 8 * (j >> 1) + (j & 1);
                const float4 qq = make_float4(tq[0], tq[2], tq[4], tq[6]);
                if (fabsf(fmaf(rq.x, qq.x, fmaf(rq.y, qq.y, fmaf(rq.z, qq.z, rq.w * qq.w)))) < P.qdot_min) continue;
                const float4 q = tile[j];
                if (fmaf(rp.x, q.x, fmaf(rp.y, q.y, fmaf(rp.z, q.z, q.w))) > rp.w) continue;
                const int pos = atomicAdd(&p_count, 1);                 // pass B runs the exact recipe with all lanes busy
                if (pos < CELL_PCAP) pairs[pos] = ((unsigned)lr << 10) | (unsigned)j;
                else if (pair_covered(ref + 14 * (int64_t)ref_sorted_ids[s0 + lr], qry + 14 * (int64_t)tile_id[j], lut, P)) { s_cov[lr] = 1; break; }
            }
        }
    };
    // this block's part of the neighbourhood (a chunk lies in one column, so x0..x1 and y0..y1 hold at most 3 values)
    int xa = x0, xb = x1, ya = y0, yb = y1;
    if (split == 3) { xa = xb = x0 + (int)blockIdx.y; }
    else if (split == 9) { xa = xb = x0 + (int)blockIdx.y / 3; ya = yb = y0 + (int)blockIdx.y % 3; }
    if (xa > x1 || ya > y1) return;                         // (whole block: nothing was staged yet)
    for (int ix = xa; ix <= xb; ++ix)
        for (int iy = ya; iy <= yb; ++iy) {
            // contiguous range of the column's cells z0..z1 (empty cells have start == end == 0)
            int b = -1, e = 0;
            for (int iz = z0; iz <= z1; ++iz) {
                const int k = (ix * g.ny + iy) * g.nz + iz;
                const int cs = __ldg(cell_start + k), ce = __ldg(cell_end + k);
                if (ce > cs) { if (b < 0) b = cs; e = ce; }
            }
            if (b < 0) continue;
            for (int t0 = b; t0 < e; t0 += CELL_TILE) {
                const int cnt = min(CELL_TILE, e - t0);
                __syncthreads();
                for (int j = threadIdx.x; j < cnt; j += CELL_THREADS) {
                    const float4 q = __ldg(q_sorted + t0 + j);
                    const float x = q.x - ccx, y = q.y - ccy, z = q.z - ccz;
                    tile[j] = make_float4(x, y, z, fmaf(x, x, fmaf(y, y, z * z)));
                    const int qi = __float_as_int(q.w);
                    tile_id[j] = qi;
                    const float4 qq = __ldg(q_quat_sorted + t0 + j);
                    float* tq = tile_quat + 8 * (j >> 1) + (j & 1);
                    tq[0] = qq.x; tq[2] = qq.y; tq[4] = qq.z; tq[6] = qq.w;
                }
                if (threadIdx.x < 4) {
                    const int j = cnt + threadIdx.x;
                    float* tq = tile_quat + 8 * (j >> 1) + (j & 1);
                    tq[0] = 0.f; tq[2] = 0.f; tq[4] = 0.f; tq[6] = 0.f;
                }
#pragma unroll
                for (int r = 0; r < CELL_RPT; ++r)
                    if (s_cov[threadIdx.x + r * CELL_THREADS]) qmin[r] = 2.0f;        // already covered: stop testing this row
                __syncthreads();
                // In a grid walk ~15 % of the candidates pass the translation test but only ~0.1 % the rotation test, so the
                // orientation bound goes FIRST here: |q1 . q2| = cos(angle / 2), one product + 3 multiply-adds on the staged
                // quaternions.  Four queries x CELL_RPT rows are evaluated branch-free; two queries share packed registers
                // (FMUL2 + 3 FFMA2 for two pairs), the compares are OR-ed into predicates, and a thread with a passing pair
                // appends the group's index to its own short list (one predicated store, one predicated add: the loop has
                // no branch, a rare path run by one lane would cost the warp its issue slots).  Round 1 recomputed the eight
                // dot products on a divergent path (8.6 issued instructions per pair), round 2's first version pushed
                // (thread, group) through a shared atomic (5.9 + the divergent push); now 4 LDS.128 + 16 packed + 8 FSETP +
                // 4 = 32 per 8 pairs.
                const unsigned slot0 = (unsigned)__cvta_generic_to_shared(&lists[0][threadIdx.x]);
                const unsigned slot_end = slot0 + CELL_LCAP * CELL_THREADS * (unsigned)sizeof(unsigned short);
                unsigned slot = slot0;
                const float4* tq4 = reinterpret_cast<const float4*>(tile_quat);
                for (int j0 = 0; j0 < cnt; j0 += 4) {
                    bool any0 = false, any1 = false;             // (|=, not ||: no short-circuit branches between the chains)
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const float4 qa = tq4[j0 + 2 * u], qb = tq4[j0 + 2 * u + 1];
                        const f32x2 qx = pack2(qa.x, qa.y), qy = pack2(qa.z, qa.w), qz = pack2(qb.x, qb.y), qw = pack2(qb.z, qb.w);
#pragma unroll
                        for (int r = 0; r < CELL_RPT; ++r) {
                            const f32x2 v = fma2(rqx[r], qx, fma2(rqy[r], qy, fma2(rqz[r], qz, mul2(rqw[r], qw))));
                            any0 |= !(fabsf(lo2(v)) < qmin[r]);       // (not >=: a NaN quaternion must pass)
                            any1 |= !(fabsf(hi2(v)) < qmin[r]);
                        }
                    }
                    const bool hit = any0 | any1;
                    if (hit && slot < slot_end) asm volatile("st.shared.u16 [%0], %1;" :: "r"(slot), "h"((unsigned short)j0) : "memory");
                    if (hit) slot += CELL_THREADS * (unsigned)sizeof(unsigned short);
                }
                // pass A, per warp with all lanes busy: prefix sum of the list lengths, entry e -> (owner lane, k)
                int n_mine = (int)((slot - slot0) / (CELL_THREADS * (unsigned)sizeof(unsigned short)));
                if (n_mine > CELL_LCAP) {                        // list overflow: this thread sorts out every group itself
                    for (int j0 = 0; j0 < cnt; j0 += 4) drain_group(threadIdx.x, j0, cnt);
                    n_mine = 0;
                }
                __syncwarp();
                int incl = n_mine;
#pragma unroll
                for (int sft = 1; sft < 32; sft <<= 1) {
                    const int v = __shfl_up_sync(FULL, incl, sft);
                    if (lane >= sft) incl += v;
                }
                const int excl = incl - n_mine, total = __shfl_sync(FULL, incl, 31);
                for (int e0 = 0; e0 < total; e0 += 32) {
                    const int e = e0 + lane;
                    int owner = 0;                               // the last lane whose list starts at or before entry e
#pragma unroll
                    for (int sft = 16; sft >= 1; sft >>= 1) {
                        const int v = __shfl_sync(FULL, excl, owner + sft);
                        if (v <= e) owner += sft;
                    }
                    const int start = __shfl_sync(FULL, excl, owner);
                    if (e < total) drain_group(wbase + owner, (int)lists[e - start][wbase + owner], cnt);
                }
                __syncthreads();
                const int n_p = min(p_count, CELL_PCAP);
                for (int i = threadIdx.x; i < n_p; i += CELL_THREADS) {
                    const unsigned code = pairs[i];
                    const int lr = (int)(code >> 10), j = (int)(code & 1023u);
                    if (!s_cov[lr] && pair_covered(ref + 14 * (int64_t)ref_sorted_ids[s0 + lr], qry + 14 * (int64_t)tile_id[j], lut, P)) s_cov[lr] = 1;
                }
                __syncthreads();
                if (threadIdx.x == 0) p_count = 0;
            }
        }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < CELL_RPT; ++r)
        if (row[r] >= 0 && s_cov[threadIdx.x + r * CELL_THREADS]) covered[row[r]] = 1;
}

__global__ void count_covered_kernel(const uint8_t* __restrict__ covered, int64_t n, unsigned long long* __restrict__ count) {
    unsigned long long c = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) c += covered[i] ? 1 : 0;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) c += __shfl_xor_sync(0xffffffffu, c, s);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}

__global__ void pairwise_kernel(const float* __restrict__ g1, int64_t n1, const float* __restrict__ g2, int64_t n2,
                                int which, float weight, const float* __restrict__ lut, float* __restrict__ out) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= n1 * n2) return;
    const float* a = g1 + 14 * (idx / n2);
    const float* b = g2 + 14 * (idx % n2);
    float d = 0.f, ang = 0.f;
    if (which != 1) {
        const float dx = __fsub_rn(a[0], b[0]), dy = __fsub_rn(a[1], b[1]), dz = __fsub_rn(a[2], b[2]);
        d = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    }
    if (which != 0) {
        float g[3];
#pragma unroll
        for (int r = 0; r < 3; ++r)
            g[r] = __fmaf_rn(a[3 + 3 * r + 2], b[3 + 3 * r + 2], __fmaf_rn(a[3 + 3 * r + 1], b[3 + 3 * r + 1], __fmul_rn(a[3 + 3 * r], b[3 + 3 * r])));
        const float c = __fdiv_rn(__fsub_rn(__fadd_rn(__fadd_rn(g[0], g[1]), g[2]), 1.0f), 2.0f);
        const float kf = rintf(__fmul_rn(c, 10000.0f));
        ang = (fabsf(kf) <= 10000.0f) ? lut[(int)kf + 10000] : __int_as_float(0x7fc00000);
    }
    out[idx] = which == 0 ? d : (which == 1 ? ang : __fadd_rn(__fmul_rn(weight, d), ang));
}

// ---- avg_gripper_point_distances (metrics.py:94-138) ------------------------------------------------------------------
struct KeyPoints { double p[32][3]; int k; };

// key points of every grasp: float32 pose entries x float64 points, products accumulated as numpy's einsum kernel does,
// in two lanes: (R0 p0 + R2 p2) + (R1 p1 + t * 1)
__global__ void keypoints_kernel(const float* __restrict__ gs, int64_t n, KeyPoints kp, double* __restrict__ out) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= n * kp.k) return;
    const int64_t i = idx / kp.k;
    const int l = (int)(idx - i * kp.k);
    const float* g = gs + 14 * i;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const double r0 = g[3 + 3 * j], r1 = g[4 + 3 * j], r2 = g[5 + 3 * j], t = g[j];
        out[3 * idx + j] = (r0 * kp.p[l][0] + r2 * kp.p[l][2]) + (r1 * kp.p[l][1] + t * 1.0);
    }
}

__global__ void avg_point_distance_kernel(const double* __restrict__ kp1, int64_t n1, const double* __restrict__ kp2, int64_t n2,
                                          int k, int symmetric, double* __restrict__ out) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= n1 * n2) return;
    const double* a = kp1 + 3 * k * (idx / n2);
    const double* b = kp2 + 3 * k * (idx % n2);
    const int flip[5] = {1, 0, 3, 2, 4};                        // metrics.py:131
    double sum = 0.0;
    for (int l = 0; l < k; ++l) {
        d3 dv = sub3(ld3(a + 3 * l), ld3(b + 3 * l));
        double d = norm3(dv);
        if (symmetric) d = fmin(d, norm3(sub3(ld3(a + 3 * l), ld3(b + 3 * flip[l]))));
        sum = (l == 0) ? d : sum + d;
    }
    out[idx] = sum / (double)k;
}

}  // namespace bt

using namespace bt;

extern "C" int bt_avg_gripper_point_distances(const float* d_gs1, int64_t n1, const float* d_gs2, int64_t n2,
                                              const double* h_points, int32_t n_points, int32_t consider_symmetry,
                                              double* d_out, void* stream) {
    BT_REQUIRE(d_gs1 && d_gs2 && h_points && d_out, "null pointer");
    BT_REQUIRE(n_points >= 1 && n_points <= 32, "1 <= n_points <= 32");
    BT_REQUIRE(!consider_symmetry || n_points == 5, "consider_symmetry needs the 5 default key points (reference metrics.py:128-131)");
    if (n1 <= 0 || n2 <= 0) return BT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    KeyPoints kp;
    kp.k = n_points;
    for (int l = 0; l < 32; ++l) for (int c = 0; c < 3; ++c) kp.p[l][c] = l < n_points ? h_points[3 * l + c] : 0.0;
    Scratch k1, k2;
    BT_CUDA(k1.alloc(sizeof(double) * 3 * n_points * n1, st));
    BT_CUDA(k2.alloc(sizeof(double) * 3 * n_points * n2, st));
    keypoints_kernel<<<(unsigned)ceil_div(n1 * n_points, 256), 256, 0, st>>>(d_gs1, n1, kp, k1.as<double>());
    keypoints_kernel<<<(unsigned)ceil_div(n2 * n_points, 256), 256, 0, st>>>(d_gs2, n2, kp, k2.as<double>());
    avg_point_distance_kernel<<<(unsigned)ceil_div(n1 * n2, 256), 256, 0, st>>>(k1.as<double>(), n1, k2.as<double>(), n2, n_points,
                                                                               consider_symmetry, d_out);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

// bounding box of both translation sets (one tiny read-back; the metric API is synchronous anyway)
static int data_bounds(const float* d_ref, int64_t n_ref, const float* d_qry, int64_t n_qry, float lo[3], float hi[3],
                       cudaStream_t st) {
    Scratch mm;
    BT_CUDA(mm.alloc(6 * sizeof(uint32_t), st));
    uint32_t h_mm[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    BT_CUDA(cudaMemcpyAsync(mm.p, h_mm, sizeof(h_mm), cudaMemcpyHostToDevice, st));
    bounds_kernel<<<148 * 4, 256, 0, st>>>(d_ref, n_ref, mm.as<uint32_t>());
    bounds_kernel<<<148 * 4, 256, 0, st>>>(d_qry, n_qry, mm.as<uint32_t>());
    BT_LAUNCH_CHECK();
    BT_CUDA(cudaMemcpyAsync(h_mm, mm.p, sizeof(h_mm), cudaMemcpyDeviceToHost, st));
    BT_CUDA(cudaStreamSynchronize(st));
    for (int k = 0; k < 3; ++k) { lo[k] = float_unflip(h_mm[k]); hi[k] = float_unflip(h_mm[3 + k]); }
    return BT_OK;
}

// ---- the query side of the uniform-grid coverage, prepared once: grid over the QUERY set's bounds (cell >= the largest
// distance that can pass the exact test), queries sorted by cell (float4 translations + unit quaternions) and the cell range
// table.  A sweep -- many reference sets, or the shards of one reference set on several GPUs, against one query set --
// bins only its reference rows per call.  Reference rows outside the query bounds fall into the clamped border cells:
// their 27-cell neighbourhood then still holds every query within reach, and the exact test rejects the rest.
struct bt_coverage_query {
    int device = 0;
    const float* d_qry = nullptr;         // the caller's rows (must outlive the handle: the exact test reads them)
    int64_t n_qry = 0;
    double epsilon = 0.0;
    bt::Grid g{};
    bool dense = false, use_cells = false;
    int key_bits = 1;
    float4* qs = nullptr;                 // [n_qry] sorted translations (+ original index)
    float4* qq = nullptr;                 // [n_qry] sorted unit quaternions (cell-tile kernel)
    int32_t *cs = nullptr, *ce = nullptr; // dense cell ranges
    uint32_t* keys_s = nullptr;           // sorted cell keys (sparse grids: binary search)
    cudaStream_t transient = nullptr;     // one-shot handles (bt_coverage, algo 1): stream-ordered memory, freed on this stream
    bool is_transient = false;
};

namespace bt {

static void free_query(bt_coverage_query* q) {
    if (!q) return;
    void* bufs[5] = {q->qs, q->qq, q->cs, q->ce, q->keys_s};
    for (void* b : bufs) {
        if (!b) continue;
        if (q->is_transient) cudaFreeAsync(b, q->transient); else cudaFree(b);
    }
    delete q;
}

static int prepare_query(const float* d_qry, int64_t n_qry, double epsilon, float r_bound, bt_coverage_query** out, cudaStream_t st,
                         bool transient = false) {
    int device = 0;
    BT_CUDA(cudaGetDevice(&device));
    float lo[3], hi[3];
    {
        int rc = data_bounds(d_qry, n_qry, nullptr, 0, lo, hi, st);   // host sync: the grid shape is decided here
        if (rc != BT_OK) return rc;
    }
    float extent = 0.f;
    for (int k = 0; k < 3; ++k) extent = fmaxf(extent, hi[k] - lo[k]);
    if (!(extent >= 0.f) || !isfinite(extent)) { set_error("bt_coverage: non-finite translations"); return BT_E_INVALID; }
    // cell size: >= the largest distance that can pass the exact test, with head-room for the float32 rounding of
    // the cell index, and at most 1023 cells per axis
    float h = fmaxf(r_bound * 1.01f, extent / 1023.0f);
    if (!(h > 0.f)) h = 1.0f;
    bt_coverage_query* q = new bt_coverage_query();
    q->device = device; q->d_qry = d_qry; q->n_qry = n_qry; q->epsilon = epsilon;
    q->is_transient = transient; q->transient = st;
    if (transient) retain_pool_memory();
    auto q_alloc = [&](void** p, size_t bytes) { return transient ? cudaMallocAsync(p, bytes ? bytes : 1, st) : cudaMalloc(p, bytes ? bytes : 1); };
    Grid& g = q->g;
    g.ox = lo[0]; g.oy = lo[1]; g.oz = lo[2];
    g.inv_h = 1.0f / h;
    g.nx = (int)fminf(1024.f, floorf((hi[0] - lo[0]) / h) + 1.f);
    g.ny = (int)fminf(1024.f, floorf((hi[1] - lo[1]) / h) + 1.f);
    g.nz = (int)fminf(1024.f, floorf((hi[2] - lo[2]) / h) + 1.f);
    const int64_t n_cells = (int64_t)g.nx * g.ny * g.nz;
    q->dense = n_cells <= (1ll << 24);
    // dense data (>= 16 queries per cell; measured at ~30 per cell: 0.79 -> 0.45 ms for 100k x 100k): block-per-chunk tiles;
    // sparse data: warp-per-reference walk
    static double cells_min = -1.0;                       // tuning knob (environment, read once): queries per cell from which the cell-tile kernel runs
    if (cells_min < 0.0) { const char* e = getenv("BT_COVERAGE_CELLS_MIN"); cells_min = e ? atof(e) : 16.0; }
    q->use_cells = q->dense && (double)n_qry / (double)n_cells >= cells_min;
    while ((1ll << q->key_bits) < n_cells) ++q->key_bits;
#define Q_TRY(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { free_query(q); return cuda_fail(_e, #call, __FILE__, __LINE__); } } while (0)
    Q_TRY(q_alloc((void**)&q->qs, sizeof(float4) * n_qry));
    Q_TRY(q_alloc((void**)&q->keys_s, sizeof(uint32_t) * n_qry));
    if (q->use_cells) Q_TRY(q_alloc((void**)&q->qq, sizeof(float4) * n_qry));
    if (q->dense) {
        Q_TRY(q_alloc((void**)&q->cs, sizeof(int32_t) * n_cells));
        Q_TRY(q_alloc((void**)&q->ce, sizeof(int32_t) * n_cells));
        Q_TRY(cudaMemsetAsync(q->cs, 0, sizeof(int32_t) * n_cells, st));
        Q_TRY(cudaMemsetAsync(q->ce, 0, sizeof(int32_t) * n_cells, st));
    }
    Scratch keys, ids, ids_s, tmp;
    Q_TRY(keys.alloc(sizeof(uint32_t) * n_qry, st));
    Q_TRY(ids.alloc(sizeof(int32_t) * n_qry, st)); Q_TRY(ids_s.alloc(sizeof(int32_t) * n_qry, st));
    size_t b1 = 0;
    Q_TRY(cub::DeviceRadixSort::SortPairs(nullptr, b1, keys.as<uint32_t>(), q->keys_s, ids.as<int32_t>(), ids_s.as<int32_t>(), (int)n_qry, 0, q->key_bits, st));
    Q_TRY(tmp.alloc(b1, st));
    cell_key_kernel<<<(unsigned)ceil_div(n_qry, 256), 256, 0, st>>>(d_qry, n_qry, g, keys.as<uint32_t>(), ids.as<int32_t>());
    Q_TRY(cudaGetLastError());
    Q_TRY(cub::DeviceRadixSort::SortPairs(tmp.p, b1, keys.as<uint32_t>(), q->keys_s, ids.as<int32_t>(), ids_s.as<int32_t>(), (int)n_qry, 0, q->key_bits, st));
    gather_sorted_kernel<<<(unsigned)ceil_div(n_qry, 256), 256, 0, st>>>(d_qry, ids_s.as<int32_t>(), q->keys_s, n_qry, q->qs, q->qq, q->cs, q->ce);
    Q_TRY(cudaGetLastError());
#undef Q_TRY
    *out = q;
    return BT_OK;
}

// the reference side, per call: bin the reference rows in the query's grid, then one of the two grid kernels
static int run_prepared(const float* d_ref, int64_t n_ref, const bt_coverage_query* q, const float* lut, CovParams P,
                        uint8_t* d_covered, cudaStream_t st) {
    const Grid g = q->g;
    const float* d_qry = q->d_qry;
    const int64_t n_qry = q->n_qry;
    Scratch rkeys, rkeys_s, rids, rids_s, tmp;
    BT_CUDA(rkeys.alloc(sizeof(uint32_t) * n_ref, st)); BT_CUDA(rkeys_s.alloc(sizeof(uint32_t) * n_ref, st));
    BT_CUDA(rids.alloc(sizeof(int32_t) * n_ref, st));   BT_CUDA(rids_s.alloc(sizeof(int32_t) * n_ref, st));
    size_t b2 = 0;
    BT_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, b2, rkeys.as<uint32_t>(), rkeys_s.as<uint32_t>(), rids.as<int32_t>(), rids_s.as<int32_t>(), (int)n_ref, 0, q->key_bits, st));
    BT_CUDA(tmp.alloc(b2, st));
    cell_key_kernel<<<(unsigned)ceil_div(n_ref, 256), 256, 0, st>>>(d_ref, n_ref, g, rkeys.as<uint32_t>(), rids.as<int32_t>());
    BT_LAUNCH_CHECK();
    BT_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, b2, rkeys.as<uint32_t>(), rkeys_s.as<uint32_t>(), rids.as<int32_t>(), rids_s.as<int32_t>(), (int)n_ref, 0, q->key_bits, st));
    if (q->use_cells) {
        const int n_cols = g.nx * g.ny;
        const int64_t max_chunks = ceil_div(n_ref, CELL_REFS) + n_cols;
        Scratch col_b, col_e, chunks, n_chunks;
        BT_CUDA(col_b.alloc(sizeof(int32_t) * n_cols, st)); BT_CUDA(col_e.alloc(sizeof(int32_t) * n_cols, st));
        BT_CUDA(chunks.alloc(sizeof(int2) * max_chunks, st)); BT_CUDA(n_chunks.alloc(sizeof(int), st));
        BT_CUDA(cudaMemsetAsync(col_b.p, 0, sizeof(int32_t) * n_cols, st));
        BT_CUDA(cudaMemsetAsync(col_e.p, 0, sizeof(int32_t) * n_cols, st));
        BT_CUDA(cudaMemsetAsync(n_chunks.p, 0, sizeof(int), st));
        column_bounds_kernel<<<(unsigned)ceil_div(n_ref, 256), 256, 0, st>>>(rkeys_s.as<uint32_t>(), n_ref, (uint32_t)g.nz, col_b.as<int32_t>(), col_e.as<int32_t>());
        emit_chunks_kernel<<<(unsigned)ceil_div(n_cols, 256), 256, 0, st>>>(col_b.as<int32_t>(), col_e.as<int32_t>(), n_cols, CELL_REFS, chunks.as<int2>(), n_chunks.as<int>());
        BT_LAUNCH_CHECK();
        // enough blocks for ~16 per SM: split the chunks' neighbourhoods when the shard is small
        static int split_env = -1;
        if (split_env < 0) { const char* e = getenv("BT_COVERAGE_SPLIT"); split_env = e ? atoi(e) : 0; }
        const int64_t est_chunks = ceil_div(n_ref, CELL_REFS);
        int split = est_chunks >= 148 * 16 ? 1 : (est_chunks * 3 >= 148 * 16 ? 3 : 9);
        if (split_env == 1 || split_env == 3 || split_env == 9) split = split_env;
        coverage_cells_kernel<<<dim3((unsigned)max_chunks, (unsigned)split), CELL_THREADS, 0, st>>>(d_ref, rids_s.as<int32_t>(), chunks.as<int2>(), n_chunks.as<int>(),
                                                                                                    d_qry, q->qs, q->qq, q->cs, q->ce, g, lut, P, split, d_covered);
    } else
        coverage_grid_kernel<<<(unsigned)ceil_div(n_ref * 32, 256), 256, 0, st>>>(d_ref, rids_s.as<int32_t>(), n_ref, d_qry, q->qs, q->keys_s,
                                                                                  (int)n_qry, q->dense ? q->cs : nullptr, q->dense ? q->ce : nullptr,
                                                                                  g, lut, P, d_covered);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

// exact-test parameters of a coverage call (the float32 compare of metrics.py:157/:219, the float64 radius of :209)
static CovParams coverage_params(double epsilon, int mode, float* r_bound_out) {
    CovParams P;
    P.eps = (float)epsilon; P.weight = 1000.0f; P.mode = mode;
    const double r = epsilon / 1000.0;
    P.r2_kd = r * r;
    // largest d with fl(1000*d) <= eps is eps/1000 * (1 + 2^-23); add margin
    const float r_bound = (float)(r * (1.0 + 1e-6)) + 1e-12f;
    P.cx = P.cy = P.cz = 0.f;
    // deg <= eps needs angle <= eps (+ the 4-decimal rounding of the cosine, < 0.03 deg at eps >= 1): |q1.q2| >= cos((eps + 0.1 deg)/2)
    P.qdot_min = epsilon >= 179.0 ? -1.0f : (float)(cos((epsilon + 0.1) * 0.5 * 3.14159265358979323846 / 180.0) - 2e-5);
    P.r2_reject = r_bound * r_bound * 1.0001f;          // direct-difference form: a few ulps of rounding only
    *r_bound_out = r_bound;
    return P;
}

}  // namespace bt

using namespace bt;

extern "C" int bt_coverage_prepare(const float* d_qry, int64_t n_qry, double epsilon, bt_coverage_query** out, void* stream) {
    BT_REQUIRE(d_qry && out, "null pointer");
    BT_REQUIRE(n_qry >= 1 && n_qry < (1ll << 31), "bad counts");
    BT_REQUIRE(epsilon >= 0.0, "epsilon must be >= 0");
    float r_bound;
    coverage_params(epsilon, 1, &r_bound);
    return prepare_query(d_qry, n_qry, epsilon, r_bound, out, (cudaStream_t)stream);
}

extern "C" int bt_coverage_query_destroy(bt_coverage_query* q) {
    if (!q) return BT_OK;
    cudaSetDevice(q->device);
    free_query(q);
    return BT_OK;
}

extern "C" int bt_coverage_prepared(const float* d_ref, int64_t n_ref, const bt_coverage_query* query, const float* d_lut,
                                    int32_t mode, uint8_t* d_covered, int64_t* d_count, void* stream) {
    BT_REQUIRE(query && d_lut && d_count, "null pointer");
    BT_REQUIRE(n_ref >= 0 && n_ref < (1ll << 31), "bad counts");
    BT_REQUIRE((d_ref && d_covered) || n_ref == 0, "null reference set");
    BT_REQUIRE(mode == 0 || mode == 1, "mode must be 0 (brute force semantics) or 1 (kd-tree semantics)");
    cudaStream_t st = (cudaStream_t)stream;
    BT_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int64_t), st));
    if (n_ref == 0) return BT_OK;
    BT_CUDA(cudaMemsetAsync(d_covered, 0, (size_t)n_ref, st));
    float r_bound;
    const CovParams P = coverage_params(query->epsilon, mode, &r_bound);
    int rc = run_prepared(d_ref, n_ref, query, d_lut, P, d_covered, st);
    if (rc != BT_OK) return rc;
    count_covered_kernel<<<148, 256, 0, st>>>(d_covered, n_ref, reinterpret_cast<unsigned long long*>(d_count));
    BT_LAUNCH_CHECK();
    return BT_OK;
}

extern "C" int bt_coverage(const float* d_ref, int64_t n_ref, const float* d_qry, int64_t n_qry, double epsilon,
                           const float* d_lut, int32_t mode, int32_t algo, uint8_t* d_covered, int64_t* d_count,
                           void* stream) {
    BT_REQUIRE(d_lut && d_count, "null pointer");
    BT_REQUIRE(n_ref >= 0 && n_qry >= 0 && n_ref < (1ll << 31) && n_qry < (1ll << 31), "bad counts");
    BT_REQUIRE((d_ref && d_covered) || n_ref == 0, "null reference set");
    BT_REQUIRE(d_qry || n_qry == 0, "null query set");
    BT_REQUIRE(mode == 0 || mode == 1, "mode must be 0 (brute force semantics) or 1 (kd-tree semantics)");
    BT_REQUIRE(algo == 0 || algo == 1, "algo must be 0 (tiles) or 1 (grid)");
    cudaStream_t st = (cudaStream_t)stream;
    BT_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int64_t), st));
    if (n_ref == 0) return BT_OK;
    BT_CUDA(cudaMemsetAsync(d_covered, 0, (size_t)n_ref, st));
    if (n_qry == 0 || !(epsilon >= 0.0)) return BT_OK;
    float r_bound;
    CovParams P = coverage_params(epsilon, mode, &r_bound);
    if (algo == 1) {
        bt_coverage_query* q = nullptr;
        int rc = prepare_query(d_qry, n_qry, epsilon, r_bound, &q, st, /*transient=*/true);
        if (rc != BT_OK) return rc;
        rc = run_prepared(d_ref, n_ref, q, d_lut, P, d_covered, st);
        free_query(q);                                      // stream-ordered: behind the kernels that read the buffers
        if (rc != BT_OK) return rc;
    } else {
        // centre on the first query row and bound the rounding of the expanded form by the data's spread; both are
        // read back once (tiny, but a host sync -- the metric API is synchronous anyway)
        float lo[3], hi[3], smax = 0.f;
        {
            int rc = data_bounds(d_ref, n_ref, d_qry, n_qry, lo, hi, st);
            if (rc != BT_OK) return rc;
        }
        P.cx = 0.5f * (lo[0] + hi[0]); P.cy = 0.5f * (lo[1] + hi[1]); P.cz = 0.5f * (lo[2] + hi[2]);
        for (int k = 0; k < 3; ++k) { const float e = 0.5f * (hi[k] - lo[k]) * 1.001f; smax += e * e; }
        if (!isfinite(smax)) { set_error("bt_coverage: non-finite translations"); return BT_E_INVALID; }
        // |computed - exact| <= ~24 eps32 * max|x|^2 for the 4-term FMA chain plus the two precomputed norms
        P.r2_reject = r_bound * r_bound * 1.0001f + 64.0f * 5.97e-8f * smax;
        const int64_t row_blocks = ceil_div(n_ref, COV_THREADS * COV_RPT);
        int64_t splits = ceil_div(148 * 8, row_blocks);
        splits = splits < 1 ? 1 : splits;
        int64_t q_per_split = ceil_div(ceil_div(n_qry, splits), COV_TILE) * COV_TILE;
        splits = ceil_div(n_qry, q_per_split);
        dim3 grid((unsigned)row_blocks, (unsigned)splits);
        static bool smem_opt_in[64] = {};                   // > 48 KB of shared memory per block: opt in once per device
        int dev = 0;
        BT_CUDA(cudaGetDevice(&dev));
        if (dev < 0 || dev >= 64 || !smem_opt_in[dev]) {
            BT_CUDA(cudaFuncSetAttribute(coverage_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CovTilesSmem)));
            if (dev >= 0 && dev < 64) smem_opt_in[dev] = true;
        }
        Scratch quats;
        BT_CUDA(quats.alloc(sizeof(float4) * (size_t)(n_ref + n_qry), st));
        float4* ref_quat = quats.as<float4>();
        float4* qry_quat = ref_quat + n_ref;
        row_quat_kernel<<<(unsigned)ceil_div(n_ref, 256), 256, 0, st>>>(d_ref, n_ref, ref_quat);
        BT_LAUNCH_CHECK();
        row_quat_kernel<<<(unsigned)ceil_div(n_qry, 256), 256, 0, st>>>(d_qry, n_qry, qry_quat);
        BT_LAUNCH_CHECK();
        coverage_tiles_kernel<<<grid, COV_THREADS, sizeof(CovTilesSmem), st>>>(d_ref, n_ref, d_qry, n_qry, ref_quat, qry_quat, q_per_split, d_lut, P, d_covered);
        BT_LAUNCH_CHECK();
    }
    count_covered_kernel<<<148, 256, 0, st>>>(d_covered, n_ref, reinterpret_cast<unsigned long long*>(d_count));
    BT_LAUNCH_CHECK();
    return BT_OK;
}

extern "C" int bt_pairwise_distances(const float* d_gs1, int64_t n1, const float* d_gs2, int64_t n2, int32_t which,
                                     float weight, const float* d_lut, float* d_out, void* stream) {
    BT_REQUIRE(d_gs1 && d_gs2 && d_out, "null pointer");
    BT_REQUIRE(which >= 0 && which <= 2, "which must be 0, 1 or 2");
    BT_REQUIRE(which == 0 || d_lut, "lookup table required for angular / combined distances");
    if (n1 <= 0 || n2 <= 0) return BT_OK;
    pairwise_kernel<<<(unsigned)ceil_div(n1 * n2, 256), 256, 0, (cudaStream_t)stream>>>(d_gs1, n1, d_gs2, n2, which, weight, d_lut, d_out);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

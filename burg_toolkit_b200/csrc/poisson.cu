// poisson.cu -- weighted sample elimination on the device: the second stage of open3d's
// TriangleMesh::SamplePointsPoissonDisk, which the reference calls for its first-contact candidates
// (burg_toolkit/mesh_processing.py:52-83 <- sampling.py:199-200).
//
// open3d (Yuksel 2015): draw init_factor * n uniform surface samples, give every sample the weight
//     w_i = sum over living samples j within r_max of (1 - max(d_ij, r_min) / r_max)^alpha ,
// and remove the sample with the largest weight (its neighbours' weights drop) until n are left -- a serial
// priority-queue loop.  Weights only ever decrease, so the serial loop removes samples in descending order of their
// weight AT REMOVAL TIME.  That order can be computed in parallel: a sample that outweighs all its living neighbours
// ("local maximum"; ties: the smaller index) keeps its weight until it is removed itself, and none of its neighbours can be
// removed before it -- so removing ALL local maxima at once, round after round, gives every sample the removal-time weight
// the serial loop would give it.  The serial result after K = n0 - n removals is then simply: the K samples with the
// largest removal-time weights.  The rounds continue (past K removals) until no living sample can still reach the K-th
// largest removal-time weight, then everything but the top K is brought back.  The same result as the serial algorithm
// up to rounding (float32 sums here, incremental float64 updates there).  The point set is binned in a uniform grid of
// cell >= r_max (CUB radix sort by cell), so a sample's neighbours are in its 27 surrounding cells.
//
// open3d seeds its generator from std::random_device, i.e. the reference's sample set is not reproducible and parity at
// this stage is distributional (blue-noise statistics against the serial restatement in the test oracle on the same
// initial samples: tests/test_gpu_poisson.py).
#include <cub/cub.cuh>
#include <math.h>
#include <algorithm>

#include "common.cuh"

namespace bt {

struct PGrid { float ox, oy, oz, inv_cell; int nx, ny, nz; };

__device__ __forceinline__ unsigned enc_f(float f) { const unsigned u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__host__ __device__ inline float dec_f(unsigned u) {
    const unsigned v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    float f;
    memcpy(&f, &v, sizeof(f));
    return f;
}

__global__ void pd_bounds_kernel(const double* __restrict__ pts, int64_t n, unsigned* __restrict__ mm) {
    float lo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, hi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        for (int k = 0; k < 3; ++k) { const float v = (float)pts[3 * i + k]; lo[k] = fminf(lo[k], v); hi[k] = fmaxf(hi[k], v); }
    for (int k = 0; k < 3; ++k) {
        for (int s = 16; s > 0; s >>= 1) { lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], s)); hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], s)); }
        if ((threadIdx.x & 31) == 0) { atomicMin(mm + k, enc_f(lo[k])); atomicMax(mm + 3 + k, enc_f(hi[k])); }
    }
}

__device__ __forceinline__ int3 pd_cell(const PGrid& g, float x, float y, float z) {
    return make_int3(min(g.nx - 1, max(0, (int)((x - g.ox) * g.inv_cell))), min(g.ny - 1, max(0, (int)((y - g.oy) * g.inv_cell))),
                     min(g.nz - 1, max(0, (int)((z - g.oz) * g.inv_cell))));
}

__global__ void pd_key_kernel(const double* __restrict__ pts, int64_t n, PGrid g, uint32_t* __restrict__ keys, int32_t* __restrict__ ids) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int3 c = pd_cell(g, (float)pts[3 * i], (float)pts[3 * i + 1], (float)pts[3 * i + 2]);
    keys[i] = (uint32_t)((c.z * g.ny + c.y) * g.nx + c.x);
    ids[i] = (int32_t)i;
}

// sorted order: positions relative to the grid origin (float32: a 7 cm object is resolved to ~4 nm) + cell ranges
__global__ void pd_gather_kernel(const double* __restrict__ pts, const int32_t* __restrict__ ids_s, const uint32_t* __restrict__ keys_s,
                                 int64_t n, PGrid g, float4* __restrict__ pos, int32_t* __restrict__ cell_begin, int32_t* __restrict__ cell_end) {
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int64_t i = ids_s[p];
    pos[p] = make_float4((float)pts[3 * i] - g.ox, (float)pts[3 * i + 1] - g.oy, (float)pts[3 * i + 2] - g.oz, 0.f);
    const uint32_t k = keys_s[p];
    if (p == 0 || keys_s[p - 1] != k) cell_begin[k] = (int32_t)p;
    if (p == n - 1 || keys_s[p + 1] != k) cell_end[k] = (int32_t)p + 1;
}

struct PdParams { float r_max, r_min, alpha; };

// f(q) for every living neighbour q of sample p (27 cells, d < r_max)
template <class F>
__device__ __forceinline__ void pd_for_neighbours(const PGrid& g, const float4* __restrict__ pos, const uint8_t* __restrict__ alive,
                                                  const int32_t* __restrict__ cell_begin, const int32_t* __restrict__ cell_end, int p,
                                                  float r_max, F f) {
    const float4 a = pos[p];
    const int cx = min(g.nx - 1, (int)(a.x * g.inv_cell)), cy = min(g.ny - 1, (int)(a.y * g.inv_cell)), cz = min(g.nz - 1, (int)(a.z * g.inv_cell));
    const float r2 = r_max * r_max;
    for (int dz = -1; dz <= 1; ++dz) {
        const int z = cz + dz;
        if (z < 0 || z >= g.nz) continue;
        for (int dy = -1; dy <= 1; ++dy) {
            const int y = cy + dy;
            if (y < 0 || y >= g.ny) continue;
            // the three cells of a row are adjacent keys, but not adjacent ranges (empty cells keep stale bounds): walk them one by one
            for (int dx = -1; dx <= 1; ++dx) {
                const int x = cx + dx;
                if (x < 0 || x >= g.nx) continue;
                const int k = (z * g.ny + y) * g.nx + x;
                const int b = cell_begin[k], e = cell_end[k];
                for (int q = b; q < e; ++q) {
                    if (q == p || !alive[q]) continue;
                    const float4 o = pos[q];
                    const float ddx = o.x - a.x, ddy = o.y - a.y, ddz = o.z - a.z;
                    const float d2 = fmaf(ddx, ddx, fmaf(ddy, ddy, ddz * ddz));
                    if (d2 < r2) f(q, d2);
                }
            }
        }
    }
}

__global__ void pd_weight_kernel(PGrid g, const float4* __restrict__ pos, const uint8_t* __restrict__ alive, const int32_t* __restrict__ cell_begin,
                                 const int32_t* __restrict__ cell_end, int n, PdParams P, float* __restrict__ weight) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    if (!alive[p]) { weight[p] = -1.f; return; }
    float w = 0.f;
    pd_for_neighbours(g, pos, alive, cell_begin, cell_end, p, P.r_max, [&](int, float d2) {
        const float d = fmaxf(sqrtf(d2), P.r_min);
        w += powf(1.f - d / P.r_max, P.alpha);
    });
    weight[p] = w;
}

// candidate = living sample that outweighs every living neighbour (ties: the smaller sample index wins) and can still belong
// to the K heaviest removals (weight >= floor).  Candidates are removed by pd_kill_kernel, which records their weight.
__global__ void pd_candidate_kernel(PGrid g, const float4* __restrict__ pos, const uint8_t* __restrict__ alive, const int32_t* __restrict__ cell_begin,
                                    const int32_t* __restrict__ cell_end, int n, PdParams P, const float* __restrict__ weight, float floor_w,
                                    const int32_t* __restrict__ ids_s, uint8_t* __restrict__ cand, int* __restrict__ n_cand) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    bool c = false;
    if (p < n && alive[p]) {
        const float w = weight[p];
        c = w >= floor_w;
        if (c)
            pd_for_neighbours(g, pos, alive, cell_begin, cell_end, p, P.r_max, [&](int q, float) {
                const float wq = weight[q];
                if (wq > w || (wq == w && ids_s[q] < ids_s[p])) c = false;      // ties: the smaller sample index goes first (a heap of (-w, i))
            });
    }
    if (p < n) cand[p] = c ? 1 : 0;
    const unsigned b = __ballot_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(n_cand, __popc(b));
}

__global__ void pd_kill_kernel(int n, const uint8_t* __restrict__ cand, const float* __restrict__ weight, uint8_t* __restrict__ alive,
                               float* __restrict__ w_removed) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n && cand[p]) { alive[p] = 0; w_removed[p] = weight[p]; }
}
__global__ void pd_init_kernel(int n, int32_t* __restrict__ idx, float* __restrict__ w_removed) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) { idx[p] = p; w_removed[p] = -1.0f; }
}
// after the sort by removal-time weight (descending): everything behind the first k comes back to life
__global__ void pd_revive_kernel(int n, int k, const int32_t* __restrict__ idx_sorted, uint8_t* __restrict__ alive) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) alive[idx_sorted[t]] = t < k ? 0 : 1;
}
__global__ void pd_scatter_kernel(int n, const int32_t* __restrict__ ids_s, const uint8_t* __restrict__ alive, uint8_t* __restrict__ keep) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) keep[ids_s[p]] = alive[p];
}

}  // namespace bt

using namespace bt;

extern "C" int bt_poisson_disk_eliminate(const double* d_pts, int64_t n0, int64_t n_target, double surface_area, double alpha,
                                         double beta, double gamma, uint8_t* d_keep, int32_t* h_rounds, void* stream) {
    BT_REQUIRE(d_pts && d_keep, "null pointer");
    BT_REQUIRE(n0 >= 1 && n0 < ((int64_t)1 << 30) && n_target >= 1 && surface_area > 0.0, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (h_rounds) *h_rounds = 0;
    if (n_target >= n0) { BT_CUDA(cudaMemsetAsync(d_keep, 1, (size_t)n0, st)); return BT_OK; }
    const int n = (int)n0;
    // open3d: r_max from the ideal hexagonal packing of n discs on the surface; r_min flattens the weight of very close pairs
    const double r_max = 2.0 * sqrt((surface_area / (double)n_target) / (2.0 * sqrt(3.0)));
    const double r_min = r_max * beta * (1.0 - pow((double)n_target / (double)n0, gamma));
    PdParams P{(float)r_max, (float)r_min, (float)alpha};

    Scratch mm, keys, ids, keys_s, ids_s, pos, alive, cand, weight, counter, tmp, fkeys, fkeys_s, fidx, fidx_s, cb, ce;
    BT_CUDA(mm.alloc(6 * sizeof(unsigned), st));
    {
        const unsigned init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
        BT_CUDA(cudaMemcpyAsync(mm.p, init, sizeof(init), cudaMemcpyHostToDevice, st));
        BT_CUDA(cudaStreamSynchronize(st));                  // (init lives on the host stack)
    }
    pd_bounds_kernel<<<148 * 4, 256, 0, st>>>(d_pts, n0, mm.as<unsigned>());
    BT_LAUNCH_CHECK();
    unsigned h_mm[6];
    BT_CUDA(cudaMemcpyAsync(h_mm, mm.p, sizeof(h_mm), cudaMemcpyDeviceToHost, st));
    BT_CUDA(cudaStreamSynchronize(st));
    float lo[3], hi[3];
    for (int k = 0; k < 3; ++k) { lo[k] = dec_f(h_mm[k]); hi[k] = dec_f(h_mm[3 + k]); }
    // cell >= r_max (a little more: float32 cell arithmetic), at most 2^25 cells
    double cell = r_max * 1.001 + 1e-12;
    int dims[3];
    while (true) {
        double cells = 1.0;
        for (int k = 0; k < 3; ++k) { dims[k] = (int)floor((double)(hi[k] - lo[k]) / cell) + 1; cells *= dims[k]; }
        if (cells <= (double)(1 << 25)) break;
        cell *= 1.26;
    }
    PGrid g{lo[0], lo[1], lo[2], (float)(1.0 / cell), dims[0], dims[1], dims[2]};
    const int64_t n_cells = (int64_t)dims[0] * dims[1] * dims[2];
    BT_CUDA(keys.alloc(sizeof(uint32_t) * n0, st)); BT_CUDA(ids.alloc(sizeof(int32_t) * n0, st));
    BT_CUDA(keys_s.alloc(sizeof(uint32_t) * n0, st)); BT_CUDA(ids_s.alloc(sizeof(int32_t) * n0, st));
    BT_CUDA(pos.alloc(sizeof(float4) * n0, st)); BT_CUDA(alive.alloc((size_t)n0, st)); BT_CUDA(cand.alloc((size_t)n0, st));
    BT_CUDA(weight.alloc(sizeof(float) * n0, st)); BT_CUDA(counter.alloc(sizeof(int), st));
    BT_CUDA(cb.alloc(sizeof(int32_t) * n_cells, st)); BT_CUDA(ce.alloc(sizeof(int32_t) * n_cells, st));
    BT_CUDA(cudaMemsetAsync(cb.p, 0, sizeof(int32_t) * n_cells, st));      // empty cells: begin == end == 0
    BT_CUDA(cudaMemsetAsync(ce.p, 0, sizeof(int32_t) * n_cells, st));
    BT_CUDA(cudaMemsetAsync(alive.p, 1, (size_t)n0, st));
    const unsigned G = (unsigned)ceil_div(n0, 256);
    pd_key_kernel<<<G, 256, 0, st>>>(d_pts, n0, g, keys.as<uint32_t>(), ids.as<int32_t>());
    BT_LAUNCH_CHECK();
    int key_bits = 1;
    while (((int64_t)1 << key_bits) < n_cells) ++key_bits;
    size_t tmp_bytes = 0, tmp_bytes2 = 0;
    BT_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys.as<uint32_t>(), keys_s.as<uint32_t>(), ids.as<int32_t>(), ids_s.as<int32_t>(), n, 0, key_bits, st));
    BT_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_bytes2, (const float*)nullptr, (float*)nullptr, (const int32_t*)nullptr, (int32_t*)nullptr, n, 0, 32, st));
    BT_CUDA(tmp.alloc(std::max(tmp_bytes, tmp_bytes2), st));
    tmp_bytes = std::max(tmp_bytes, tmp_bytes2);
    BT_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys.as<uint32_t>(), keys_s.as<uint32_t>(), ids.as<int32_t>(), ids_s.as<int32_t>(), n, 0, key_bits, st));
    pd_gather_kernel<<<G, 256, 0, st>>>(d_pts, ids_s.as<int32_t>(), keys_s.as<uint32_t>(), n0, g, pos.as<float4>(), cb.as<int32_t>(), ce.as<int32_t>());
    BT_LAUNCH_CHECK();

    // removal-time weights (-1 = never removed), their sorted copy and the index permutation of the final selection
    BT_CUDA(fkeys.alloc(sizeof(float) * n0, st)); BT_CUDA(fkeys_s.alloc(sizeof(float) * n0, st));
    BT_CUDA(fidx.alloc(sizeof(int32_t) * n0, st)); BT_CUDA(fidx_s.alloc(sizeof(int32_t) * n0, st));
    pd_init_kernel<<<G, 256, 0, st>>>(n, fidx.as<int32_t>(), fkeys.as<float>());
    BT_LAUNCH_CHECK();
    const int64_t K = n0 - n_target;                       // removals of the serial algorithm
    int64_t removed = 0;
    float floor_w = -1.0f;                                 // K-th largest removal-time weight so far (a lower bound of the final one)
    int rounds = 0;
    while (true) {
        ++rounds;
        BT_REQUIRE(rounds <= 100000, "sample elimination does not converge");
        pd_weight_kernel<<<G, 256, 0, st>>>(g, pos.as<float4>(), alive.as<uint8_t>(), cb.as<int32_t>(), ce.as<int32_t>(), n, P, weight.as<float>());
        BT_LAUNCH_CHECK();
        BT_CUDA(cudaMemsetAsync(counter.p, 0, sizeof(int), st));
        pd_candidate_kernel<<<G, 256, 0, st>>>(g, pos.as<float4>(), alive.as<uint8_t>(), cb.as<int32_t>(), ce.as<int32_t>(), n, P, weight.as<float>(),
                                               floor_w, ids_s.as<int32_t>(), cand.as<uint8_t>(), counter.as<int>());
        BT_LAUNCH_CHECK();
        int n_cand = 0;
        BT_CUDA(cudaMemcpyAsync(&n_cand, counter.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        BT_CUDA(cudaStreamSynchronize(st));
        if (n_cand == 0) {
            BT_REQUIRE(removed >= K, "sample elimination found no candidate (NaN coordinates?)");
            break;                                         // no living sample can reach the K heaviest removals any more
        }
        pd_kill_kernel<<<G, 256, 0, st>>>(n, cand.as<uint8_t>(), weight.as<float>(), alive.as<uint8_t>(), fkeys.as<float>());
        BT_LAUNCH_CHECK();
        removed += n_cand;
        if (removed >= K) {
            BT_CUDA(cub::DeviceRadixSort::SortPairsDescending(tmp.p, tmp_bytes, fkeys.as<float>(), fkeys_s.as<float>(), fidx.as<int32_t>(), fidx_s.as<int32_t>(), n, 0, 32, st));
            BT_CUDA(cudaMemcpyAsync(&floor_w, fkeys_s.as<float>() + (K - 1), sizeof(float), cudaMemcpyDeviceToHost, st));
            BT_CUDA(cudaStreamSynchronize(st));
        }
    }
    // the K heaviest removals stand (the last sort saw every removal: no sample was removed after it), the rest come back
    pd_revive_kernel<<<G, 256, 0, st>>>(n, (int)K, fidx_s.as<int32_t>(), alive.as<uint8_t>());
    BT_LAUNCH_CHECK();
    pd_scatter_kernel<<<G, 256, 0, st>>>(n, ids_s.as<int32_t>(), alive.as<uint8_t>(), d_keep);
    BT_LAUNCH_CHECK();
    BT_CUDA(cudaStreamSynchronize(st));
    if (h_rounds) *h_rounds = rounds;
    return BT_OK;
}

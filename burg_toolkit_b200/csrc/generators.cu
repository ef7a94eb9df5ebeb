// generators.cu -- device-side producers/consumers of grasp sets that feed the hot path:
//   random_poses            (reference burg_toolkit/sampling.py:457-469)   distributional parity (different RNG)
//   grasp_perturbations     (sampling.py:400-454)                          deterministic, float32 rows
//   farthest_point_sampling (sampling.py:472-495)                          deterministic, index-exact
#include <math.h>

#include "common.cuh"

namespace bt {

// ---- random poses: uniform random rotations (normalised 4-d Gaussian quaternions), translations U[0,1)^3 -----------
__global__ void random_poses_kernel(int64_t n, uint64_t seed, double* __restrict__ out /* [n,4,4] */) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t a[4], b[4], c[4];
    philox4x32(seed, (uint64_t)i, 0x52504f53ull, a);
    philox4x32(seed, (uint64_t)i, 0x52504f54ull, b);
    philox4x32(seed, (uint64_t)i, 0x52504f55ull, c);
    const double tp = 2.0 * 3.14159265358979323846;
    const double r1 = sqrt(-2.0 * log(1.0 - u01(a[0], a[1]))), r2 = sqrt(-2.0 * log(1.0 - u01(a[2], a[3])));
    const double t1 = tp * u01(b[0], b[1]), t2 = tp * u01(b[2], b[3]);
    double x = r1 * cos(t1), y = r1 * sin(t1), z = r2 * cos(t2), w = r2 * sin(t2);
    const double inv = 1.0 / sqrt(x * x + y * y + z * z + w * w);
    x *= inv; y *= inv; z *= inv; w *= inv;
    double* m = out + 16 * i;
    m[0] = 1 - 2 * (y * y + z * z); m[1] = 2 * (x * y - z * w);     m[2] = 2 * (x * z + y * w);     m[3] = u01(c[0], c[1]);
    m[4] = 2 * (x * y + z * w);     m[5] = 1 - 2 * (x * x + z * z); m[6] = 2 * (y * z - x * w);     m[7] = u01(c[2], c[3]);
    m[8] = 2 * (x * z - y * w);     m[9] = 2 * (y * z + x * w);     m[10] = 1 - 2 * (x * x + y * y);
    uint32_t d[4];
    philox4x32(seed, (uint64_t)i, 0x52504f56ull, d);
    m[11] = u01(d[0], d[1]);
    m[12] = 0; m[13] = 0; m[14] = 0; m[15] = 1;
}

// ---- grasp perturbations --------------------------------------------------------------------------------------------
struct Radii { double r[16]; int n; };

// scipy Rotation.from_rotvec(v).as_matrix(): rotvec -> quaternion (x, y, z, w) -> matrix
__device__ __forceinline__ void rotvec_matrix(d3 v, double R[9]) {
    const double angle = norm3(v);
    double scale;
    if (angle <= 1e-3) { const double a2 = angle * angle; scale = 0.5 - a2 / 48.0 + a2 * a2 / 3840.0; }
    else scale = sin(angle / 2.0) / angle;
    const double x = scale * v.x, y = scale * v.y, z = scale * v.z, w = cos(angle / 2.0);
    const double x2 = x * x, y2 = y * y, z2 = z * z, w2 = w * w;
    const double xy = x * y, zw = z * w, xz = x * z, yw = y * w, yz = y * z, xw = x * w;
    R[0] = x2 - y2 - z2 + w2; R[1] = 2 * (xy - zw);       R[2] = 2 * (xz + yw);
    R[3] = 2 * (xy + zw);       R[4] = -x2 + y2 - z2 + w2; R[5] = 2 * (yz - xw);
    R[6] = 2 * (xz - yw);       R[7] = 2 * (yz + xw);       R[8] = -x2 - y2 + z2 + w2;
}

__global__ void perturbations_kernel(const float* __restrict__ gs, int64_t n, Radii radii, int include_original,
                                     float* __restrict__ out) {
    const int per = radii.n * 12 + (include_original ? 1 : 0);
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= n * per) return;
    const int64_t g = idx / per;
    int k = (int)(idx - g * per);
    const float* src = gs + 14 * g;
    float* dst = out + 14 * idx;
    if (include_original) {
        if (k == 0) {
#pragma unroll
            for (int c = 0; c < 14; ++c) dst[c] = src[c];
            return;
        }
        --k;
    }
    const int ri = k / 12, which = k % 12;          // 0..5 translations (axis 0 +,-, 1 +,-, 2 +,-), 6..11 rotations
    const double radius = radii.r[ri];
    double R[9], t[3];
#pragma unroll
    for (int c = 0; c < 9; ++c) R[c] = src[3 + c];
#pragma unroll
    for (int c = 0; c < 3; ++c) t[c] = src[c];
    const int axis_idx = (which % 6) / 2;
    const double sign = (which % 2 == 0) ? 1.0 : -1.0;
    const d3 axis = mk3(R[axis_idx], R[3 + axis_idx], R[6 + axis_idx]);          // column of the pose
    if (which < 6) {
        const double shift = radius / 1000;
        t[0] = t[0] + sign * shift * axis.x; t[1] = t[1] + sign * shift * axis.y; t[2] = t[2] + sign * shift * axis.z;
    } else {
        const double rot_rad = radius * (3.14159265358979323846 / 180.0);          // np.deg2rad
        double Q[9], N[9];
        rotvec_matrix(mk3(sign * rot_rad * axis.x, sign * rot_rad * axis.y, sign * rot_rad * axis.z), Q);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)           // 3x3 float64 matmul, ascending-k FMA chain like numpy's
                N[3 * i + j] = fma(Q[3 * i + 2], R[6 + j], fma(Q[3 * i + 1], R[3 + j], Q[3 * i] * R[j]));
#pragma unroll
        for (int c = 0; c < 9; ++c) R[c] = N[c];
    }
    dst[0] = (float)t[0]; dst[1] = (float)t[1]; dst[2] = (float)t[2];
#pragma unroll
    for (int c = 0; c < 9; ++c) dst[3 + c] = (float)R[c];
    dst[12] = 0.f; dst[13] = 0.f;
}

// ---- farthest point sampling ---------------------------------------------------------------------------------------
constexpr int FPS_T = 256;

template <typename T>
__global__ void fps_update_kernel(const T* __restrict__ pts, int stride, int64_t n, const int64_t* __restrict__ last,
                                  T* __restrict__ dist, T* __restrict__ blk_val, int64_t* __restrict__ blk_idx) {
    __shared__ T s_val[FPS_T];
    __shared__ int64_t s_idx[FPS_T];
    const int64_t i = blockIdx.x * (int64_t)FPS_T + threadIdx.x;
    T best = (T)-1;
    int64_t best_i = 0x7fffffffffffffffll;
    if (i < n) {
        const T* q = pts + (int64_t)stride * (*last);
        const T* p = pts + (int64_t)stride * i;
        const T dx = q[0] - p[0], dy = q[1] - p[1], dz = q[2] - p[2];
        const T d = (dx * dx + dy * dy) + dz * dz;                 // ((q - p) ** 2).sum(axis=1)
        const T cur = dist[i];
        const T m = (d < cur || cur != cur) ? d : cur;             // np.minimum (NaN propagates from dist only at start: inf)
        dist[i] = m;
        best = m; best_i = i;
    }
    s_val[threadIdx.x] = best; s_idx[threadIdx.x] = best_i;
    __syncthreads();
    for (int s = FPS_T / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            const T v = s_val[threadIdx.x + s];
            const int64_t j = s_idx[threadIdx.x + s];
            if (v > s_val[threadIdx.x] || (v == s_val[threadIdx.x] && j < s_idx[threadIdx.x])) { s_val[threadIdx.x] = v; s_idx[threadIdx.x] = j; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { blk_val[blockIdx.x] = s_val[0]; blk_idx[blockIdx.x] = s_idx[0]; }
}

template <typename T>
__global__ void fps_pick_kernel(const T* __restrict__ blk_val, const int64_t* __restrict__ blk_idx, int n_blocks,
                                int64_t* __restrict__ out, int it, int64_t* __restrict__ last) {
    __shared__ T s_val[FPS_T];
    __shared__ int64_t s_idx[FPS_T];
    T best = (T)-1;
    int64_t best_i = 0x7fffffffffffffffll;
    for (int b = threadIdx.x; b < n_blocks; b += FPS_T) {
        const T v = blk_val[b];
        const int64_t j = blk_idx[b];
        if (v > best || (v == best && j < best_i)) { best = v; best_i = j; }
    }
    s_val[threadIdx.x] = best; s_idx[threadIdx.x] = best_i;
    __syncthreads();
    for (int s = FPS_T / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            const T v = s_val[threadIdx.x + s];
            const int64_t j = s_idx[threadIdx.x + s];
            if (v > s_val[threadIdx.x] || (v == s_val[threadIdx.x] && j < s_idx[threadIdx.x])) { s_val[threadIdx.x] = v; s_idx[threadIdx.x] = j; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[it] = s_idx[0]; *last = s_idx[0]; }
}

template <typename T>
__global__ void fps_init_kernel(T* __restrict__ dist, int64_t n, int64_t* __restrict__ out, int64_t* __restrict__ last) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) dist[i] = (T)INFINITY;
    if (i == 0) { out[0] = 0; *last = 0; }          // first chosen point is index 0 (sampling.py:485)
}

template <typename T>
static int fps_run(const T* d_pts, int stride, int64_t n, int64_t k, int64_t* d_out, cudaStream_t st) {
    const int n_blocks = (int)ceil_div(n, FPS_T);
    Scratch dist, bv, bi, last;
    BT_CUDA(dist.alloc(sizeof(T) * n, st));
    BT_CUDA(bv.alloc(sizeof(T) * n_blocks, st));
    BT_CUDA(bi.alloc(sizeof(int64_t) * n_blocks, st));
    BT_CUDA(last.alloc(sizeof(int64_t), st));
    fps_init_kernel<T><<<n_blocks, FPS_T, 0, st>>>(dist.as<T>(), n, d_out, last.as<int64_t>());
    for (int64_t it = 1; it < k; ++it) {
        fps_update_kernel<T><<<n_blocks, FPS_T, 0, st>>>(d_pts, stride, n, last.as<int64_t>(), dist.as<T>(), bv.as<T>(), bi.as<int64_t>());
        fps_pick_kernel<T><<<1, FPS_T, 0, st>>>(bv.as<T>(), bi.as<int64_t>(), n_blocks, d_out, (int)it, last.as<int64_t>());
    }
    BT_LAUNCH_CHECK();
    return BT_OK;
}

}  // namespace bt

using namespace bt;

extern "C" int bt_random_poses(int64_t n, uint64_t seed, double* d_out, void* stream) {
    BT_REQUIRE(d_out && n >= 0, "bad arguments");
    if (n == 0) return BT_OK;
    random_poses_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(n, seed, d_out);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

extern "C" int bt_grasp_perturbations(const float* d_gs, int64_t n, const double* h_radii, int32_t n_radii,
                                      int32_t include_original, float* d_out, void* stream) {
    BT_REQUIRE(d_gs && h_radii && d_out, "null pointer");
    BT_REQUIRE(n >= 0 && n_radii >= 0 && n_radii <= 16, "0 <= n_radii <= 16");
    const int per = n_radii * 12 + (include_original ? 1 : 0);
    if (n == 0 || per == 0) return BT_OK;
    Radii r;
    r.n = n_radii;
    for (int i = 0; i < 16; ++i) r.r[i] = i < n_radii ? h_radii[i] : 0.0;
    perturbations_kernel<<<(unsigned)ceil_div(n * per, 256), 256, 0, (cudaStream_t)stream>>>(d_gs, n, r, include_original, d_out);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

extern "C" int bt_farthest_point_sampling(const void* d_points, int32_t is_float64, int32_t stride, int64_t n, int64_t k,
                                          int64_t* d_indices, void* stream) {
    BT_REQUIRE(d_points && d_indices, "null pointer");
    BT_REQUIRE(stride >= 3 && n >= 1 && k >= 1 && k <= n, "need stride >= 3 and 1 <= k <= n");
    cudaStream_t st = (cudaStream_t)stream;
    return is_float64 ? fps_run<double>((const double*)d_points, stride, n, k, d_indices, st)
                      : fps_run<float>((const float*)d_points, stride, n, k, d_indices, st);
}

// peer.cu -- peer-visible device buffers for the multi-GPU gather.
//
// The path shards by reference point; the only exchange is "compacted collision-free grasps -> rank 0".  Instead of a
// send/recv collective after the compaction, rank 0 owns one buffer with a slot per rank, every other rank maps it
// through CUDA IPC, and compact_scatter_kernel (collide.cu) stores its kept rows straight into its slot over
// NVLink/NVSwitch: the gather IS the scatter of the compaction, row by row, while the kernel runs.  One process per
// GPU (torchrun), so the mapping has to cross process boundaries: cudaIpcGetMemHandle on a plain cudaMalloc block
// (pool / stream-ordered memory cannot be exported), 64 opaque bytes that travel through torch.distributed.
#include "common.cuh"

namespace bt {
// One-sided completion protocol of the gather (no collective in the step).  Every rank, after its compaction kernel has
// stored its rows into its slot of the destination's buffer, publishes its count and then an epoch flag there; the
// destination polls the flags of all ranks.  A kernel's stores -- peer stores included -- are complete when the kernel is,
// so the flag written by the NEXT kernel on the same stream cannot overtake them (the fence orders count before flag).
__global__ void peer_signal_kernel(const int64_t* __restrict__ local_count, int64_t* count_slot, int64_t* flag_slot, int64_t epoch) {
    *reinterpret_cast<volatile int64_t*>(count_slot) = *local_count;
    __threadfence_system();
    *reinterpret_cast<volatile int64_t*>(flag_slot) = epoch;
}

// Staged form of the gather: the compaction has written this rank's kept rows into LOCAL memory (at HBM speed, inside the
// step's graph); this kernel -- on the rank's exchange stream, off the step's critical path -- streams exactly *d_n rows
// (56-byte grasp rows, then 48-byte contact records) into the rank's slot of the destination's buffer as 16-byte peer
// stores, grid-stride, with a few blocks only: it lives as long as the link needs and must not crowd the SMs.
// Source and destination are 16-byte aligned (bt_peer_push checks); both row sizes are multiples of 8 bytes.
__global__ void __launch_bounds__(256) peer_push_kernel(const int64_t* __restrict__ d_n, int64_t cap, const float* __restrict__ gs,
                                                        const double* __restrict__ contacts, float* dst_gs, double* dst_contacts) {
    const int64_t n = min(cap, *d_n);
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, stride = gridDim.x * (int64_t)blockDim.x;
    auto push = [&](const void* src, void* dst, int64_t bytes) {
        const uint4* s4 = reinterpret_cast<const uint4*>(src);
        uint4* d4 = reinterpret_cast<uint4*>(dst);
        const int64_t n4 = bytes >> 4;
        for (int64_t i = tid; i < n4; i += stride) {
            const uint4 v = __ldcs(s4 + i);
            asm volatile("st.global.v4.b32 [%0], {%1,%2,%3,%4};" :: "l"(d4 + i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        }
        if ((bytes & 8) && tid == 0) reinterpret_cast<uint2*>(dst)[2 * n4] = reinterpret_cast<const uint2*>(src)[2 * n4];
    };
    push(gs, dst_gs, n * 56);
    if (contacts && dst_contacts) push(contacts, dst_contacts, n * 48);
}

// lane r polls word r until it has reached `epoch`; gives up after `timeout_ns` (status[0] = 1) so that a lost peer
// can never hang the device.  With `ack` the waiter then publishes that it has seen (consumed) `ack_value`.
__global__ void peer_wait_kernel(const int64_t* words, int n_words, int64_t epoch, unsigned long long timeout_ns,
                                 int32_t* status, int64_t* ack, int64_t ack_value) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    bool ok = true;
    for (int r = threadIdx.x; r < n_words; r += blockDim.x) {
        while (*reinterpret_cast<const volatile int64_t*>(words + r) < epoch) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            if (t - t0 > timeout_ns) { ok = false; break; }
            __nanosleep(200);
        }
    }
    if (!ok) *reinterpret_cast<volatile int32_t*>(status) = 1;       // (device or mapped pinned host memory)
    __syncthreads();
    __threadfence_system();
    if (threadIdx.x == 0 && ack) *reinterpret_cast<volatile int64_t*>(ack) = ack_value;
}
}  // namespace bt

using namespace bt;

static_assert(sizeof(cudaIpcMemHandle_t) == BT_IPC_HANDLE_BYTES, "bt_ipc_handle must hold a cudaIpcMemHandle_t");

extern "C" int bt_peer_alloc(int device, int64_t bytes, void** d_ptr, bt_ipc_handle* handle) {
    BT_REQUIRE(d_ptr && handle, "null pointer");
    BT_REQUIRE(bytes > 0, "size must be positive");
    BT_CUDA(cudaSetDevice(device));
    void* p = nullptr;
    BT_CUDA(cudaMalloc(&p, (size_t)bytes));
    cudaError_t e = cudaMemset(p, 0, (size_t)bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return cuda_fail(e, "cudaIpcGetMemHandle", __FILE__, __LINE__); }
    memcpy(handle->bytes, &h, sizeof(h));
    *d_ptr = p;
    return BT_OK;
}

extern "C" int bt_peer_open(int device, const bt_ipc_handle* handle, void** d_ptr) {
    BT_REQUIRE(d_ptr && handle, "null pointer");
    BT_CUDA(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle->bytes, sizeof(h));
    BT_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return BT_OK;
}

extern "C" int bt_peer_close(int device, void* d_ptr) {
    if (!d_ptr) return BT_OK;
    BT_CUDA(cudaSetDevice(device));
    BT_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return BT_OK;
}

extern "C" int bt_peer_free(int device, void* d_ptr) {
    if (!d_ptr) return BT_OK;
    BT_CUDA(cudaSetDevice(device));
    BT_CUDA(cudaFree(d_ptr));
    return BT_OK;
}

extern "C" int bt_peer_signal(const int64_t* d_local_count, int64_t* d_count_slot, int64_t* d_flag_slot, int64_t epoch, void* stream) {
    BT_REQUIRE(d_local_count && d_count_slot && d_flag_slot, "null pointer");
    peer_signal_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(d_local_count, d_count_slot, d_flag_slot, epoch);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

extern "C" int bt_peer_push(const int64_t* d_n, int64_t cap_rows, const float* d_gs, const double* d_contacts, float* d_dst_gs,
                            double* d_dst_contacts, int32_t n_blocks, void* stream) {
    BT_REQUIRE(d_n && d_gs && d_dst_gs, "null pointer");
    BT_REQUIRE(cap_rows >= 0 && n_blocks >= 1 && n_blocks <= 4096, "bad arguments");
    BT_REQUIRE(((reinterpret_cast<uintptr_t>(d_gs) | reinterpret_cast<uintptr_t>(d_dst_gs) | reinterpret_cast<uintptr_t>(d_contacts) |
                 reinterpret_cast<uintptr_t>(d_dst_contacts)) & 15) == 0, "buffers must be 16-byte aligned");
    if (cap_rows == 0) return BT_OK;
    peer_push_kernel<<<(unsigned)n_blocks, 256, 0, (cudaStream_t)stream>>>(d_n, cap_rows, d_gs, d_contacts, d_dst_gs, d_dst_contacts);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

extern "C" int bt_peer_copy(void* d_dst, const void* d_src, int64_t bytes, void* stream) {
    BT_REQUIRE(d_dst && d_src && bytes >= 0, "bad arguments");
    if (bytes == 0) return BT_OK;
    BT_CUDA(cudaMemcpyAsync(d_dst, d_src, (size_t)bytes, cudaMemcpyDefault, (cudaStream_t)stream));     // peer destination: a copy engine over NVLink, no SM
    return BT_OK;
}

extern "C" int bt_peer_wait(const int64_t* d_words, int32_t n_words, int64_t epoch, double timeout_s, int32_t* d_status,
                            int64_t* d_ack, int64_t ack_value, void* stream) {
    BT_REQUIRE(d_words && d_status && n_words >= 1, "bad arguments");
    BT_REQUIRE(timeout_s > 0.0 && timeout_s <= 60.0, "timeout must be in (0, 60] seconds");
    peer_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(d_words, n_words, epoch, (unsigned long long)(timeout_s * 1e9), d_status, d_ack, ack_value);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

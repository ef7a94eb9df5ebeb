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

extern "C" int bt_peer_wait(const int64_t* d_words, int32_t n_words, int64_t epoch, double timeout_s, int32_t* d_status,
                            int64_t* d_ack, int64_t ack_value, void* stream) {
    BT_REQUIRE(d_words && d_status && n_words >= 1, "bad arguments");
    BT_REQUIRE(timeout_s > 0.0 && timeout_s <= 60.0, "timeout must be in (0, 60] seconds");
    peer_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(d_words, n_words, epoch, (unsigned long long)(timeout_s * 1e9), d_status, d_ack, ack_value);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

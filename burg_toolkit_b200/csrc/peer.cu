// peer.cu -- peer-visible device buffers for the multi-GPU gather.
//
// The path shards by reference point; the only exchange is "compacted collision-free grasps -> rank 0".  Instead of a
// send/recv collective after the compaction, rank 0 owns one buffer with a slot per rank, every other rank maps it
// through CUDA IPC, and compact_scatter_kernel (collide.cu) stores its kept rows straight into its slot over
// NVLink/NVSwitch: the gather IS the scatter of the compaction, row by row, while the kernel runs.  One process per
// GPU (torchrun), so the mapping has to cross process boundaries: cudaIpcGetMemHandle on a plain cudaMalloc block
// (pool / stream-ordered memory cannot be exported), 64 opaque bytes that travel through torch.distributed.
#include "common.cuh"

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

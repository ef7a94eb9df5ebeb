// pipeline.cu -- the whole AGS batch pipeline behind one C-ABI call, plus per-stage CUDA-event timers.
//
// Replaces the body of the reference's hot loop (burg_toolkit/sampling.py:220-350) followed by check_collisions
// (:354-397) for a batch of reference points: the host enqueues ~15 launches on one stream and returns; nothing is
// copied to or from the host and no stage waits for the CPU.
#include "common.cuh"

#include <algorithm>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

struct bt_timer {
    int device = 0;
    cudaEvent_t ev[BT_N_STAGES + 1][2];
    bool ran[BT_N_STAGES];
};

namespace bt {
thread_local Arena* g_arena = nullptr;
struct ArenaScope {               // installs the call's scratch arena for the duration of bt_ags_pipeline
    Arena a;
    ArenaScope(void* base, int64_t bytes) {
        if (base && bytes > 0) { a.base = static_cast<char*>(base); a.size = (size_t)bytes; g_arena = &a; }
    }
    ~ArenaScope() { g_arena = nullptr; }
};
// number of valid rows inside the range [row0, row0 + n_cap) of a list whose valid length is *d_n
__global__ void range_count_kernel(const int64_t* __restrict__ d_n, int64_t row0, int64_t n_cap, int64_t* __restrict__ out) {
    const int64_t v = *d_n - row0;
    *out = v < 0 ? 0 : (v > n_cap ? n_cap : v);
}

// Side stream, events and counters of the split tail (cfg->tail_splits > 1): the posed grasps are cut into row ranges and
// range s+1 is collision-checked on the other stream while range s is being compacted -- which matters when the
// compaction stores over PCIe (mapped host memory) or NVLink and is bound by the link, not by the SMs.
struct TailState {
    cudaStream_t side = nullptr;
    cudaEvent_t ready = nullptr, done[2] = {nullptr, nullptr};
    int64_t* counts = nullptr;            // [2 * BT_MAX_TAIL_SPLITS]: running kept-row counts, then valid rows per range
};
constexpr int BT_MAX_TAIL_SPLITS = 8;
// one state per (device, caller stream): two pipelines in flight on different streams of a device must not share the side
// stream, the events or the running counts
static std::map<std::pair<int, cudaStream_t>, TailState> g_tail;
static std::mutex g_tail_mutex;

static int tail_state(int device, cudaStream_t caller, TailState** out) {
    BT_REQUIRE(device >= 0 && device < 64, "device index out of range");
    std::lock_guard<std::mutex> lock(g_tail_mutex);
    TailState& t = g_tail[std::make_pair(device, caller)];
    if (!t.side) {
        BT_CUDA(cudaStreamCreateWithFlags(&t.side, cudaStreamNonBlocking));
        BT_CUDA(cudaEventCreateWithFlags(&t.ready, cudaEventDisableTiming));
        for (int k = 0; k < 2; ++k) BT_CUDA(cudaEventCreateWithFlags(&t.done[k], cudaEventDisableTiming));
        BT_CUDA(cudaMalloc(&t.counts, sizeof(int64_t) * 2 * BT_MAX_TAIL_SPLITS));
    }
    *out = &t;
    return BT_OK;
}
}  // namespace bt

using namespace bt;

extern "C" int bt_timer_create(int device, bt_timer** out) {
    BT_REQUIRE(out, "null pointer");
    BT_CUDA(cudaSetDevice(device));
    bt_timer* t = new bt_timer();
    t->device = device;
    for (int i = 0; i < BT_N_STAGES; ++i) {
        t->ran[i] = false;
        for (int k = 0; k < 2; ++k) {
            cudaError_t e = cudaEventCreate(&t->ev[i][k]);
            if (e != cudaSuccess) { delete t; return cuda_fail(e, "cudaEventCreate", __FILE__, __LINE__); }
        }
    }
    *out = t;
    return BT_OK;
}

extern "C" int bt_timer_destroy(bt_timer* t) {
    if (!t) return BT_OK;
    cudaSetDevice(t->device);
    for (int i = 0; i < BT_N_STAGES; ++i) for (int k = 0; k < 2; ++k) cudaEventDestroy(t->ev[i][k]);
    delete t;
    return BT_OK;
}

extern "C" int bt_timer_read(bt_timer* t, float* h_stage_ms) {
    BT_REQUIRE(t && h_stage_ms, "null pointer");
    BT_CUDA(cudaSetDevice(t->device));
    for (int i = 0; i < BT_N_STAGES; ++i) {
        h_stage_ms[i] = -1.0f;
        if (!t->ran[i]) continue;
        BT_CUDA(cudaEventSynchronize(t->ev[i][1]));
        BT_CUDA(cudaEventElapsedTime(&h_stage_ms[i], t->ev[i][0], t->ev[i][1]));
    }
    return BT_OK;
}

#define STAGE_BEGIN(i) do { if (timer) { timer->ran[i] = true; BT_CUDA(cudaEventRecord(timer->ev[i][0], st)); } } while (0)
#define STAGE_END(i)   do { if (timer) BT_CUDA(cudaEventRecord(timer->ev[i][1], st)); } while (0)
#define STAGE_CALL(expr) do { int _rc = (expr); if (_rc != BT_OK) return _rc; } while (0)

extern "C" int bt_ags_pipeline(const bt_mesh* target, const bt_mesh* scene, const bt_gripper* gripper,
                               const bt_pipeline_config* cfg, const bt_pipeline_buffers* b, int64_t n_ref,
                               uint64_t seed, bt_timer* timer, void* stream) {
    BT_REQUIRE(target && cfg && b, "null pointer");
    BT_REQUIRE(n_ref >= 0, "negative count");
    BT_REQUIRE(b->ref_pts && b->dirs && b->hit_count && b->overflow && b->pair_count && b->pair_offset &&
               b->pair_ref && b->pair_q && b->gs && b->n_grasps, "missing buffer");
    BT_REQUIRE(b->hits || (b->hit_t && b->valid_mask), "either hit records or the lean pair (hit_t, valid_mask) is required");
    BT_REQUIRE(b->ref_nrm || (!b->sample_surface && !b->generate_dirs), "reference normals required to generate rays");
    BT_REQUIRE(cfg->halfspace || b->frame_vecs, "frame vector buffer required unless halfspace");
    BT_REQUIRE(b->cap_pairs >= n_ref * (int64_t)cfg->max_targets, "cap_pairs too small");
    BT_REQUIRE(!cfg->check_collisions || (scene && gripper && b->colliding), "collision stage needs scene, gripper and mask buffer");
    BT_REQUIRE(!cfg->compact || (cfg->check_collisions && b->free_gs && b->n_free), "compaction needs the collision stage and its buffers");
    BT_REQUIRE(cfg->surface_total == 0 || (cfg->surface_total > 0 && b->seed_add && b->sample_surface), "surface_total needs sample_surface and the device-side parameters (seed_add)");
    cudaStream_t st = (cudaStream_t)stream;
    BT_CUDA(cudaSetDevice(target->m.device));
    BT_REQUIRE((reinterpret_cast<uintptr_t>(b->scratch) & 255) == 0, "scratch must be 256-byte aligned");
    ArenaScope arena(b->scratch, b->scratch_bytes);
    if (timer) for (int i = 0; i < BT_N_STAGES; ++i) timer->ran[i] = false;
    const int n_o = cfg->n_orientations;

    // S1 + S2 + frame vectors: whatever the caller did not supply, in one launch
    const int gen_frame = (b->generate_frame_vecs && !cfg->halfspace) ? 1 : 0;
    if (b->sample_surface || b->generate_dirs || gen_frame) {
        STAGE_BEGIN(0);
        STAGE_CALL(generate_impl(target, n_ref, cfg->n_rays, cfg->ags.angle_max, seed, b->seed_add, cfg->surface_total, b->sample_surface, b->generate_dirs,
                                 gen_frame, b->ref_pts, b->ref_nrm, b->frame_vecs, b->dirs, stream));
        STAGE_END(0);
    }
    STAGE_BEGIN(2);
    STAGE_CALL(ags_cast_impl(target, b->ref_pts, b->dirs, n_ref, cfg->n_rays, &cfg->ags, b->hit_count, b->hits, b->hits ? nullptr : b->hit_t,
                             b->valid_mask, b->overflow, !b->hits && cfg->cone_culling, stream));
    STAGE_END(2);
    STAGE_BEGIN(3);
    STAGE_CALL(ags_select_fused_impl(b->hit_count, b->hits, b->hit_t, b->ref_pts, b->dirs, b->valid_mask, n_ref, cfg->n_rays,
                                     cfg->ags.max_hits_per_ray, cfg->max_targets,
                                     seed ? seed + 0x2000 : 0 /* seed 0: first valid hit in canonical order */, b->seed_add, b->pair_count,
                                     b->pair_offset, b->pair_ref, b->pair_q, b->cap_pairs, stream));
    STAGE_END(3);
    STAGE_BEGIN(5);
    STAGE_CALL(expand_impl(b->ref_pts, b->pair_offset, b->pair_ref, b->pair_q, b->frame_vecs, n_ref, b->cap_pairs, n_o, cfg->halfspace,
                           cfg->orient_cos, cfg->orient_sin, b->gs, b->contacts, b->n_grasps, stream));
    STAGE_END(5);
    if (cfg->check_collisions) {
        const int64_t cap_rows = b->cap_pairs * n_o;
        const int splits = (cfg->compact && !timer) ? std::min(std::max(cfg->tail_splits, 1), BT_MAX_TAIL_SPLITS) : 1;
        const int64_t rows_per = ceil_div(ceil_div(cap_rows, splits), 256) * 256;      // ranges start on compaction-block boundaries
        if (splits == 1 || rows_per >= cap_rows) {
            STAGE_BEGIN(6);
            STAGE_CALL(bt_collide(scene, gripper, b->gs, cap_rows, b->n_grasps, cfg->ags.opening_width,
                                  cfg->collision_width_tolerance, cfg->use_width, b->colliding, stream));
            STAGE_END(6);
            if (cfg->compact) {
                if (b->compact_after) BT_CUDA(cudaStreamWaitEvent(st, (cudaEvent_t)b->compact_after, 0));
                STAGE_BEGIN(7);
                STAGE_CALL(bt_compact(b->colliding, 0, b->gs, b->contacts, cap_rows, b->n_grasps, b->free_gs,
                                      b->contacts ? b->free_contacts : nullptr, b->n_free, b->free_base, stream));
                STAGE_END(7);
            }
        } else {
            // split tail: collide(range s+1) on one stream overlaps compact(range s) on the other; every compaction appends
            // behind the previous range's rows (device-side running count), so the output is the same contiguous list
            TailState* t = nullptr;
            STAGE_CALL(tail_state(target->m.device, st, &t));
            BT_CUDA(cudaEventRecord(t->ready, st));
            BT_CUDA(cudaStreamWaitEvent(t->side, t->ready, 0));
            int last = 0;
            for (int s = 0; s * rows_per < cap_rows; ++s) {
                const int64_t r0 = s * rows_per, n_s = std::min(rows_per, cap_rows - r0);
                cudaStream_t x = (s & 1) ? t->side : st;
                int64_t* range_n = t->counts + BT_MAX_TAIL_SPLITS + s;
                const bool final_range = (s + 1) * rows_per >= cap_rows;
                range_count_kernel<<<1, 1, 0, x>>>(b->n_grasps, r0, n_s, range_n);
                BT_LAUNCH_CHECK();
                STAGE_CALL(bt_collide(scene, gripper, b->gs + 14 * r0, n_s, range_n, cfg->ags.opening_width,
                                      cfg->collision_width_tolerance, cfg->use_width, b->colliding + r0, x));
                if (s > 0) BT_CUDA(cudaStreamWaitEvent(x, t->done[(s - 1) & 1], 0));
                else if (b->compact_after) BT_CUDA(cudaStreamWaitEvent(x, (cudaEvent_t)b->compact_after, 0));
                STAGE_CALL(bt_compact(b->colliding + r0, 0, b->gs + 14 * r0, b->contacts ? b->contacts + 6 * r0 : nullptr, n_s, range_n,
                                      b->free_gs, b->contacts ? b->free_contacts : nullptr, final_range ? b->n_free : t->counts + s,
                                      s > 0 ? t->counts + s - 1 : b->free_base, x));
                BT_CUDA(cudaEventRecord(t->done[s & 1], x));
                last = s;
            }
            if (last & 1) BT_CUDA(cudaStreamWaitEvent(st, t->done[1], 0));      // the caller's stream owns the result again
        }
    }
    return BT_OK;
}

// ---- CUDA graphs: capture a sequence of library calls on a stream once, replay it with one launch -------------------------
// bt_ags_pipeline enqueues ~10 kernels, a handful of memsets and stream-ordered scratch allocations: ~0.2 ms of host time
// per call, which is what a small batch (config 1) and the end-to-end call pay.  Captured into a graph the same work is one
// cudaGraphLaunch (~10 us); the seed of a replay comes from the device (bt_pipeline_buffers.seed_add, set with bt_set_u64).
struct bt_graph {
    int device = 0;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
};

namespace bt {
__global__ void set_u64_kernel(uint64_t* p, uint64_t v) { *p = v; }
__global__ void set_u64x3_kernel(uint64_t* p, uint64_t a, uint64_t b, uint64_t c) { p[0] = a; p[1] = b; p[2] = c; }
}

extern "C" int bt_set_u64x3(uint64_t* d_ptr, uint64_t v0, uint64_t v1, uint64_t v2, void* stream) {
    BT_REQUIRE(d_ptr, "null pointer");
    bt::set_u64x3_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(d_ptr, v0, v1, v2);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

extern "C" int bt_set_u64(uint64_t* d_ptr, uint64_t value, void* stream) {
    BT_REQUIRE(d_ptr, "null pointer");
    bt::set_u64_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(d_ptr, value);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

extern "C" int bt_graph_begin_capture(void* stream) {
    BT_REQUIRE(stream, "capture needs a non-default stream");
    bt::retain_pool_memory();
    BT_CUDA(cudaStreamBeginCapture((cudaStream_t)stream, cudaStreamCaptureModeThreadLocal));
    return BT_OK;
}

extern "C" int bt_graph_end_capture(void* stream, bt_graph** out) {
    BT_REQUIRE(stream && out, "null pointer");
    cudaGraph_t g = nullptr;
    BT_CUDA(cudaStreamEndCapture((cudaStream_t)stream, &g));
    BT_REQUIRE(g, "the capture produced no graph (a call failed while capturing)");
    bt_graph* h = new bt_graph();
    cudaGetDevice(&h->device);
    h->graph = g;
    cudaError_t e = cudaGraphInstantiate(&h->exec, g, 0);
    if (e != cudaSuccess) { cudaGraphDestroy(g); delete h; return bt::cuda_fail(e, "cudaGraphInstantiate", __FILE__, __LINE__); }
    *out = h;
    return BT_OK;
}

extern "C" int bt_graph_launch(bt_graph* g, void* stream) {
    BT_REQUIRE(g && g->exec, "null graph");
    BT_CUDA(cudaGraphLaunch(g->exec, (cudaStream_t)stream));
    return BT_OK;
}

extern "C" int bt_graph_destroy(bt_graph* g) {
    if (!g) return BT_OK;
    cudaSetDevice(g->device);
    if (g->exec) cudaGraphExecDestroy(g->exec);
    if (g->graph) cudaGraphDestroy(g->graph);
    delete g;
    return BT_OK;
}

// what a graph holds: number of nodes, number of kernel nodes, and the kernels' names ('\n'-separated, in node order) --
// bench.py counts its launches with this instead of keeping a list by hand
extern "C" int bt_graph_info(const bt_graph* g, int32_t* n_nodes, int32_t* n_kernel_nodes, char* names, int64_t names_cap) {
    BT_REQUIRE(g && g->graph, "null graph");
    size_t n = 0;
    BT_CUDA(cudaGraphGetNodes(g->graph, nullptr, &n));
    std::vector<cudaGraphNode_t> nodes(n);
    if (n) BT_CUDA(cudaGraphGetNodes(g->graph, nodes.data(), &n));
    int kernels = 0;
    int64_t used = 0;
    if (names && names_cap > 0) names[0] = 0;
    for (size_t i = 0; i < n; ++i) {
        cudaGraphNodeType type;
        BT_CUDA(cudaGraphNodeGetType(nodes[i], &type));
        if (type != cudaGraphNodeTypeKernel) continue;
        ++kernels;
        cudaKernelNodeParams kp;
        const char* nm = "?";
        if (cudaGraphKernelNodeGetParams(nodes[i], &kp) == cudaSuccess && kp.func) {
            const char* raw = nullptr;
            if (cudaFuncGetName(&raw, kp.func) == cudaSuccess && raw) nm = raw;
        } else cudaGetLastError();
        if (names && names_cap > 0) {
            const int64_t len = (int64_t)strlen(nm);
            if (used + len + 2 <= names_cap) { memcpy(names + used, nm, len); names[used + len] = '\n'; used += len + 1; names[used] = 0; }
        }
    }
    if (n_nodes) *n_nodes = (int32_t)n;
    if (n_kernel_nodes) *n_kernel_nodes = kernels;
    return BT_OK;
}


/*
 * burg_b200.h -- C ABI of libburgb200.so: the B200 (sm_100a) implementation of burg_toolkit's
 * antipodal-grasp-sampling and grasp-set-metric hot path.
 *
 * The reference (mrudorfer/burg-toolkit) is pure Python and has no FFI of its own; its boundary for
 * this path is the Python API (burg_toolkit/sampling.py:51-397, burg_toolkit/metrics.py:9-229).  Each
 * entry point below names the reference interface whose arithmetic it replaces; the Python mirror
 * (burg_toolkit_b200/sampling.py, metrics.py) binds them with ctypes (see INTEGRATION.md for the stub a
 * reference maintainer would add).
 *
 * Conventions
 *   - every function returns 0 on success, <0 on failure (BT_E_*); bt_last_error() gives the message
 *     (thread-local).  No function ever falls back to the CPU.
 *   - pointers named h_* are HOST memory, d_* are DEVICE memory on the mesh's / current device; plain
 *     C types only.  Arrays are row-major and densely packed.
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).  Calls are
 *     asynchronous w.r.t. the host unless stated ("synchronises").
 *   - scratch memory comes from the stream-ordered allocator (cudaMallocAsync) on `stream`.
 */
#ifndef BURG_B200_H
#define BURG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define BT_API __attribute__((visibility("default")))
#else
#define BT_API
#endif

#define BT_OK            0
#define BT_E_INVALID    -1   /* bad argument                                   */
#define BT_E_CUDA       -2   /* CUDA runtime error (message has the details)   */
#define BT_E_CAPACITY   -3   /* an output capacity was exceeded                */
#define BT_E_NODEVICE   -4   /* no CUDA device / wrong architecture            */

/* hit flags written by bt_ags_cast: the FIRST mask of the reference's sequence (sampling.py:239-319) that rejects the hit;
 * a hit is a valid antipodal target iff flags == 0.  nrm / angle of hits rejected by the origin or width mask are 0 / NaN
 * (the reference never computes them). */
#define BT_HIT_ORIGIN    1u  /* sampling.py:239   hit coincides with the reference point          */
#define BT_HIT_WIDTH     2u  /* sampling.py:249   fails (d <= ow - tol) | (d >= min_width)  (an OR) */
#define BT_HIT_SIGN      4u  /* sampling.py:295   sign(d . n) <= 0                                 */
#define BT_HIT_CONE      8u  /* sampling.py:305   angle > atan(mu)                                 */
#define BT_HIT_BELOW_Z  16u  /* sampling.py:316   z <= no_contact_below_z                          */

typedef struct bt_mesh bt_mesh;     /* device-resident triangle soup + LBVH (one object, or a merged scene) */

/* One ray/triangle crossing as the reference sees it after trimesh's intersects_location
 * (sampling.py:231-232) plus the per-hit quantities of sampling.py:239-319.  72 bytes. */
typedef struct bt_hit {
    double   loc[3];    /* intersection location                                  */
    double   nrm[3];    /* interpolated vertex normal (mesh_processing.py:110-153) */
    double   t;         /* ray parameter  proj_ori / proj_dir                      */
    double   angle;     /* atan(|d x n| / d . n)   (util.py:46-82), radians        */
    int32_t  tri;       /* triangle index in the caller's face array               */
    uint32_t flags;     /* BT_HIT_* bits                                           */
} bt_hit;

typedef struct bt_ags_params {      /* AntipodalGraspSampler attributes, sampling.py:57-70 */
    double  angle_max;              /* np.arctan(mu), computed by the caller (sampling.py:211)   */
    double  opening_width;          /* gripper.opening_width                                      */
    double  width_tolerance;        /* default 0.005                                              */
    double  min_grasp_width;        /* default 0.002                                              */
    double  no_contact_below_z;     /* used iff use_below_z                                       */
    int32_t use_below_z;
    int32_t max_hits_per_ray;       /* capacity of the per-ray hit list (>=1, <= 32)              */
} bt_ags_params;

typedef struct bt_mesh_info {
    int64_t n_vertices, n_triangles, n_nodes;
    int64_t device_bytes;
    double  surface_area;
    double  build_ms;               /* LBVH build time (CUDA events)                              */
    float   bounds_min[3], bounds_max[3];
    int32_t device;
    int32_t max_depth;              /* deepest leaf of the LBVH                                   */
} bt_mesh_info;

typedef struct bt_gripper_part {    /* a group of gripper triangles and their TCP-frame AABB */
    double  box_min[3], box_max[3];
    int32_t tri_begin, tri_end;     /* [begin, end) into the gripper triangle array */
} bt_gripper_part;

/* ---- library ------------------------------------------------------------------------------------ */
BT_API int         bt_version(void);
BT_API const char* bt_last_error(void);
BT_API int         bt_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, int64_t* total_mem);
BT_API int         bt_malloc(int device, int64_t bytes, void** d_ptr);   /* plain cudaMalloc for non-torch callers */
BT_API int         bt_free(void* d_ptr);
BT_API int         bt_memcpy_h2d(void* d_dst, const void* h_src, int64_t bytes, void* stream);
BT_API int         bt_memcpy_d2h(void* h_dst, const void* d_src, int64_t bytes, void* stream);
BT_API int         bt_stream_sync(void* stream);

/* instrumentation: install (or remove, with NULL) a device array of 8 uint64 counters; while installed, bt_ags_cast and
 * bt_collide run instrumented instantiations of their kernels that add to it:
 *   [0] cast LBVH nodes visited  [1] cast float32 leaf pre-tests  [2] cast exact float64 tests
 *   [3] collide LBVH nodes visited  [4] collide leaves classified  [5] collide leaves sent to the exact test
 *   [6] cast LBVH nodes fetched but skipped by the normal-cone test
 * (used by bench.py, outside the timed region, for the L2 traffic model of the roofline) */
BT_API int bt_debug_set_counters(uint64_t* d_counters);
/* measured L2 read bandwidth [GB/s]: all SMs stream one L2-resident buffer of `bytes` (1 MiB ..) `iters` times (synchronises).
 * The mesh + LBVH of the ray caster live in L2, so this is the memory roof of its traffic model (bench.py roofline.binding) */
BT_API int bt_debug_l2_probe(int device, int64_t bytes, int32_t iters, double* gbytes_per_s);

/* ---- mesh / scene ---------------------------------------------------------------------------------
 * replaces mesh_processing.as_trimesh + trimesh.Trimesh construction + RayMeshIntersector / collision
 * manager set-up (sampling.py:191-192, 372-380).  h_VN may be NULL (area-weighted vertex normals are
 * then computed, the io.load_mesh convention io.py:96).  Face normals are always recomputed as
 * normalise(cross(v1-v0, v2-v0)) (trimesh).  Builds the LBVH; synchronises. */
BT_API int bt_mesh_create(const double* h_V, int64_t n_vertices, const int32_t* h_F, int64_t n_triangles,
                   const double* h_VN, int device, bt_mesh** out);
BT_API int bt_mesh_destroy(bt_mesh* mesh);
BT_API int bt_mesh_get_info(const bt_mesh* mesh, bt_mesh_info* info);
/* debug/validation: copy the LBVH out (nodes: 16 floats each; leaf_tri: sorted leaf -> caller's triangle id) */
BT_API int bt_mesh_copy_bvh(const bt_mesh* mesh, float* h_nodes, int32_t* h_leaf_tri);
/* Topologies that bt_mesh_create builds from now on (process-wide).  The library keeps two trees per mesh: the RAY tree under
 * the 4-wide nodes of the ray caster and the WALK tree of the binary nodes (collision walks, mesh pairs).
 *   1 (default)  ray tree by PLOC (bottom-up clustering of the Morton order by merged surface area: ~12 % fewer node fetches
 *                per ray); walk tree = Karras' Morton hierarchy below 2^18 triangles (shallower: those walks are bound by
 *                their longest chain of passes), PLOC from there on (fewer fetches win once the scene outgrows L2)
 *   0            Karras' hierarchy of the Morton codes for both (round 1; also the automatic fall-back for a PLOC tree deeper
 *                than the traversal stacks)
 *   2            PLOC for both
 * Results never depend on the trees -- they are the broad phase of trimesh's intersects_location / FCL's collide
 * (sampling.py:231, :354-397).  The environment variable BT_BVH_BUILD=lbvh|ploc|ploc_all sets the initial value. */
BT_API int bt_set_bvh_build(int32_t kind);
/* what this mesh was built with: ray_kind / walk_kind (0 = Morton hierarchy, 1 = PLOC; walk_kind may be NULL);
 * build_iterations (may be NULL): merge rounds of the last PLOC build */
BT_API int bt_mesh_bvh_kind(const bt_mesh* mesh, int32_t* ray_kind, int32_t* walk_kind, int32_t* build_iterations);

/* ---- S1: area-weighted surface samples (replaces mesh_processing.poisson_disk_sampling's uniform
 * stage, mesh_processing.py:52-83 -> open3d SamplePointsUniformlyImpl).  d_pts/d_nrm: [n,3] f64,
 * d_tri: [n] i32 (may be NULL). */
BT_API int bt_sample_surface(const bt_mesh* mesh, int64_t n, uint64_t seed, double* d_pts, double* d_nrm,
                      int32_t* d_tri, void* stream);

/* ---- S1, second stage: weighted sample elimination (replaces the elimination loop of open3d's
 * TriangleMesh::SamplePointsPoissonDisk behind mesh_processing.poisson_disk_sampling, mesh_processing.py:77-81; Yuksel 2015
 * with open3d's alpha = 8, beta = 0.5, gamma = 1.5).  d_pts [n0,3] f64: the init_factor * n uniform samples of
 * bt_sample_surface; d_keep [n0] u8 out: 1 for the n_target samples that stay.  Parallel rounds (every sample that outweighs
 * all its living neighbours goes at once) instead of open3d's serial priority queue; h_rounds (nullable, host): rounds taken.
 * Synchronises. */
BT_API int bt_poisson_disk_eliminate(const double* d_pts, int64_t n0, int64_t n_target, double surface_area, double alpha,
                                     double beta, double gamma, uint8_t* d_keep, int32_t* h_rounds, void* stream);

/* ---- S2: friction-cone ray directions (replaces sampling.rays_within_cone, sampling.py:17-48, for
 * axis = -normal).  d_normals: [n_ref,3] f64, d_dirs: [n_ref,n_rays,3] f64. */
BT_API int bt_cone_rays(const double* d_normals, int64_t n_ref, int32_t n_rays, double angle, uint64_t seed,
                 double* d_dirs, void* stream);
/* the same with the reference's `uniform_on_plane` option (sampling.py:33-34): inclination = arctan(U(0, tan(angle))),
 * i.e. the rays are uniform on the ground plane of the cone instead of uniform in the angle */
BT_API int bt_cone_rays_ex(const double* d_normals, int64_t n_ref, int32_t n_rays, double angle, int32_t uniform_on_plane,
                 uint64_t seed, double* d_dirs, void* stream);

/* random unit vectors (util.generate_random_unit_vector, util.py:32-43): normalised N(0,1)^3; [n,3] f64 */
BT_API int bt_random_unit_vectors(int64_t n, uint64_t seed, double* d_out, void* stream);

/* ---- R1..R6: cast every ray of every reference point against the mesh, ALL crossings, and evaluate the
 * antipodal filters per hit (replaces intersector.intersects_location + sampling.py:239-319).
 *   d_ref_pts [n_ref,3] f64, d_dirs [n_ref,n_rays,3] f64
 *   d_hit_count [n_ref*n_rays] i32      number of hits of each ray (after trimesh's 1e-8 de-duplication)
 *   d_hits      [n_ref*n_rays*cap] bt_hit   a ray's hits sorted by (t, tri); cap = max_hits_per_ray
 *   d_valid_mask [n_ref*n_rays] u32     bit h set iff hit h of the ray has flags == 0 (may be NULL; lets bt_ags_select
 *                                       work without re-reading the hit records)
 *   d_overflow  [1] i32                 set to the number of rays whose hit list overflowed (may be NULL) */
BT_API int bt_ags_cast(const bt_mesh* mesh, const double* d_ref_pts, const double* d_dirs, int64_t n_ref,
                int32_t n_rays, const bt_ags_params* params, int32_t* d_hit_count, bt_hit* d_hits,
                uint32_t* d_valid_mask, int32_t* d_overflow, void* stream);

/* ---- R7: per reference point pick up to max_targets of its valid hits (replaces sampling.py:328-331).
 *   seed == 0 : deterministic, the first max_targets valid hits in canonical order (ray, t)
 *   seed != 0 : uniformly random subset keyed by (seed, reference point)
 *   d_pair_count [n_ref] i32 out; d_pair_offset [n_ref+1] i64 out (exclusive scan);
 *   d_pair_ref [cap_pairs] i32, d_pair_q [cap_pairs,3] f64 out; cap_pairs >= n_ref*max_targets is always enough */
BT_API int bt_ags_select(const int32_t* d_hit_count, const bt_hit* d_hits, const uint32_t* d_valid_mask /* may be NULL */,
                  int64_t n_ref, int32_t n_rays,
                  int32_t max_hits_per_ray, int32_t max_targets, uint64_t seed, int32_t* d_pair_count,
                  int64_t* d_pair_offset, int32_t* d_pair_ref, double* d_pair_q, int64_t cap_pairs,
                  void* stream);

/* ---- O1/O2: expand contact pairs into n_orientations gripper poses (replaces construct_grasp_set,
 * sampling.py:132-183, and construct_halfspace_grasp_set, :72-130, incl. the reference's row order:
 * within a reference point's group of k pairs, row r has the pose of (orientation r / k, target r % k)
 * but width / contacts of target r / n_orientations).
 *   d_ref_pts [n_ref,3]; d_pair_offset [n_ref+1]; d_pair_q [n_pairs,3];
 *   d_frame_vecs [n_ref,3] f64 (the random unit vector of sampling.py:157; ignored for halfspace)
 *   h_cos/h_sin [n_orientations] f64 host tables of the orientation angles (computed with numpy by the
 *   caller so they are bit-identical to the reference's np.cos/np.sin)
 *   d_gs [n_pairs*n_orientations,14] f32 out, d_contacts [n_pairs*n_orientations,2,3] f64 out (may be NULL)
 *   the total pair count is read on the device from d_pair_offset[n_ref] (no host sync); the launch is
 *   sized for cap_pairs; n_orientations <= 64 */
BT_API int bt_expand_orientations(const double* d_ref_pts, const int64_t* d_pair_offset, const int32_t* d_pair_ref,
                           const double* d_pair_q, const double* d_frame_vecs, int64_t n_ref, int64_t cap_pairs,
                           int32_t n_orientations, int32_t halfspace, const double* h_cos,
                           const double* h_sin, float* d_gs, double* d_contacts, void* stream);

/* ---- K1: gripper-vs-scene collision (replaces AntipodalGraspSampler.check_collisions,
 * sampling.py:354-397 -> trimesh CollisionManager -> FCL).  colliding[i] = any triangle of the gripper
 * mesh, mapped by pose_i * diag(s_i,1,1,1) with s_i = (width_i + tol)/opening_width (float32
 * arithmetic as numpy>=2 evaluates sampling.py:394) intersects any scene triangle (FCL's 17-axis
 * separating-axis test; touching counts; containment does not).
 *   bt_gripper_create: h_tris [n_tris,3,3] f64 in the TCP frame (gripper.mesh after tf_base_to_TCP,
 *   sampling.py:382-385); h_parts [n_parts] groups of consecutive triangles with their TCP-frame AABB
 *   bt_collide: d_gs [n,14] f32; d_n (may be NULL) device pointer to the actual row count (<= n);
 *   d_colliding [n] u8 out (4-byte aligned, capacity padded to a multiple of 4 bytes) */
typedef struct bt_gripper bt_gripper;
BT_API int bt_gripper_create(const double* h_tris, int32_t n_tris, const bt_gripper_part* h_parts, int32_t n_parts,
                      int device, bt_gripper** out);
BT_API int bt_gripper_destroy(bt_gripper* gripper);
BT_API int bt_collide(const bt_mesh* scene, const bt_gripper* gripper, const float* d_gs, int64_t n,
               const int64_t* d_n, double opening_width, double width_tolerance, int32_t use_width,
               uint8_t* d_colliding, void* stream);

/* ---- mesh-vs-mesh collision (replaces mesh_processing.collisions, mesh_processing.py:266-281, the
 * pair test behind Scene.colliding_instances core.py:666-687 and sampling.sample_scene sampling.py:498-569 ->
 * trimesh CollisionManager.in_collision_internal -> FCL).  Both meshes are in world coordinates.  d_flag[0] (i32, device)
 * becomes 1 iff some triangle of `a` intersects some triangle of `b` (FCL's 17-axis test with a's triangle in the P role;
 * touching counts, containment does not), else 0.  Only enqueues work on `stream`. */
BT_API int bt_mesh_pair_collide(const bt_mesh* a, const bt_mesh* b, int32_t* d_flag, void* stream);

/* ---- stream compaction of the surviving grasps: keeps rows with mask[i] == keep_value.
 *   d_out_gs [n,14], d_out_contacts [n,2,3] (may be NULL with d_contacts), d_n_out [1] i64
 *   d_base (nullable, [1] i64 on the device): number of rows ALREADY in the outputs -- the kept rows are appended
 *   behind them and d_n_out becomes *d_base + kept (batches of one call chained on the device, no host round trip) */
BT_API int bt_compact(const uint8_t* d_mask, uint8_t keep_value, const float* d_gs, const double* d_contacts,
               int64_t n, const int64_t* d_n, float* d_out_gs, double* d_out_contacts, int64_t* d_n_out,
               const int64_t* d_base, void* stream);
/* The outputs of bt_compact (and free_gs / free_contacts / n_free of bt_ags_pipeline) may be ANY memory the device can
 * store to: local HBM, a peer GPU's buffer mapped with bt_peer_open (the multi-GPU gather of SURVEY 8e: the kept rows
 * travel over NVLink while the scatter kernel runs, no collective afterwards), or mapped pinned host memory (the
 * device-to-host copy of the end-to-end call).  The kernel emits each block's run as consecutive aligned 32-byte stores. */

/* ---- peer-visible buffers for the gather to rank 0 (one process per GPU, so the mapping crosses processes: CUDA IPC).
 * bt_peer_alloc: plain cudaMalloc block of `bytes` (zeroed) on `device` + its 64-byte IPC handle; the handle travels
 * through any host channel (torch.distributed broadcast).  bt_peer_open maps it into another process on that process's
 * `device` (peer access enabled lazily); bt_peer_close / bt_peer_free undo them.  All four synchronise. */
#define BT_IPC_HANDLE_BYTES 64
typedef struct bt_ipc_handle { unsigned char bytes[BT_IPC_HANDLE_BYTES]; } bt_ipc_handle;
BT_API int bt_peer_alloc(int device, int64_t bytes, void** d_ptr, bt_ipc_handle* handle);
BT_API int bt_peer_open(int device, const bt_ipc_handle* handle, void** d_ptr);
BT_API int bt_peer_close(int device, void* d_ptr);
BT_API int bt_peer_free(int device, void* d_ptr);
/* One-sided completion of the gather (no collective in the step).  bt_peer_signal: enqueued behind the compaction on
 * `stream`, stores *d_local_count to d_count_slot and then `epoch` to d_flag_slot (both usually peer addresses).
 * bt_peer_wait: one small kernel that polls d_words[0 .. n_words) until each is >= epoch, for at most timeout_s seconds
 * (then d_status[0] = 1 and the kernel ends: a lost peer cannot hang the device); afterwards, if d_ack is not NULL,
 * stores ack_value there (the destination telling the peers which epoch it has consumed). */
/* Staged gather: stream min(*d_n, cap_rows) compacted rows (56 B) and contact records (48 B, optional) from local memory
 * into a peer-mapped destination with a thin grid of `n_blocks` blocks (16-byte stores); enqueue it on an exchange stream
 * behind the step, followed by bt_peer_signal.  All four buffers 16-byte aligned. */
BT_API int bt_peer_push(const int64_t* d_n, int64_t cap_rows, const float* d_gs, const double* d_contacts, float* d_dst_gs,
                        double* d_dst_contacts, int32_t n_blocks, void* stream);
/* The same transfer by a copy engine (cudaMemcpyAsync, no SM): `bytes` of a padded buffer, for callers that accept moving
 * the whole capacity instead of the counted rows. */
BT_API int bt_peer_copy(void* d_dst, const void* d_src, int64_t bytes, void* stream);
BT_API int bt_peer_signal(const int64_t* d_local_count, int64_t* d_count_slot, int64_t* d_flag_slot, int64_t epoch, void* stream);
BT_API int bt_peer_wait(const int64_t* d_words, int32_t n_words, int64_t epoch, double timeout_s, int32_t* d_status,
                 int64_t* d_ack, int64_t ack_value, void* stream);

/* ---- the whole AGS pipeline in ONE call (what AntipodalGraspSampler.sample's loop body + check_collisions do
 * per batch, sampling.py:220-350 and :354-397): [S1] -> [S2] -> R1..R6 -> R7 -> [frame vectors] -> O1/O2 ->
 * [K1 -> compaction].  All buffers are caller-allocated device memory (torch tensors in the Python mirror); the call
 * only enqueues kernels on `stream` and never synchronises.  bt_timer (optional) records one CUDA event per stage
 * boundary so a benchmark can read per-stage device times afterwards.  seed == 0 makes the target choice deterministic
 * (first valid hit in canonical order), any other seed keys the device RNG streams. */
typedef struct bt_timer bt_timer;
#define BT_N_STAGES 8        /* generate (S1 + S2 + frame vectors), -, cast, select, -, expand, collide, compact */
BT_API int bt_timer_create(int device, bt_timer** out);
BT_API int bt_timer_destroy(bt_timer* timer);
BT_API int bt_timer_read(bt_timer* timer, float* h_stage_ms /* [BT_N_STAGES], -1 for stages that did not run */);  /* synchronises */

typedef struct bt_pipeline_config {
    bt_ags_params ags;
    int32_t n_rays, n_orientations, max_targets, halfspace;
    int32_t check_collisions, use_width, compact;
    int32_t tail_splits;     /* > 1: collide + compact run over that many row ranges on two streams, so that the compaction of
                                one range (link-bound when it stores to mapped host / peer memory) overlaps the collision
                                walk of the next; same output, bit for bit.  0 / 1 = off.  Max 8.                       */
    int32_t cone_culling;    /* lean mode (buffers->hits == NULL) only: the ray caster skips triangles and LBVH subtrees whose
                                vertex normals rule out a VALID antipodal hit for the ray (interpolated normal within
                                atan(mu) of the ray); valid_mask / contacts / grasps are unchanged, hit_count then counts
                                only the hits that were looked at.  0 = find every crossing                               */
    int32_t reserved;
    double  collision_width_tolerance;
    double  orient_cos[64], orient_sin[64];   /* numpy-evaluated cos/sin of the orientation angles */
    int64_t surface_total;   /* > 0 (with buffers->sample_surface and ->seed_add): the batch's reference points are taken from ONE set
                                of surface_total area-weighted surface samples in a keyed pseudo-random order -- reference point i is
                                sample perm(seed_add[1] + i mod surface_total), the samples keyed by seed_add[2] -- i.e. stage S1 and the
                                np.random.shuffle of a whole sample() call (sampling.py:199-201) without materialising the set.
                                0: the batch draws its own n_ref samples keyed by `seed`                                          */
} bt_pipeline_config;

typedef struct bt_pipeline_buffers {
    double*  ref_pts;        /* [n_ref,3]   in, or out of S1 when sample_surface != 0       */
    double*  ref_nrm;        /* [n_ref,3]                                                    */
    double*  dirs;           /* [n_ref,n_rays,3] in, or out of S2 when generate_dirs != 0    */
    double*  frame_vecs;     /* [n_ref,3]   in, or generated when generate_frame_vecs != 0   */
    int32_t  sample_surface, generate_dirs, generate_frame_vecs, reserved;
    int32_t* hit_count;      /* [n_ref*n_rays]                                               */
    uint32_t* valid_mask;    /* [n_ref*n_rays]                                               */
    bt_hit*  hits;           /* [n_ref*n_rays*max_hits_per_ray], or NULL = lean mode: no hit records are
                                materialised (they are an intermediate: 144 MB per million rays), the caster
                                keeps only valid_mask and hit_t, from which the target choice rebuilds the
                                contact o + t d with the caster's own expression                          */
    int32_t* overflow;       /* [1]                                                          */
    int32_t* pair_count;     /* [n_ref]                                                      */
    int64_t* pair_offset;    /* [n_ref+1]                                                    */
    int32_t* pair_ref;       /* [cap_pairs]                                                  */
    double*  pair_q;         /* [cap_pairs,3]                                                */
    int64_t  cap_pairs;      /* >= n_ref * max_targets                                       */
    float*   gs;             /* [cap_pairs*n_orientations,14]                                */
    double*  contacts;       /* [cap_pairs*n_orientations,2,3]                               */
    int64_t* n_grasps;       /* [1] number of posed grasps = pairs * n_orientations          */
    uint8_t* colliding;      /* [cap rows, padded to 4]   (check_collisions)                 */
    float*   free_gs;        /* [cap rows,14]             (compact)                          */
    double*  free_contacts;  /* [cap rows,2,3]            (compact)                          */
    int64_t* n_free;         /* [1]                       (compact)                          */
    const int64_t* free_base; /* nullable [1]: rows already in free_gs / free_contacts (see bt_compact d_base)  */
    double*  hit_t;          /* [max_hits_per_ray, n_ref*n_rays] slot-major ray parameters (lean mode only)   */
    void*    compact_after;  /* nullable cudaEvent_t: the compaction stage waits for it (the previous batch's
                                compaction on another stream, whose count is free_base)                       */
    const uint64_t* seed_add; /* nullable [4] on the device, what changes between replays of a captured call: [0] added to `seed` by
                                every kernel that draws random numbers, [1] / [2] position in and key of the shuffled surface
                                samples (config->surface_total), [3] spare.  Set with bt_set_u64 / bt_set_u64x3.           */
    void*    scratch;        /* nullable, 256-byte aligned device memory the call may use for its intermediates (candidate
                                lists, scan states, collision queues) instead of stream-ordered allocations; what does not
                                fit is allocated as before.  ~80 B per ray + 200 B per posed grasp holds everything.     */
    int64_t  scratch_bytes;
} bt_pipeline_buffers;

BT_API int bt_ags_pipeline(const bt_mesh* target, const bt_mesh* scene, const bt_gripper* gripper,
                    const bt_pipeline_config* config, const bt_pipeline_buffers* buffers, int64_t n_ref,
                    uint64_t seed, bt_timer* timer, void* stream);

/* ---- CUDA graphs.  Every call of this library only enqueues work on the caller's stream, so a sequence of calls can be
 * captured once (bt_graph_begin_capture ... calls ... bt_graph_end_capture; `stream` must not be the default stream) and
 * replayed with bt_graph_launch: one launch instead of ~25 driver calls per bt_ags_pipeline.  The buffers are baked into
 * the graph; a replay's random numbers are keyed by bt_pipeline_buffers.seed_add (device memory, set with bt_set_u64 on
 * the same stream before the launch).  bt_graph_info: node counts and the '\n'-separated names of the kernel nodes. */
typedef struct bt_graph bt_graph;
BT_API int bt_graph_begin_capture(void* stream);
BT_API int bt_graph_end_capture(void* stream, bt_graph** out);
BT_API int bt_graph_launch(bt_graph* graph, void* stream);
BT_API int bt_graph_info(const bt_graph* graph, int32_t* n_nodes, int32_t* n_kernel_nodes, char* names, int64_t names_cap);
BT_API int bt_graph_destroy(bt_graph* graph);
BT_API int bt_set_u64(uint64_t* d_ptr, uint64_t value, void* stream);
BT_API int bt_set_u64x3(uint64_t* d_ptr, uint64_t v0, uint64_t v1, uint64_t v2, void* stream);

/* ---- M1..M5: coverage (replaces metrics.coverage / coverage_brute_force, metrics.py:141-229, and the
 * distance functions they call, :9-77).  covered[i] = any_j ( 1000*|t_i - t_j| + deg(R_i, R_j) <= eps ),
 * all float32 in the reference's operation order; deg() goes through the caller-supplied 20001-entry
 * table of np.rad2deg(np.arccos(k/1e4)) so the transcendental is bit-identical to numpy.
 *   mode 0: brute-force semantics (coverage_brute_force)
 *   mode 1: kd-tree semantics (coverage): additionally requires the float64 squared translation distance
 *           to be <= (eps/1000)^2, the cKDTree.query_ball_tree candidate test
 *   algo 0: all-pairs tiles; algo 1: uniform-grid culled (queries binned in cells of eps/1000)
 *   d_covered [n_ref] u8 out, d_count [1] i64 out */
BT_API int bt_coverage(const float* d_ref, int64_t n_ref, const float* d_qry, int64_t n_qry, double epsilon,
                const float* d_acos_deg_lut, int32_t mode, int32_t algo, uint8_t* d_covered,
                int64_t* d_count, void* stream);

/* The query side of the uniform-grid coverage, prepared once and reused: a sweep (many reference sets, or the shards of one
 * reference set on several GPUs, against ONE query set -- BASELINE config 5) bins only its reference rows per call.
 * bt_coverage_prepare: grid over the query set's bounds for `epsilon`, queries sorted by cell (synchronises; d_qry must
 * outlive the handle).  bt_coverage_prepared: bt_coverage(algo = 1) against the prepared set, same masks. */
typedef struct bt_coverage_query bt_coverage_query;
BT_API int bt_coverage_prepare(const float* d_qry, int64_t n_qry, double epsilon, bt_coverage_query** out, void* stream);
BT_API int bt_coverage_prepared(const float* d_ref, int64_t n_ref, const bt_coverage_query* query, const float* d_acos_deg_lut,
                                int32_t mode, uint8_t* d_covered, int64_t* d_count, void* stream);
BT_API int bt_coverage_query_destroy(bt_coverage_query* query);

/* pairwise distance matrices (metrics.py:9-77), [n1,n2] f32 out; which: 0 euclidean, 1 angular [rad lut
 * supplied by caller as d_lut], 2 combined (lut = degrees) */
BT_API int bt_pairwise_distances(const float* d_gs1, int64_t n1, const float* d_gs2, int64_t n2, int32_t which,
                          float weight, const float* d_lut, float* d_out, void* stream);

/* average key-point distance between every pair of grasps (replaces metrics.avg_gripper_point_distances,
 * metrics.py:94-138; h_points [n_points,3] f64 key points in the gripper frame, default = metrics.py:80-91);
 * [n1,n2] f64 out.  consider_symmetry also tries the second gripper flipped (points [1,0,3,2,4], needs n_points == 5). */
BT_API int bt_avg_gripper_point_distances(const float* d_gs1, int64_t n1, const float* d_gs2, int64_t n2,
                                          const double* h_points, int32_t n_points, int32_t consider_symmetry,
                                          double* d_out, void* stream);

/* ---- generators feeding the path -------------------------------------------------------------
 * bt_random_poses: [n,4,4] f64 homogeneous transforms, uniform random rotation + translation U[0,1)^3 (replaces
 *   sampling.random_poses, sampling.py:457-469; device RNG, distributional parity only)
 * bt_grasp_perturbations: for every grasp [original +] per radius 6 translations (+-radius mm along each pose axis) and
 *   6 rotations (+-radius deg about each pose axis); rows float32 [n * (12 * n_radii + include_original), 14] in the
 *   reference's order (replaces sampling.grasp_perturbations, sampling.py:400-454)
 * bt_farthest_point_sampling: k indices, the first one is 0, then repeatedly the point farthest from all chosen ones
 *   (replaces sampling.farthest_point_sampling, sampling.py:472-495); points [n, stride >= 3] f32 or f64 (xyz first) */
BT_API int bt_random_poses(int64_t n, uint64_t seed, double* d_out, void* stream);
BT_API int bt_grasp_perturbations(const float* d_gs, int64_t n, const double* h_radii, int32_t n_radii,
                                  int32_t include_original, float* d_out, void* stream);
BT_API int bt_farthest_point_sampling(const void* d_points, int32_t is_float64, int32_t stride, int64_t n, int64_t k,
                                      int64_t* d_indices, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BURG_B200_H */

#!/usr/bin/env python
"""bench.py -- the headline benchmark of the AGS + coverage hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload ags_c3|ags_c4|coverage_c2|coverage_c5]

Prints ONE JSON line (rank 0).  Workloads (BASELINE.json `configs`):
  ags_c3       config 3: AGS on the 99 904-triangle 'YCB-like' mesh, 10 000 reference points x 100 rays = 1 M
               candidates per step per GPU, max_targets_per_ref_point = 1, 12 orientations, gripper-vs-object
               collision + compaction.  metric: candidates/s.  (default; the config the metric's target is quoted on)
  ags_c1       config 1: AntipodalGraspSampler.sample(1000) + check_collisions on a 20 480-triangle icosphere (public API call).
  ags_c4       config 4: 20 objects (~1 M triangles) + ground slab, the target object changes every step, gripper vs scene.
  coverage_c2  config 2: coverage of two 100k grasp sets (perturbed pair).  metric: pairs/s.
  coverage_c5  config 5: 1 M x 1 M grasps; with several ranks the reference rows are sharded (strong scaling).

A "step" is one pass of the device pipeline over one batch of synthetic candidates.  The timed region holds
exactly K steps, each bracketed by CUDA events on the launching stream; L2 is flushed between steps (outside the
event pairs); the reported time is the max over ranks of the summed step times.

--impl reference times the reference's CPU algorithm on the host cores instead: the reference is pure Python on
un-installable third-party leaves, so this arm runs the oracle port (oracle/coracle.c: same arithmetic, BVH broad
phase, OpenMP over all host threads) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_REF = 10_000
N_RAYS = 100
N_ORIENT = 12
COV_N = 100_000
# own kernels launched by one bt_ags_pipeline call, in order (CUB scans and memsets not counted)
AGS_KERNELS = ['sample_surface_kernel', 'cone_rays_kernel', 'traverse_kernel', 'cast_kernel', 'zero_first_kernel', 'select_count_kernel',
               'select_fill_kernel', 'unit_vectors_kernel', 'expand_kernel', 'scale_count_kernel', 'collide_kernel<lean>',
               'collide_exact_kernel', 'collide_kernel<rescue>', 'compact_count_kernel', 'compact_scatter_kernel']


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe): NVML every ~2 ms when the
    bindings are importable (nvidia-smi takes ~100 ms per query, longer than a whole timed region here), else nvidia-smi."""
    QUERY = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
    NAMES = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']

    def __init__(self, gpu_index):
        self.gpu, self.sm, self.mx, self.reasons = gpu_index, [], [], set()
        self.stop_flag, self.thread, self.source = False, None, 'nvidia-smi'
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # LOCAL_RANK indexes CUDA_VISIBLE_DEVICES; NVML indexes physical devices
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            idx = gpu_index
            if vis:
                ids = [v.strip() for v in vis.split(',') if v.strip()]
                if gpu_index < len(ids) and ids[gpu_index].isdigit():
                    idx = int(ids[gpu_index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nvml, self.source = pynvml, 'nvml'
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        self.sm.append(int(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
        self.mx.append(int(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)))
        r = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)) if hasattr(n, 'nvmlDeviceGetCurrentClocksEventReasons') \
            else int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        for name, bit in (('hw_slowdown', 0x8), ('hw_thermal_slowdown', 0x40), ('sw_thermal_slowdown', 0x20), ('sw_power_cap', 0x4)):
            if r & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(['nvidia-smi', f'--query-gpu={self.QUERY}', '--format=csv,noheader,nounits', '-i',
                              str(self.gpu)], capture_output=True, text=True, timeout=5).stdout.strip()
        r = [c.strip() for c in out.split(',')] if out else []
        if r and r[0].isdigit():
            self.sm.append(int(r[0]))
        if len(r) > 1 and r[1].isdigit():
            self.mx.append(int(r[1]))
        for i in range(4):
            if len(r) >= 6 and r[2 + i].lower().startswith('active'):
                self.reasons.add(self.NAMES[i])

    def _run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            time.sleep(0.002 if self.nvml is not None else 0.1)

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=6)
        return {'sm_mhz': int(statistics.median(self.sm)) if self.sm else None, 'sm_max_mhz': max(self.mx) if self.mx else None,
                'reasons': sorted(self.reasons), 'samples': len(self.sm), 'source': self.source}


# ---------------------------------------------------------------------------------------------------------------
# synthetic inputs
# ---------------------------------------------------------------------------------------------------------------
from burg_toolkit_b200.synthetic import c3_mesh, c4_scene, coverage_sets, perturbed_rows, random_rows  # noqa: E402,F401


def host_candidates(mesh, n_ref, seed):
    """host-side reference points, normals, cone rays and frame vectors (CPU baseline / parity inputs)"""
    from oracle import leaves, port
    rng = np.random.default_rng(seed)
    m = port.Mesh(mesh.vertices, mesh.triangles, mesh.vertex_normals)
    tn = leaves.unitize(np.cross(m.triangles[:, 1] - m.triangles[:, 0], m.triangles[:, 2] - m.triangles[:, 0]))
    pts, nrm, _ = leaves.sample_points_uniformly(m.vertices, m.faces, tn, n_ref, rng)
    sel = rng.permutation(len(pts))
    pts, nrm = pts[sel], nrm[sel]
    angle = np.arctan(0.25)
    dirs = np.stack([port.cone_rays(-nrm[i], rng.uniform(0, 2 * np.pi, N_RAYS), rng.uniform(0, angle, N_RAYS)) for i in range(len(pts))])
    fv = rng.normal(size=(len(pts), 3))
    fv /= np.linalg.norm(fv, axis=1, keepdims=True)
    return pts, nrm, dirs, fv


def host_candidates_fast(mesh, n_ref, seed):
    """the same distribution as host_candidates, vectorised over reference points (CPU timing samples only; the parity
    tests use host_candidates / the oracle's own generator)"""
    from oracle import leaves, port
    rng = np.random.default_rng(seed)
    m = port.Mesh(mesh.vertices, mesh.triangles, mesh.vertex_normals)
    tn = leaves.unitize(np.cross(m.triangles[:, 1] - m.triangles[:, 0], m.triangles[:, 2] - m.triangles[:, 0]))
    pts, nrm, _ = leaves.sample_points_uniformly(m.vertices, m.faces, tn, n_ref, rng)
    sel = rng.permutation(len(pts))
    pts, nrm = pts[sel], nrm[sel]
    n = len(pts)
    az, inc = rng.uniform(0, 2 * np.pi, (n, N_RAYS)), rng.uniform(0, np.arctan(0.25), (n, N_RAYS))
    cart = np.stack([np.sin(inc) * np.cos(az), np.sin(inc) * np.sin(az), np.cos(inc)], axis=-1)       # (n, rays, 3)
    k = -nrm / np.linalg.norm(nrm, axis=1, keepdims=True) + np.array([0.0, 0.0, 1.0])                 # util.py:7-29
    k[np.linalg.norm(k, axis=1) == 0] = [1.0, 0.0, 0.0]
    kc = np.einsum('nk,nrk->nr', k, cart)
    dirs = 2.0 * kc[..., None] * k[:, None, :] / np.einsum('nk,nk->n', k, k)[:, None, None] - cart
    fv = rng.normal(size=(n, 3))
    fv /= np.linalg.norm(fv, axis=1, keepdims=True)
    return pts, nrm, np.ascontiguousarray(dirs), fv


# ---------------------------------------------------------------------------------------------------------------
# CPU arm (oracle port, all host threads)
# ---------------------------------------------------------------------------------------------------------------
_CPU_MESHES = {}


def merged_scene(meshes):
    """vertices / faces of a list of meshes concatenated into one (the CPU port's scene for config 4)"""
    v = np.concatenate([m.vertices for m in meshes])
    off = np.cumsum([0] + [len(m.vertices) for m in meshes[:-1]])
    f = np.concatenate([m.triangles + o for m, o in zip(meshes, off)])
    return v, f


def cpu_ags(mesh, n_ref, seed=100, scene_meshes=None):
    """one bounded CPU sample of the AGS workload -> (candidates, seconds, threads); the BVH builds are not timed"""
    from oracle import cport
    from burg_toolkit_b200.gripper import ParallelJawGripper
    key = (id(mesh), id(scene_meshes))
    if key not in _CPU_MESHES:
        cm = cport.Mesh(mesh.vertices, mesh.triangles, mesh.vertex_normals)
        cs = cm if scene_meshes is None else cport.Mesh(*merged_scene(scene_meshes))
        _CPU_MESHES[key] = (cm, cs)
    cm, cs = _CPU_MESHES[key]
    cport.set_num_threads(os.cpu_count() or 1)            # all host threads, also under torchrun (OMP_NUM_THREADS=1 there)
    pts, nrm, dirs, fv = host_candidates_fast(mesh, n_ref, seed)
    gt = ParallelJawGripper().tcp_triangles()
    t0 = time.perf_counter()
    out = cport.pipeline(cm, cs, pts, dirs, fv, gt, n_orientations=N_ORIENT, coll_tol=0.01, max_hits_per_ray=16)
    dt = time.perf_counter() - t0
    return len(pts) * N_RAYS, dt, cport.num_threads(), out


def cpu_coverage(n, seed=0):
    """the reference's kd-tree coverage (scipy cKDTree + Python loop, metrics.py:161-229) restated in oracle/port.py"""
    from oracle import port
    ref, qry = coverage_sets(n)
    t0 = time.perf_counter()
    cov = port.covered_kd(ref, qry)
    dt = time.perf_counter() - t0
    return n * len(qry), dt, 1, float(cov.mean())


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count()
    if args.workload in ('ags_c1', 'ags_c3', 'ags_c4'):
        scene_meshes = None
        if args.workload == 'ags_c3':
            mesh = c3_mesh()
        elif args.workload == 'ags_c1':
            from burg_toolkit_b200 import mesh as bmesh
            mesh = bmesh.create_icosphere(5, 0.03)
        else:
            scene_meshes = c4_scene().get_mesh_list()
            mesh = scene_meshes[1]
        n_ref = args.cpu_ref_points
        for _ in range(max(args.warmup, 0)):
            cpu_ags(mesh, max(8, n_ref // 8), scene_meshes=scene_meshes)
        times, cands, threads = [], 0, 1
        for s in range(args.steps):
            c, dt, threads, _ = cpu_ags(mesh, n_ref, seed=100 + s, scene_meshes=scene_meshes)
            times.append(dt)
            cands += c
        total = sum(times)
        value, unit, metric = cands / total, 'candidates/s', 'AGS candidates/sec (ray+cone+collision)'
        sample = f'{n_ref} reference points x {N_RAYS} rays per step ({n_ref * N_RAYS} candidates) of the {args.workload} workload'
        config = {'workload': args.workload, 'mesh_triangles': len(mesh.triangles), 'ref_points_per_step': n_ref,
                  'n_rays': N_RAYS, 'n_orientations': N_ORIENT, 'max_targets_per_ref_point': 1,
                  'collision': 'gripper-vs-object' if scene_meshes is None else
                  f'gripper-vs-scene ({sum(len(m.triangles) for m in scene_meshes)} triangles)'}
    else:
        n = args.cpu_coverage_n
        times, pairs, threads = [], 0, 1
        for s in range(args.steps):
            p, dt, threads, _ = cpu_coverage(n)
            times.append(dt)
            pairs += p
        total = sum(times)
        value, unit, metric = pairs / total, 'pairs/s', 'coverage pairs/sec'
        sample = f'{n} x {n} grasps per step (kd-tree coverage, single-threaded like the reference)'
        config = {'workload': args.workload, 'n_ref': n, 'n_query': n}
    line = {'impl': 'reference', 'metric': metric, 'value': value, 'unit': unit, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': 1e3 * total / max(args.steps, 1), 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64' if args.workload.startswith('ags') else 'f32',
            'data': 'synthetic', 'config': config,
            'cpu_baseline': {'value': value, 'unit': unit, 'cores': threads, 'kind': 'port', 'sample': sample,
                             'host_cpus': cores},
            'e2e': {'value': value, 'unit': unit, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
def event_ms(pairs):
    return [a.elapsed_time(b) for a, b in pairs]


def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local_rank}'))
    from burg_toolkit_b200 import metrics, sampling
    from burg_toolkit_b200.gripper import ParallelJawGripper

    peak_gbs, peak_src = measured_peaks()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')       # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    clocks = ClockSampler(local_rank)
    line = {}
    if args.workload in ('ags_c3', 'ags_c4'):
        ags = sampling.AntipodalGraspSampler()
        ags.gripper, ags.verbose, ags.device = ParallelJawGripper(), False, local_rank
        ags.n_orientations, ags.n_rays = N_ORIENT, N_RAYS
        if args.workload == 'ags_c3':
            mesh = c3_mesh()
            targets, scene_meshes, pipe_kw = [mesh], None, {}
        else:       # config 4: 20 objects + ground slab; every step samples on one object, gripper vs the WHOLE scene
            scene_meshes = c4_scene().get_mesh_list()
            targets = scene_meshes[1:]
            mesh = targets[0]
            pipe_kw = {'additional_objects': scene_meshes, 'exclude_shape': True}
        scene_tris = sum(len(m.triangles) for m in scene_meshes) if scene_meshes else len(mesh.triangles)
        from burg_toolkit_b200 import mesh as bmesh
        ags.mesh = bmesh.create_icosphere(1, 0.03)               # first build in the process also loads the CUDA module
        ags._device_mesh()
        build_ms = []
        for t in targets:                                       # build every target's LBVH before the timed region
            ags.mesh = t
            build_ms.append(ags._device_mesh().info.build_ms)
        ags.mesh = mesh
        dm = ags._device_mesh()
        info = dm.info
        # the path's only exchange: compacted grasp rows + contacts + counts to rank 0.  Every rank's compaction
        # kernel stores its kept rows straight into its slot of rank 0's buffer over NVLink (distributed.PeerGather, CUDA IPC);
        # NCCL carries the 8-byte counts, which is also the completion barrier.  If the box refuses the IPC mapping the
        # bench says so and uses a plain NCCL gather of the padded rows.
        peer, gather_buf, exchange = None, None, 'none (1 GPU)'
        if world > 1:
            from burg_toolkit_b200 import distributed as bdist
            from burg_toolkit_b200._native import NativeError
            try:
                mode = os.environ.get('BT_PEER_ONE_SIDED')                  # A/B knob; default: one-sided from 4 ranks on
                peer = bdist.PeerGather(N_REF * N_ORIENT, with_contacts=True, device=local_rank, one_sided=None if mode is None else mode != '0', timeout_s=30.0)
                exchange = ('compaction kernel stores rows + contacts into rank 0 over NVLink (CUDA IPC peer buffer); ' +
                            ('counts + epoch flags stored the same way, rank 0 polls them (no collective in the step)' if peer.one_sided
                             else 'NCCL all_gather of counts'))
            except NativeError as ex:
                exchange = f'NCCL all_reduce + gather of padded rows (peer mapping unavailable: {ex})'
                gather_buf = torch.empty((world, N_REF * N_ORIENT, 14), dtype=torch.float32, device='cuda') if rank == 0 else None

        from burg_toolkit_b200.device import StageTimer
        stage_timer = StageTimer(local_rank)

        def step(seed, timer=None):
            ags.mesh = targets[seed % len(targets)]
            out = ags.run_pipeline(n_ref=N_REF, seed=seed, check_collisions=True, timer=timer,
                                   destinations=peer.destinations() if peer else None, **pipe_kw)
            if peer is not None:
                peer.complete(out['n_free'])
            elif world > 1:
                dist.all_reduce(out['n_free'], op=dist.ReduceOp.SUM)
                dist.gather(out['free_gs'], list(gather_buf.unbind(0)) if rank == 0 else None, dst=0)
            return out

        clocks.start()                                          # samples cover warm-up + timed region (all under load)
        for w in range(args.warmup):
            step(1000 * rank + w + 1)
        barrier()
        stage_acc, step_events, stats = {}, [], []
        for s in range(args.steps):
            flush.fill_(s & 0xff)                                                    # L2 flush, outside the event pair
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = step(7919 * (rank + 1) + s)                                        # no per-stage events inside the timed region
            e1.record()
            step_events.append((e0, e1))
            stats.append((out['n_grasps'].clone(), out['n_free'].clone(), out['overflow'].clone()))
        barrier()
        clk = clocks.stop()
        # per-stage device times: a few extra, untimed steps with the stage timer (its 16 event records and the host wait
        # per step are instrumentation, so they stay out of the timed region above)
        for s in range(min(args.steps, 5)):
            flush.fill_(s & 0xff)
            step(104729 * (rank + 1) + s, stage_timer)
            for name, ms in stage_timer.read().items():
                stage_acc.setdefault(name, []).append(ms)
        step_ms = event_ms(step_events)
        total_ms = max_over_ranks(sum(step_ms))
        candidates = sum_over_ranks(N_REF * N_RAYS * args.steps)
        value = candidates / (total_ms * 1e-3)
        stage_ms = {k: statistics.mean(v) for k, v in stage_acc.items()}
        n_grasps = int(torch.stack([a.reshape(()) for a, _, _ in stats]).double().mean().item())
        n_free = int(torch.stack([b.reshape(()) for _, b, _ in stats]).double().mean().item()) // (1 if (peer is not None or world == 1) else world)
        overflow = int(torch.stack([c.reshape(()) for _, _, c in stats]).sum().item())

        # work counters (instrumented kernel instantiations, one extra untimed step) for the L2 traffic model
        from burg_toolkit_b200 import _native as nat
        counters = torch.zeros(8, dtype=torch.int64, device='cuda')
        nat.call('bt_debug_set_counters', counters.data_ptr())
        step(424242 + rank)
        torch.cuda.synchronize()
        nat.call('bt_debug_set_counters', None)
        cnt = [int(x) for x in counters.cpu().tolist()]
        n_cand = N_REF * N_RAYS
        per_ray = {'nodes': cnt[0] / n_cand, 'leaf_pretests': cnt[1] / n_cand, 'exact_tests': cnt[2] / n_cand}
        per_grasp = {'nodes': cnt[3] / max(n_grasps, 1), 'leaves_classified': cnt[4] / max(n_grasps, 1),
                     'leaves_exact': cnt[5] / max(n_grasps, 1)}

        # roofline of the dominant stage (the ray caster = traverse_kernel + cast_kernel): algorithmic HBM bytes per launch
        # (64 B per candidate + 80 B per triangle) over its event-timed duration, against the measured HBM peak;
        # plus the traversal-traffic model against L2 (the mesh + LBVH are L2-resident by design, HBM is not the binding roof)
        cast_ms = stage_ms['cast']
        alg_bytes = N_REF * N_RAYS * 64 + 80 * info.n_triangles
        achieved = alg_bytes / (cast_ms * 1e-3) / 1e9
        l2_bytes = cnt[0] * 64 + cnt[1] * 48 + cnt[2] * 128 + n_cand * (24 + 48 + 2 * 8 + 8)
        prof = {}
        ppath = os.path.join(ROOT, 'profiles', 'roofline_inputs.json')
        if os.path.exists(ppath):
            with open(ppath) as fh:
                prof = json.load(fh)
        roofline = {'bound': 'hbm', 'kernel': 'traverse_kernel + cast_kernel (ray caster stage)', 'achieved': achieved, 'peak': peak_gbs,
                    'unit': 'GB/s', 'frac': achieved / peak_gbs, 'traffic': prof.get('cast_stage', {}).get('dram_bytes_per_launch'), 'traffic_source': prof.get('cast_stage', {}).get('source'),
                    'peak_source': peak_src, 'algorithmic_bytes_per_launch': alg_bytes, 'kernel_ms': cast_ms,
                    'rays_per_s': N_REF * N_RAYS / (cast_ms * 1e-3),
                    'l2_model': {'bytes_per_launch': l2_bytes, 'achieved_gbs': l2_bytes / (cast_ms * 1e-3) / 1e9,
                                 'definition': 'wide_nodes*64 + leaf_pretests*48 + exact_tests*128 + rays*(24 B float32 ray in traverse + 48 B float64 ray in resolve + 2*8 B ray parameters + 8 B count and mask out)',
                                 'per_ray': per_ray},
                    'collide_per_grasp': per_grasp,
                    'note': 'mesh + LBVH + per-leaf records (~50 MB) are L2-resident by design, so HBM is not the binding roof: the stage is bound by L1 data-pipe wavefronts of divergent node fetches; see DESIGN.md section 4'}

        # end-to-end through the public API with HOST buffers (pinned), H2D + D2H inside the timed region
        ags.mesh = mesh
        pts, nrm, _ = [t.cpu().numpy() for t in ags.sample_surface(N_REF, 4242 + rank)]
        pinned = {}
        for w in range(max(args.warmup, 1)):
            ags.evaluate_host(pts, nrm, seed=w + 1, pinned=pinned, **pipe_kw)
        barrier()
        e2e_events, h2d, d2h = [], 0, 0
        for s in range(args.steps):
            flush.fill_(s & 0xff)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            res = ags.evaluate_host(pts, nrm, seed=s + 11, pinned=pinned, **pipe_kw)
            e1.record()
            e2e_events.append((e0, e1))
            h2d, d2h = res['h2d_bytes'], res['d2h_bytes']
        barrier()
        e2e_ms = max_over_ranks(sum(event_ms(e2e_events)))
        e2e_value = candidates / (e2e_ms * 1e-3)
        exchange_timed_out = False
        if peer is not None:                                   # a timed-out wait kernel must not cost the JSON line: it is reported in it
            exchange_timed_out = sum_over_ranks(1.0 if peer.close(raise_on_timeout=False) else 0.0) > 0

        line = {'metric': 'AGS candidates/sec (ray+cone+collision)', 'value': value, 'unit': 'candidates/s',
                'ms_per_step': total_ms / args.steps, 'dtype': 'f64',
                'config': {'workload': args.workload,
                           'mesh': 'bumpy sphere 224x224 (YCB-like, seed 1)' if scene_meshes is None else
                                   '20 bumpy objects (158x158 grids, seeds 10-29) on a 1.0 x 0.5 m ground slab; target object changes every step',
                           'mesh_triangles': int(info.n_triangles), 'scene_triangles': int(scene_tris),
                           'ref_points_per_step_per_gpu': N_REF, 'n_rays': N_RAYS,
                           'candidates_per_step_per_gpu': N_REF * N_RAYS, 'n_orientations': N_ORIENT,
                           'max_targets_per_ref_point': 1,
                           'collision': ('gripper-vs-object' if scene_meshes is None else 'gripper-vs-scene') + ', use_width, tol 0.01',
                           'mode': 'reference mode (one pair per reference point)', 'l2': 'flushed between steps (256 MB write)',
                           'bvh_build_ms': float(np.mean(build_ms)), 'bvh_max_depth': int(info.max_depth),
                           'posed_grasps_per_step': n_grasps, 'collision_free_per_step': n_free, 'hit_list_overflows': overflow,
                           'exchange': exchange, 'exchange_timed_out': exchange_timed_out},
                'roofline': roofline, 'stage_ms': stage_ms,
                'e2e': {'value': e2e_value, 'unit': 'candidates/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                        'ms_per_step': e2e_ms / args.steps},
                'gpu_launches': len(AGS_KERNELS) * args.steps, 'kernels': AGS_KERNELS}
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            c, dt, threads, _ = cpu_ags(mesh, args.cpu_ref_points, scene_meshes=scene_meshes)
            line['cpu_baseline'] = {'value': c / dt, 'unit': 'candidates/s', 'cores': threads, 'kind': 'port',
                                    'sample': f'{args.cpu_ref_points} reference points x {N_RAYS} rays ({c} candidates) of the same '
                                              f'workload, oracle/coracle.c with OpenMP, {dt:.1f} s', 'host_cpus': os.cpu_count()}
            # secondary metric of BASELINE.json: coverage pairs/s (config 2), same JSON line
            if args.workload == 'ags_c3':
                line['secondary'] = coverage_numbers(args, torch, metrics, flush, barrier, short=True)
    elif args.workload == 'ags_c1':
        # config 0 of BASELINE.json: the reference's own CPU-runnable case through the public API --
        # AntipodalGraspSampler.sample(1000) on a 20 480-triangle icosphere, 12 orientations, then check_collisions
        from burg_toolkit_b200 import mesh as bmesh
        ags = sampling.AntipodalGraspSampler()
        ags.mesh, ags.gripper, ags.verbose, ags.device = bmesh.create_icosphere(5, 0.03), ParallelJawGripper(), False, local_rank
        ags.n_orientations, ags.n_rays = N_ORIENT, N_RAYS
        np.random.seed(1234 + rank)
        clocks.start()
        for _ in range(args.warmup):
            gs, _ = ags.sample(1000)
            ags.check_collisions(gs)
        barrier()
        ev, cands, grasps = [], 0, 0
        for s in range(args.steps):
            flush.fill_(s & 0xff)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            gs, contacts = ags.sample(1000)                    # host call: numpy results come back every time
            coll = ags.check_collisions(gs)
            e1.record()
            ev.append((e0, e1))
            cands += ags.last_stats['candidates']
            grasps += len(gs)
        barrier()
        clk = clocks.stop()
        total_ms = max_over_ranks(sum(event_ms(ev)))
        cand_all = sum_over_ranks(cands)
        value = cand_all / (total_ms * 1e-3)
        line = {'metric': 'AGS candidates/sec (ray+cone+collision)', 'value': value, 'unit': 'candidates/s',
                'ms_per_step': total_ms / args.steps, 'dtype': 'f64',
                'config': {'workload': 'ags_c1', 'mesh': 'icosphere, 5 subdivisions, r = 0.03', 'mesh_triangles': 20480,
                           'call': 'AntipodalGraspSampler.sample(1000) + check_collisions(gs), numpy in / numpy out',
                           'n_rays': N_RAYS, 'n_orientations': N_ORIENT, 'max_targets_per_ref_point': 1,
                           'candidates_per_step_per_gpu': cands / args.steps, 'grasps_per_step': grasps / args.steps,
                           'l2': 'flushed between steps (256 MB write)'},
                'roofline': None,
                'e2e': {'value': value, 'unit': 'candidates/s', 'ms_per_step': total_ms / args.steps,
                        'h2d_bytes_per_step': int(grasps / args.steps * 56), 'd2h_bytes_per_step': int(grasps / args.steps * (56 + 48 + 1)),
                        'note': 'this workload IS the end-to-end call: value and e2e coincide'},
                'gpu_launches': None, 'note': 'launch-latency bound: ~100 reference points per call; the throughput '
                                              'configurations are ags_c3 / ags_c4'}
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            c, dt, threads, _ = cpu_ags(ags.mesh, 2000)
            line['cpu_baseline'] = {'value': c / dt, 'unit': 'candidates/s', 'cores': threads, 'kind': 'port',
                                    'sample': f'2000 reference points x {N_RAYS} rays of the same mesh, oracle/coracle.c with OpenMP, {dt:.2f} s'}
    else:
        line = coverage_numbers(args, torch, metrics, flush, barrier, short=False, clocks=clocks, rank=rank, world=world,
                                dist=dist if world > 1 else None)
        clk = line.pop('clocks')
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            p, dt, threads, cov = cpu_coverage(args.cpu_coverage_n)
            line['cpu_baseline'] = {'value': p / dt, 'unit': 'pairs/s', 'cores': threads, 'kind': 'port',
                                    'sample': f'{args.cpu_coverage_n} x {args.cpu_coverage_n} grasps, kd-tree coverage '
                                              f'(scipy cKDTree + Python loop as metrics.py:161-229), {dt:.1f} s'}
    line.update({'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'higher_is_better': True,
                 'scaling': 'weak' if args.workload.startswith('ags') else 'strong',
                 'vs_baseline': None, 'data': 'synthetic', 'clocks': clk})
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def coverage_numbers(args, torch, metrics, flush, barrier, short, clocks=None, rank=0, world=1, dist=None):
    """coverage (perturbed pair, eps 15): device-resident pairs/s and e2e (host rows in, fraction out).  With several
    ranks the REFERENCE rows are sharded (query set replicated, one all-reduce of the covered count inside the step):
    the total work is fixed, i.e. strong scaling."""
    n_cov = 1_000_000 if args.workload == 'coverage_c5' else COV_N
    ref_all, qry = coverage_sets(n_cov)
    pairs = float(len(ref_all)) * float(len(qry))
    from burg_toolkit_b200.distributed import shard_range
    b, e = shard_range(len(ref_all), rank, world)
    ref = ref_all[b:e]
    dref, dqry = torch.from_numpy(ref).cuda(), torch.from_numpy(qry).cuda()

    def reduce_max(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    res = {}
    steps = max(3, args.steps // 2) if short else args.steps
    for algo in (('grid',) if n_cov > 200_000 else ('grid', 'tiles')):
        for _ in range(max(args.warmup, 3)):
            metrics.covered_mask(dref, dqry, algo=algo, return_device=True)
        barrier()
        if clocks is not None and algo == 'grid':
            clocks.start()
        ev = []
        for s in range(steps):
            flush.fill_(s & 0xff)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _, count = metrics.covered_mask(dref, dqry, algo=algo, return_device=True)
            if world > 1:
                dist.all_reduce(count, op=dist.ReduceOp.SUM)           # the path's only exchange: one int64 per rank
            e1.record()
            ev.append((e0, e1))
        barrier()
        ms = reduce_max(statistics.mean(event_ms(ev)))
        res[algo] = {'ms': ms, 'pairs_per_s': pairs / (ms * 1e-3), 'coverage': int(count.item()) / len(ref_all)}
    clk = clocks.stop() if clocks is not None else None
    ev = []
    for s in range(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        frac = metrics.coverage(ref, qry)                       # host rows in (pageable), python float out
        if world > 1:
            t = torch.tensor([frac * len(ref)], dtype=torch.float64, device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            frac = float(t.item()) / len(ref_all)
        e1.record()
        ev.append((e0, e1))
    barrier()
    e2e_ms = reduce_max(statistics.mean(event_ms(ev)))
    # FP32-issue roofline: 4 lane-ops per tested pair (3 FFMA + 1 compare).  The all-pairs kernel tests
    # every pair; the grid kernels test only the pairs of the 27 neighbouring cells, counted here on the host.
    lane_ops = 148 * 128 * 1.965e9
    if 'tiles' in res:
        roof_kernel, tested, roof_ms = 'coverage_tiles_kernel', pairs, res['tiles']['ms']
    else:
        cell = 1.01 * 15.0 / 1000.0
        lo = np.minimum(ref_all[:, :3].min(0), qry[:, :3].min(0))
        cr, cq = np.floor((ref_all[:, :3] - lo) / cell).astype(np.int64), np.floor((qry[:, :3] - lo) / cell).astype(np.int64)
        dims = np.maximum(cr.max(0), cq.max(0)) + 3
        occ = np.zeros(tuple(dims), dtype=np.int64)
        np.add.at(occ, tuple((cq + 1).T), 1)
        nb = sum(np.roll(occ, (dx, dy, dz), axis=(0, 1, 2)) for dx in (-1, 0, 1) for dy in (-1, 0, 1) for dz in (-1, 0, 1))
        tested = float(nb[tuple((cr + 1).T)].sum())
        cells = len(qry) / float(np.prod(dims - 2)) >= 16.0         # the library's kernel choice (queries per cell)
        roof_kernel = ('coverage_cells_kernel' if cells else 'coverage_grid_kernel') + ' (27-cell neighbourhood)'
        roof_ms = res['grid']['ms']
    # lane-ops per tested pair: 3 FFMA + 1 compare (translation reject first: tiles / grid kernels), 3 FFMA + 1 FMUL + 1 compare
    # (orientation bound first: cell-tile kernel)
    ops_per_pair = 5 if roof_kernel.startswith('coverage_cells') else 4
    achieved_ops = tested / world * ops_per_pair / (roof_ms * 1e-3)            # per GPU (every rank tests 1/world of the pairs)
    out = {'metric': 'coverage pairs/sec', 'value': res['grid']['pairs_per_s'], 'unit': 'pairs/s', 'dtype': 'f32',
           'ms_per_step': res['grid']['ms'],
           'config': {'workload': args.workload if not args.workload.startswith('ags') else 'coverage_c2', 'n_ref': len(ref_all), 'n_query': len(qry), 'epsilon': 15.0,
                      'sharding': f'reference rows split over {world} rank(s), query set replicated, all-reduce of the covered count',
                      'sets': 'random_poses x0.2 m cube; query = 50% uniform + 50% perturbed (5 mm / 5 deg)',
                      'semantics': 'metrics.coverage (kd-tree radius test)', 'algo': 'uniform grid (27 cells)',
                      'coverage': res['grid']['coverage'], 'l2': 'flushed between steps (256 MB write)'},
           'tiles_all_pairs': res.get('tiles'),
           'roofline': {'bound': 'fp32-issue', 'kernel': roof_kernel, 'achieved': achieved_ops / 1e12,
                        'peak': lane_ops / 1e12, 'unit': 'Tlane-op/s', 'frac': achieved_ops / lane_ops,
                        'pairs_tested_per_launch': tested, 'traffic': None,
                        'lane_ops_per_pair': ops_per_pair,
                        'note': 'FP32 lane-ops per tested pair: 4 (3 FFMA + compare) or 5 (quaternion dot + compare, cell-tile kernel); value counts nominal N*M pairs, the roofline '
                                'counts the pairs the kernel actually tests; HBM bytes are negligible'},
           'e2e': {'value': pairs / (e2e_ms * 1e-3), 'unit': 'pairs/s', 'h2d_bytes_per_step': int(ref.nbytes + qry.nbytes) * world,
                   'd2h_bytes_per_step': 8, 'ms_per_step': e2e_ms, 'coverage': frac},
           'gpu_launches': (10 if roof_kernel.startswith('coverage_cells') else 8) * steps, 'clocks': clk}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='ags_c3', choices=['ags_c1', 'ags_c3', 'ags_c4', 'coverage_c2', 'coverage_c5'])
    ap.add_argument('--cpu-ref-points', type=int, default=100_000, help='reference points per CPU sample (x100 rays)')
    ap.add_argument('--cpu-coverage-n', type=int, default=100_000)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)
    run_b200(args, rank, world, local_rank)


if __name__ == '__main__':
    main()

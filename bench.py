#!/usr/bin/env python
"""bench.py -- the headline benchmark of the AGS + coverage hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload ags_c3|ags_c1|ags_c4|ags_c3_exhaustive|coverage_c2|coverage_c5]

Prints ONE JSON line (rank 0).  The default workload (ags_c3, the configuration the metric is quoted on) also carries a
`secondary` list with one block per other BASELINE configuration (each with its own roofline / e2e / cpu_baseline), at
every N: coverage_c2, ags_c4, coverage_c5, ags_c3_exhaustive and (N = 1) ags_c1.  Workloads (BASELINE.json `configs`):
  ags_c3       config 3: AGS on the 99 904-triangle 'YCB-like' mesh, 10 000 reference points x 100 rays = 1 M
               candidates per step per GPU, max_targets_per_ref_point = 1, 12 orientations, gripper-vs-object
               collision + compaction.  metric: candidates/s.  (default; the config the metric's target is quoted on)
  ags_c1       config 1: AntipodalGraspSampler.sample(1000) + check_collisions on a 20 480-triangle icosphere (public API call).
  ags_c4       config 4: 20 objects (~1 M triangles) + ground slab, the target object changes every step, gripper vs scene.
  coverage_c2  config 2: coverage of two 100k grasp sets (perturbed pair).  metric: pairs/s.
  coverage_c5  config 5: 1 M x 1 M grasps; with several ranks the reference rows are sharded (strong scaling).
  ags_c3_exhaustive  config 3 with EVERY valid hit of a reference point expanded (max_targets_per_ref_point = 128, reference
               sampling.py:328-338): ~0.9 M contact pairs = ~11 M posed grippers per 1 M candidates -- the collision stage under load.

A "step" is one pass of the device pipeline over one batch of synthetic candidates.  The timed region holds
exactly K steps between one CUDA-event pair on the launching stream (the steps are issued round-robin on --streams
streams and joined), bracketed by a barrier + synchronize; the reported time is the max over ranks.  L2 between steps
(--l2): `rotate` (default) = inputs larger than L2, consecutive steps work on different objects (4 x ~50 MB at config 3);
`flush` = round 1's 256 MB fill before every step inside the timed region; the line carries both.

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
C3_OBJECTS = 4                          # distinct config-3 objects the steps rotate over (4 x ~50 MB > L2)
COV_N = 100_000
EXHAUSTIVE_TARGETS = 128


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
# CPU arm (oracle port: the reference's arithmetic in C with a BVH and OpenMP)
# ---------------------------------------------------------------------------------------------------------------
_CPU_MESHES = {}


def merged_scene(meshes):
    """vertices / faces of a list of meshes concatenated into one (the CPU port's scene for config 4)"""
    v = np.concatenate([m.vertices for m in meshes])
    off = np.cumsum([0] + [len(m.vertices) for m in meshes[:-1]])
    f = np.concatenate([m.triangles + o for m, o in zip(meshes, off)])
    return v, f


def cpu_ags(mesh, n_ref, seed=100, scene_meshes=None, threads=None):
    """one bounded CPU sample of the AGS workload -> (candidates, seconds, threads); the BVH builds are not timed"""
    from oracle import cport
    from burg_toolkit_b200.gripper import ParallelJawGripper
    key = (id(mesh), id(scene_meshes))
    if key not in _CPU_MESHES:
        cm = cport.Mesh(mesh.vertices, mesh.triangles, mesh.vertex_normals)
        cs = cm if scene_meshes is None else cport.Mesh(*merged_scene(scene_meshes))
        _CPU_MESHES[key] = (cm, cs)
    cm, cs = _CPU_MESHES[key]
    cport.set_num_threads(threads or os.cpu_count() or 1)   # all host threads by default, also under torchrun (OMP_NUM_THREADS=1 there)
    pts, nrm, dirs, fv = host_candidates_fast(mesh, n_ref, seed)
    gt = ParallelJawGripper().tcp_triangles()
    t0 = time.perf_counter()
    out = cport.pipeline(cm, cs, pts, dirs, fv, gt, n_orientations=N_ORIENT, coll_tol=0.01, max_hits_per_ray=16)
    dt = time.perf_counter() - t0
    return len(pts) * N_RAYS, dt, cport.num_threads(), out


def cpu_ags_record(mesh, n_ref_all, n_ref_one, scene_meshes=None, what='the same workload'):
    """cpu_baseline record of an AGS workload: all host threads (the line's value) and ONE thread (the reference itself is
    single-threaded Python), both on bounded samples"""
    c, dt, threads, _ = cpu_ags(mesh, n_ref_all, scene_meshes=scene_meshes)
    c1, dt1, _, _ = cpu_ags(mesh, n_ref_one, seed=101, scene_meshes=scene_meshes, threads=1)
    return {'value': c / dt, 'unit': 'candidates/s', 'cores': threads, 'kind': 'port',
            'sample': f'{n_ref_all} reference points x {N_RAYS} rays ({c} candidates) of {what}, oracle/coracle.c with OpenMP on '
                      f'{threads} threads, {dt:.1f} s',
            'one_thread': {'value': c1 / dt1, 'unit': 'candidates/s', 'cores': 1,
                           'sample': f'{n_ref_one} reference points x {N_RAYS} rays ({c1} candidates), 1 thread, {dt1:.1f} s'},
            'host_cpus': os.cpu_count()}


def cpu_coverage(n, n_ref_sample=None):
    """the reference's kd-tree coverage (scipy cKDTree + Python loop, metrics.py:161-229) restated in oracle/port.py; with
    ``n_ref_sample`` only that many reference rows are evaluated against the FULL query set (config 5: the whole job would
    take the better part of an hour and tens of GB of Python lists) and the pairs/s of the sample is reported"""
    from oracle import port
    ref, qry = coverage_sets(n)
    if n_ref_sample is not None and n_ref_sample < len(ref):
        ref = ref[np.random.default_rng(5).choice(len(ref), n_ref_sample, replace=False)]
    t0 = time.perf_counter()
    cov = port.covered_kd(ref, qry)
    dt = time.perf_counter() - t0
    return float(len(ref)) * len(qry), dt, 1, float(cov.mean())


def cpu_coverage_record(workload, n_sample_c2=100_000, n_ref_sample_c5=4000):
    if workload == 'coverage_c5':
        p, dt, threads, cov = cpu_coverage(1_000_000, n_ref_sample_c5)
        sample = (f'{n_ref_sample_c5} of the 1 000 000 reference rows against all 1 000 000 query rows (kd-tree coverage as metrics.py:161-229: scipy '
                  f'cKDTree + Python loop, 1 thread), {dt:.1f} s; the rows are independent, so the whole 1 M x 1 M job extrapolates to '
                  f'{dt * 1_000_000 / n_ref_sample_c5 / 60:.0f} min')
    else:
        p, dt, threads, cov = cpu_coverage(n_sample_c2)
        sample = f'{n_sample_c2} x {n_sample_c2} grasps, kd-tree coverage (scipy cKDTree + Python loop as metrics.py:161-229), 1 thread, {dt:.1f} s'
    return {'value': p / dt, 'unit': 'pairs/s', 'cores': threads, 'kind': 'port', 'sample': sample, 'coverage_of_sample': cov,
            'host_cpus': os.cpu_count()}


def ags_workload_meshes(workload):
    """-> (target meshes, scene meshes or None)"""
    if workload in ('ags_c3', 'ags_c3_exhaustive'):
        # C3_OBJECTS distinct objects of the same kind, one per step in turn: with the per-leaf records each is ~50 MB on the
        # device, so consecutive uses of one object are separated by > 126 MB (L2) of other objects and per-step buffers
        from burg_toolkit_b200 import mesh as bmesh
        return [c3_mesh()] + [bmesh.create_bumpy_sphere(224, 224, radius=0.035, amplitude=0.15, seed=1 + k) for k in range(1, C3_OBJECTS)], None
    if workload == 'ags_c1':
        from burg_toolkit_b200 import mesh as bmesh
        return [bmesh.create_icosphere(5, 0.03)], None
    scene_meshes = c4_scene().get_mesh_list()
    return scene_meshes[1:], scene_meshes


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm on the host cores.  The reference is pure Python on third-party
    leaves that cannot be installed here, so this arm runs the oracle port (oracle/coracle.c: same arithmetic, BVH broad
    phase, OpenMP over all host threads -- far faster than the reference's own Python would be) on bounded samples."""
    if rank != 0:
        return
    cores = os.cpu_count()
    if args.workload.startswith('ags'):
        targets, scene_meshes = ags_workload_meshes(args.workload)
        mesh = targets[0]
        n_ref = args.cpu_ref_points if args.workload != 'ags_c1' else 2000
        for _ in range(max(args.warmup, 0)):
            cpu_ags(mesh, max(8, n_ref // 8), scene_meshes=scene_meshes)
        times, cands, threads = [], 0, 1
        for s in range(args.steps):
            c, dt, threads, _ = cpu_ags(mesh, n_ref, seed=100 + s, scene_meshes=scene_meshes)
            times.append(dt)
            cands += c
        total = sum(times)
        c1, dt1, _, _ = cpu_ags(mesh, max(100, n_ref // 10), seed=99, scene_meshes=scene_meshes, threads=1)
        value, unit, metric = cands / total, 'candidates/s', 'AGS candidates/sec (ray+cone+collision)'
        sample = f'{n_ref} reference points x {N_RAYS} rays per step ({n_ref * N_RAYS} candidates) of the {args.workload} workload, {threads} threads'
        one_thread = {'value': c1 / dt1, 'unit': unit, 'cores': 1, 'sample': f'{c1} candidates on 1 thread, {dt1:.1f} s'}
        config = {'workload': args.workload, 'mesh_triangles': len(mesh.triangles), 'ref_points_per_step': n_ref,
                  'n_rays': N_RAYS, 'n_orientations': N_ORIENT, 'max_targets_per_ref_point': 1,
                  'collision': 'gripper-vs-object' if scene_meshes is None else
                  f'gripper-vs-scene ({sum(len(m.triangles) for m in scene_meshes)} triangles)'}
    else:
        times, pairs, threads = [], 0, 1
        for s in range(args.steps):
            if args.workload == 'coverage_c5':
                p, dt, threads, _ = cpu_coverage(1_000_000, 2000)
            else:
                p, dt, threads, _ = cpu_coverage(args.cpu_coverage_n)
            times.append(dt)
            pairs += p
        total = sum(times)
        value, unit, metric = pairs / total, 'pairs/s', 'coverage pairs/sec'
        sample = ('2000 reference rows x 1 000 000 query rows per step' if args.workload == 'coverage_c5' else
                  f'{args.cpu_coverage_n} x {args.cpu_coverage_n} grasps per step') + ' (kd-tree coverage, single-threaded like the reference)'
        one_thread = None
        config = {'workload': args.workload}
    line = {'impl': 'reference', 'metric': metric, 'value': value, 'unit': unit, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': 1e3 * total / max(args.steps, 1), 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64' if args.workload.startswith('ags') else 'f32',
            'data': 'synthetic', 'config': config,
            'cpu_baseline': {'value': value, 'unit': unit, 'cores': threads, 'kind': 'port', 'sample': sample,
                             'one_thread': one_thread, 'host_cpus': cores},
            'e2e': {'value': value, 'unit': unit, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
def event_ms(pairs):
    return [a.elapsed_time(b) for a, b in pairs]


class Ctx:
    """what every block needs: torch, the process group, the L2 flush buffer, reductions over ranks"""

    def __init__(self, args, rank, world, local_rank):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.rank, self.world, self.local_rank = rank, world, local_rank
        torch.cuda.set_device(local_rank)
        if world > 1:
            dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local_rank}'))
        self.peak_gbs, self.peak_src = measured_peaks()
        self.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')       # > 126 MB L2
        self._l2_gbs = None

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def _reduce(self, x, op):
        t = self.torch.tensor([x], dtype=self.torch.float64, device='cuda')
        if self.world > 1:
            self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(self, x):
        return self._reduce(x, self.dist.ReduceOp.MAX)

    def sum_over_ranks(self, x):
        return self._reduce(x, self.dist.ReduceOp.SUM)

    def l2_gbs(self):
        """measured L2 read bandwidth of this GPU (bt_debug_l2_probe: 32 MB buffer, all SMs, ld.global.cg)"""
        if self._l2_gbs is None:
            import ctypes
            from burg_toolkit_b200 import _native as nat
            out = ctypes.c_double()
            nat.call('bt_debug_l2_probe', self.local_rank, 32 << 20, 20, ctypes.byref(out))
            self._l2_gbs = float(out.value)
        return self._l2_gbs


def ncu_inputs():
    path = os.path.join(ROOT, 'profiles', 'roofline_inputs.json')
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh)
    return {}


def ags_block(ctx, workload, steps, warmup, with_cpu):
    """one AGS workload (config 3, config 3 exhaustive, config 4): device-resident candidates/s, per-stage times, roofline of
    the ray caster, end-to-end through the host API, multi-GPU gather (+ its bit-for-bit verification), CPU baseline"""
    torch, dist, rank, world, local_rank = ctx.torch, ctx.dist, ctx.rank, ctx.world, ctx.local_rank
    from burg_toolkit_b200 import sampling
    from burg_toolkit_b200 import mesh as bmesh
    from burg_toolkit_b200.gripper import ParallelJawGripper
    exhaustive = workload == 'ags_c3_exhaustive'
    ags = sampling.AntipodalGraspSampler()
    ags.gripper, ags.verbose, ags.device = ParallelJawGripper(), False, local_rank
    ags.n_orientations, ags.n_rays = N_ORIENT, N_RAYS
    ags.max_targets_per_ref_point = EXHAUSTIVE_TARGETS if exhaustive else 1
    if os.environ.get('BT_HOST_TRANSPORT'):                 # A/B knob: 'store' | 'dma'
        ags.host_transport = os.environ['BT_HOST_TRANSPORT']
    targets, scene_meshes = ags_workload_meshes(workload)
    mesh = targets[0]
    pipe_kw = {} if scene_meshes is None else {'additional_objects': scene_meshes, 'exclude_shape': True}
    scene_tris = sum(len(m.triangles) for m in scene_meshes) if scene_meshes else len(mesh.triangles)
    ags.mesh = bmesh.create_icosphere(1, 0.03)               # first build in the process also loads the CUDA module
    ags._device_mesh()
    build_ms = []
    for t in targets:                                       # build every target's LBVH before the timed region
        ags.mesh = t
        build_ms.append(ags._device_mesh().info.build_ms)
    ags.mesh = mesh
    info = ags._device_mesh().info
    bvh_kind, bvh_rounds, walk_kind = ags._device_mesh().bvh_kind
    cap_rows = N_REF * ags.max_targets_per_ref_point * N_ORIENT
    # the path's only exchange: compacted grasp rows + contacts + counts to rank 0.  Every rank's compaction
    # kernel stores its kept rows straight into its slot of rank 0's buffer over NVLink (distributed.PeerGather, CUDA IPC).
    # If the box refuses the IPC mapping the bench says so and uses a plain NCCL gather of the padded rows.
    peer, gather_buf, exchange = None, None, 'none (1 GPU)'
    if world > 1 and exhaustive:
        exchange = 'none: every rank keeps its ~10 M rows (the gather buffer of this mode would hold 3 x world x 1.6 GB on rank 0)'
    elif world > 1 and os.environ.get('BT_BENCH_NO_EXCHANGE') == '1':              # diagnostic: what the exchange costs
        exchange = 'none: disabled by BT_BENCH_NO_EXCHANGE (diagnostic run, every rank keeps its rows)'
    elif world > 1:
        from burg_toolkit_b200 import distributed as bdist
        from burg_toolkit_b200._native import NativeError
        try:
            mode = os.environ.get('BT_PEER_ONE_SIDED')                  # A/B knob; default: one-sided from 4 ranks on
            peer = bdist.PeerGather(cap_rows, with_contacts=True, device=local_rank, one_sided=None if mode is None else mode != '0', timeout_s=30.0)
            exchange = (('compaction into local memory, then the exchange stream moves rows + contacts into rank 0 over NVLink (CUDA IPC peer buffer; ' + ('copy engine, padded to the capacity' if peer.push_blocks == 0 else f'bt_peer_push, {peer.push_blocks} blocks, exactly the kept rows') + '); '
                         if peer.staged else 'compaction kernel stores rows + contacts into rank 0 over NVLink (CUDA IPC peer buffer); ') +
                        ('counts + epoch flags stored the same way, rank 0 polls them (no collective in the step)' if peer.one_sided
                         else 'NCCL all_gather of counts'))
        except NativeError as ex:
            exchange = f'NCCL all_reduce + gather of padded rows (peer mapping unavailable: {ex})'
            gather_buf = torch.empty((world, cap_rows, 14), dtype=torch.float32, device='cuda') if rank == 0 else None

    from burg_toolkit_b200.device import StageTimer
    stage_timer = StageTimer(local_rank)

    # Steps are independent batches: they are issued round-robin on `n_streams` streams, so the latency-bound tail of one
    # batch (target choice, expansion, compaction -- and a compaction whose stores cross NVLink / PCIe) overlaps the ray
    # casting of the next.  Each stream owns a workspace (slot) and an L2-flush buffer.
    n_streams = max(1, int(ctx.args.streams))
    streams = [torch.cuda.Stream() for _ in range(n_streams)]
    # L2 between steps (timing rules: "flush L2 or use inputs larger than L2"):
    #   rotate (default): every step works on the next of several distinct objects (config 3: C3_OBJECTS x ~50 MB of mesh,
    #           LBVH and per-leaf records; config 4: 20 targets in a 340 MB scene), so an object's data has been displaced by
    #           > 126 MB of other objects' data and per-step buffers before it is used again -- nothing but the path is timed;
    #   flush:  round 1's method, a 256 MB fill on the step's stream inside the timed region (reported as l2_flush_variant).
    l2_mode = ctx.args.l2
    l2_desc = ('a 256 MB buffer is written before every step, on the step\'s stream, inside the timed region' if l2_mode == 'flush' else
               f'inputs larger than L2: consecutive steps work on different objects ({len(targets)} objects in turn' +
               (f', ~50 MB of mesh + LBVH + per-leaf records each' if scene_meshes is None else f', inside a scene of {scene_tris} triangles = 340 MB') +
               '), so an object is displaced from L2 by the other objects and > 100 MB of per-step buffers before it is used again; no flush kernel in the timed region')
    flushes = [ctx.flush] + [torch.empty_like(ctx.flush) for _ in range(n_streams - 1)]

    def step(seed, timer=None, graph=True, to_peer=True, slot=0):
        ags.mesh = targets[seed % len(targets)]
        dest = peer.destinations(slot=slot) if (peer is not None and to_peer) else None
        out = ags.run_pipeline(n_ref=N_REF, seed=seed, check_collisions=True, timer=timer, destinations=dest, slot=slot,
                               graph=graph and timer is None, **pipe_kw)
        if peer is not None and to_peer:
            peer.complete(out['n_free'], rows=out.get('free_gs'), contacts=out.get('free_contacts'), slot=slot)
        elif world > 1 and gather_buf is not None and to_peer:
            dist.all_reduce(out['n_free'], op=dist.ReduceOp.SUM)
            dist.gather(out['free_gs'], list(gather_buf.unbind(0)) if rank == 0 else None, dst=0)
        return out

    issue_s = [0.0]

    def run_steps(n, timed, flush=(l2_mode == 'flush')):
        """n steps round-robin over the streams (flush: a 256 MB fill before every step on the step's stream) -> (events, stats)"""
        main = torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stats = []
        e0.record(main)
        for st in streams:
            st.wait_event(e0)
        t_issue = time.perf_counter()
        for s in range(n):
            k = s % n_streams
            with torch.cuda.stream(streams[k]):
                if flush:
                    flushes[k].fill_(s & 0xff)                                   # > L2, written between consecutive steps
                out = step(7919 * (rank + 1) + s, slot=k)
                if timed:
                    stats.append((out['n_grasps'].clone(), (out['n_free'] if out.get('n_free') is not None else out['n_grasps']).clone(), out['overflow'].clone()))
        issue_s[0] = time.perf_counter() - t_issue          # host time spent issuing the n steps (no waits in there)
        for st in streams:
            main.wait_stream(st)
        e1.record(main)
        return (e0, e1), stats

    clocks = ClockSampler(local_rank)
    clocks.start()                                          # samples cover warm-up + timed region (all under load)
    # warm-up = the timed sequence itself (same seeds, hence same target meshes and stream slots), a multiple of the gather
    # ring's depth long, so that every CUDA graph the timed region replays has been captured
    ring = peer.depth if peer is not None else 1
    n_warm = max(warmup, steps)
    n_warm = ((n_warm + ring * n_streams - 1) // (ring * n_streams)) * (ring * n_streams)
    run_steps(n_warm, False)
    ctx.barrier()
    (e0, e1), stats = run_steps(steps, True)
    ctx.barrier()
    clk = clocks.stop()
    my_ms = e0.elapsed_time(e1)
    total_ms = ctx.max_over_ranks(my_ms)
    host_issue_ms = ctx.max_over_ranks(issue_s[0] * 1e3) / steps
    rank_ms, clk_by_rank = [my_ms / steps], [clk.get('sm_mhz')]
    if world > 1:
        rank_ms, clk_by_rank = [None] * world, [None] * world
        dist.all_gather_object(rank_ms, my_ms / steps)
        dist.all_gather_object(clk_by_rank, clk.get('sm_mhz'))
    # the other L2 method on the same steps, for continuity with round 1 (first re-align the gather ring with the step
    # sequence, so that the same captured graphs are replayed)
    while peer is not None and peer.epoch % ring:
        step(1)
    (f0, f1), _ = run_steps(steps, False, flush=(l2_mode != 'flush'))
    ctx.barrier()
    other_ms = ctx.max_over_ranks(f0.elapsed_time(f1))
    # the same steps one at a time on one stream (continuity with round 1, and what the per-stage times below add up to)
    single_events = []
    for s in range(min(steps, 5)):
        ctx.flush.fill_(s & 0xff)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step(31337 * (rank + 1) + s, to_peer=False)
        b.record()
        single_events.append((a, b))
    torch.cuda.synchronize()
    single_ms = ctx.max_over_ranks(statistics.median(event_ms(single_events)))
    # launches of one step, counted from the captured graph (kernel nodes), not kept by hand
    ws = ags._workspaces[0]
    launch = {'nodes': None, 'kernels': None, 'names': None}
    if getattr(ws, 'last_graph', None) is not None:
        n_nodes, n_kernels, names = ws.last_graph.info()
        launch = {'nodes': n_nodes, 'kernels': n_kernels, 'names': names}
    # per-stage device times: a few extra, untimed steps with the stage timer (its event records and the host wait per
    # step are instrumentation, so they stay out of the timed region above)
    stage_acc = {}
    for s in range(min(steps, 5) + 1):
        ctx.flush.fill_(s & 0xff)
        torch.cuda.synchronize()
        step(104729 * (rank + 1) + s, stage_timer, to_peer=False)
        for name, ms in stage_timer.read().items():
            if s > 0:                                       # the first plain (not graph-replayed) call grows the scratch pool
                stage_acc.setdefault(name, []).append(ms)
    candidates = ctx.sum_over_ranks(N_REF * N_RAYS * steps)
    value = candidates / (total_ms * 1e-3)
    stage_ms = {k: statistics.median(v) for k, v in stage_acc.items()}
    n_grasps = int(torch.stack([a.reshape(()) for a, _, _ in stats]).double().mean().item())
    n_free = int(torch.stack([b.reshape(()) for _, b, _ in stats]).double().mean().item()) // (world if (peer is None and gather_buf is not None) else 1)
    overflow = int(torch.stack([c.reshape(()) for _, _, c in stats]).sum().item())

    # work counters (instrumented kernel instantiations, one extra untimed step) for the traffic model
    from burg_toolkit_b200 import _native as nat
    counters = torch.zeros(8, dtype=torch.int64, device='cuda')
    nat.call('bt_debug_set_counters', counters.data_ptr())
    step(424242 + rank, graph=False, to_peer=False)
    torch.cuda.synchronize()
    nat.call('bt_debug_set_counters', None)
    cnt = [int(x) for x in counters.cpu().tolist()]
    n_cand = N_REF * N_RAYS
    per_ray = {'nodes_fetched': cnt[0] / n_cand, 'nodes_skipped_by_normal_cone': cnt[6] / n_cand, 'leaf_pretests': cnt[1] / n_cand,
               'exact_tests': cnt[2] / n_cand}
    per_grasp = {'nodes': cnt[3] / max(n_grasps, 1), 'leaves_classified': cnt[4] / max(n_grasps, 1),
                 'leaves_exact': cnt[5] / max(n_grasps, 1)}

    # roofline of the dominant stage (the ray caster = traverse_kernel + cast_kernel): algorithmic HBM bytes per launch
    # (64 B per candidate + 80 B per triangle) over its event-timed duration, against the measured HBM peak.
    # The mesh + LBVH are L2-resident by design, so HBM is not what binds: `binding` states the roofs that do.
    cast_ms = stage_ms['cast']
    alg_bytes = N_REF * N_RAYS * 64 + 80 * info.n_triangles
    achieved = alg_bytes / (cast_ms * 1e-3) / 1e9
    model_bytes = cnt[0] * 64 + cnt[1] * (48 + 4) + cnt[2] * 128 + n_cand * (24 + 48 + 2 * 8 + 8)
    prof = ncu_inputs()
    l2_peak = ctx.l2_gbs()
    trav = prof.get('traverse_kernel', {})
    roofline = {'bound': 'hbm', 'kernel': 'traverse_kernel + cast_kernel (ray caster stage)', 'achieved': achieved, 'peak': ctx.peak_gbs,
                'unit': 'GB/s', 'frac': achieved / ctx.peak_gbs, 'traffic': prof.get('cast_stage', {}).get('dram_bytes_per_launch'),
                'traffic_source': prof.get('cast_stage', {}).get('source'),
                'peak_source': ctx.peak_src, 'algorithmic_bytes_per_launch': alg_bytes, 'kernel_ms': cast_ms,
                'rays_per_s': N_REF * N_RAYS / (cast_ms * 1e-3),
                'binding': {'what': 'L1 data-pipe wavefronts and ALU issue of divergent 64-byte node fetches over an L2-resident tree (not HBM)',
                            'traffic_model_bytes_per_launch': model_bytes, 'traffic_model_gbs': model_bytes / (cast_ms * 1e-3) / 1e9,
                            'l2_peak_gbs_measured': l2_peak, 'frac_of_l2_peak': model_bytes / (cast_ms * 1e-3) / 1e9 / l2_peak,
                            'definition': 'nodes_fetched*64 + leaf_pretests*(48 + 4) + exact_tests*128 + rays*(24 B float32 ray in traverse + 48 B float64 ray in '
                                          'resolve + 2*8 B ray parameters + 8 B count and mask out); l2 peak = bt_debug_l2_probe (32 MB, ld.global.cg, all SMs)',
                            'ncu_traverse_kernel': {k: trav.get(k) for k in ('l1tex_data_pipe_lsu_wavefronts_pct', 'alu_pipe_pct', 'issue_active_pct',
                                                                            'duration_us', 'source')},
                            'per_ray': per_ray},
                'collide_per_grasp': per_grasp,
                'note': 'mesh + LBVH + per-leaf records (~50 MB at 100k triangles) are L2-resident by design, so the HBM fraction is small by construction; '
                        'see binding and DESIGN.md section 4'}

    # end-to-end through the public API with HOST buffers: pinned staging, H2D, pipeline, the compaction kernel's stores
    # into pinned host memory, all inside the timed region; timed on the host's wall clock around synchronised calls.
    #   serial:    evaluate_host, one batch at a time (L2 flushed between batches, outside the clock)
    #   pipelined: submit_host on two slots, batch i+1 submitted before batch i's result is awaited (its PCIe stores overlap
    #              the next batch's ray casting); L2 is NOT flushed between overlapping batches (they stream > 100 MB each)
    e2e = None
    if not exhaustive:
        # batch b works on object b % len(targets) (the same rotation as the device-resident steps)
        n_hb = 4 if len(targets) == 1 else len(targets)
        host_batches = []
        for b in range(n_hb):
            ags.mesh = targets[b % len(targets)]
            host_batches.append([t.cpu().numpy() for t in ags.sample_surface(N_REF, 4242 + rank + 17 * b)[:2]])

        def host_call(fn, b, **kw):
            ags.mesh = targets[b % len(targets)]
            return fn(*host_batches[b % n_hb], **kw, **pipe_kw)
        n_slots = max(2, int(ctx.args.e2e_slots))
        for w in range(max(warmup, 2, n_hb) * n_slots):            # every (object, slot) pair of the timed loops: graphs captured
            if w % n_slots == 0:
                host_call(ags.evaluate_host, w // n_slots, seed=w + 1)
            host_call(ags.submit_host, w // n_slots, seed=w + 1, slot=w % n_slots).result()
        ctx.barrier()
        serial_s, h2d, d2h = 0.0, 0, 0
        for s in range(steps):
            if l2_mode == 'flush':
                ctx.flush.fill_(s & 0xff)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res = host_call(ags.evaluate_host, s, seed=s + 11)
            serial_s += time.perf_counter() - t0
            h2d, d2h = res['h2d_bytes'], res['d2h_bytes']
        ctx.barrier()
        from collections import deque
        t0 = time.perf_counter()
        pending, got = deque(), 0
        for s in range(steps):
            if len(pending) == n_slots:                             # a slot is resubmitted only after its result has been taken
                got += pending.popleft().result()['n_free']
            pending.append(host_call(ags.submit_host, s, seed=s + 11, slot=s % n_slots))
        while pending:
            got += pending.popleft().result()['n_free']
        pipelined_s = time.perf_counter() - t0
        ctx.barrier()
        pipe_ms = ctx.max_over_ranks(pipelined_s * 1e3)
        serial_ms = ctx.max_over_ranks(serial_s * 1e3)
        e2e = {'value': candidates / (pipe_ms * 1e-3), 'unit': 'candidates/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
               'ms_per_step': pipe_ms / steps, 'clock': 'host wall clock (time.perf_counter) around submit ... result',
               'mode': f'AntipodalGraspSampler.submit_host on {n_slots} slots in turn (a slot is resubmitted after its result has been taken, so up to {n_slots} batches are in flight); '
                       'every batch: pinned staging copy, H2D, graph-replayed pipeline, ' + ('result buffers (padded to the capacity) moved into pinned host memory by a copy engine' if ags.host_transport == 'dma' else 'kept rows stored into pinned host memory by the compaction kernel'),
               'serial': {'value': candidates / (serial_ms * 1e-3), 'ms_per_step': serial_ms / steps,
                          'mode': 'AntipodalGraspSampler.evaluate_host, one synchronous call per batch (' + l2_desc + ')'},
               'rows_received_per_step': got / steps}

    # multi-GPU: does rank 0 hold, bit for bit, what the ranks computed?  One more step into the peer buffer, the same step
    # (same seed: the pipeline is deterministic) into local memory, the local rows sent through a plain NCCL gather.
    gather_verified = None
    exchange_timed_out = False
    if peer is not None:
        from burg_toolkit_b200 import distributed as bdist
        vseed = 555 + 31 * rank
        step(vseed)
        got_rows, got_contacts = peer.result()
        local = step(vseed, to_peer=False)
        k = int(local['n_free'].item())
        want_rows = bdist.gather_rows(local['free_gs'][:k].contiguous(), 0)
        want_contacts = bdist.gather_rows(local['free_contacts'][:k].contiguous(), 0)
        ok = 1.0
        if rank == 0:
            ok = float(got_rows.shape == want_rows.shape and torch.equal(got_rows, want_rows) and torch.equal(got_contacts, want_contacts)
                       and want_rows.shape[0] > 0)
        gather_verified = ctx.sum_over_ranks(ok if rank == 0 else 0.0) == 1.0
        exchange_timed_out = ctx.sum_over_ranks(1.0 if peer.close(raise_on_timeout=False) else 0.0) > 0

    mesh_desc = {'ags_c3': f'bumpy spheres 224x224 (YCB-like, seeds 1-{C3_OBJECTS}), one per step in turn', 'ags_c3_exhaustive': f'bumpy spheres 224x224 (YCB-like, seeds 1-{C3_OBJECTS}), one per step in turn',
                 'ags_c4': '20 bumpy objects (158x158 grids, seeds 10-29) on a 1.0 x 0.5 m ground slab; target object changes every step'}[workload]
    block = {'metric': 'AGS candidates/sec (ray+cone+collision)', 'value': value, 'unit': 'candidates/s',
             'ms_per_step': total_ms / steps, 'dtype': 'f64', 'scaling': 'weak',
             'config': {'workload': workload, 'mesh': mesh_desc,
                        'mesh_triangles': int(info.n_triangles), 'scene_triangles': int(scene_tris),
                        'ref_points_per_step_per_gpu': N_REF, 'n_rays': N_RAYS,
                        'candidates_per_step_per_gpu': N_REF * N_RAYS, 'n_orientations': N_ORIENT,
                        'max_targets_per_ref_point': int(ags.max_targets_per_ref_point),
                        'collision': ('gripper-vs-object' if scene_meshes is None else 'gripper-vs-scene') + ', use_width, tol 0.01',
                        'mode': 'exhaustive (every valid hit of a reference point becomes a contact pair)' if exhaustive else
                                'reference mode (one pair per reference point)',
                        'l2': l2_desc,
                        'l2_other_method': {'method': 'rotate' if l2_mode == 'flush' else 'flush (256 MB written before every step, on the step\'s stream, inside the timed region: round 1\'s method)',
                                            'ms_per_step': other_ms / steps, 'value': candidates / (other_ms * 1e-3)},
                        'launch': 'bt_ags_pipeline replayed as a CUDA graph (bt_graph_*)', 'streams': n_streams,
                        'timing': f'one CUDA-event pair around all {steps} steps (issued round-robin on {n_streams} streams, joined on the launching stream)',
                        'single_stream_ms_per_step': single_ms, 'host_issue_ms_per_step': host_issue_ms,
                        'ms_per_step_by_rank': [round(x, 4) for x in rank_ms], 'sm_mhz_by_rank': clk_by_rank,
                        'bvh_build_ms': float(np.mean(build_ms)), 'bvh_max_depth': int(info.max_depth),
                        'bvh_build': 'ray tree (target object): ' + {'lbvh': 'Karras hierarchy of the Morton codes', 'ploc': f'PLOC (radius 8, {bvh_rounds} merge rounds) on the Morton order'}[bvh_kind]
                                     + ', 4-wide collapse of every other level, normal cones; walk tree of the target object (binary nodes): ' + walk_kind
                                     + ' (the collision scene of config 4 is a mesh of its own: PLOC from 2^18 triangles on)',
                        'posed_grasps_per_step': n_grasps, 'collision_free_per_step': n_free, 'hit_list_overflows': overflow,
                        'exchange': exchange, 'exchange_timed_out': exchange_timed_out},
             'roofline': roofline, 'stage_ms': stage_ms, 'e2e': e2e,
             'gpu_launches': None if launch['kernels'] is None else launch['kernels'] * steps,
             'launches_per_step': {'kernel_nodes': launch['kernels'], 'graph_nodes': launch['nodes'], 'kernels': launch['names'],
                                   'source': 'bt_graph_info of the captured step (kernel nodes of the CUDA graph)'},
             'clocks': clk}
    if exhaustive:
        block['collision_stage'] = {'posed_grippers_per_s_per_gpu': n_grasps / (stage_ms['collide'] * 1e-3), 'ms': stage_ms['collide'],
                                    'posed_grippers_per_step': n_grasps}
    if gather_verified is not None:
        block['gather_verified'] = bool(gather_verified)
    if with_cpu and rank == 0 and world == 1:
        n_all = ctx.args.cpu_ref_points if scene_meshes is None else ctx.args.cpu_ref_points // 2
        block['cpu_baseline'] = cpu_ags_record(mesh, n_all, max(1000, n_all // 20), scene_meshes, what=f'the {workload} workload (one pair per reference point)')
    del ags
    return block


def ags_c1_block(ctx, steps, warmup, with_cpu):
    """config 0 of BASELINE.json: the reference's own CPU-runnable case through the public API --
    AntipodalGraspSampler.sample(1000) on a 20 480-triangle icosphere, 12 orientations, then check_collisions"""
    torch, rank, world, local_rank = ctx.torch, ctx.rank, ctx.world, ctx.local_rank
    from burg_toolkit_b200 import mesh as bmesh, sampling
    from burg_toolkit_b200.gripper import ParallelJawGripper
    ags = sampling.AntipodalGraspSampler()
    ags.mesh, ags.gripper, ags.verbose, ags.device = bmesh.create_icosphere(5, 0.03), ParallelJawGripper(), False, local_rank
    ags.n_orientations, ags.n_rays = N_ORIENT, N_RAYS
    np.random.seed(1234 + rank)
    clocks = ClockSampler(local_rank)
    clocks.start()
    for _ in range(warmup):
        gs, _ = ags.sample(1000)
        ags.check_collisions(gs)
    ctx.barrier()
    wall, cands, grasps = 0.0, 0, 0
    for s in range(steps):
        ctx.flush.fill_(s & 0xff)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        gs, contacts = ags.sample(1000)                    # host call: numpy results come back every time
        coll = ags.check_collisions(gs)
        wall += time.perf_counter() - t0
        cands += ags.last_stats['candidates']
        grasps += len(gs)
    ctx.barrier()
    clk = clocks.stop()
    total_ms = ctx.max_over_ranks(wall * 1e3)
    cand_all = ctx.sum_over_ranks(cands)
    value = cand_all / (total_ms * 1e-3)
    block = {'metric': 'AGS candidates/sec (ray+cone+collision)', 'value': value, 'unit': 'candidates/s',
             'ms_per_step': total_ms / steps, 'dtype': 'f64', 'scaling': 'weak',
             'config': {'workload': 'ags_c1', 'mesh': 'icosphere, 5 subdivisions, r = 0.03', 'mesh_triangles': 20480,
                        'call': 'AntipodalGraspSampler.sample(1000) + check_collisions(gs), numpy in / numpy out',
                        'n_rays': N_RAYS, 'n_orientations': N_ORIENT, 'max_targets_per_ref_point': 1,
                        'candidates_per_step_per_gpu': cands / steps, 'grasps_per_step': grasps / steps,
                        'l2': 'flushed between steps (256 MB write)', 'clock': 'host wall clock around the two calls'},
             'roofline': None,
             'e2e': {'value': value, 'unit': 'candidates/s', 'ms_per_step': total_ms / steps,
                     'h2d_bytes_per_step': int(grasps / steps * 56), 'd2h_bytes_per_step': int(grasps / steps * (56 + 48 + 1)),
                     'note': 'this workload IS the end-to-end call: value and e2e coincide'},
             'gpu_launches': None, 'clocks': clk,
             'note': 'launch-latency bound: ~100 reference points per call; the throughput configurations are ags_c3 / ags_c4'}
    if with_cpu and rank == 0 and world == 1:
        block['cpu_baseline'] = cpu_ags_record(ags.mesh, 2000, 500, what='the same mesh')
    return block


def coverage_block(ctx, workload, steps, warmup, with_cpu):
    """coverage (perturbed pair, eps 15): device-resident pairs/s and e2e (host rows in, fraction out).  With several
    ranks the REFERENCE rows are sharded (query set replicated, one all-reduce of the covered count inside the step):
    the total work is fixed, i.e. strong scaling."""
    torch, dist, rank, world = ctx.torch, ctx.dist, ctx.rank, ctx.world
    from burg_toolkit_b200 import metrics
    from burg_toolkit_b200.distributed import shard_range, spatial_order
    n_cov = 1_000_000 if workload == 'coverage_c5' else COV_N
    ref_all, qry = coverage_sets(n_cov)
    pairs = float(len(ref_all)) * float(len(qry))
    b, e = shard_range(len(ref_all), rank, world)
    # shards = spatial tiles of the pair matrix: the reference rows in grid-cell order, one contiguous slice per rank
    # (input partitioning, done once on the host; coverage does not depend on the row order)
    ref = ref_all[spatial_order(ref_all)[b:e]] if world > 1 else ref_all
    dref, dqry = torch.from_numpy(ref).cuda(), torch.from_numpy(qry).cuda()
    clocks = ClockSampler(ctx.local_rank)
    res = {}
    # 'grid': the query set prepared once (resident, binned, sorted: metrics.prepare_query) -- the sweep of BASELINE config 5
    # evaluates many reference sets / shards against one query set; 'grid_unprepared': every call bins the queries again
    t0 = time.perf_counter()
    prepared = metrics.prepare_query(dqry, 15.0)
    torch.cuda.synchronize()
    prepare_ms = (time.perf_counter() - t0) * 1e3
    for algo in (('grid', 'grid_unprepared') if n_cov > 200_000 else ('grid', 'grid_unprepared', 'tiles')):
        def call():
            if algo == 'grid':
                return metrics.covered_mask(dref, prepared, return_device=True)
            return metrics.covered_mask(dref, dqry, algo='grid' if algo == 'grid_unprepared' else algo, return_device=True)
        for _ in range(max(warmup, 3)):
            call()
        ctx.barrier()
        if algo == 'grid':
            clocks.start()
        ev = []
        for s in range(steps):
            ctx.flush.fill_(s & 0xff)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _, count = call()
            if world > 1:
                dist.all_reduce(count, op=dist.ReduceOp.SUM)           # the path's only exchange: one int64 per rank
            e1.record()
            ev.append((e0, e1))
        ctx.barrier()
        if algo == 'grid':
            clk = clocks.stop()
        ms = ctx.max_over_ranks(statistics.mean(event_ms(ev)))
        res[algo] = {'ms': ms, 'pairs_per_s': pairs / (ms * 1e-3), 'coverage': int(count.item()) / len(ref_all)}
    # end to end: host reference rows in (this rank's shard, through reusable pinned staging), python float out, against
    # the prepared query set; 'unprepared' = metrics.coverage(ref, qry) with both sets as pageable host arrays
    pinned, wall = {}, 0.0
    metrics.coverage(ref, prepared, pinned=pinned)
    ctx.barrier()
    for s in range(steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        frac = metrics.coverage(ref, prepared, pinned=pinned)
        if world > 1:
            t = torch.tensor([frac * len(ref)], dtype=torch.float64, device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            frac = float(t.item()) / len(ref_all)
        wall += time.perf_counter() - t0
    ctx.barrier()
    e2e_ms = ctx.max_over_ranks(wall * 1e3 / steps)
    wall = 0.0
    for s in range(min(steps, 3)):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        metrics.coverage(ref, qry)
        wall += time.perf_counter() - t0
    ctx.barrier()
    e2e_unprepared_ms = ctx.max_over_ranks(wall * 1e3 / min(steps, 3))
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
    block = {'metric': 'coverage pairs/sec', 'value': res['grid']['pairs_per_s'], 'unit': 'pairs/s', 'dtype': 'f32',
             'ms_per_step': res['grid']['ms'], 'scaling': 'strong',
             'config': {'workload': workload, 'n_ref': len(ref_all), 'n_query': len(qry), 'epsilon': 15.0,
                        'sharding': f'reference rows split over {world} rank(s) as spatial tiles (contiguous slices of the grid-cell order, distributed.spatial_order), '
                                    'query set replicated, all-reduce of the covered count',
                        'sets': 'random_poses x0.2 m cube; query = 50% uniform + 50% perturbed (5 mm / 5 deg)',
                        'semantics': 'metrics.coverage (kd-tree radius test)', 'algo': 'uniform grid (27 cells), query set prepared once (metrics.prepare_query)',
                        'unprepared_ms_per_step': res['grid_unprepared']['ms'],
                        'coverage': res['grid']['coverage'], 'l2': 'flushed between steps (256 MB write)'},
             'tiles_all_pairs': res.get('tiles'),
             'roofline': {'bound': 'fp32-issue', 'kernel': roof_kernel, 'achieved': achieved_ops / 1e12,
                          'peak': lane_ops / 1e12, 'unit': 'Tlane-op/s', 'frac': achieved_ops / lane_ops,
                          'pairs_tested_per_launch': tested, 'traffic': None,
                          'lane_ops_per_pair': ops_per_pair,
                          'note': 'FP32 lane-ops per tested pair: 4 (3 FFMA + compare) or 5 (quaternion dot + compare, cell-tile kernel); value counts nominal N*M pairs, the roofline '
                                  'counts the pairs the kernel actually tests; HBM bytes are negligible'},
             'e2e': {'value': pairs / (e2e_ms * 1e-3), 'unit': 'pairs/s', 'h2d_bytes_per_step': int(ref.nbytes) * world,
                     'd2h_bytes_per_step': 8, 'ms_per_step': e2e_ms, 'coverage': frac,
                     'clock': 'host wall clock around metrics.coverage(host reference rows, prepared query set)',
                     'query_prepare_ms_once': prepare_ms,
                     'unprepared': {'ms_per_step': e2e_unprepared_ms, 'h2d_bytes_per_step': int(ref.nbytes + qry.nbytes) * world,
                                    'mode': 'metrics.coverage(ref, qry): both sets pageable host arrays, queries uploaded and binned per call'}},
             'gpu_launches': (10 if roof_kernel.startswith('coverage_cells') else 8) * steps, 'clocks': clk}
    if with_cpu and rank == 0 and world == 1:
        block['cpu_baseline'] = cpu_coverage_record(workload, ctx.args.cpu_coverage_n)
    del dref, dqry
    return block


def run_block(ctx, workload, steps, warmup, with_cpu):
    if workload == 'ags_c1':
        return ags_c1_block(ctx, steps, warmup, with_cpu)
    if workload.startswith('ags'):
        return ags_block(ctx, workload, steps, warmup, with_cpu)
    return coverage_block(ctx, workload, steps, warmup, with_cpu)


def run_b200(args, rank, world, local_rank):
    ctx = Ctx(args, rank, world, local_rank)
    with_cpu = not args.no_cpu_baseline
    line = run_block(ctx, args.workload, args.steps, args.warmup, with_cpu)
    clk = line.pop('clocks')
    if args.workload == 'ags_c3' and not args.no_secondary:
        # the other BASELINE configurations, every N (the driver's scaling runs carry config 4 and config 5 this way)
        short = max(3, args.steps // 2)
        line['secondary'] = []
        for wl in ['coverage_c2', 'ags_c4', 'coverage_c5', 'ags_c3_exhaustive'] + (['ags_c1'] if world == 1 else []):
            ctx.torch.cuda.empty_cache()
            try:
                line['secondary'].append(run_block(ctx, wl, short, min(args.warmup, 3), with_cpu))
            except Exception as ex:             # a secondary block must not cost the headline line
                line['secondary'].append({'config': {'workload': wl}, 'error': f'{type(ex).__name__}: {ex}'})
                if world > 1:
                    raise
    line.update({'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'higher_is_better': True,
                 'vs_baseline': None, 'data': 'synthetic', 'clocks': clk})
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        ctx.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=40)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='ags_c3', choices=['ags_c1', 'ags_c3', 'ags_c3_exhaustive', 'ags_c4', 'coverage_c2', 'coverage_c5'])
    ap.add_argument('--cpu-ref-points', type=int, default=100_000, help='reference points per CPU sample (x100 rays)')
    ap.add_argument('--cpu-coverage-n', type=int, default=100_000)
    ap.add_argument('--streams', type=int, default=4, help='AGS workloads: batches in flight per GPU (round-robin streams)')
    ap.add_argument('--e2e-slots', type=int, default=8, help='AGS workloads: host batches in flight in the pipelined end-to-end measurement')
    ap.add_argument('--l2', default='rotate', choices=['rotate', 'flush'], help='AGS workloads: how L2 is kept cold between steps')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-secondary', action='store_true', help='ags_c3 only: skip the blocks of the other configurations')
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

"""Antipodal grasp sampling on the B200.

Host-side mirror of the reference's ``burg_toolkit/sampling.py`` for the hot path: the same class
(``AntipodalGraspSampler`` with the same public attributes and defaults, sampling.py:57-70), the same
methods (``sample`` :185-352, ``check_collisions`` :354-397, ``construct_grasp_set`` :132-183,
``construct_halfspace_grasp_set`` :72-130) and ``rays_within_cone`` (:17-48).  All arithmetic runs in the
CUDA kernels of ``csrc/`` through the C ABI (``include/burg_b200.h``); there is no CPU fallback.

Reference quirks that are kept on purpose (SURVEY.md TL;DR 3):
  * the width filter is an OR and therefore a no-op for sane parameters (sampling.py:249-251);
  * ``sample(n)`` returns >= n grasps: whole reference points are consumed (sampling.py:220);
  * every surface crossing of a ray is a candidate; the origin triangle is removed by
    ``np.isclose(loc, p_r, atol=1e-11)`` whose rtol (1e-5) is relative to p_r (sampling.py:239);
  * with ``max_targets_per_ref_point > 1`` poses are orientation-major while widths / contacts are
    target-major (sampling.py:177 vs :181, :344).

Two execution modes:
  * batch (default): reference points are processed thousands at a time; random numbers come from the
    device (Philox) and are keyed by a seed drawn from ``np.random`` (so ``np.random.seed`` makes runs
    reproducible), results are cut at the first reference point where >= n grasps exist -- the same
    stopping rule as the reference's loop.
  * ``reference_rng = True``: one reference point per launch, consuming the legacy global numpy RNG in
    exactly the reference's order (ray angles, target shuffle, frame vector).  With identical surface
    samples (``surface_samples``) this reproduces the reference's output; used by the parity tests.
"""
import ctypes as C

import numpy as np

from . import _native as nat
from . import device as dev
from . import util
from .grasp import GraspSet
from .mesh import as_mesh


def rays_within_cone(axis, angle, n=10, uniform_on_plane=False):
    """``n`` unit ray directions inside the cone around ``axis`` with half-angle ``angle`` [rad]
    (reference sampling.py:17-48).  Host version: consumes ``np.random`` like the reference (azimuth
    draws first, then inclinations); the batch path uses the device kernel ``bt_cone_rays`` instead."""
    azimuth = np.random.uniform(0, 2 * np.pi, n)
    if uniform_on_plane:
        inclination = np.arctan(np.random.uniform(0, np.tan(angle), n))
    else:
        inclination = np.random.uniform(0, angle, n)
    cartesian = np.empty((n, 3, 1))
    cartesian[:, 0, 0] = np.sin(inclination) * np.cos(azimuth)
    cartesian[:, 1, 0] = np.sin(inclination) * np.sin(azimuth)
    cartesian[:, 2, 0] = np.cos(inclination)
    rot_mat = util.rotation_to_align_vectors([0, 0, 1], axis)
    return np.matmul(rot_mat, cartesian)[:, :, 0]


def _orientation_tables(n_orientations, halfspace):
    """cos / sin of the orientation angles, evaluated with numpy exactly as the reference does
    (sampling.py:169-174: arange(0, 2pi, 2pi/n); :106: linspace(0, pi, n))."""
    if halfspace:
        theta = np.linspace(0, np.pi, num=n_orientations)
    else:
        theta = np.arange(0, 2 * np.pi, 2 * np.pi / n_orientations)[:n_orientations]
    return np.ascontiguousarray(np.cos(theta)), np.ascontiguousarray(np.sin(theta))


class CastResult:
    """Per-ray hit lists of one ``bt_ags_cast`` launch, resident on the device."""

    def __init__(self, hit_count, hits, overflow, n_ref, n_rays, cap, valid_mask=None):
        self.hit_count, self.hits, self.overflow, self.valid_mask = hit_count, hits, overflow, valid_mask
        self.n_ref, self.n_rays, self.cap = n_ref, n_rays, cap

    def to_host(self):
        """-> (hit_count int32 [n_ref, n_rays], hits structured [n_ref, n_rays, cap] (nat.HIT_DTYPE))"""
        cnt = self.hit_count.cpu().numpy().reshape(self.n_ref, self.n_rays)
        raw = self.hits.cpu().numpy().view(nat.HIT_DTYPE).reshape(self.n_ref, self.n_rays, self.cap)
        return cnt, raw

    def flat_hits(self):
        """canonical flat list: (ref, ray, hit records) ordered by (ref, ray, t)"""
        cnt, raw = self.to_host()
        sel = np.arange(self.cap)[None, None, :] < cnt[:, :, None]
        ref_idx, ray_idx, _ = np.nonzero(sel)
        return ref_idx, ray_idx, raw[sel]


class PendingHostBatch:
    """A batch in flight (``AntipodalGraspSampler.submit_host``); ``result()`` waits for it and returns host views."""

    def __init__(self, event, pinned, n_ref, n_candidates, retry=None, padded_rows=None):
        self.event, self.pinned, self.n_ref, self.n_candidates, self.retry = event, pinned, n_ref, n_candidates, retry
        self.padded_rows = padded_rows                 # 'dma' transport: rows the copy engine moved (the capacity), else None

    def done(self):
        return self.event.query()

    def result(self):
        self.event.synchronize()
        p = self.pinned
        n_over = int(p['out_n'][3])
        if n_over and self.retry is not None:
            again = self.retry()                       # a longer hit list (up to 32 crossings per ray)
            if again is not None:
                return again
            import warnings
            warnings.warn(f'{n_over} rays cross more than 32 surfaces of the mesh: their farthest crossings were not considered')
        k = int(p['out_n'][0])
        return dict(gs=GraspSet.from_array(p['out_gs'][:k].numpy()), contacts=p['out_c'][:k].numpy(), n_grasps=int(p['out_n'][2]),
                    n_free=k, h2d_bytes=2 * self.n_ref * 24, d2h_bytes=(k if self.padded_rows is None else self.padded_rows) * (56 + 48) + 16,
                    n_candidates=self.n_candidates, chunks=1,
                    hit_list_overflows=n_over)


class AntipodalGraspSampler:
    """A sampler for antipodal grasps: looks for two contact points that satisfy the antipodal
    constraint for a friction coefficient ``mu`` (reference sampling.py:51-397)."""

    def __init__(self):
        # -- the reference's configuration surface (sampling.py:57-70) --
        self.gripper = None
        self.mu = 0.25
        self.n_orientations = 12
        self.n_rays = 100
        self.min_grasp_width = 0.002
        self.width_tolerance = 0.005
        self.max_targets_per_ref_point = 1
        self.only_grasp_from_above = False
        self.no_contact_below_z = None
        self.verbose = True
        self.verbose_debug = False
        self.mesh = None
        # -- B200 extensions --
        self.device = None                  # CUDA device index (default: current)
        self.reference_rng = False          # True: sequential mode that consumes np.random like the reference
        self.host_transport = 'dma'         # submit_host / evaluate_host results: 'dma' (local compaction, copy engine moves the padded buffers) or
                                            # 'store' (the compaction kernel stores exactly the kept rows into pinned host memory)
        self.surface_samples = None         # optional (n_sample, 6) [point | normal] override of stage S1
        self.max_hits_per_ray = 8           # per-ray hit list capacity (doubled automatically on overflow)
        self.max_passes = 10                # cap on wrap-arounds of the reference-point list (the reference spins forever)
        self.poisson_disk = False           # first-contact candidates: weighted sample elimination after the uniform samples (open3d's
                                            # Poisson-disk stage, mesh_processing.py:77-81); off: the uniform stage only
        self.cone_culling = True            # lean pipeline: skip triangles / LBVH subtrees whose vertex normals rule out a valid hit
        self.last_stats = {}

    # -----------------------------------------------------------------------------------------------
    # device plumbing
    # -----------------------------------------------------------------------------------------------
    def _device(self):
        return dev.device_index(self.device)

    def _device_mesh(self):
        if self.mesh is None:
            raise ValueError('AntipodalGraspSampler.mesh is not set')
        return dev.cached_device_mesh(self.mesh, self._device())

    def _params(self, cap=None):
        p = nat.AgsParams()
        p.angle_max = float(np.arctan(self.mu))
        p.opening_width = float(self.gripper.opening_width) if self.gripper is not None else 0.08
        p.width_tolerance = float(self.width_tolerance)
        p.min_grasp_width = float(self.min_grasp_width)
        p.use_below_z = 0 if self.no_contact_below_z is None else 1
        p.no_contact_below_z = 0.0 if self.no_contact_below_z is None else float(self.no_contact_below_z)
        p.max_hits_per_ray = int(cap or self.max_hits_per_ray)
        return p

    def _tensor(self, array, dtype):
        torch = nat.require_cuda()
        if isinstance(array, torch.Tensor):
            return array.to(device=f'cuda:{self._device()}', dtype=dtype).contiguous()
        return torch.from_numpy(np.ascontiguousarray(array)).to(device=f'cuda:{self._device()}', dtype=dtype)

    # -----------------------------------------------------------------------------------------------
    # stages (each is one C-ABI call; inputs may be numpy arrays or CUDA tensors)
    # -----------------------------------------------------------------------------------------------
    def sample_surface(self, n, seed):
        """S1: ``n`` area-weighted surface points + triangle normals -> (pts, nrm, tri) CUDA tensors."""
        torch = nat.require_cuda()
        dm = self._device_mesh()
        with torch.cuda.device(dm.device):
            pts = torch.empty((n, 3), dtype=torch.float64, device='cuda')
            nrm = torch.empty((n, 3), dtype=torch.float64, device='cuda')
            tri = torch.empty((n,), dtype=torch.int32, device='cuda')
            nat.call('bt_sample_surface', dm.handle, n, int(seed), nat.ptr(pts), nat.ptr(nrm), nat.ptr(tri),
                     nat.current_stream())
        return pts, nrm, tri

    def surface_points(self, n, seed, poisson_disk=False, init_factor=5):
        """S1 as the reference calls it (mesh_processing.poisson_disk_sampling): ``n`` surface points + triangle normals.
        ``poisson_disk``: draw ``init_factor * n`` uniform samples and keep the ``n`` that weighted sample elimination
        leaves (open3d's alpha = 8, beta = 0.5, gamma = 1.5); else the uniform stage only.  -> (pts, nrm) CUDA tensors."""
        torch = nat.require_cuda()
        if not poisson_disk or init_factor <= 1:
            pts, nrm, _ = self.sample_surface(n, seed)
            return pts, nrm
        dm = self._device_mesh()
        n0 = int(init_factor) * int(n)
        pts, nrm, _ = self.sample_surface(n0, seed)
        with torch.cuda.device(dm.device):
            keep = torch.empty((n0,), dtype=torch.uint8, device='cuda')
            rounds = C.c_int32()
            nat.call('bt_poisson_disk_eliminate', nat.ptr(pts), n0, int(n), float(dm.info.surface_area), 8.0, 0.5, 1.5,
                     nat.ptr(keep), C.byref(rounds), nat.current_stream())
            self.last_stats = dict(self.last_stats, poisson_disk_rounds=int(rounds.value), poisson_disk_initial=n0)
            sel = keep.bool()
            return pts[sel].contiguous(), nrm[sel].contiguous()

    def cone_rays(self, ref_normals, seed, uniform_on_plane=False):
        """S2: [n_ref, n_rays, 3] directions in the friction cone about ``-normal`` (device RNG).  ``uniform_on_plane``:
        the reference's option of ``rays_within_cone`` (sampling.py:33-34) -- rays uniform on the cone's ground plane
        instead of uniform in the angle (the sampler itself uses the default, like the reference's sample())."""
        torch = nat.require_cuda()
        nrm = self._tensor(ref_normals, torch.float64)
        with torch.cuda.device(nrm.device):
            dirs = torch.empty((nrm.shape[0], self.n_rays, 3), dtype=torch.float64, device=nrm.device)
            nat.call('bt_cone_rays_ex', nat.ptr(nrm), nrm.shape[0], int(self.n_rays), float(np.arctan(self.mu)),
                     int(bool(uniform_on_plane)), int(seed), nat.ptr(dirs), nat.current_stream())
        return dirs

    def cast(self, ref_points, directions, cap=None, retry_on_overflow=True):
        """R1-R6: all crossings of every ray with ``self.mesh`` + the per-hit antipodal filters.
        ``retry_on_overflow`` reads the overflow counter back (one host sync) and re-runs with a larger hit
        list if some ray crossed more surfaces than fit; the asynchronous pipeline passes False and reports
        the counter instead."""
        torch = nat.require_cuda()
        dm = self._device_mesh()
        pts = self._tensor(ref_points, torch.float64).reshape(-1, 3)
        dirs = self._tensor(directions, torch.float64)
        n_ref = pts.shape[0]
        n_rays = dirs.numel() // (3 * n_ref) if n_ref else self.n_rays
        cap = int(cap or self.max_hits_per_ray)
        with torch.cuda.device(dm.device):
            while True:
                params = self._params(cap)
                hit_count = torch.empty((n_ref * n_rays,), dtype=torch.int32, device='cuda')
                hits = torch.empty((n_ref * n_rays * cap, 72), dtype=torch.uint8, device='cuda')
                overflow = torch.zeros((1,), dtype=torch.int32, device='cuda')
                valid_mask = torch.empty((n_ref * n_rays,), dtype=torch.int32, device='cuda')
                nat.call('bt_ags_cast', dm.handle, nat.ptr(pts), nat.ptr(dirs), n_ref, n_rays, C.byref(params),
                         nat.ptr(hit_count), nat.ptr(hits), nat.ptr(valid_mask), nat.ptr(overflow), nat.current_stream())
                if not retry_on_overflow:
                    break
                n_over = int(overflow.item())
                if n_over == 0:
                    break
                if cap >= 32:
                    raise RuntimeError(f'{n_over} rays cross more than 32 surfaces of the mesh; the hit list cannot '
                                       f'hold them (bt_ags_cast max_hits_per_ray <= 32)')
                cap = min(32, cap * 2)                   # a ray crossed more surfaces than the list holds
        return CastResult(hit_count, hits, overflow, n_ref, n_rays, cap, valid_mask)

    def select(self, cast, seed):
        """R7: per reference point up to ``max_targets_per_ref_point`` valid hits -> contact pairs."""
        torch = nat.require_cuda()
        k = int(self.max_targets_per_ref_point)
        cap_pairs = max(1, cast.n_ref * k)
        with torch.cuda.device(cast.hits.device):
            pair_count = torch.empty((max(cast.n_ref, 1),), dtype=torch.int32, device='cuda')
            pair_offset = torch.zeros((cast.n_ref + 1,), dtype=torch.int64, device='cuda')
            pair_ref = torch.empty((cap_pairs,), dtype=torch.int32, device='cuda')
            pair_q = torch.empty((cap_pairs, 3), dtype=torch.float64, device='cuda')
            nat.call('bt_ags_select', nat.ptr(cast.hit_count), nat.ptr(cast.hits), nat.ptr(cast.valid_mask), cast.n_ref, cast.n_rays, cast.cap,
                     k, int(seed), nat.ptr(pair_count), nat.ptr(pair_offset), nat.ptr(pair_ref), nat.ptr(pair_q),
                     cap_pairs, nat.current_stream())
        return pair_count, pair_offset, pair_ref, pair_q, cap_pairs

    def expand(self, ref_points, pair_offset, pair_ref, pair_q, frame_vecs, cap_pairs):
        """O1/O2: contact pairs -> n_orientations grasps each (float32 [*,14] rows + float64 contacts)."""
        torch = nat.require_cuda()
        pts = self._tensor(ref_points, torch.float64).reshape(-1, 3)
        n_o = int(self.n_orientations)
        half = 1 if self.only_grasp_from_above else 0
        cos_t, sin_t = _orientation_tables(n_o, half)
        fv = None if frame_vecs is None else self._tensor(frame_vecs, torch.float64)
        with torch.cuda.device(pts.device):
            gs = torch.zeros((cap_pairs * n_o, 14), dtype=torch.float32, device='cuda')
            contacts = torch.zeros((cap_pairs * n_o, 2, 3), dtype=torch.float64, device='cuda')
            nat.call('bt_expand_orientations', nat.ptr(pts), nat.ptr(pair_offset), nat.ptr(pair_ref), nat.ptr(pair_q),
                     nat.ptr(fv), pts.shape[0], cap_pairs, n_o, half, cos_t.ctypes.data, sin_t.ctypes.data,
                     nat.ptr(gs), nat.ptr(contacts), nat.current_stream())
        return gs, contacts

    def frame_vectors(self, n, seed):
        torch = nat.require_cuda()
        with torch.cuda.device(self._device()):
            out = torch.empty((n, 3), dtype=torch.float64, device='cuda')
            nat.call('bt_random_unit_vectors', n, int(seed), nat.ptr(out), nat.current_stream())
        return out

    def evaluate_candidates(self, ref_points, ref_normals=None, directions=None, seed=1, frame_vecs=None,
                            choice_seed=0):
        """Whole device pipeline for a batch of reference points: rays -> cast + filters -> target choice ->
        orientation expansion.  ``directions`` / ``frame_vecs`` may be host-seeded (parity runs) or None
        (generated on the device).  -> dict of CUDA tensors; nothing is copied to the host."""
        torch = nat.require_cuda()
        pts = self._tensor(ref_points, torch.float64).reshape(-1, 3)
        n_ref = pts.shape[0]
        if directions is None:
            if ref_normals is None:
                raise ValueError('either directions or ref_normals is required')
            directions = self.cone_rays(ref_normals, seed)
        cast = self.cast(pts, directions)
        pair_count, pair_offset, pair_ref, pair_q, cap_pairs = self.select(cast, choice_seed)
        if frame_vecs is None and not self.only_grasp_from_above:
            frame_vecs = self.frame_vectors(n_ref, seed + 0x9E3779B9)
        gs, contacts = self.expand(pts, pair_offset, pair_ref, pair_q, frame_vecs, cap_pairs)
        return dict(cast=cast, pair_count=pair_count, pair_offset=pair_offset, pair_ref=pair_ref, pair_q=pair_q,
                    gs=gs, contacts=contacts, n_ref=n_ref, cap_pairs=cap_pairs, directions=directions)

    def run_pipeline(self, ref_points=None, ref_normals=None, n_ref=None, seed=1, directions=None, frame_vecs=None,
                     check_collisions=True, collision_width_tolerance=0.01, use_width=True, additional_objects=None,
                     exclude_shape=False, compact=True, timer=None, destinations=None, slot=0, tail_splits=1, keep_hits=False,
                     graph=False, surface=None):
        """The whole device pipeline for one batch of candidates in ONE C-ABI call (``bt_ags_pipeline``); nothing
        leaves the GPU and the host never waits:

            [S1 surface samples] -> [S2 cone rays] -> R1-R6 cast + filters -> R7 target choice -> O1/O2 orientation
            expansion -> [K1 gripper-vs-scene collision -> ballot/scan compaction of the collision-free grasps]

        ``ref_points`` / ``ref_normals``: [n_ref, 3] numpy arrays or CUDA tensors; if None, ``n_ref`` points are
        drawn on the device (S1).  ``directions`` / ``frame_vecs``: host-seeded inputs for parity runs, else device
        RNG keyed by ``seed``.  Candidates evaluated = n_ref * n_rays.  ``timer``: optional ``device.StageTimer``.
        ``destinations``: optional raw addresses for the compaction outputs (``{'free_gs': .., 'free_contacts': ..,
        'n_free': ..}``) -- a peer GPU's slot (``distributed.PeerGather``) or mapped pinned host memory; the scatter
        kernel stores there directly and the workspace tensors of those names are then NOT written.  Two more keys chain
        batches on the device: ``'free_base'`` (address of an int64 = rows already at the destination; this batch's rows are
        appended behind them) and ``'compact_after'`` (a CUDA event handle the compaction stage waits for).
        ``slot``: which cached workspace to use (batches in flight on different streams need different ones).
        ``keep_hits``: materialise the per-ray hit records (``out['cast']``, as ``cast()`` returns them).  They are an
        intermediate of the path -- the reference's ``locations / index_ray / index_tri`` arrays -- and cost 144 MB of stores
        per million rays, so by default the caster keeps only the valid-hit masks and the ray parameters, from which the
        target choice rebuilds each contact with the caster's own expression (same bits).  In this lean mode, with
        ``self.cone_culling`` (default), the caster also skips every triangle and LBVH subtree whose vertex normals rule out a
        VALID hit for the ray (the interpolated normal has to lie within atan(mu) of the ray, sampling.py:289-308): the valid
        masks, contacts and grasps are the same bits, ``hit_count`` then counts only the crossings that were looked at.
        ``tail_splits`` > 1: collision check + compaction run over that many row ranges on two streams, so a link-bound
        compaction (host / peer destination) overlaps the next range's collision walk; the output is identical.
        ``graph``: replay the call as a captured CUDA graph (``bt_graph_*``): the first call with a given shape, mesh and
        configuration captures it, later ones cost one graph launch instead of ~25 driver calls (the seed reaches the
        kernels through device memory); same results, bit for bit.  Ignored with a ``timer`` or a ``compact_after`` event.
        ``surface`` = (total, offset, key), with ``ref_points=None``: the batch's reference points are entries ``offset ..
        offset + n_ref - 1`` (mod total) of a keyed pseudo-random ordering of ONE set of ``total`` area-weighted surface
        samples -- stage S1 plus the shuffle of a whole ``sample()`` call (sampling.py:199-201), generated on the fly.
        -> dict of CUDA tensors (views into the cached workspace: valid until the next call)."""
        torch = nat.require_cuda()
        device = self._device()
        dm = self._device_mesh()
        if ref_points is not None:
            n_ref = len(ref_points)
        n_ref = int(n_ref)
        n_o, k, cap = int(self.n_orientations), int(self.max_targets_per_ref_point), int(self.max_hits_per_ray)
        if not 1 <= n_o <= 64:
            raise ValueError('n_orientations must be in [1, 64]')
        key = (n_ref, int(self.n_rays), cap, n_o, k, device, bool(check_collisions), bool(check_collisions and compact), bool(keep_hits))
        spaces = self.__dict__.setdefault('_workspaces', {})
        ws = spaces.get(slot)
        if ws is None or ws.key != key:
            ws = spaces[slot] = dev.PipelineWorkspace(*key)
        with torch.cuda.device(device):
            # inputs go straight into the workspace (one copy each; a pinned host tensor becomes one asynchronous H2D)
            for buf, src, per_ref in ((ws.ref_pts, ref_points, 3), (ws.ref_nrm, ref_normals if ref_points is not None else None, 3),
                                      (ws.dirs, directions, 3 * int(self.n_rays)), (ws.frame_vecs, frame_vecs, 3)):
                if src is not None:
                    if not isinstance(src, torch.Tensor):
                        src = torch.from_numpy(np.ascontiguousarray(src, dtype=np.float64))
                    if src.numel() != n_ref * per_ref:
                        raise ValueError(f'expected {n_ref} x {per_ref} values, got {src.numel()}')
                    if n_ref:                          # (the workspace buffers hold at least one row)
                        buf.reshape(-1)[:n_ref * per_ref].copy_(src.reshape(-1), non_blocking=True)
            if surface is not None and ref_points is not None:
                raise ValueError('surface=(total, offset, key) generates the reference points: ref_points must be None')
            surface_total = 0 if surface is None else int(surface[0])
            sig = (surface_total, float(self.mu), float(self.gripper.opening_width) if self.gripper is not None else None, float(self.width_tolerance),
                   float(self.min_grasp_width), self.no_contact_below_z, cap, int(self.n_rays), n_o, k, bool(self.only_grasp_from_above),
                   bool(check_collisions), bool(use_width), bool(compact), float(collision_width_tolerance), int(tail_splits), bool(self.cone_culling))
            cached = getattr(ws, 'config_cache', None)
            if cached is not None and cached[0] == sig:
                cfg = cached[1]
            else:
                cfg = nat.PipelineConfig()
                cfg.ags = self._params(cap)
                cfg.n_rays, cfg.n_orientations, cfg.max_targets = int(self.n_rays), n_o, k
                cfg.halfspace = 1 if self.only_grasp_from_above else 0
                cfg.check_collisions, cfg.use_width = int(bool(check_collisions)), int(bool(use_width))
                cfg.compact = int(bool(check_collisions and compact))
                cfg.collision_width_tolerance = float(collision_width_tolerance)
                cfg.tail_splits = int(tail_splits)
                cfg.cone_culling = int(bool(self.cone_culling))
                cfg.surface_total = surface_total
                cos_t, sin_t = _orientation_tables(n_o, cfg.halfspace)
                cfg.orient_cos[:n_o] = cos_t.tolist()
                cfg.orient_sin[:n_o] = sin_t.tolist()
                ws.config_cache = (sig, cfg)
            scene = gripper = None
            if check_collisions:
                meshes = ([] if exclude_shape else [self.mesh]) + list(additional_objects or [])
                scene = self._scene(meshes, device)
                gripper = self._device_gripper(device)
            ws.redirect = dict(destinations or {})
            gen = (ref_points is None, directions is None, frame_vecs is None)
            stream = nat.current_stream()
            seed = int(seed) & 0xFFFFFFFFFFFFFFFF
            if graph and timer is None and tail_splits <= 1 and 'compact_after' not in ws.redirect:
                # the captured call carries the seed's "zero or not" (canonical vs random target choice); the rest of the
                # seed reaches the kernels through ws.seed_add, so one graph serves every seed
                base = 0 if seed == 0 else 1
                gkey = (sig, dm.handle.value, None if scene is None else scene.handle.value,
                        None if gripper is None else gripper.handle.value, gen, tuple(sorted(ws.redirect.items())), base, n_ref)
                g = ws.graphs.get(gkey)
                if g is None:
                    if len(ws.graphs) >= 64:
                        ws.graphs.pop(next(iter(ws.graphs)))
                    cap = self.__dict__.setdefault('_capture_streams', {}).setdefault(device, torch.cuda.Stream(device))
                    bufs = ws.buffers(*gen)
                    nat.call('bt_graph_begin_capture', cap.cuda_stream)
                    try:
                        nat.call('bt_ags_pipeline', dm.handle, None if scene is None else scene.handle,
                                 None if gripper is None else gripper.handle, C.byref(cfg), C.byref(bufs), n_ref, base, None, cap.cuda_stream)
                    finally:
                        handle = C.c_void_p()
                        nat.call('bt_graph_end_capture', cap.cuda_stream, C.byref(handle))
                    g = ws.graphs[gkey] = dev.DeviceGraph(handle, keep=(dm, scene, gripper, ws))
                if surface is None:
                    nat.call('bt_set_u64', ws.seed_add.data_ptr(), (seed - base) & 0xFFFFFFFFFFFFFFFF, stream)
                else:
                    nat.call('bt_set_u64x3', ws.seed_add.data_ptr(), (seed - base) & 0xFFFFFFFFFFFFFFFF, int(surface[1]) & 0xFFFFFFFFFFFFFFFF,
                             int(surface[2]) & 0xFFFFFFFFFFFFFFFF, stream)
                g.launch(stream)
                ws.last_graph, ws.seed_add_dirty = g, True
            else:
                if surface is not None:
                    nat.call('bt_set_u64x3', ws.seed_add.data_ptr(), 0, int(surface[1]) & 0xFFFFFFFFFFFFFFFF, int(surface[2]) & 0xFFFFFFFFFFFFFFFF, stream)
                    ws.seed_add_dirty = False
                elif getattr(ws, 'seed_add_dirty', False):             # a graph replay left its offset there
                    ws.seed_add.zero_()
                    ws.seed_add_dirty = False
                bufs = ws.buffers(*gen)
                nat.call('bt_ags_pipeline', dm.handle, None if scene is None else scene.handle,
                         None if gripper is None else gripper.handle, C.byref(cfg), C.byref(bufs), n_ref, seed,
                         None if timer is None else timer.handle, stream)
        out = dict(n_ref=n_ref, n_candidates=n_ref * int(self.n_rays), pair_count=ws.pair_count, pair_offset=ws.pair_offset,
                   pair_ref=ws.pair_ref, pair_q=ws.pair_q, gs=ws.gs, contacts=ws.contacts, n_grasps=ws.n_grasps,
                   overflow=ws.overflow, ref_pts=ws.ref_pts, ref_nrm=ws.ref_nrm, directions=ws.dirs, frame_vecs=ws.frame_vecs,
                   cast=CastResult(ws.hit_count, ws.hits, ws.overflow, n_ref, int(self.n_rays), cap, ws.valid_mask) if keep_hits else None,
                   hit_count=ws.hit_count, valid_mask=ws.valid_mask, hit_t=ws.hit_t, cap_pairs=ws.cap_pairs)
        if check_collisions:
            out['colliding'] = ws.colliding[:ws.rows]
            if compact:
                out.update(free_gs=ws.free_gs, free_contacts=ws.free_contacts, n_free=ws.n_free)
                for name in ('free_gs', 'free_contacts', 'n_free'):       # stored elsewhere: the workspace copy is stale
                    if name in ws.redirect:
                        out[name] = None
        return out

    def evaluate_host(self, ref_points, ref_normals, seed=1, pinned=None, chunks=None, **kwargs):
        """End-to-end call with HOST buffers: uploads the reference points / normals, runs ``run_pipeline`` and
        returns host results (GraspSet of collision-free grasps, contacts, counts).  ``pinned``: optional dict of
        reusable pinned staging tensors (created on first use).

        With the collision stage on, there is no separate device-to-host copy: the pinned result buffers are mapped
        into the device's address space (every ``cudaHostAlloc`` block is, under unified addressing) and the compaction
        kernel stores the kept rows, their contacts and the count straight into them over PCIe -- exactly the rows that
        survive, no count read-back before sizing a copy, one host wait per call.

        Two ways of overlapping those PCIe stores (0.25 ms for 11.4 MB at config 3) with compute exist and are tested for
        identical output, but both measured SLOWER on B200 and are therefore off by default: ``tail_splits`` > 1 (collision
        check + compaction over row ranges on two streams inside ``bt_ags_pipeline``: 1.40 / 1.46 / 1.66 ms at 1 / 2 / 4
        ranges -- the fixed cost of a collision launch exceeds what the overlap returns) and

        ``chunks`` > 1, which splits the REFERENCE POINTS into batches on two alternating streams (whole pipelines
        overlap; each compaction appends behind the previous batch's rows through a device-side running count): 1.41 /
        1.47 / 1.75 ms at 1 / 2 / 4 chunks (the ~20 launches per batch cost more host time than the overlap wins).  Both
        are kept for callers with far larger batches.  The device RNG streams are keyed per batch, so for ``seed != 0`` the
        random draws (not their distribution) depend on ``chunks``."""
        torch = nat.require_cuda()
        device = self._device()
        caller_pinned = pinned
        pinned = {} if pinned is None else pinned
        n = len(ref_points)
        cap = max(1, n * int(self.max_targets_per_ref_point)) * int(self.n_orientations)
        push = bool(kwargs.get('check_collisions', True) and kwargs.get('compact', True))
        chunks = int(chunks or 1)
        kwargs.setdefault('tail_splits', int(getattr(self, 'host_tail_splits', 1)))
        chunks = max(1, min(chunks, n // 256, 64)) if push else 1
        while n % chunks:                                   # equal batches only (one workspace shape per stream)
            chunks -= 1
        if push and chunks == 1 and kwargs['tail_splits'] <= 1 and n > 0:
            return self.submit_host(ref_points, ref_normals, seed=seed, slot=0, pinned=caller_pinned, **kwargs).result()      # (staging buffers: the caller's, else the slot's own)
        kwargs.pop('host_transport', None)                  # (the split / chunked forms below always store from the kernel)
        with torch.cuda.device(device):
            for key, arr in (('pts', ref_points), ('nrm', ref_normals)):
                if key not in pinned or pinned[key].shape[0] != n:
                    pinned[key] = torch.empty((n, 3), dtype=torch.float64).pin_memory()
                pinned[key].numpy()[...] = arr
            if 'out_gs' not in pinned or pinned['out_gs'].shape[0] != cap:
                pinned['out_gs'] = torch.empty((cap, 14), dtype=torch.float32).pin_memory()
                pinned['out_c'] = torch.empty((cap, 2, 3), dtype=torch.float64).pin_memory()
                pinned['out_n'] = torch.zeros((66,), dtype=torch.int64).pin_memory()      # [0] kept rows, [1] spare, [2 + c] posed grasps of batch c
            pts, nrm = pinned['pts'], pinned['nrm']               # run_pipeline uploads them into its workspace (async H2D)
            out_n = pinned['out_n']
            dest = {'free_gs': pinned['out_gs'].data_ptr(), 'free_contacts': pinned['out_c'].data_ptr(),
                    'n_free': out_n.data_ptr()} if push else None
            main = torch.cuda.current_stream()
            if chunks == 1:
                out = self.run_pipeline(pts, nrm, seed=seed, destinations=dest, **kwargs)
                out_n[2:3].copy_(out['n_grasps'], non_blocking=True)
                n_candidates = out['n_candidates']
            else:
                side = self.__dict__.setdefault('_host_streams', {}).setdefault(device, [torch.cuda.Stream(device), torch.cuda.Stream(device)])
                events = self.__dict__.setdefault('_host_events', [torch.cuda.Event() for _ in range(2)])
                counts = self.__dict__.setdefault('_host_counts', {}).setdefault(device, torch.zeros((2,), dtype=torch.int64, device=f'cuda:{device}'))
                sliced = {k: kwargs.pop(k) for k in ('directions', 'frame_vecs') if k in kwargs}
                n_candidates = 0
                for c in range(chunks):
                    b, e = (n * c) // chunks, (n * (c + 1)) // chunks
                    st = side[c & 1]
                    if c < 2:
                        st.wait_stream(main)
                    last = c == chunks - 1
                    d = dict(dest)
                    d['n_free'] = out_n.data_ptr() if last else counts[c & 1:].data_ptr()      # the running count stays on the device until the last batch
                    if c > 0:
                        d['free_base'] = counts[(c - 1) & 1:].data_ptr()
                        d['compact_after'] = events[(c - 1) & 1].cuda_event
                    with torch.cuda.stream(st):
                        out = self.run_pipeline(pts[b:e], nrm[b:e], seed=seed if seed == 0 else seed + 0x10000 * c, destinations=d, slot=1 + (c & 1),
                                                **{k: (None if v is None else v[b:e]) for k, v in sliced.items()}, **kwargs)
                        out_n[2 + c:3 + c].copy_(out['n_grasps'], non_blocking=True)
                        events[c & 1].record(st)
                    n_candidates += out['n_candidates']
                main.wait_stream(side[0])
                main.wait_stream(side[1])
            if push:
                main.synchronize()
                k = int(out_n[0])
            else:
                out_n[0:1].copy_(out['n_grasps'], non_blocking=True)
                main.synchronize()
                k = int(out_n[0])                          # read the count, then copy exactly that many rows
                pinned['out_gs'][:k].copy_(out['gs'][:k], non_blocking=True)
                pinned['out_c'][:k].copy_(out['contacts'][:k], non_blocking=True)
                main.synchronize()
        h2d = 2 * n * 24
        d2h = k * (56 + 48) + 8 * (1 + chunks)
        return dict(gs=GraspSet.from_array(pinned['out_gs'][:k].numpy()), contacts=pinned['out_c'][:k].numpy(),
                    n_grasps=int(out_n[2:2 + chunks].sum()), n_free=k, h2d_bytes=h2d, d2h_bytes=d2h, n_candidates=n_candidates,
                    chunks=chunks)

    def submit_host(self, ref_points, ref_normals, seed=1, slot=0, pinned=None, **kwargs):
        """Asynchronous form of ``evaluate_host`` (collision check + compaction on): copies the reference points into
        pinned staging memory, enqueues upload + pipeline (replayed as a CUDA graph) + the transfer of the results on the
        private stream of ``slot`` and returns at once; ``.result()`` of the returned handle waits for that batch only.
        ``host_transport`` (keyword or ``self.host_transport``): 'dma' (default) -- the compaction writes device memory and a
        copy engine moves the padded result buffers into the pinned arrays, so no SM is held while the link works; 'store' --
        the compaction kernel stores exactly the kept rows into the (device-addressable) pinned arrays itself: fewer bytes, no
        second pass, but a thin kernel that lives 0.25 ms per 11 MB.  Measured at config 3 with eight batches in flight:
        0.50 ms per batch with 'dma' (= the device-resident rate), 0.60 with 'store'.  With several slots used in turn, a
        batch's PCIe transfer overlaps the next batches' ray casting:

            pending = ags.submit_host(pts0, nrm0, seed=1, slot=0)
            for i in range(1, n_batches):
                nxt = ags.submit_host(pts[i], nrm[i], seed=i + 1, slot=i & 1)
                use(pending.result())            # views into slot (i - 1) & 1's pinned arrays: valid until that slot is resubmitted
                pending = nxt

        Results are the bits ``evaluate_host`` returns for the same arguments."""
        torch = nat.require_cuda()
        device = self._device()
        n = len(ref_points)
        cap = max(1, n * int(self.max_targets_per_ref_point)) * int(self.n_orientations)
        if not (kwargs.get('check_collisions', True) and kwargs.get('compact', True)):
            raise ValueError('submit_host needs the collision check and the compaction (use evaluate_host otherwise)')
        slots = self.__dict__.setdefault('_host_slots', {})
        hs = slots.get((device, slot))
        if hs is None:
            hs = slots[(device, slot)] = dict(stream=torch.cuda.Stream(device), event=torch.cuda.Event(), pinned={})
        pinned = hs['pinned'] if pinned is None else pinned
        with torch.cuda.device(device):
            for key, arr in (('pts', ref_points), ('nrm', ref_normals)):
                if key not in pinned or pinned[key].shape[0] != n:
                    pinned[key] = torch.empty((n, 3), dtype=torch.float64).pin_memory()
                pinned[key].numpy()[...] = arr
            if 'out_gs' not in pinned or pinned['out_gs'].shape[0] != cap:
                pinned['out_gs'] = torch.empty((cap, 14), dtype=torch.float32).pin_memory()
                pinned['out_c'] = torch.empty((cap, 2, 3), dtype=torch.float64).pin_memory()
                pinned['out_n'] = torch.zeros((66,), dtype=torch.int64).pin_memory()      # [0] kept rows, [2] posed grasps
            out_n = pinned['out_n']
            transport = kwargs.pop('host_transport', None) or self.host_transport
            if transport not in ('store', 'dma'):
                raise ValueError("host_transport must be 'store' or 'dma'")
            dest = {'free_gs': pinned['out_gs'].data_ptr(), 'free_contacts': pinned['out_c'].data_ptr(), 'n_free': out_n.data_ptr()}
            st = hs['stream']
            st.wait_stream(torch.cuda.current_stream())
            kwargs.setdefault('graph', True)
            with torch.cuda.stream(st):
                if transport == 'store':
                    # the compaction kernel's own stores cross PCIe (exactly the kept rows; one thin kernel per batch)
                    out = self.run_pipeline(pinned['pts'], pinned['nrm'], seed=seed, destinations=dest, slot=('host', slot), **kwargs)
                else:
                    # local compaction, then the copy engine moves the (padded) result buffers: no SM is held while the
                    # link works, which is worth more than the ~10 % of padding once several batches are in flight
                    out = self.run_pipeline(pinned['pts'], pinned['nrm'], seed=seed, slot=('host', slot), **kwargs)
                    pinned['out_gs'].copy_(out['free_gs'][:cap], non_blocking=True)
                    pinned['out_c'].copy_(out['free_contacts'][:cap], non_blocking=True)
                    out_n[0:1].copy_(out['n_free'].reshape(1), non_blocking=True)
                out_n[2:3].copy_(out['n_grasps'], non_blocking=True)
                out_n[3:4].copy_(out['overflow'], non_blocking=True)
                hs['event'].record(st)

        def retry():
            # some ray crossed more surfaces than the hit list holds (the reference has no cap): again with a longer list
            cap_hits = int(self.max_hits_per_ray)
            if cap_hits >= 32:
                return None
            self.max_hits_per_ray = min(32, 2 * cap_hits)
            try:
                return self.submit_host(ref_points, ref_normals, seed=seed, slot=slot, pinned=pinned, host_transport=transport, **kwargs).result()
            finally:
                self.max_hits_per_ray = cap_hits
        return PendingHostBatch(hs['event'], pinned, n, out['n_candidates'], retry, padded_rows=cap if transport == 'dma' else None)

    # -----------------------------------------------------------------------------------------------
    # reference API
    # -----------------------------------------------------------------------------------------------
    @staticmethod
    def construct_grasp_set(reference_point, target_points, n_orientations, frame_vec=None, device=None):
        """Grasps at the centre between ``reference_point`` and each target, x-axis towards the target,
        ``n_orientations`` rotations about it (reference sampling.py:132-183).  The random unit vector of
        sampling.py:157 is drawn from ``np.random`` unless ``frame_vec`` is given."""
        return _construct(reference_point, target_points, n_orientations, False, frame_vec, device)

    @staticmethod
    def construct_halfspace_grasp_set(reference_point, target_points, n_orientations, device=None):
        """Same, approaching only from the upper half space (reference sampling.py:72-130)."""
        return _construct(reference_point, target_points, n_orientations, True, None, device)

    def _ref_points(self, n):
        """stage S1 + shuffle + z filter (sampling.py:199-205) -> (n_sample, 6) float64 CUDA tensor"""
        torch = nat.require_cuda()
        m = as_mesh(self.mesh)
        n_sample = int(np.max([n, 1000, len(m.triangles)]))
        if self.surface_samples is not None:
            ref = np.array(self.surface_samples, dtype=np.float64)
            np.random.shuffle(ref)                                  # sampling.py:201 (legacy global RNG)
            ref = torch.from_numpy(ref).to(f'cuda:{self._device()}')
        else:
            seed = int(np.random.randint(1, 2 ** 62))
            pts, nrm = self.surface_points(n_sample, seed, poisson_disk=bool(self.poisson_disk))
            ref = torch.cat([pts, nrm], dim=1)
            gen = torch.Generator(device=ref.device)
            gen.manual_seed(seed)
            ref = ref[torch.randperm(n_sample, device=ref.device, generator=gen)]
        if self.no_contact_below_z is not None:
            ref = ref[ref[:, 2] > self.no_contact_below_z]
        return ref.contiguous()

    def sample(self, n=10):
        """-> (GraspSet with >= n grasps, contacts float64 (N, 2, 3))  (reference sampling.py:185-352)."""
        if self.verbose:
            print('preparing to sample grasps...')
        torch = nat.require_cuda()
        if not self.reference_rng and self.surface_samples is None and self.no_contact_below_z is None and not self.poisson_disk:
            return self._sample_on_device(n)
        ref = self._ref_points(n)
        if self.verbose:
            print(f'sampled {len(ref)} first contact point candidates, beginning to find grasps.')
        if len(ref) == 0:
            raise ValueError('no reference points left after the z filter')
        if self.reference_rng:
            return self._sample_sequential(n, ref.cpu().numpy())
        n_o, k = int(self.n_orientations), int(self.max_targets_per_ref_point)
        seed = int(np.random.randint(1, 2 ** 62))
        rows, contacts, total = [], [], 0
        cursor, consumed, n_refs_used, overflowed = 0, 0, 0, 0
        batch = max(64, int(np.ceil(n / (n_o * 0.8))) + 16)
        while total < n and consumed < self.max_passes * len(ref):
            idx = (cursor + torch.arange(batch, device=ref.device)) % len(ref)
            pts, nrm = ref[idx, 0:3].contiguous(), ref[idx, 3:6].contiguous()
            out = self.run_pipeline(pts, nrm, seed=seed + consumed, check_collisions=False)
            n_over = int(out['overflow'].item())
            while n_over and self.max_hits_per_ray < 32:                # the reference has no cap on crossings per ray:
                self.max_hits_per_ray = min(32, 2 * int(self.max_hits_per_ray))        # redo the batch with a longer hit list
                out = self.run_pipeline(pts, nrm, seed=seed + consumed, check_collisions=False)
                n_over = int(out['overflow'].item())
            overflowed += n_over
            per_ref = out['pair_count'][:batch].to(torch.int64) * n_o
            cum = torch.cumsum(per_ref, 0)
            need = n - total
            reached = torch.nonzero(cum >= need)
            take_refs = int(reached[0].item()) + 1 if len(reached) else batch
            take_rows = int(cum[take_refs - 1].item())
            rows.append(out['gs'][:take_rows].cpu().numpy())
            contacts.append(out['contacts'][:take_rows].cpu().numpy())
            total += take_rows
            n_refs_used += take_refs
            cursor = (cursor + take_refs) % len(ref)
            consumed += take_refs
            batch = min(1 << 16, max(64, int((n - total) / max(total / max(n_refs_used, 1), 1e-3)) + 16))
        self.last_stats = dict(ref_points_used=n_refs_used, candidates=n_refs_used * self.n_rays, grasps=total,
                               exhausted=total < n, hit_list_overflows=overflowed)
        if overflowed:
            import warnings
            warnings.warn(f'{overflowed} rays cross more than 32 surfaces of the mesh: their farthest crossings were not considered')
        if total < n and self.verbose:
            print(f'warning: only {total} grasps found after {self.max_passes} passes over the reference points')
        gs = GraspSet.from_array(np.concatenate(rows) if rows else np.zeros((0, 14), dtype=np.float32))
        return gs, (np.concatenate(contacts) if contacts else np.empty((0, 2, 3)))

    def _sample_on_device(self, n):
        """``sample(n)`` without materialising the first-contact candidates: batch b of reference points is a slice of a
        keyed pseudo-random ordering of the ``max(n, 1000, #triangles)`` surface samples (S1 + shuffle, sampling.py:199-201),
        generated inside the graph-replayed pipeline; per batch ONE wait for the device -- pair counts, rows and contacts come
        back in three asynchronous copies into pinned memory and are cut on the host at the first reference point where
        >= n grasps exist (the reference's stopping rule)."""
        torch = nat.require_cuda()
        device = self._device()
        m = as_mesh(self.mesh)
        n_sample = int(np.max([n, 1000, len(m.triangles)]))
        n_o, k = int(self.n_orientations), int(self.max_targets_per_ref_point)
        key = int(np.random.randint(1, 2 ** 62))
        seed = int(np.random.randint(1, 2 ** 62))
        if self.verbose:
            print(f'sampled {n_sample} first contact point candidates, beginning to find grasps.')
        rows, contacts, total = [], [], 0
        consumed, n_refs_used, overflowed = 0, 0, 0
        batch = -(-max(64, int(np.ceil(n / (n_o * 0.8))) + 16) // 64) * 64            # (multiples of 64: few distinct workspace shapes)
        stage = self.__dict__.setdefault('_sample_pinned', {})
        with torch.cuda.device(device):
            while total < n and consumed < self.max_passes * n_sample:
                cap_rows = batch * k * n_o
                if stage.get('cap', 0) < cap_rows or stage.get('batch', 0) < batch:
                    stage.update(cap=cap_rows, batch=batch, gs=torch.empty((cap_rows, 14), dtype=torch.float32).pin_memory(),
                                 contacts=torch.empty((cap_rows, 2, 3), dtype=torch.float64).pin_memory(),
                                 counts=torch.empty((batch + 1,), dtype=torch.int32).pin_memory())
                while True:
                    out = self.run_pipeline(n_ref=batch, seed=seed + consumed, check_collisions=False, graph=True,
                                            surface=(n_sample, consumed % n_sample, key), slot=('sample', batch))
                    stage['counts'][:batch].copy_(out['pair_count'][:batch], non_blocking=True)
                    stage['counts'][batch:batch + 1].copy_(out['overflow'], non_blocking=True)
                    stage['gs'][:cap_rows].copy_(out['gs'][:cap_rows], non_blocking=True)
                    stage['contacts'][:cap_rows].copy_(out['contacts'][:cap_rows], non_blocking=True)
                    torch.cuda.current_stream().synchronize()               # the one wait of this batch
                    n_over = int(stage['counts'][batch])
                    if not n_over or self.max_hits_per_ray >= 32:
                        break
                    self.max_hits_per_ray = min(32, 2 * int(self.max_hits_per_ray))      # the reference has no cap on crossings per ray
                overflowed += n_over
                cum = np.cumsum(stage['counts'][:batch].numpy().astype(np.int64) * n_o)
                reached = np.nonzero(cum >= n - total)[0]
                take_refs = int(reached[0]) + 1 if len(reached) else batch
                take_rows = int(cum[take_refs - 1])
                rows.append(stage['gs'][:take_rows].numpy().copy())
                contacts.append(stage['contacts'][:take_rows].numpy().copy())
                total += take_rows
                n_refs_used += take_refs
                consumed += take_refs
                batch = -(-min(1 << 16, max(64, int((n - total) / max(total / max(n_refs_used, 1), 1e-3)) + 16)) // 64) * 64
        self.last_stats = dict(ref_points_used=n_refs_used, candidates=n_refs_used * self.n_rays, grasps=total,
                               exhausted=total < n, hit_list_overflows=overflowed)
        if overflowed:
            import warnings
            warnings.warn(f'{overflowed} rays cross more than 32 surfaces of the mesh: their farthest crossings were not considered')
        if total < n and self.verbose:
            print(f'warning: only {total} grasps found after {self.max_passes} passes over the reference points')
        gs = GraspSet.from_array(np.concatenate(rows) if len(rows) != 1 else rows[0]) if rows else GraspSet.from_array(np.zeros((0, 14), dtype=np.float32))
        return gs, (np.concatenate(contacts) if len(contacts) != 1 else contacts[0]) if contacts else np.empty((0, 2, 3))

    def _sample_sequential(self, n, ref_points):
        """the reference's loop verbatim in structure (sampling.py:215-352), one launch per reference point,
        with the host consuming np.random in the reference's order."""
        angle = np.arctan(self.mu)
        gs = GraspSet()
        gs_contacts = np.empty((0, 2, 3))
        i_ref, iterations = 0, 0
        while len(gs) < n and iterations < self.max_passes * len(ref_points):
            iterations += 1
            p_r, n_r = ref_points[i_ref, 0:3], ref_points[i_ref, 3:6]
            i_ref = (i_ref + 1) % len(ref_points)
            ray_directions = rays_within_cone(-n_r, angle, self.n_rays)
            _, _, hits = self.cast(p_r.reshape(1, 3), ray_directions.reshape(1, -1, 3)).flat_hits()
            locations = hits['loc'][hits['flags'] == 0]
            if len(locations) == 0:
                continue
            if len(locations) > self.max_targets_per_ref_point:
                indices = np.arange(len(locations))
                np.random.shuffle(indices)
                locations = locations[indices[:self.max_targets_per_ref_point]]
            if self.only_grasp_from_above:
                grasps = self.construct_halfspace_grasp_set(p_r, locations, self.n_orientations, device=self._device())
            else:
                grasps = self.construct_grasp_set(p_r, locations, self.n_orientations, device=self._device())
            contacts = np.empty((len(locations), 2, 3))
            contacts[:, 0] = p_r
            contacts[:, 1] = locations
            gs_contacts = np.concatenate([gs_contacts, np.repeat(contacts, self.n_orientations, axis=0)], axis=0)
            gs.add(grasps)
        self.last_stats = dict(ref_points_used=iterations, candidates=iterations * self.n_rays, grasps=len(gs),
                               exhausted=len(gs) < n)
        return gs, gs_contacts

    def check_collisions(self, graspset, use_width=True, width_tolerance=0.01, additional_objects=None,
                         exclude_shape=False):
        """bool[N]: does the gripper mesh, posed at each grasp (and squeezed in x to the grasp width plus
        ``width_tolerance`` if ``use_width``), intersect the object and/or ``additional_objects``
        (reference sampling.py:354-397)."""
        if not additional_objects and exclude_shape:
            raise ValueError('no collision objects specified.')
        torch = nat.require_cuda()
        device = self._device()
        meshes = ([] if exclude_shape else [self.mesh]) + list(additional_objects or [])
        scene = self._scene(meshes, device)
        gripper = self._device_gripper(device)
        rows = graspset if isinstance(graspset, torch.Tensor) else \
            torch.from_numpy(np.ascontiguousarray(graspset._gs_array, dtype=np.float32)).to(f'cuda:{device}')
        if self.verbose:
            print('checking collisions...')
        colliding = collide_rows(scene, gripper, rows, float(self.gripper.opening_width), float(width_tolerance),
                                 bool(use_width))
        return colliding.cpu().numpy().astype(bool)

    def _scene(self, meshes, device):
        """the collision scene of ``meshes`` on ``device``: a single mesh shares the ray caster's cached DeviceMesh (one
        LBVH, never two views of one object); several are merged, keyed by every member's identity AND geometry tag"""
        if len(meshes) == 1:
            return dev.cached_device_mesh(meshes[0], device)
        key = tuple((id(m), dev.mesh_tag(m)) for m in meshes)
        cache = getattr(self, '_scene_cache', None)      # (the cache entry keeps the mesh objects alive, so their ids cannot be reused)
        if cache is not None and cache[0] == key and cache[1].device == device and all(a is b for a, b in zip(cache[2], meshes)):
            return cache[1]
        scene = dev.DeviceMesh.from_meshes(meshes, device)
        for m in meshes:
            if hasattr(m, 'freeze'):
                m.freeze()
        self._scene_cache = (key, scene, meshes)
        return scene

    def _device_gripper(self, device):
        cache = getattr(self, '_gripper_cache', None)
        if cache is not None and cache[0] is self.gripper and cache[1].device == device:
            return cache[1]
        g = dev.DeviceGripper.from_gripper(self.gripper, device)
        self._gripper_cache = (self.gripper, g)
        return g


def collide_rows(scene, gripper, rows, opening_width, width_tolerance, use_width, n_valid=None):
    """K1 on device-resident float32 [N,14] rows -> uint8 CUDA tensor [N] (1 = colliding)."""
    torch = nat.require_cuda()
    n = rows.shape[0]
    with torch.cuda.device(rows.device):
        out = torch.empty((((n + 3) // 4) * 4,), dtype=torch.uint8, device=rows.device)
        nat.call('bt_collide', scene.handle, gripper.handle, nat.ptr(rows), n, nat.ptr(n_valid), opening_width,
                 width_tolerance, 1 if use_width else 0, nat.ptr(out), nat.current_stream())
    return out[:n]


def compact_rows(mask, keep_value, rows, contacts=None, n_valid=None):
    """ordered stream compaction of grasp rows (and contacts) -> (rows, contacts, count tensor)"""
    torch = nat.require_cuda()
    n = rows.shape[0]
    with torch.cuda.device(rows.device):
        out_rows = torch.empty_like(rows)
        out_contacts = None if contacts is None else torch.empty_like(contacts)
        count = torch.zeros((1,), dtype=torch.int64, device=rows.device)
        nat.call('bt_compact', nat.ptr(mask), int(keep_value), nat.ptr(rows), nat.ptr(contacts), n, nat.ptr(n_valid),
                 nat.ptr(out_rows), nat.ptr(out_contacts), nat.ptr(count), None, nat.current_stream())
    return out_rows, out_contacts, count


def _construct(reference_point, target_points, n_orientations, halfspace, frame_vec, device):
    torch = nat.require_cuda()
    d = dev.device_index(device)
    p_r = np.reshape(np.asarray(reference_point, dtype=np.float64), (1, 3))
    targets = np.reshape(np.asarray(target_points, dtype=np.float64), (-1, 3))
    k = len(targets)
    if not halfspace and frame_vec is None:
        x_axis = (targets - p_r) / np.linalg.norm(targets - p_r, axis=-1)[:, np.newaxis]
        y_axis = np.zeros(x_axis.shape)
        while (np.linalg.norm(y_axis, axis=-1) == 0).any():          # sampling.py:156-158
            frame_vec = util.generate_random_unit_vector()
            y_axis = np.cross(x_axis, frame_vec)
    cos_t, sin_t = _orientation_tables(n_orientations, halfspace)
    with torch.cuda.device(d):
        pts = torch.from_numpy(p_r).cuda()
        q = torch.from_numpy(np.ascontiguousarray(targets)).cuda()
        off = torch.tensor([0, k], dtype=torch.int64, device='cuda')
        ref = torch.zeros((max(k, 1),), dtype=torch.int32, device='cuda')
        fv = None if halfspace else torch.from_numpy(np.asarray(frame_vec, dtype=np.float64).reshape(1, 3)).cuda()
        gs = torch.zeros((k * n_orientations, 14), dtype=torch.float32, device='cuda')
        if k:
            nat.call('bt_expand_orientations', nat.ptr(pts), nat.ptr(off), nat.ptr(ref), nat.ptr(q), nat.ptr(fv), 1, k,
                     int(n_orientations), 1 if halfspace else 0, cos_t.ctypes.data, sin_t.ctypes.data, nat.ptr(gs),
                     None, nat.current_stream())
        return GraspSet.from_array(gs.cpu().numpy())


# ---------------------------------------------------------------------------------------------------------------------
# generators around the sampler (reference sampling.py:400-495), device-side
# ---------------------------------------------------------------------------------------------------------------------
def grasp_perturbations(grasps, radii=None, include_original_grasp=True, device=None):
    """Perturbed grasp poses on 6-d 'spheres' (1 mm translation = 1 deg rotation): per radius and grasp 6 translations
    (+-radius mm along each pose axis) and 6 rotations (+-radius deg about each pose axis), optionally preceded by the
    original grasp; all perturbations of the first grasp come first, and so on (reference sampling.py:400-454)."""
    from .grasp import Grasp
    torch = nat.require_cuda()
    if radii is None:
        radii = [5, 10, 15]
    elif not isinstance(radii, list):
        raise ValueError('radii must be a list (or None)')
    if not isinstance(grasps, GraspSet):
        if isinstance(grasps, Grasp):
            grasps = grasps.as_grasp_set()
        else:
            raise ValueError('g must be a grasp.Grasp or a grasp.GraspSet')
    if len(radii) > 16:
        raise ValueError('at most 16 radii are supported')
    per = len(radii) * 12 + int(include_original_grasp)
    d = dev.device_index(device)
    r = np.ascontiguousarray(radii, dtype=np.float64)
    with torch.cuda.device(d):
        rows = torch.from_numpy(grasps.as_contiguous()).cuda()
        out = torch.zeros((len(grasps) * per, 14), dtype=torch.float32, device='cuda')
        nat.call('bt_grasp_perturbations', nat.ptr(rows), len(grasps), r.ctypes.data, len(radii),
                 int(include_original_grasp), nat.ptr(out), nat.current_stream())
        return GraspSet.from_array(out.cpu().numpy())


def random_poses(n, seed=None, device=None):
    """(n, 4, 4) float64 random poses: uniform random orientations, positions in [0, 1) (reference sampling.py:457-469).
    The random numbers come from the device generator keyed by ``seed`` (default: drawn from ``np.random``)."""
    torch = nat.require_cuda()
    if seed is None:
        seed = int(np.random.randint(1, 2 ** 62))
    with torch.cuda.device(dev.device_index(device)):
        out = torch.empty((int(n), 4, 4), dtype=torch.float64, device='cuda')
        nat.call('bt_random_poses', int(n), int(seed), nat.ptr(out), nat.current_stream())
        return out.cpu().numpy()


def farthest_point_sampling(point_cloud, k, device=None):
    """(k,) indices of an approximate farthest point sampling of ``point_cloud`` ((n, c >= 3), xyz first): index 0, then
    repeatedly the point with the largest distance to all points chosen so far (reference sampling.py:472-495)."""
    torch = nat.require_cuda()
    if len(point_cloud) < k:
        raise ValueError(f'given point cloud has only {len(point_cloud)} elements, cannot sample {k} points')
    with torch.cuda.device(dev.device_index(device)):
        if isinstance(point_cloud, torch.Tensor):
            pts = point_cloud.cuda()
        else:
            arr = np.asarray(point_cloud)
            if arr.dtype not in (np.float32, np.float64):
                arr = arr.astype(np.float64)
            pts = torch.from_numpy(np.ascontiguousarray(arr)).cuda()
        if pts.dtype not in (torch.float32, torch.float64):
            pts = pts.to(torch.float64)
        pts = pts.contiguous()
        out = torch.zeros((int(k),), dtype=torch.int64, device='cuda')
        if k > 0:
            nat.call('bt_farthest_point_sampling', nat.ptr(pts), 1 if pts.dtype == torch.float64 else 0, pts.shape[1],
                     pts.shape[0], int(k), nat.ptr(out), nat.current_stream())
        return out.cpu().numpy()


def sample_scene(object_library, ground_area, instances_per_scene, instances_per_object=1, max_tries=20, rng=None,
                 device=None):
    """Samples a physically plausible scene from the objects of ``object_library`` (reference sampling.py:498-569):
    every object gets a random rotation about z, one of its stable poses and a random xy offset that keeps its bounding
    box on the ground area; a placement is kept if the new object's mesh does not collide with the meshes already
    placed (``bt_mesh_pair_collide`` instead of a trimesh CollisionManager), otherwise it is retried up to ``max_tries``
    times.  ``rng``: optional ``np.random.Generator`` for reproducible scenes (the reference is unseeded)."""
    import logging
    from scipy.spatial.transform import Rotation as R
    from . import core, mesh_processing
    scene = core.Scene(ground_area=ground_area)
    rng = np.random.default_rng() if rng is None else rng
    population = [obj_type for obj_type in object_library.values() for _ in range(instances_per_object)]
    obj_types = rng.choice(np.array(population, dtype=object), instances_per_scene, replace=False)
    placed = []                                     # posed meshes already in the scene (mutually collision-free)
    for object_type in obj_types:
        success = False
        for _ in range(max_tries):
            angle = rng.random() * np.pi * 2
            tf_rot = np.eye(4)
            tf_rot[:3, :3] = R.from_rotvec(angle * np.array([0, 0, 1])).as_matrix()
            pose = tf_rot @ object_type.stable_poses.sample_pose(uniformly=True, rng=rng)
            instance = core.ObjectInstance(object_type, pose)
            m = instance.get_mesh()
            min_x, min_y, _ = m.get_min_bound()
            max_x, max_y, _ = m.get_max_bound()
            range_x, range_y = ground_area[0] - (max_x - min_x), ground_area[1] - (max_y - min_y)
            if range_x < 0 or range_y < 0:
                continue                            # the ground area is too small for this object
            x, y = rng.random() * range_x - min_x, rng.random() * range_y - min_y
            instance.pose[0, 3] = x + pose[0, 3]
            instance.pose[1, 3] = y + pose[1, 3]
            candidate = instance.get_mesh()
            # the meshes already placed do not collide with each other, so only pairs with the new mesh can
            if not any(len(placed) in pair for pair in mesh_processing.collisions(placed + [candidate], device=device)):
                scene.objects.append(instance)
                placed.append(candidate)
                success = True
                break
        if not success:
            logging.warning(f'Could not add object to scene, exceeded number of max_tries ({max_tries}). Returning '
                            f'scene with fewer object instances than requested.')
    return scene

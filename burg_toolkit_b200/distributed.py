"""Multi-GPU sharding of the hot path: one process per GPU, ``torch.distributed`` for the plumbing.

The path shards naturally: reference points (hence candidates) and the rows of the reference
grasp set are independent units.  Every rank builds the mesh + LBVH locally (a deterministic build, cheaper than a
broadcast) and holds the full query set; there is NO collective inside the computation.  The only exchanges are

  * AGS:      gather of each rank's compacted grasp rows / contacts (variable length) -- rank order = candidate order,
              so the gathered result equals the single-GPU result on the same candidates;
  * coverage: all-reduce(SUM) of one int64 covered-count per rank (+ optional gather of the mask).

All helpers work with any backend (NCCL over NVLink on the B200 box, gloo on CPU for the host-logic tests).

``PeerGather`` is the B200 form of the AGS exchange: the destination rank owns one buffer with a slot per rank, the
others map it through CUDA IPC, and every rank's compaction kernel stores its kept rows straight into its slot over
NVLink while it runs (``bt_compact`` / ``bt_ags_pipeline`` write to whatever address they are given).  What is left for
NCCL is one 8-byte-per-rank ``all_gather`` of the counts, which is also the completion barrier.
"""
import ctypes as C

import numpy as np


def shard_range(n, rank, world):
    """contiguous [begin, end) of ``n`` units owned by ``rank`` (sizes differ by at most one)."""
    base, rem = divmod(int(n), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def _dist():
    import torch.distributed as dist
    return dist


def gather_rows(local_rows, dst=0, group=None):
    """Variable-length gather of per-rank [n_r, ...] tensors to ``dst`` in rank order (-> concatenated tensor on
    ``dst``, None elsewhere).  Counts travel first (all_gather of one int64), then padded payloads."""
    import torch
    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n_local = torch.tensor([local_rows.shape[0]], dtype=torch.int64, device=local_rows.device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)
    padded = torch.zeros((cap,) + tuple(local_rows.shape[1:]), dtype=local_rows.dtype, device=local_rows.device)
    padded[:local_rows.shape[0]] = local_rows
    bufs = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)


def sharded_coverage(reference_rows, query_rows, epsilon=15.0, kd_semantics=True, covered_fn=None, group=None,
                     device=None):
    """coverage of ``reference_rows`` by ``query_rows`` with the reference rows split over the ranks.
    ``covered_fn(ref_shard, qry) -> bool[n_shard]`` defaults to the CUDA kernel (``metrics.covered_mask``); the CPU
    tests inject the oracle here.  Every rank returns the same float."""
    import torch
    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = len(reference_rows)
    b, e = shard_range(n, rank, world)
    if covered_fn is None:
        from . import metrics
        covered_fn = lambda r, q: metrics.covered_mask(r, q, epsilon, kd_semantics)   # noqa: E731
    local = covered_fn(reference_rows[b:e], query_rows) if e > b else np.zeros(0, dtype=bool)
    count = torch.tensor([int(np.count_nonzero(local))], dtype=torch.int64, device=device or 'cpu')
    dist.all_reduce(count, op=dist.ReduceOp.SUM, group=group)
    return int(count.item()) / n


def sharded_candidates(ref_points, ref_normals, evaluate_fn, dst=0, group=None, device=None):
    """Evaluate antipodal candidates with the reference points split over the ranks.
    ``evaluate_fn(points, normals, first_index) -> (rows float32 [m,14], contacts float64 [m,2,3])`` for the shard
    (CUDA pipeline by default caller; oracle in the CPU tests).  -> (rows, contacts) gathered on ``dst`` in
    candidate order, (None, None) elsewhere."""
    import torch
    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    b, e = shard_range(len(ref_points), rank, world)
    rows, contacts = evaluate_fn(ref_points[b:e], ref_normals[b:e], b)
    dev = device or 'cpu'
    rows_t = torch.as_tensor(np.ascontiguousarray(rows, dtype=np.float32)).to(dev).reshape(-1, 14)
    cont_t = torch.as_tensor(np.ascontiguousarray(contacts, dtype=np.float64)).to(dev).reshape(-1, 2, 3)
    all_rows = gather_rows(rows_t, dst, group)
    all_contacts = gather_rows(cont_t, dst, group)
    if rank != dst:
        return None, None
    return all_rows.cpu().numpy(), all_contacts.cpu().numpy()


class _DeviceArray:
    """``__cuda_array_interface__`` wrapper so torch can view memory that libburgb200 allocated (bt_peer_alloc)."""

    def __init__(self, address, shape, typestr):
        self.__cuda_array_interface__ = {'shape': tuple(shape), 'typestr': typestr, 'data': (int(address), False),
                                         'version': 2, 'strides': None}


class PeerGather:
    """Gather of the compacted grasp rows (and contacts) to ``dst`` WITHOUT a payload collective.

    ``dst`` allocates ``[2, world, cap_rows, 14] float32`` (+ ``[2, world, cap_rows, 2, 3] float64``) with ``bt_peer_alloc``;
    the IPC handle is broadcast once; every rank passes ``destinations()`` to ``AntipodalGraspSampler.run_pipeline`` so
    that its ``compact_scatter_kernel`` writes slot ``rank`` directly (peer stores over NVLink / NVSwitch; local stores on
    ``dst`` itself).  ``complete(n_free)`` all-gathers the per-rank counts: each rank's contribution is enqueued behind
    its scatter kernel, so when the collective has finished on ``dst`` every slot is complete.  Rank order = candidate
    order, so ``result()`` equals the single-GPU output on the same candidates.

    The slots are double-buffered by call parity: ``complete()`` flips the half that ``destinations()`` hands out, and
    ``result()`` reads the half just completed.  A peer can only reach the compaction of call s+2 (same half as s) after
    the count exchange of call s+1 has finished, and ``dst`` enqueues its contribution to that exchange behind its own
    reads of call s -- so nobody overwrites rows that ``dst`` is still reading."""

    def __init__(self, cap_rows, with_contacts=True, device=None, dst=0, group=None):
        import torch
        from . import _native as nat
        dist = _dist()
        self.group, self.dst = group, dst
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.cap_rows, self.with_contacts = int(cap_rows), bool(with_contacts)
        self.device = torch.cuda.current_device() if device is None else int(device)
        self._nat = nat
        gs_bytes = 2 * self.world * self.cap_rows * 56
        total = gs_bytes + (2 * self.world * self.cap_rows * 48 if with_contacts else 0)
        base, handle = C.c_void_p(), (C.c_ubyte * 64)()
        self.owner = self.rank == dst
        error = None
        try:
            if self.owner:
                nat.call('bt_peer_alloc', self.device, total, C.byref(base), handle)
        except nat.NativeError as ex:
            error = str(ex)
        box = [bytes(handle) if self.owner and error is None else None]
        dist.broadcast_object_list(box, src=dst, group=group)
        try:
            if not self.owner and box[0] is not None:
                handle = (C.c_ubyte * 64).from_buffer_copy(box[0])
                nat.call('bt_peer_open', self.device, handle, C.byref(base))
        except nat.NativeError as ex:
            error = str(ex)
        errors = [None] * self.world                      # every rank takes the same decision
        dist.all_gather_object(errors, error if (error or box[0] is not None) else 'owner could not allocate', group=group)
        if any(errors):
            if base.value:
                nat.call('bt_peer_free' if self.owner else 'bt_peer_close', self.device, base)
            raise nat.NativeError('PeerGather: no peer mapping (%s)' % '; '.join(f'rank {r}: {e}' for r, e in enumerate(errors) if e))
        self.base = int(base.value)
        self.gs_bytes = gs_bytes
        self.half = 0                                     # which half the next run_pipeline writes
        self.counts = torch.zeros((self.world,), dtype=torch.int64, device=f'cuda:{self.device}')
        self.rows = self.contacts = None
        if self.owner:
            dev = f'cuda:{self.device}'
            self.rows = torch.as_tensor(_DeviceArray(self.base, (2, self.world, self.cap_rows, 14), '<f4'), device=dev)
            if with_contacts:
                self.contacts = torch.as_tensor(_DeviceArray(self.base + gs_bytes, (2, self.world, self.cap_rows, 2, 3), '<f8'), device=dev)

    def destinations(self):
        """``run_pipeline(destinations=...)``: where this rank's compaction stores its output (current half, slot ``rank``)"""
        slot = self.half * self.world + self.rank
        d = {'free_gs': self.base + slot * self.cap_rows * 56}
        if self.with_contacts:
            d['free_contacts'] = self.base + self.gs_bytes + slot * self.cap_rows * 48
        return d

    def complete(self, n_free):
        """enqueue the count exchange behind the scatter kernel of the current stream (-> counts [world] on every rank)"""
        import torch
        dist = _dist()
        if dist.get_backend(self.group) == 'nccl':
            dist.all_gather_into_tensor(self.counts, n_free.reshape(1), group=self.group)
        else:                           # host-side backends (the single-GPU two-process test): order by a stream wait
            torch.cuda.current_stream().synchronize()
            parts = [torch.zeros(1, dtype=torch.int64) for _ in range(self.world)]
            dist.all_gather(parts, n_free.reshape(1).cpu(), group=self.group)
            self.counts.copy_(torch.cat(parts))
        self.done_half, self.half = self.half, self.half ^ 1
        return self.counts

    def result(self):
        """on ``dst``: (rows [sum n_r, 14], contacts [sum n_r, 2, 3] or None) in rank order; (None, None) elsewhere"""
        import torch
        if not self.owner:
            return None, None
        counts = [int(c) for c in self.counts.cpu().tolist()]
        h = self.done_half
        rows = torch.cat([self.rows[h, r, :c] for r, c in enumerate(counts)], dim=0)
        contacts = torch.cat([self.contacts[h, r, :c] for r, c in enumerate(counts)], dim=0) if self.with_contacts else None
        return rows, contacts

    def close(self):
        import torch
        if not self.base:
            return
        torch.cuda.synchronize(self.device)
        self.rows = self.contacts = None
        _dist().barrier(group=self.group)                     # nobody is still storing into the buffer
        if self.owner:
            _dist().barrier(group=self.group)                 # the peers have unmapped it
            self._nat.call('bt_peer_free', self.device, C.c_void_p(self.base))
        else:
            self._nat.call('bt_peer_close', self.device, C.c_void_p(self.base))
            _dist().barrier(group=self.group)
        self.base = 0

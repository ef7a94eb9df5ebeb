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
NVLink while it runs (``bt_compact`` / ``bt_ags_pipeline`` write to whatever address they are given).  Completion is
signalled the same way (count + epoch flag stored into the destination's control block, one polling kernel there), or --
below 4 ranks -- by one 8-byte-per-rank ``all_gather`` of the counts.
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


def spatial_order(rows, epsilon=15.0):
    """permutation that sorts grasp rows by the cell of the coverage grid (cell = 1.01 * epsilon / 1000, x-major) their
    translation falls into.  Coverage shards should be SPATIAL tiles: contiguous slices of this order give every rank a
    compact block of cells (and each rank's grid kernel the same cell-dense blocks of reference rows as a single GPU has);
    contiguous slices of an arbitrary row order spread every rank's rows thinly over the whole grid, which makes the
    per-rank work shrink much more slowly than 1 / world.  Masks and counts do not depend on the row order."""
    t = np.asarray(rows, dtype=np.float32).reshape(-1, 14)[:, 0:3].astype(np.float64)
    cell = 1.01 * float(epsilon) / 1000.0
    c = np.floor((t - t.min(axis=0)) / cell).astype(np.int64)
    dims = c.max(axis=0) + 1
    return np.argsort((c[:, 0] * dims[1] + c[:, 1]) * dims[2] + c[:, 2], kind='stable')


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

    ``dst`` allocates ``[depth, world, cap_rows, 14] float32`` (+ ``[depth, world, cap_rows, 2, 3] float64``) with ``bt_peer_alloc``;
    the IPC handle is broadcast once; every rank passes ``destinations()`` to ``AntipodalGraspSampler.run_pipeline`` so
    that its ``compact_kernel`` writes slot ``rank`` directly (peer stores over NVLink / NVSwitch; local stores on
    ``dst`` itself).  Rank order = candidate order, so ``result()`` equals the single-GPU output on the same candidates.

    Per step:  ``d = peer.destinations()``  ->  ``run_pipeline(..., destinations=d)``  ->  ``peer.complete(out['n_free'])``,
    all three on the step's stream.  Steps may be issued on DIFFERENT streams (several batches in flight per rank, so that
    a link-bound compaction overlaps the next batch's ray casting); the calls only have to be made in step order.

    Completion, ``one_sided=True`` (the default with NCCL): no collective at all.  ``complete`` enqueues
    ``bt_peer_signal`` behind the compaction -- the rank's count and then an epoch flag are stored into ``dst``'s control
    block.  ``dst`` additionally runs ``bt_peer_wait`` for the epoch on its private gather stream (behind the step's own
    compaction): one small kernel that polls the flags of all ranks, bounded by ``timeout_s`` (a lost peer sets a status
    word in pinned host memory -- ``check()`` raises -- instead of hanging the device).  The other ranks never wait for
    each other, so a step does not cost the slowest rank's time on every rank.
    Slot reuse: every rank owns a ring of ``depth`` slots (epoch e uses slot e % depth).  ``dst`` acknowledges epoch e-2 in
    the wait kernel of epoch e, i.e. ``result()`` of an epoch has to be called before the second ``complete`` after it;
    a rank's step e waits (``destinations()``: a polling kernel on a side stream + an event the step's stream waits for) until
    e-depth is acknowledged -- with the default depth of 6 that is satisfied long before, so nothing stalls.
    ``one_sided=False``: ``complete`` all-gathers the counts with the process group behind the scatter kernel (NCCL), or --
    host-side backends, the two-process single-GPU test -- after a stream synchronisation.

    ``transport`` (one-sided mode; env ``BT_PEER_TRANSPORT``):
      'direct'  the compaction kernel's own stores cross NVLink (``destinations()`` hands it the peer addresses).  Nothing is
                written twice, but the kernel then lives as long as the link needs -- with seven ranks storing into one
                destination ~5x its local 16 us -- at the end of the step's dependency chain and on the step's SMs: measured
                on 8 B200s the sending ranks ran 0.545-0.56 ms per step against rank 0's 0.514.
      'staged'  (default) every rank but ``dst`` compacts into its own workspace as on one GPU; ``complete(n_free, rows,
                contacts, slot)`` enqueues, on the rank's exchange stream behind the step, the ring-slot acknowledgement wait,
                the transfer and ``bt_peer_signal``.  The transfer is a copy engine moving the padded local buffers
                (``push_blocks=0``, default: no SM at all) or ``bt_peer_push`` (a thin grid of ``push_blocks`` blocks streaming
                exactly ``n_free`` rows).  The step's stream starts its next batch at once: 0.52-0.536 ms on the senders.
    Pass ``slot`` (the pipeline workspace the step uses) to ``destinations`` / ``complete`` in staged mode: the next step on
    that workspace waits for the transfer that still reads it."""

    ACK_LAG = 2

    def __init__(self, cap_rows, with_contacts=True, device=None, dst=0, group=None, one_sided=None, timeout_s=10.0, depth=6,
                 transport=None, push_blocks=None):
        import os
        import torch
        from . import _native as nat
        dist = _dist()
        self.group, self.dst = group, dst
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.cap_rows, self.with_contacts = (int(cap_rows) + 31) // 32 * 32, bool(with_contacts)       # slot bases stay 256-byte aligned
        transport = transport or os.environ.get('BT_PEER_TRANSPORT') or 'staged'
        if transport not in ('direct', 'staged'):
            raise ValueError("transport must be 'direct' or 'staged'")
        self.push_blocks = int(push_blocks if push_blocks is not None else os.environ.get('BT_PEER_PUSH_BLOCKS', 0))      # 0: copy engine, padded (default); > 0: bt_peer_push with that many blocks
        self.push_done = {}                               # staged: workspace slot -> event behind the last push that read it
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.one_sided = (dist.get_backend(group) == 'nccl') if one_sided is None else bool(one_sided)
        self.timeout_s = float(timeout_s)
        self.staged = transport == 'staged' and self.one_sided
        self.depth = D = max(self.ACK_LAG + 2, int(depth))   # slots per rank (ring): how many steps a rank may run ahead of dst
        self._nat = nat
        gs_bytes = D * self.world * self.cap_rows * 56
        data_bytes = gs_bytes + (D * self.world * self.cap_rows * 48 if with_contacts else 0)
        ctrl_off = (data_bytes + 255) // 256 * 256                 # control block: flags [D, world], counts [D, world], ack [1] (int64)
        total = ctrl_off + 8 * (2 * D * self.world + 1)
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
        self.flags_ptr = self.base + ctrl_off
        self.counts_ptr = self.flags_ptr + 8 * D * self.world
        self.ack_ptr = self.counts_ptr + 8 * D * self.world
        self.epoch = 0                                    # steps issued so far; the next one writes ring slot (epoch + 1) % depth
        self.completed = 0
        dev = f'cuda:{self.device}'
        self.counts = torch.zeros((self.world,), dtype=torch.int64, device=dev)
        self.status = torch.zeros((1,), dtype=torch.int32).pin_memory()       # written by a timed-out wait kernel; the host reads it without a sync
        self.side = torch.cuda.Stream(self.device)        # dst: the wait kernels, in epoch order; others: the acknowledgement polls
        self.done_events = {}                             # dst: epoch -> event behind its wait kernel
        self.rows = self.contacts = self.slot_counts = None
        if self.owner:
            self.rows = torch.as_tensor(_DeviceArray(self.base, (D, self.world, self.cap_rows, 14), '<f4'), device=dev)
            if with_contacts:
                self.contacts = torch.as_tensor(_DeviceArray(self.base + gs_bytes, (D, self.world, self.cap_rows, 2, 3), '<f8'), device=dev)
            self.slot_counts = torch.as_tensor(_DeviceArray(self.counts_ptr, (D, self.world), '<i8'), device=dev)

    def destinations(self, slot=None):
        """begin the next step on the current stream: -> ``run_pipeline(destinations=...)``, where this rank's compaction
        stores its output (ring slot of the step, slot ``rank``).  One-sided mode, from step depth + 1 on: the current stream
        first waits until ``dst`` has acknowledged the step that used this ring slot before.
        ``transport='staged'`` (ranks other than ``dst``): -> {} -- the compaction writes the pipeline workspace ``slot`` as
        on one GPU and ``complete`` pushes the rows from there; the current stream only waits for the previous push that
        read this workspace."""
        import torch
        self.epoch += 1
        e, h = self.epoch, self.epoch % self.depth
        if self.staged and not self.owner:
            done = self.push_done.get(slot)
            if done is not None:
                torch.cuda.current_stream().wait_event(done)
            if e % self.depth == 0:
                self.check()
            return {}
        slot = h * self.world + self.rank
        d = {'free_gs': self.base + slot * self.cap_rows * 56}
        if self.with_contacts:
            d['free_contacts'] = self.base + self.gs_bytes + slot * self.cap_rows * 48
        if self.one_sided and not self.owner and e > self.depth:
            ev = torch.cuda.Event()
            with torch.cuda.stream(self.side):
                self._nat.call('bt_peer_wait', self.ack_ptr, 1, e - self.depth, self.timeout_s, self.status.data_ptr(), None, 0,
                               self.side.cuda_stream)
                ev.record(self.side)
            torch.cuda.current_stream().wait_event(ev)
        if e % self.depth == 0:
            self.check()
        return d

    def complete(self, n_free, rows=None, contacts=None, slot=None):
        """enqueue the completion of the step begun by the last ``destinations()`` behind its scatter kernel (current stream).
        ``transport='staged'``: ``rows`` / ``contacts`` = the step's locally compacted output (``out['free_gs']``,
        ``out['free_contacts']`` of workspace ``slot``); on the exchange stream, behind the step: wait for the ring slot's
        acknowledgement, ``bt_peer_push`` of exactly ``n_free`` rows into ``dst``'s buffer, ``bt_peer_signal``.  The step's
        stream is free for the next batch at once; the link time is neither on its critical path nor on its SMs."""
        import torch
        dist = _dist()
        self.completed += 1
        e = self.completed
        if e > self.epoch:
            raise RuntimeError('PeerGather.complete() without a matching destinations()')
        h = e % self.depth
        if self.staged and not self.owner:
            if rows is None or (self.with_contacts and contacts is None):
                raise ValueError("transport='staged' needs the step's local rows (and contacts)")
            if rows.shape[0] > self.cap_rows:
                raise ValueError('more local rows than the gather slot holds')
            cur = torch.cuda.current_stream()
            ev = torch.cuda.Event()
            ev.record(cur)
            self.side.wait_event(ev)
            ring = h * self.world + self.rank
            done = torch.cuda.Event()
            with torch.cuda.stream(self.side):
                if e > self.depth:
                    self._nat.call('bt_peer_wait', self.ack_ptr, 1, e - self.depth, self.timeout_s, self.status.data_ptr(), None, 0,
                                   self.side.cuda_stream)
                if self.push_blocks > 0:
                    self._nat.call('bt_peer_push', n_free.data_ptr(), int(rows.shape[0]), rows.data_ptr(),
                                   contacts.data_ptr() if self.with_contacts else None, self.base + ring * self.cap_rows * 56,
                                   (self.base + self.gs_bytes + ring * self.cap_rows * 48) if self.with_contacts else None,
                                   self.push_blocks, self.side.cuda_stream)
                else:                                     # push_blocks = 0: the whole (padded) local buffers by copy engine
                    self._nat.call('bt_peer_copy', self.base + ring * self.cap_rows * 56, rows.data_ptr(), int(rows.shape[0]) * 56, self.side.cuda_stream)
                    if self.with_contacts:
                        self._nat.call('bt_peer_copy', self.base + self.gs_bytes + ring * self.cap_rows * 48, contacts.data_ptr(),
                                       int(rows.shape[0]) * 48, self.side.cuda_stream)
                self._nat.call('bt_peer_signal', n_free.data_ptr(), self.counts_ptr + 8 * ring, self.flags_ptr + 8 * ring, e,
                               self.side.cuda_stream)
                done.record(self.side)
            self.push_done[slot] = done
        elif self.one_sided:
            cur = torch.cuda.current_stream()
            self._nat.call('bt_peer_signal', n_free.data_ptr(), self.counts_ptr + 8 * (h * self.world + self.rank),
                           self.flags_ptr + 8 * (h * self.world + self.rank), e, cur.cuda_stream)
            if self.owner:
                ev = torch.cuda.Event()
                ev.record(cur)
                self.side.wait_event(ev)
                done = torch.cuda.Event()
                with torch.cuda.stream(self.side):
                    self._nat.call('bt_peer_wait', self.flags_ptr + 8 * h * self.world, self.world, e, self.timeout_s,
                                   self.status.data_ptr(), self.ack_ptr, max(0, e - self.ACK_LAG), self.side.cuda_stream)
                    done.record(self.side)
                self.done_events[e] = done
                self.done_events.pop(e - self.depth, None)
        elif dist.get_backend(self.group) == 'nccl':
            dist.all_gather_into_tensor(self.counts, n_free.reshape(1), group=self.group)
        else:                           # host-side backends (the single-GPU two-process test): order by a stream wait
            torch.cuda.current_stream().synchronize()
            parts = [torch.zeros(1, dtype=torch.int64) for _ in range(self.world)]
            dist.all_gather(parts, n_free.reshape(1).cpu(), group=self.group)
            self.counts.copy_(torch.cat(parts))
        return e

    def check(self):
        """raise if a wait kernel of this rank has timed out (no device synchronisation: the status word is host memory)"""
        if int(self.status[0]):
            raise self._nat.NativeError(f'PeerGather: rank {self.rank} waited more than {self.timeout_s} s for a peer')

    def result(self, epoch=None):
        """on ``dst``: (rows [sum n_r, 14], contacts [sum n_r, 2, 3] or None) of step ``epoch`` (default: the last completed
        one) in rank order, as new tensors; (None, None) elsewhere.  Waits for that step's completion only."""
        import torch
        if not self.owner:
            return None, None
        e = self.completed if epoch is None else int(epoch)
        h = e % self.depth
        if self.one_sided:
            self.done_events[e].synchronize()
            self.check()
            counts = [int(c) for c in self.slot_counts[h].cpu().tolist()]
        else:
            counts = [int(c) for c in self.counts.cpu().tolist()]
        rows = torch.cat([self.rows[h, r, :c] for r, c in enumerate(counts)], dim=0)
        contacts = torch.cat([self.contacts[h, r, :c] for r, c in enumerate(counts)], dim=0) if self.with_contacts else None
        return rows, contacts

    def close(self, raise_on_timeout=True):
        """release the mapping (collective: every rank calls it).  -> True if a wait kernel of this rank had timed out
        (raises instead unless ``raise_on_timeout`` is False)"""
        import torch
        if not self.base:
            return False
        torch.cuda.synchronize(self.device)
        status = int(self.status[0])
        self.rows = self.contacts = self.slot_counts = None
        _dist().barrier(group=self.group)                     # nobody is still storing into the buffer
        if self.owner:
            _dist().barrier(group=self.group)                 # the peers have unmapped it
            self._nat.call('bt_peer_free', self.device, C.c_void_p(self.base))
        else:
            self._nat.call('bt_peer_close', self.device, C.c_void_p(self.base))
            _dist().barrier(group=self.group)
        self.base = 0
        if status and raise_on_timeout:
            raise self._nat.NativeError(f'PeerGather: rank {self.rank} waited more than {self.timeout_s} s for a peer')
        return bool(status)

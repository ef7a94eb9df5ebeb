"""Multi-GPU sharding of the hot path: one process per GPU, ``torch.distributed`` for the plumbing.

The path shards naturally: reference points (hence candidates) and the rows of the reference
grasp set are independent units.  Every rank builds the mesh + LBVH locally (a deterministic build, cheaper than a
broadcast) and holds the full query set; there is NO collective inside the computation.  The only exchanges are

  * AGS:      gather of each rank's compacted grasp rows / contacts (variable length) -- rank order = candidate order,
              so the gathered result equals the single-GPU result on the same candidates;
  * coverage: all-reduce(SUM) of one int64 covered-count per rank (+ optional gather of the mask).

All helpers work with any backend (NCCL over NVLink on the B200 box, gloo on CPU for the host-logic tests).
"""
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

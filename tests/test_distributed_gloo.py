"""CPU tests of the multi-rank host logic (world_size 2 and 3, gloo): sharding, variable-length gather, count
all-reduce.  The per-shard compute is the oracle here (no GPU in this container); on the B200 box the same helpers run
over NCCL with the CUDA kernels (bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_golden
from burg_toolkit_b200 import distributed as bdist


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, fn, queue):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        queue.put((rank, fn(rank, world)))
    finally:
        dist.destroy_process_group()


def run_ranks(fn, world):
    ctx = mp.get_context('spawn')
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fn, queue)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(queue.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return results


def test_shard_range_partitions_everything():
    for n in (0, 1, 7, 100, 10_001):
        for world in (1, 2, 3, 8):
            spans = [bdist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _coverage_job(rank, world):
    from oracle import port
    g = load_golden('metrics_small')
    return bdist.sharded_coverage(g['ref'], g['qry'], 15.0, covered_fn=lambda r, q: port.covered_kd(r, q))


@pytest.mark.parametrize('world', [2, 3])
def test_sharded_coverage_equals_single_rank(world):
    g = load_golden('metrics_small')
    res = run_ranks(_coverage_job, world)
    assert all(v == float(g['coverage_kd']) for v in res.values())


def _candidates_job(rank, world):
    from oracle import port
    g = load_golden('ags_icosphere')
    mesh = port.Mesh(g['vertices'], g['triangles'], g['vertex_normals'])
    pts, dirs = g['p_r'][:9], g['dirs'][:9]

    def evaluate(p, n, first):
        rows, contacts = [], []
        for i in range(len(p)):
            rec = port.evaluate_ref_point(mesh, p[i], dirs[first + i])
            valid = rec['location'][rec['flags'] == 0]
            if len(valid):
                rows.append(port.construct_grasp_set(p[i], valid[:1], 12, np.array([0.0, 0.6, 0.8])))
                c = np.empty((12, 2, 3)); c[:, 0] = p[i]; c[:, 1] = valid[0]
                contacts.append(c)
        return (np.concatenate(rows) if rows else np.zeros((0, 14), np.float32),
                np.concatenate(contacts) if contacts else np.zeros((0, 2, 3)))

    rows, contacts = bdist.sharded_candidates(pts, pts, evaluate)
    return None if rows is None else (rows, contacts)


@pytest.mark.parametrize('world', [2, 3])
def test_sharded_candidates_gather_in_candidate_order(world):
    single = run_ranks(_candidates_job, 1)[0]
    res = run_ranks(_candidates_job, world)
    assert all(res[r] is None for r in range(1, world))
    assert np.array_equal(res[0][0], single[0]) and np.array_equal(res[0][1], single[1])
    assert len(single[0]) > 0 and len(single[0]) % 12 == 0


def _ragged_job(rank, world):
    t = torch.arange(rank * 3, dtype=torch.float32).reshape(-1, 1) + 100 * rank      # rank 0 contributes nothing
    out = bdist.gather_rows(t)
    return None if out is None else out.numpy()


def test_gather_rows_handles_empty_and_ragged_shards():
    res = run_ranks(_ragged_job, 3)
    want = np.concatenate([np.arange(r * 3, dtype=np.float32).reshape(-1, 1) + 100 * r for r in range(3)])
    assert np.array_equal(res[0], want) and res[1] is None and res[2] is None

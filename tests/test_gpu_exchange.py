"""GPU tests of the two places where the compaction kernel stores OUTSIDE local HBM: mapped pinned host memory (the
end-to-end call, ``evaluate_host``) and a peer-visible buffer mapped through CUDA IPC (``distributed.PeerGather``, the
multi-GPU gather).  Both must deliver exactly the bytes the device-resident pipeline produces."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _inputs(n_ref, seed):
    from burg_toolkit_b200 import mesh as bmesh
    from test_gpu_fullsize import host_candidates
    mesh = bmesh.create_bumpy_sphere(64, 64, radius=0.035, amplitude=0.15, seed=1)      # some grasps collide, some do not
    return (mesh,) + host_candidates(mesh, n_ref, seed)


def _sampler(mesh, device=0):
    from test_gpu_fullsize import make_sampler
    ags = make_sampler(mesh)
    ags.device = device
    return ags


def _device_result(ags, pts, nrm, dirs, fv):
    out = ags.run_pipeline(pts, nrm, directions=dirs, frame_vecs=fv, seed=0)
    k = int(out['n_free'].item())
    return out['free_gs'][:k].cpu().numpy(), out['free_contacts'][:k].cpu().numpy(), int(out['n_grasps'].item())


def test_evaluate_host_equals_device_pipeline():
    mesh, pts, nrm, dirs, fv = _inputs(1500, 3)
    ags = _sampler(mesh)
    rows, contacts, n_grasps = _device_result(ags, pts, nrm, dirs, fv)
    assert 0 < len(rows) < n_grasps                                     # some grasps collide, some do not
    pinned = {}
    for _ in range(2):                                                  # second call reuses the pinned buffers
        res = ags.evaluate_host(pts, nrm, seed=0, pinned=pinned, directions=dirs, frame_vecs=fv, host_transport='store')
        assert res['n_free'] == len(rows) and res['n_grasps'] == n_grasps
        assert np.array_equal(res['gs']._gs_array, rows)
        assert np.array_equal(res['contacts'], contacts)
        assert res['d2h_bytes'] == len(rows) * 104 + 16                # the compaction kernel stored exactly the kept rows
    # the copy-engine transport (local compaction + padded D2H copies; the default) returns the same rows
    assert ags.host_transport == 'dma'                                 # the default
    res = ags.evaluate_host(pts, nrm, seed=0, directions=dirs, frame_vecs=fv)
    assert res['n_free'] == len(rows) and res['n_grasps'] == n_grasps and res['d2h_bytes'] == len(pts) * 12 * 104 + 16
    assert np.array_equal(res['gs']._gs_array, rows) and np.array_equal(res['contacts'], contacts)
    # collide + compact over row ranges on two streams: same bytes, same order
    for splits in (1, 2, 3, 8):
        res = ags.evaluate_host(pts, nrm, seed=0, pinned=pinned, tail_splits=splits, directions=dirs, frame_vecs=fv)
        assert res['n_free'] == len(rows) and res['n_grasps'] == n_grasps
        assert np.array_equal(res['gs']._gs_array, rows) and np.array_equal(res['contacts'], contacts)
    out = ags.run_pipeline(pts, nrm, directions=dirs, frame_vecs=fv, seed=0, tail_splits=4)          # device-resident destination
    k = int(out['n_free'].item())
    assert k == len(rows) and np.array_equal(out['free_gs'][:k].cpu().numpy(), rows)
    assert np.array_equal(out['free_contacts'][:k].cpu().numpy(), contacts)
    assert np.array_equal(out['colliding'][:n_grasps].cpu().numpy().astype(bool).sum(), n_grasps - len(rows))
    # batches on two streams, each compaction appended behind the previous one's rows: same bytes, same order
    for chunks in (2, 3, 5):
        res = ags.evaluate_host(pts, nrm, seed=0, pinned=pinned, chunks=chunks, directions=dirs, frame_vecs=fv)
        assert res['chunks'] == chunks and res['n_free'] == len(rows) and res['n_grasps'] == n_grasps
        assert np.array_equal(res['gs']._gs_array, rows) and np.array_equal(res['contacts'], contacts)
    # without the collision stage the rows come back through plain copies
    res = ags.evaluate_host(pts, nrm, seed=0, directions=dirs, frame_vecs=fv, check_collisions=False)
    assert res['n_free'] == n_grasps


def test_compact_to_mapped_host_memory_ragged():
    """bt_compact with host-mapped outputs at sizes around the 256-row block: nothing kept, ragged tails, all kept."""
    import torch
    from burg_toolkit_b200 import _native as nat
    rng = np.random.default_rng(0)
    for n in (1, 255, 256, 257, 5000):
        rows = torch.from_numpy(rng.normal(size=(n, 14)).astype(np.float32)).cuda()
        contacts = torch.from_numpy(rng.normal(size=(n, 2, 3))).cuda()
        out_rows = torch.full((n, 14), -1.0, dtype=torch.float32).pin_memory()
        out_contacts = torch.full((n, 2, 3), -1.0, dtype=torch.float64).pin_memory()
        count = torch.full((1,), -1, dtype=torch.int64).pin_memory()
        for frac in (0.0, 0.5, 1.0):
            mask = torch.from_numpy((rng.random(n) < frac).astype(np.uint8)).cuda()
            nat.call('bt_compact', nat.ptr(mask), 0, nat.ptr(rows), nat.ptr(contacts), n, None, nat.ptr(out_rows),
                     nat.ptr(out_contacts), nat.ptr(count), None, nat.current_stream())
            torch.cuda.synchronize()
            keep = ~mask.cpu().numpy().astype(bool)
            k = int(count[0])
            assert k == keep.sum()
            assert np.array_equal(out_rows.numpy()[:k], rows.cpu().numpy()[keep])
            assert np.array_equal(out_contacts.numpy()[:k], contacts.cpu().numpy()[keep])


def _peer_job(rank, world, port, queue, one_sided):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from burg_toolkit_b200 import distributed as bdist
        torch.cuda.set_device(0)                                        # both ranks share the one GPU of the test box
        mesh, pts, nrm, dirs, fv = _inputs(1200, 5)
        ags = _sampler(mesh)
        b, e = bdist.shard_range(len(pts), rank, world)
        peer = bdist.PeerGather((e - b + 1) * 12, with_contacts=True, device=0, one_sided=one_sided, timeout_s=30.0)
        ok = True
        for rep in range(4):                                            # the slots are reused call after call
            out = ags.run_pipeline(pts[b:e], nrm[b:e], directions=dirs[b:e], frame_vecs=fv[b:e], seed=0,
                                   destinations=peer.destinations())
            assert out['free_gs'] is None and out['n_free'] is not None
            peer.complete(out['n_free'])
            rows, contacts = peer.result()
            if rank == 0:
                want_rows, want_contacts, _ = _device_result(ags, pts, nrm, dirs, fv)
                ok = ok and np.array_equal(rows.cpu().numpy(), want_rows) and np.array_equal(contacts.cpu().numpy(), want_contacts)
                ok = ok and len(want_rows) > 1000
            dist.barrier()
        peer.close()
        queue.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('one_sided', [False])
def test_peer_gather_two_processes_share_a_buffer_through_ipc(one_sided):
    """two ranks (processes) on one GPU: rank 1's compaction stores into rank 0's buffer through the IPC mapping; the
    gathered rows / contacts equal the single-process result on all reference points, bit for bit.  Completion here: an
    all_gather of the counts after a stream synchronisation.  (The one-sided completion -- count + epoch-flag stores and a
    polling kernel on rank 0 -- needs the ranks' GPUs to run at the same time: kernels that wait for one another must not
    share a GPU, so that mode is verified on real multi-GPU runs, tests/multi_gpu_check.py and bench.py's gather_verified.)"""
    import torch.multiprocessing as mp
    from test_distributed_gloo import _free_port
    ctx = mp.get_context('spawn')
    queue, port = ctx.Queue(), _free_port()
    procs = [ctx.Process(target=_peer_job, args=(r, 2, port, queue, one_sided)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(queue.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results == {0: True, 1: True}


def test_peer_push_copies_exactly_the_counted_rows():
    """bt_peer_push (the staged gather's transfer kernel) with a device-side count: rows and contact records up to the
    count arrive bit for bit, nothing behind them is touched; odd counts exercise the 8-byte tail of the 16-byte copy."""
    import torch
    from burg_toolkit_b200 import _native as nat
    rng = np.random.default_rng(1)
    cap = 4096
    rows = torch.from_numpy(rng.normal(size=(cap, 14)).astype(np.float32)).cuda()
    contacts = torch.from_numpy(rng.normal(size=(cap, 2, 3))).cuda()
    for n, blocks in ((0, 1), (1, 1), (255, 3), (1001, 32), (cap, 8), (cap + 77, 4)):
        dst_rows = torch.full((cap, 14), -1.0, dtype=torch.float32, device='cuda')
        dst_contacts = torch.full((cap, 2, 3), -1.0, dtype=torch.float64, device='cuda')
        count = torch.tensor([n], dtype=torch.int64, device='cuda')
        nat.call('bt_peer_push', nat.ptr(count), cap, nat.ptr(rows), nat.ptr(contacts), nat.ptr(dst_rows), nat.ptr(dst_contacts),
                 blocks, nat.current_stream())
        torch.cuda.synchronize()
        k = min(n, cap)
        assert torch.equal(dst_rows[:k], rows[:k]) and torch.equal(dst_contacts[:k], contacts[:k])
        assert bool((dst_rows[k:] == -1).all()) and bool((dst_contacts[k:] == -1).all())

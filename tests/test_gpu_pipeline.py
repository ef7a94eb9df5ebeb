"""GPU tests of the fused / graph-replayed pipeline against its stand-alone stages: the one-launch generator, target
selection and compaction, the CUDA-graph replay and the asynchronous host submission must all produce the bits of the
stage-by-stage calls (which the other test files pin against the oracle and the reference-recorded fixtures)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _setup(n_grid=48, **attrs):
    from burg_toolkit_b200 import mesh as bmesh
    from test_gpu_fullsize import make_sampler
    mesh = bmesh.create_bumpy_sphere(n_grid, n_grid, radius=0.035, amplitude=0.15, seed=3)
    return mesh, make_sampler(mesh, **attrs)


@pytest.mark.parametrize('n_ref', [1, 7, 8, 333])
def test_fused_generator_equals_the_three_kernels(n_ref):
    """generate_kernel (S1 + S2 + frame vectors in one launch) draws the same numbers as bt_sample_surface,
    bt_cone_rays and bt_random_unit_vectors with the seeds the pipeline always used"""
    _, ags = _setup()
    seed = 1234567
    out = ags.run_pipeline(n_ref=n_ref, seed=seed, check_collisions=False)
    pts, nrm, _ = ags.sample_surface(n_ref, seed)
    dirs = ags.cone_rays(nrm, seed + 0x1000)
    fv = ags.frame_vectors(n_ref, seed + 0x3000)
    assert np.array_equal(out['ref_pts'][:n_ref].cpu().numpy(), pts.cpu().numpy())
    assert np.array_equal(out['ref_nrm'][:n_ref].cpu().numpy(), nrm.cpu().numpy())
    assert np.array_equal(out['directions'][:n_ref].cpu().numpy(), dirs.cpu().numpy())
    assert np.array_equal(out['frame_vecs'][:n_ref].cpu().numpy(), fv.cpu().numpy())


@pytest.mark.parametrize('seed,k', [(0, 1), (5, 1), (5, 4), (0, 4)])
def test_fused_selection_equals_count_scan_fill(seed, k):
    """select_fused_kernel (chained scan across blocks) against bt_ags_select (count, scan, fill) on the same hit records,
    canonical and random-subset choice, more reference points than one block holds"""
    _, ags = _setup(max_targets_per_ref_point=k)
    n_ref = 1500
    full = ags.run_pipeline(n_ref=n_ref, seed=seed, check_collisions=False, keep_hits=True)
    pair_count, pair_offset, pair_ref, pair_q, _ = ags.select(full['cast'], seed + 0x2000 if seed else 0)
    n_pairs = int(pair_offset[-1].item())
    assert n_pairs > n_ref // 2
    assert np.array_equal(full['pair_count'][:n_ref].cpu().numpy(), pair_count[:n_ref].cpu().numpy())
    assert np.array_equal(full['pair_offset'].cpu().numpy(), pair_offset.cpu().numpy())
    assert np.array_equal(full['pair_ref'][:n_pairs].cpu().numpy(), pair_ref[:n_pairs].cpu().numpy())
    assert np.array_equal(full['pair_q'][:n_pairs].cpu().numpy(), pair_q[:n_pairs].cpu().numpy())
    assert int(full['n_grasps'].item()) == n_pairs * ags.n_orientations


def test_single_pass_compaction_is_ordered_and_complete():
    import torch
    from burg_toolkit_b200 import sampling
    rng = np.random.default_rng(0)
    for n in (1, 255, 256, 257, 100_003):
        rows = torch.from_numpy(rng.random((n, 14)).astype(np.float32)).cuda()
        contacts = torch.from_numpy(rng.random((n, 2, 3))).cuda()
        mask = torch.from_numpy((rng.random(n) < 0.37).astype(np.uint8)).cuda()
        out_rows, out_contacts, count = sampling.compact_rows(mask, 0, rows, contacts)
        keep = (mask == 0).cpu().numpy()
        k = int(count.item())
        assert k == keep.sum()
        assert np.array_equal(out_rows[:k].cpu().numpy(), rows.cpu().numpy()[keep])
        assert np.array_equal(out_contacts[:k].cpu().numpy(), contacts.cpu().numpy()[keep])


@pytest.mark.parametrize('seed', [0, 99])
def test_graph_replay_equals_direct_call(seed):
    """a captured bt_ags_pipeline, replayed with seeds it was not captured with, against the plain call"""
    import torch
    _, ags = _setup()
    n_ref = 600
    names = ('ref_pts', 'directions', 'pair_count', 'pair_q', 'gs', 'contacts', 'colliding', 'free_gs', 'free_contacts')
    for s in ([0, 0] if seed == 0 else [seed, seed + 1, 3]):
        direct = ags.run_pipeline(n_ref=n_ref, seed=s)
        want = {n: direct[n].cpu().numpy().copy() for n in names}
        want_n = (int(direct['n_grasps'].item()), int(direct['n_free'].item()))
        got = ags.run_pipeline(n_ref=n_ref, seed=s, graph=True)
        torch.cuda.synchronize()
        n, nf = int(got['n_grasps'].item()), int(got['n_free'].item())
        assert (n, nf) == want_n and n > 1000
        for name in names:
            lim = {'gs': n, 'contacts': n, 'colliding': n, 'free_gs': nf, 'free_contacts': nf, 'pair_q': n // 12}.get(name)
            a, b = got[name].cpu().numpy(), want[name]
            assert np.array_equal(a[:lim], b[:lim]), name
    ws = ags._workspaces[0]
    assert len(ws.graphs) == 1                               # one graph served every seed of its kind
    n_nodes, n_kernels, kernel_names = ws.last_graph.info()
    assert n_kernels <= 10 and n_nodes >= n_kernels, kernel_names
    assert any(k.startswith('traverse_kernel') for k in kernel_names) and any(k.startswith('compact_kernel') for k in kernel_names)
    # a plain call after replays sees a zero seed offset again
    again = ags.run_pipeline(n_ref=n_ref, seed=s)
    assert np.array_equal(again['gs'].cpu().numpy()[:n], want['gs'][:n])


def test_pipelined_host_submission_equals_evaluate_host():
    mesh, ags = _setup()
    from test_gpu_fullsize import host_candidates
    batches = [host_candidates(mesh, 300, 20 + i)[:2] for i in range(5)]
    want = []
    for i, (pts, nrm) in enumerate(batches):
        r = ags.evaluate_host(pts, nrm, seed=i + 1)
        want.append((r['gs']._gs_array.copy(), r['contacts'].copy(), r['n_grasps'], r['n_free']))
    assert all(w[3] > 1000 for w in want)
    got = []
    pending = ags.submit_host(*batches[0], seed=1, slot=0)
    for i in range(1, len(batches)):
        nxt = ags.submit_host(*batches[i], seed=i + 1, slot=i & 1)
        r = pending.result()
        got.append((r['gs']._gs_array.copy(), r['contacts'].copy(), r['n_grasps'], r['n_free']))
        pending = nxt
    r = pending.result()
    got.append((r['gs']._gs_array.copy(), r['contacts'].copy(), r['n_grasps'], r['n_free']))
    for w, g in zip(want, got):
        assert w[2:] == g[2:] and np.array_equal(w[0], g[0]) and np.array_equal(w[1], g[1])


def test_shuffled_surface_samples_are_a_permutation_of_stage_s1():
    """run_pipeline(surface=(total, offset, key)): reference point i of the batch is surface sample perm(offset + i) of ONE
    set of `total` samples -- all of them exactly once over a full cycle, whatever the batch boundaries, in an order that
    depends on the key; and sample() built on it keeps the reference's contract"""
    mesh, ags = _setup()
    total, key = 1000, 424242
    want = ags.sample_surface(total, key)[0].cpu().numpy()
    got = []
    for off, n in ((0, 333), (333, 600), (933, 67)):
        out = ags.run_pipeline(n_ref=n, seed=5, check_collisions=False, surface=(total, off, key), graph=(n != 600))
        got.append(out['ref_pts'][:n].cpu().numpy().copy())
    got = np.concatenate(got)
    assert not np.array_equal(got, want)                                   # shuffled ...
    assert np.array_equal(got[np.lexsort(got.T)], want[np.lexsort(want.T)])     # ... but the same 1000 points, each once
    other = ags.run_pipeline(n_ref=333, seed=5, check_collisions=False, surface=(total, 0, key + 1))['ref_pts'][:333].cpu().numpy()
    assert not np.array_equal(other, got[:333])
    wrap = ags.run_pipeline(n_ref=100, seed=5, check_collisions=False, surface=(total, 950, key))['ref_pts'][:100].cpu().numpy()
    assert np.array_equal(wrap[:50], got[950:]) and np.array_equal(wrap[50:], got[:50])
    np.random.seed(3)
    gs, contacts = ags.sample(500)
    assert len(gs) >= 500 and len(gs) % 12 == 0 and contacts.shape == (len(gs), 2, 3)
    assert ags.last_stats['grasps'] == len(gs) and ags.last_stats['candidates'] == ags.last_stats['ref_points_used'] * 100
    np.random.seed(3)
    gs2, _ = ags.sample(500)
    assert np.array_equal(gs._gs_array, gs2._gs_array)                      # np.random.seed makes runs reproducible
    # widths of the returned grasps = distance of their contact pairs (row r belongs to contacts r)
    assert np.allclose(gs._gs_array[:, 13], np.linalg.norm(contacts[:, 1] - contacts[:, 0], axis=1), rtol=1e-6)


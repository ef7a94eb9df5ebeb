"""GPU parity tests (through the C ABI) for the metric path: bit-exact against the golden fixtures that
the reference's own metrics.py produced, and against the oracle port on seeded inputs."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import port

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def metrics():
    from burg_toolkit_b200 import metrics as m
    return m


def _gs(rows):
    from burg_toolkit_b200.grasp import GraspSet
    return GraspSet.from_array(np.ascontiguousarray(rows, dtype=np.float32))


def random_rows(n, seed, cube=0.2):
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(seed)
    rows = np.zeros((n, 14), dtype=np.float32)
    rows[:, 0:3] = rng.random((n, 3)) * cube
    rows[:, 3:12] = Rotation.random(n, random_state=seed).as_matrix().reshape(n, 9)
    return rows


def perturbed_rows(rows, seed, frac=0.5, sigma_t=0.005, sigma_deg=5.0):
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(seed)
    idx = rng.choice(len(rows), int(len(rows) * frac), replace=False)
    out = rows[idx].copy()
    out[:, 0:3] += rng.normal(0, sigma_t, (len(idx), 3)).astype(np.float32)
    rv = rng.normal(0, np.deg2rad(sigma_deg), (len(idx), 3))
    out[:, 3:12] = (Rotation.from_rotvec(rv).as_matrix() @ rows[idx, 3:12].reshape(-1, 3, 3).astype(np.float64)).reshape(-1, 9)
    return out


def test_known_answer_3mm_15deg(metrics):
    k = load_golden('known_answers')
    gs = _gs(k['gs'])
    comb = metrics.combined_distances(gs[0], gs[1])
    assert comb.dtype == np.float32 and comb.shape == (1, 1)
    assert comb[0, 0] == np.float32(18.005714416503906)
    assert np.array_equal(metrics.euclidean_distances(gs[0], gs[1]), k['euclid_3mm'])
    assert np.array_equal(np.rad2deg(metrics.angular_distances(gs[0], gs[1])), k['angular_15deg'])
    assert metrics.coverage_brute_force(gs, gs[0:20]) == 0.4 == metrics.coverage(gs, gs[0:20])


def test_pairwise_distances_bit_exact(metrics):
    g = load_golden('metrics_small')
    ref, qry = _gs(g['ref']), _gs(g['qry'])
    assert np.array_equal(metrics.euclidean_distances(ref, qry), g['euclidean'])
    assert np.array_equal(metrics.angular_distances(ref, qry), g['angular'])
    assert np.array_equal(metrics.combined_distances(ref, qry), g['combined'])


@pytest.mark.parametrize('algo', ['tiles', 'grid'])
def test_coverage_golden(metrics, algo):
    g = load_golden('metrics_small')
    ref, qry = g['ref'], g['qry']
    for kd in (True, False):
        want = port.covered_kd(ref, qry) if kd else port.covered_brute_force(ref, qry)
        got = metrics.covered_mask(_gs(ref), _gs(qry), kd_semantics=kd, algo=algo)
        assert np.array_equal(got, want)
    assert metrics.coverage(_gs(ref), _gs(qry)) == float(g['coverage_kd'])
    assert metrics.coverage_brute_force(_gs(ref), _gs(qry)) == float(g['coverage_brute'])
    assert metrics.precision(_gs(ref), _gs(qry)) == float(g['precision_kd'])
    assert metrics.coverage(_gs(ref), _gs(qry), epsilon=5.0) == float(g['coverage_eps5'])
    assert metrics.coverage_brute_force(_gs(ref), _gs(qry), epsilon=30.0) == float(g['coverage_eps30'])


@pytest.mark.parametrize('algo', ['tiles', 'grid'])
def test_coverage_mid_golden(metrics, algo):
    g = load_golden('metrics_mid')
    for q, key in [('qry_uniform', 'coverage_uniform'), ('qry_perturbed', 'coverage_perturbed')]:
        got = metrics.covered_mask(_gs(g['ref']), _gs(g[q]), algo=algo)
        assert got.mean() == float(g[key])
        assert np.array_equal(got, port.covered_kd(g['ref'], g[q]))
    assert metrics.precision(_gs(g['ref']), _gs(g['qry_perturbed'])) == float(g['precision_perturbed'])


def test_avg_gripper_point_distances_bit_exact(metrics):
    g = load_golden('metrics_keypoints')
    ref, qry = _gs(g['ref']), _gs(g['qry'])
    out = metrics.avg_gripper_point_distances(ref, qry)
    assert out.dtype == np.float64 and np.array_equal(out, g['avg'])
    assert np.array_equal(metrics.avg_gripper_point_distances(ref, qry, consider_symmetry=True), g['avg_symmetric'])
    assert np.array_equal(metrics.avg_gripper_point_distances(ref, qry, gripper_points=g['custom_points']), g['avg_custom'])
    assert np.array_equal(metrics.avg_gripper_point_distances(ref[0], qry), g['avg'][0:1])
    with pytest.raises(IndexError):
        metrics.avg_gripper_point_distances(ref, qry, gripper_points=g['custom_points'], consider_symmetry=True)


def test_edge_cases(metrics):
    rows = random_rows(37, 4, cube=0.01)
    gs = _gs(rows)
    assert metrics.coverage(gs, gs) == 1.0                                  # identical sets
    assert metrics.coverage(gs, gs, epsilon=0.0) == 1.0                     # 0 <= 0
    empty = _gs(np.zeros((0, 14), dtype=np.float32))
    assert metrics.coverage(gs, empty) == 0.0
    far = rows.copy()
    far[:, 0] += 10.0
    assert metrics.coverage(gs, _gs(far)) == 0.0
    one = _gs(rows[:1])
    assert metrics.coverage(one, gs) == 1.0 and metrics.combined_distances(one, one)[0, 0] == 0.0
    for algo in ('tiles', 'grid'):
        m = metrics.covered_mask(gs, _gs(rows[5:6]), algo=algo)
        assert m[5] and np.array_equal(m, port.covered_kd(rows, rows[5:6]))


@pytest.mark.parametrize('n,m,cube', [(1025, 2049, 0.05), (5000, 3000, 0.08), (3000, 5000, 1.0)])
def test_coverage_ragged_sizes_vs_oracle(metrics, n, m, cube):
    ref = random_rows(n, 21, cube)
    qry = np.concatenate([random_rows(m - m // 3, 22, cube), perturbed_rows(ref, 23, frac=(m // 3) / n)])[:m]
    want_kd, want_bf = port.covered_kd(ref, qry), port.covered_brute_force(ref, qry)
    for algo in ('tiles', 'grid'):
        assert np.array_equal(metrics.covered_mask(_gs(ref), _gs(qry), algo=algo), want_kd)
        assert np.array_equal(metrics.covered_mask(_gs(ref), _gs(qry), kd_semantics=False, algo=algo), want_bf)
    assert 0 < want_kd.sum() < n


def test_threshold_band(metrics):
    """pairs placed exactly on / one float32 ulp around the translation threshold (angular part 0)."""
    base = np.zeros((1, 14), dtype=np.float32)
    base[0, 3], base[0, 7], base[0, 11] = 1, 1, 1
    qs = []
    d = np.float32(0.015)
    for delta in [np.nextafter(d, np.float32(0)), d, np.nextafter(d, np.float32(1)), np.float32(0.0150001)]:
        q = base.copy()
        q[0, 0] = delta
        qs.append(q)
    for q in qs:
        for kd in (True, False):
            want = port.covered_kd(base, q) if kd else port.covered_brute_force(base, q)
            for algo in ('tiles', 'grid'):
                got = metrics.covered_mask(_gs(base), _gs(q), kd_semantics=kd, algo=algo)
                assert np.array_equal(got, want), (float(q[0, 0]), kd, algo)


def test_angular_threshold_band(metrics):
    """pairs whose angular part sits around epsilon (and around epsilon minus the weighted translation): the orientation
    bound of both kernels (a conservative quaternion test before the exact recipe) must never rule out a pair the
    reference's recipe accepts -- every case against the oracle, both semantics, both algorithms."""
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(5)
    n = 64
    base = random_rows(n, 31, cube=0.05)
    refs, qrys = [], []
    for k, ang in enumerate([14.0, 14.9, 14.97, 14.99, 14.995, 15.0, 15.005, 15.01, 15.03, 15.1, 16.0, 9.99, 10.0, 10.01, 179.9, 180.0]):
        for d_mm in (0.0, 5.0):                                  # deg + 1000 d against 15: d = 5 mm moves the band to 10 degrees
            r = base[(2 * k) % n].copy()
            q = r.copy()
            axis = rng.normal(size=3)
            axis /= np.linalg.norm(axis)
            rot = Rotation.from_rotvec(np.deg2rad(ang) * axis).as_matrix()
            q[3:12] = (rot @ r[3:12].reshape(3, 3).astype(np.float64)).reshape(9)
            q[0] += np.float32(d_mm / 1000.0)
            r[1] += np.float32(1.0 * len(refs))                    # every pair far from all the others
            q[1] += np.float32(1.0 * len(qrys))
            refs.append(r)
            qrys.append(q)
    ref, qry = np.stack(refs).astype(np.float32), np.stack(qrys).astype(np.float32)
    for kd in (True, False):
        want = port.covered_kd(ref, qry) if kd else port.covered_brute_force(ref, qry)
        assert 0 < want.sum() < len(ref)
        for algo in ('tiles', 'grid'):
            got = metrics.covered_mask(_gs(ref), _gs(qry), kd_semantics=kd, algo=algo)
            assert np.array_equal(got, want), (kd, algo, np.flatnonzero(got != want))


def test_rows_that_are_not_rotations(metrics):
    """the reference's recipe takes any 3 x 3 block (trace of R1 R2^T through arccos, NaN beyond 1); blocks that are not
    rotation matrices -- scaled, sheared, reflected, zero, NaN -- bypass the kernels' orientation bound and must give
    the oracle's answer"""
    rng = np.random.default_rng(6)
    ref = random_rows(600, 41, cube=0.03)
    qry = np.concatenate([random_rows(300, 42, cube=0.03), perturbed_rows(ref, 43)])
    for rows, seed in ((ref, 1), (qry, 2)):
        r = np.random.default_rng(seed)
        idx = r.choice(len(rows), len(rows) // 3, replace=False)
        kinds = r.integers(0, 6, len(idx))
        for i, kind in zip(idx, kinds):
            m = rows[i, 3:12].reshape(3, 3).astype(np.float64)
            if kind == 0:
                m = m * r.uniform(0.5, 1.02)
            elif kind == 1:
                m = m + r.normal(0, 0.05, (3, 3))
            elif kind == 2:
                m = m @ np.diag([1.0, 1.0, -1.0])
            elif kind == 3:
                m = np.zeros((3, 3))
            elif kind == 4:
                m = r.normal(0, 1, (3, 3))
            else:
                m = m.copy()
                m[r.integers(0, 3), r.integers(0, 3)] = np.nan
            rows[i, 3:12] = m.reshape(9).astype(np.float32)
    for kd in (True, False):
        want = port.covered_kd(ref, qry) if kd else port.covered_brute_force(ref, qry)
        for algo in ('tiles', 'grid'):
            got = metrics.covered_mask(_gs(ref), _gs(qry), kd_semantics=kd, algo=algo)
            assert np.array_equal(got, want), (kd, algo, np.flatnonzero(got != want)[:10])
    assert 0 < want.sum() < len(ref)


def test_full_size_properties(metrics):
    """BASELINE config 2 (100k x 100k): size-independent properties instead of an oracle run."""
    import torch
    n = 100_000
    ref = random_rows(n, 1)
    qry = np.concatenate([random_rows(n // 2, 2), perturbed_rows(ref, 3)])
    dref, dqry = torch.from_numpy(ref).cuda(), torch.from_numpy(qry).cuda()
    grid = metrics.covered_mask(dref, dqry, algo='grid')
    tiles = metrics.covered_mask(dref, dqry, algo='tiles')
    assert np.array_equal(grid, tiles)                                        # two algorithms, one answer
    assert metrics.coverage(dref, dref) == 1.0                                # idempotence
    sub = metrics.covered_mask(dref, dqry[: n // 4], algo='grid')
    assert not (sub & ~grid).any()                                            # monotone in the query set
    idx = np.random.default_rng(0).choice(n, 300, replace=False)
    assert np.array_equal(grid[idx], port.covered_kd(ref[idx], qry))           # spot check against the oracle
    assert 0.2 < grid.mean() < 0.8


def test_prepared_query_gives_the_same_masks():
    """metrics.prepare_query (query set resident + binned once, reference rows binned per call) against the one-shot grid
    call and the all-pairs kernel: dense and sparse grids, both semantics, reference rows outside the query set's bounds,
    shards of the reference set, pinned staging"""
    from burg_toolkit_b200 import metrics
    from burg_toolkit_b200.synthetic import coverage_sets
    for n, cube in ((6000, 0.05), (3000, 2.0)):                      # ~30 queries per cell (cell-tile kernel) / sparse (grid walk)
        ref, qry = coverage_sets(n, cube)
        ref = ref.copy()
        ref[:40, 0:3] += np.float32(1.5 * cube)                       # far outside the query bounds
        ref[40:80, 0:3] -= np.float32(0.004)                          # just outside / at the border
        pq = metrics.prepare_query(qry, 15.0)
        for kd in (True, False):
            want = metrics.covered_mask(ref, qry, kd_semantics=kd, algo='tiles')
            assert np.array_equal(metrics.covered_mask(ref, qry, kd_semantics=kd, algo='grid'), want)
            assert np.array_equal(metrics.covered_mask(ref, pq, kd_semantics=kd), want)
            pinned = {}
            got = np.concatenate([metrics.covered_mask(ref[b:e], pq, kd_semantics=kd, pinned=pinned) for b, e in ((0, 1000), (1000, 1001), (1001, len(ref)))])
            assert np.array_equal(got, want)
        assert want.sum() > 10 and not want[:40].any()
        assert metrics.coverage(ref, pq) == metrics.coverage(ref, qry)
    with pytest.raises(ValueError):
        metrics.covered_mask(ref, pq, epsilon=10.0)


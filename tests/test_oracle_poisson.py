"""CPU tests of the oracle's restatement of open3d's weighted sample elimination (oracle/leaves.poisson_disk_eliminate,
SURVEY.md Appendix A.2): exact count, blue-noise property, known tiny case."""
import numpy as np
from scipy.spatial import cKDTree

from oracle import leaves


def test_elimination_keeps_exactly_n_and_spreads_them():
    rng = np.random.default_rng(0)
    pts = np.concatenate([rng.random((4000, 2)), np.zeros((4000, 1))], axis=1)         # unit square, area 1
    keep = leaves.poisson_disk_eliminate(pts, 800, 1.0)
    assert keep.sum() == 800
    d_kept = cKDTree(pts[keep]).query(pts[keep], k=2)[0][:, 1]
    d_rand = cKDTree(pts[:800]).query(pts[:800], k=2)[0][:, 1]
    # a uniform random subset has nearest neighbours all the way down to zero; the eliminated set keeps them apart
    assert np.median(d_kept) > 1.5 * np.median(d_rand)
    assert d_kept.min() > 5 * d_rand.min()
    r_max = 2.0 * np.sqrt((1.0 / 800) / (2.0 * np.sqrt(3.0)))
    assert d_kept.min() > 0.35 * r_max


def test_elimination_removes_the_crowded_sample_first():
    # five points on a line: the middle one of the tight triple has the largest weight and goes first
    pts = np.array([[0.0, 0, 0], [0.1, 0, 0], [0.2, 0, 0], [1.0, 0, 0], [2.0, 0, 0]])        # spacing above r_min = 0.076
    keep = leaves.poisson_disk_eliminate(pts, 4, 1.0)
    assert keep.tolist() == [True, False, True, True, True]
    assert leaves.poisson_disk_eliminate(pts, 5, 1.0).all()

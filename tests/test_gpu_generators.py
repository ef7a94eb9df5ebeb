"""GPU tests (through the C ABI) of the device-side generators around the sampler."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _gs(rows):
    from burg_toolkit_b200.grasp import GraspSet
    return GraspSet.from_array(np.ascontiguousarray(rows, dtype=np.float32))


def test_grasp_perturbations_match_reference():
    from burg_toolkit_b200 import sampling
    g = load_golden('generators')
    got = sampling.grasp_perturbations(_gs(g['gs']))
    want = g['perturbed_default']
    assert got._gs_array.shape == want.shape == (20 * 37, 14)
    # translations are exact; rotations go through Rodrigues instead of scipy's code path: float32 rows within 1e-6
    assert np.allclose(got._gs_array, want, rtol=1e-5, atol=1e-6)
    assert (got._gs_array == want).mean() > 0.95
    assert np.array_equal(got._gs_array[::37], g['gs'])                         # originals, incl. score and width
    got = sampling.grasp_perturbations(_gs(g['gs'])[3], radii=[2], include_original_grasp=False)
    assert np.allclose(got._gs_array, g['perturbed_custom'], rtol=1e-5, atol=1e-6)
    with pytest.raises(ValueError):
        sampling.grasp_perturbations(_gs(g['gs']), radii=(5, 10))
    with pytest.raises(ValueError):
        sampling.grasp_perturbations(np.zeros((3, 14)))


def test_farthest_point_sampling_index_exact():
    from burg_toolkit_b200 import sampling
    g = load_golden('generators')
    assert np.array_equal(sampling.farthest_point_sampling(g['cloud'], 200), g['fps_f64'])
    assert np.array_equal(sampling.farthest_point_sampling(g['cloud'].astype(np.float32), 150), g['fps_f32'])
    assert np.array_equal(sampling.farthest_point_sampling(g['cloud'][:10], 10)[:1], [0])
    with pytest.raises(ValueError):
        sampling.farthest_point_sampling(g['cloud'][:5], 6)


def test_random_poses_distribution():
    from burg_toolkit_b200 import sampling
    tf = sampling.random_poses(200_000, seed=3)
    assert tf.shape == (200_000, 4, 4) and tf.dtype == np.float64
    r, t = tf[:, :3, :3], tf[:, :3, 3]
    assert np.allclose(r @ r.transpose(0, 2, 1), np.eye(3), atol=1e-12) and np.allclose(np.linalg.det(r), 1.0)
    assert np.array_equal(tf[:, 3], np.tile([0, 0, 0, 1.0], (len(tf), 1)))
    assert t.min() >= 0 and t.max() < 1 and np.allclose(t.mean(axis=0), 0.5, atol=5e-3)
    # Haar measure: E[trace] = 0, P(angle <= theta) = (theta - sin theta) / pi
    ang = np.arccos(np.clip((np.trace(r, axis1=1, axis2=2) - 1) / 2, -1, 1))
    assert abs(np.trace(r, axis1=1, axis2=2).mean()) < 0.01
    for theta in (0.5, 1.0, 2.0):
        assert abs((ang <= theta).mean() - (theta - np.sin(theta)) / np.pi) < 5e-3
    assert np.array_equal(tf, sampling.random_poses(200_000, seed=3)) and not np.array_equal(tf, sampling.random_poses(200_000, seed=4))

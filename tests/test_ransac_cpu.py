"""Oracle-side checks of the correspondence-RANSAC restatement (no GPU): the draws, the rigid fit and the selection rule."""
import numpy as np
import torch

from oracle import diffreg_oracle as orc


def test_draws_are_in_range_uniform_and_counter_based():
    d = orc.ransac_draws(42, 0, 20000, 3, 57)
    assert d.shape == (20000, 3) and d.min() >= 0 and d.max() == 56
    hist = np.bincount(d.ravel(), minlength=57)
    assert hist.min() > 0.8 * 60000 / 57 and hist.max() < 1.2 * 60000 / 57
    # a trial's draws depend on (seed, b, trial, draw) only
    assert (orc.ransac_draws(42, 0, 100, 3, 57) == d[:100]).all()
    assert (orc.ransac_draws(42, 1, 100, 3, 57) != d[:100]).any()
    assert (orc.ransac_draws(43, 0, 100, 3, 57) != d[:100]).any()
    # known answer of the generator (splitmix64 of the counter, then of seed ^ that): pins the CUDA side's constants too
    assert orc._splitmix64(np.array([0], dtype=np.uint64))[0] == np.uint64(0xE220A8397B1DCDAF)


def test_rigid_fit_recovers_a_motion_and_flags_degenerate_samples():
    g = torch.Generator().manual_seed(0)
    R = orc.random_rotation(g).numpy().astype(np.float64)
    t = np.array([0.3, -0.2, 0.1])
    x = np.random.default_rng(0).normal(size=(5, 3, 3)).astype(np.float32)
    y = (x @ R.T + t).astype(np.float32)
    x[4, 1] = x[4, 0]  # repeated point: two distinct points only
    y[4, 1] = y[4, 0]
    Rf, tf, valid = orc.rigid_fit(x, y)
    assert valid.tolist() == [True, True, True, True, False]
    assert np.abs(Rf[:4] - R).max() < 1e-5 and np.abs(tf[:4] - t).max() < 1e-5
    assert np.allclose(np.linalg.det(Rf[:4].astype(np.float64)), 1.0, atol=1e-5)
    # a mirrored target still gives a proper rotation (det fix)
    Rm, _, _ = orc.rigid_fit(x[:1], (y[:1] * np.array([1, 1, -1], dtype=np.float32)))
    assert np.linalg.det(Rm[0].astype(np.float64)) > 0.99


def test_selection_rule_and_planted_pose():
    assert orc.ransac_best_trial([3, 7, 7, -1], [0.1, 0.5, 0.2, 0.0]) == 2
    assert orc.ransac_best_trial([3, 7, 7, 7], [0.1, 0.2, 0.2, 0.3]) == 1
    assert orc.ransac_best_trial([0, -1, 0], [0, 0, 0]) == -1
    rng = np.random.default_rng(1)
    src = rng.uniform(-1, 1, size=(2, 50, 3)).astype(np.float32)
    tgt = rng.uniform(-1, 1, size=(2, 60, 3)).astype(np.float32)
    g = torch.Generator().manual_seed(3)
    R = orc.random_rotation(g).numpy().astype(np.float32)
    t = np.array([0.1, 0.2, -0.3], dtype=np.float32)
    i = rng.permutation(50)[:40]
    j = rng.permutation(60)[:40]
    tgt[0, j[:20]] = src[0, i[:20]] @ R.T + t
    match = np.concatenate([np.stack([np.zeros(40, np.int64), i, j], 1), np.array([[1, 0, 0], [1, 1, 1]])])
    out = orc.ransac_correspondence(src, tgt, match, 0.05, 3, 2000, seed=5)
    assert out["inlier_count"][0] >= 20 and abs(out["fitness"][0] - out["inlier_count"][0] / 40) < 1e-6
    assert np.abs(out["pose"][0, :3, :3] - R).max() < 1e-4 and np.abs(out["pose"][0, :3, 3] - t).max() < 1e-4
    assert (out["pose"][1] == np.eye(4)).all() and out["best_trial"][1] == -1  # two matches: identity (loss.py:384-387)


def test_reference_signatures_of_the_ransac_entry_points():
    """registration.ransac_pose_estimation / ransac_regist_coarse accept every call the reference's functions accept
    (Diff-Reg-4dmatch/models/loss.py:13, 366): same leading parameter names, order and defaults.  loss.py imports open3d and
    cannot be imported here, so its signatures are read from the source with ast; skipped when the reference is absent."""
    import ast
    import inspect
    import os

    import pytest
    path = "/root/reference/Diff-Reg-4dmatch/models/loss.py"
    if not os.path.exists(path):
        pytest.skip("reference sources not present")
    from diffreg_b200 import registration
    tree = ast.parse(open(path).read())
    found = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in ("ransac_pose_estimation", "ransac_regist_coarse"):
            names = [a.arg for a in node.args.args]
            defaults = [ast.literal_eval(d) for d in node.args.defaults]
            found[node.name] = (names, dict(zip(names[len(names) - len(defaults):], defaults)))
    assert set(found) == {"ransac_pose_estimation", "ransac_regist_coarse"}
    for name, (names, defaults) in found.items():
        ours = inspect.signature(getattr(registration, name)).parameters
        assert list(ours)[:len(names)] == names, (name, list(ours), names)
        for k, v in defaults.items():
            assert ours[k].default == v, (name, k)

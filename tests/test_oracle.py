"""CPU tests of the oracle (the restatement) against the only known-answer material the reference
has and against the committed golden vectors."""
import hashlib
import os

import numpy as np
import pytest

from mad_icp_b200 import synth
from util import bits_equal, pose_error

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_four_walls_matches_reference_generator_digest():
    # digest recorded by tests/golden/make_golden.py after asserting equality with the reference's
    # own generator (apps/utils/tools/tools_utils.py:3-21) under np.random.seed(42)
    g = np.load(os.path.join(GOLD, "four_walls_registration.npz"))
    np.random.seed(42)
    cloud = synth.four_walls(points_per_wall=1000)
    assert hashlib.sha256(cloud.tobytes()).hexdigest() == str(g["cloud_sha256"])


def test_kat_self_query_error_is_exactly_zero(oracle):
    """apps/utils/tools/nn_search.py:36-61 + tools/README.md:9-10: b_max=1e-5, query the cloud against
    itself -> total matching error == 0."""
    np.random.seed(42)
    cloud = synth.four_walls()  # 5 x 10000 points
    tree = oracle.OracleTree(cloud, b_max=1e-5, b_min=0.1)
    idx = tree.search(cloud)
    means, normals, _, npts = tree.leaves()
    assert tree.num_leaves == cloud.shape[0] and (npts == 1).all()
    assert np.linalg.norm(means[idx] - cloud, axis=1).sum() == 0.0
    assert np.isfinite(normals).all()


def test_kat_self_registration_converges_to_identity(oracle):
    """apps/utils/tools/mad_registration.py:48-69: ground truth is the identity."""
    g = np.load(os.path.join(GOLD, "four_walls_registration.npz"))
    np.random.seed(42)
    cloud = synth.four_walls(points_per_wall=1000)
    ref = oracle.OracleTree(cloud)
    qry = oracle.OracleTree(cloud.copy())
    r = oracle.icp_run([ref], qry, g["T_guess"], iters=15)
    ang, dt = pose_error(r["X"], np.eye(4))
    assert ang < 1e-6 and dt < 1e-6
    assert r["matched"].all()
    # golden pin of the restatement itself
    assert bits_equal(r["X"], g["X"]) and bits_equal(r["H_hist"], g["H_hist"]) and bits_equal(r["b_hist"], g["b_hist"])
    assert (r["idx_hist"] == g["idx_hist"]).all()
    assert ref.num_leaves == int(g["num_leaves"]) and ref.num_nodes == int(g["num_nodes"])


def test_golden_lidar_small(oracle):
    g = np.load(os.path.join(GOLD, "lidar_small_registration.npz"))
    c = synth.registration_case(K=2, beams=16, azimuths=512, seed=3)
    assert hashlib.sha256(c["query"].tobytes()).hexdigest() == str(g["query_sha256"])
    trees = [oracle.OracleTree(s) for s in c["scans"]]
    for t, P in zip(trees, c["kf_poses"]):
        t.apply_transform(P)
    q = oracle.OracleTree(c["query"])
    for nthreads in (1, 2):  # thread-order sum of per-thread adders: 1 vs 2 threads may differ in the last bits
        r = oracle.icp_run(trees, q, c["T_guess"], iters=10, num_threads=nthreads)
        ang, dt = pose_error(r["X"], g["X"])
        assert ang < 1e-9 and dt < 1e-9
        if nthreads == 2:
            assert bits_equal(r["X"], g["X"]) and (r["idx_hist"] == g["idx_hist"]).all()
    ang, dt = pose_error(g["X"], g["T_true"])
    assert ang < 2e-3 and dt < 2e-2  # converges to the simulated truth within sensor noise


def test_eig3_against_numpy(oracle):
    rs = np.random.RandomState(0)
    for _ in range(2000):
        A = rs.standard_normal((rs.randint(3, 30), 3)) * rs.uniform(0.01, 10, 3)
        cov = np.cov(A.T)
        w, V = oracle.eig3(cov)
        w_np, _ = np.linalg.eigh(cov)
        assert np.allclose(w, w_np, rtol=1e-9, atol=1e-12 * abs(w_np).max())
        assert np.allclose(V.T @ V, np.eye(3), atol=1e-9)
        assert np.allclose(cov @ V, V * w, atol=1e-9 * abs(w_np).max())


def test_eig3_degenerate(oracle):
    w, V = oracle.eig3(np.zeros((3, 3)))
    assert (V == np.eye(3)).all() and (w == 0).all()
    w, V = oracle.eig3(np.eye(3) * 2.5)
    assert (V == np.eye(3)).all()


def test_solve_update_against_numpy(oracle):
    rs = np.random.RandomState(1)
    for _ in range(200):
        J = rs.standard_normal((40, 6))
        H = J.T @ J
        b = rs.standard_normal(6)
        dx, Xn = oracle.solve_update(H, b, np.eye(4))
        assert np.allclose(dx, np.linalg.solve(H, -b), rtol=1e-8, atol=1e-10)
        w = dx[3:]
        th = np.linalg.norm(w)
        K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]) / th
        R = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K
        assert np.allclose(Xn[:, :3], R, atol=1e-12) and np.allclose(Xn[:, 3], dx[:3])


def test_solve_rank_deficient_is_finite(oracle):
    dx, Xn = oracle.solve_update(np.zeros((6, 6)), np.zeros(6), np.eye(4))
    assert (dx == 0).all() and bits_equal(Xn, np.eye(4)[:3])
    H = np.diag([1.0, 2.0, 0.0, 0.0, 3.0, 0.0])
    dx, _ = oracle.solve_update(H, np.ones(6), np.eye(4))
    assert np.isfinite(dx).all() and np.allclose(dx, [-1, -0.5, 0, 0, -1 / 3, 0])


def test_tree_structure_invariants(oracle):
    c = synth.registration_case(K=1, beams=16, azimuths=512, seed=5)
    t = oracle.OracleTree(c["scans"][0])
    e = t.export()
    internal = e["left"] >= 0
    assert ((e["left"] >= 0) == (e["right"] >= 0)).all()  # internal nodes always have both children
    assert t.num_nodes == 2 * t.num_leaves - 1
    assert (e["bbox"][~internal, 2] < 0.2).all() and (e["bbox"][internal, 2] >= 0.2).all()
    assert e["num_points"][0] == c["scans"][0].shape[0]
    assert e["num_points"][~internal].sum() == c["scans"][0].shape[0]
    # leaf means are cloud points; single-point leaves have zero extent
    one = (~internal) & (e["num_points"] == 1)
    assert (e["bbox"][one] == 0).all()
    assert np.isfinite(e["mean"]).all() and np.isfinite(e["eivecs"][:, :3][~internal]).all()

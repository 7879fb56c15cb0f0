"""Device-resident keyframe lifecycle, `-m gpu`: MADtree::applyTransform fused into the upload
(tools/mad_tree.cpp:165-172), the 4-ary layout built on the device, moving leaves taken from a device tree,
Frame::weight_ = det(H^-1) from the solve thread (odometry/pipeline.cpp:223), the budget-limited loop
(pipeline.cpp:167-176).  Everything is compared BIT FOR BIT with the host path / the CPU oracle."""
import numpy as np
import pytest

from mad_icp_b200 import FlatTree, MadIcpError, Registrar, synth
from mad_icp_b200 import _capi as capi
from util import bits_equal, pose_error

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def case():
    return synth.registration_case(K=3, beams=32, azimuths=1024, seed=11)


@pytest.fixture(scope="module")
def host_path(case):
    """Reference arrangement: transform on the HOST, upload the map-frame tree, host leaf means."""
    reg = Registrar(device=0, max_keyframes=4)
    for k, (scan, P) in enumerate(zip(case["scans"], case["kf_poses"])):
        ft = FlatTree(scan)
        ft.apply_transform(P)
        reg.put_keyframe(k, ft)
    q = FlatTree(case["query"])
    reg.set_moving(q.leaf_means())
    return reg, q


def _same_registration(a, b, X0, iters=10):
    ia, ib = a.search(X0), b.search(X0)
    assert (ia == ib).all()
    ra, rb = a.register(X0, iters=iters), b.register(X0, iters=iters)
    for k in ("X", "H", "b"):
        assert bits_equal(ra[k], rb[k]), k
    assert (ra["matched"] == rb["matched"]).all() and ra["n_matched"] == rb["n_matched"]
    return ra


def test_apply_transform_on_the_device_is_bit_identical(case, host_path):
    reg_h, _ = host_path
    reg_d = Registrar(device=0, max_keyframes=4)
    for k, (scan, P) in enumerate(zip(case["scans"], case["kf_poses"])):
        reg_d.put_keyframe(k, FlatTree(scan), T=P)  # sensor-frame tree + pose: transformed during the upload
    reg_d.set_moving(host_path[1].leaf_means())
    _same_registration(reg_h, reg_d, case["T_guess"])


def test_device_tree_upload_tables_and_moving_leaves(case, host_path):
    reg_h, q = host_path
    reg = Registrar(device=0, max_keyframes=4)
    dt = reg.upload_tree(q)
    assert (dt.num_nodes, dt.num_leaves) == (q.num_nodes, q.num_leaves)
    assert dt.records().tobytes() == q.records().tobytes()
    recs = q.records()
    leaf = np.nonzero(recs["link"] < 0)[0]
    want = np.empty(q.num_leaves, np.int32)
    want[-1 - recs["link"][leaf]] = leaf
    assert (dt.leaf_records() == want).all()
    # promotion from device trees (with the pose applied on the device) + moving leaves from a device tree
    for k, (scan, P) in enumerate(zip(case["scans"], case["kf_poses"])):
        reg.put_keyframe(k, reg.upload_tree(FlatTree(scan)), T=P)
    reg.set_moving_tree(dt)
    assert bits_equal(reg.get_moving(), q.leaf_means())
    _same_registration(reg_h, reg, case["T_guess"])
    # a device tree can be promoted more than once and survives the slot being overwritten
    reg.put_keyframe(3, dt, T=case["kf_poses"][0])
    reg.drop_keyframe(3)
    _same_registration(reg_h, reg, case["T_guess"], iters=3)


def _inv_det_partial_pivot(H):
    A = np.array(H, dtype=np.float64).copy()
    det = 1.0
    for k in range(6):
        p = k
        for i in range(k + 1, 6):
            if abs(A[i, k]) > abs(A[p, k]):
                p = i
        if p != k:
            A[[k, p]] = A[[p, k]]
            det = -det
        det = det * A[k, k]
        for i in range(k + 1, 6):
            f = A[i, k] / A[k, k]
            for c in range(k, 6):
                A[i, c] = A[i, c] - f * A[k, c]
    return 1.0 / det


def test_frame_weight_from_the_solve_thread(case, host_path):
    reg, _ = host_path
    reg.register_async(case["T_guess"], 15)
    out = reg.register_fetch()
    want = _inv_det_partial_pivot(out["H"])
    assert out["weight"] == want or abs(out["weight"] - want) <= 4e-16 * abs(want), (out["weight"], want)
    assert np.isfinite(out["weight"]) and out["weight"] > 0


def test_budget_limited_loop_keeps_the_union_of_matches(case, host_path):
    """pipeline.cpp:167-176: when `realtime` breaks the loop before round MAX_ICP_ITS-1 the matched flags were
    never cleared, so they are the union over the rounds that ran."""
    reg, _ = host_path
    reg.register_async(case["T_guess"], 4, partial=True)
    out = reg.register_fetch(want_matched=True)
    tr = reg.register_trace()
    union = np.zeros(reg.L, bool)
    for r in range(4):
        union |= reg.linearize(tr[r])[2] != 0
    assert (out["matched"] != 0).tolist() == union.tolist()
    assert out["n_matched"] == int(union.sum())
    full = reg.register(case["T_guess"], iters=4)  # the normal loop: last round only
    assert (full["matched"] != 0).sum() <= union.sum()
    assert bits_equal(full["X"], out["X"])


def test_zero_and_many_rounds(case, host_path, oracle):
    reg, _ = host_path
    X0 = capi.pose12(case["T_guess"])
    out = reg.register(X0, iters=0)
    assert bits_equal(out["X"], X0) and out["n_matched"] == 0      # the reference's loop with 0 rounds returns T
    out = reg.register(X0, iters=100)                                  # more than one launch holds (64)
    ref = reg.register(X0, iters=64)
    ref2 = reg.register(ref["X"], iters=36)
    assert bits_equal(out["X"], ref2["X"]) and (out["matched"] == ref2["matched"]).all()


def test_put_keyframe_records_validation(host_path):
    _, q = host_path
    reg = Registrar(device=0, max_keyframes=1)
    recs = q.records()
    reg.put_keyframe_records(0, recs, q.num_leaves)  # well-formed
    bad = recs.copy()
    internal = np.nonzero(bad["link"] >= 0)[0]
    bad["link"][internal[5]] = bad["link"][internal[4]]  # two parents for one pair: a node becomes unreachable
    with pytest.raises(MadIcpError, match="breadth-first"):
        reg.put_keyframe_records(0, bad, q.num_leaves)
    bad = recs.copy()
    leaves = np.nonzero(bad["link"] < 0)[0]
    bad["link"][leaves[0]] = bad["link"][leaves[1]]      # duplicate leaf ordinal
    with pytest.raises(MadIcpError, match="permutation"):
        reg.put_keyframe_records(0, bad, q.num_leaves)


def test_calibration_keeps_results(case, host_path):
    reg, _ = host_path
    before = reg.register(case["T_guess"], iters=10)
    assert reg.calibrate(case["T_guess"]) >= 1
    after = reg.register(case["T_guess"], iters=10)
    ang, dt = pose_error(before["X"], after["X"])
    assert ang < 1e-10 and dt < 1e-10  # another shape sums in another order: last bits only
    assert (before["matched"] == after["matched"]).mean() > 0.9999


def test_path_memo_changes_nothing(oracle):
    """From round 1 on the persistent kernel skips the walks it can prove unchanged (kernels.cuh, descend_t).  The
    proof must be airtight: with the memo off every pair is walked in every round -- poses, H, b, matched flags and
    the per-round trace must be bit-identical, from a far initial guess (large moves between rounds), from the
    optimum (no move), and at BASELINE size."""
    for kw, guess_shift in ((dict(K=3, beams=32, azimuths=1024, seed=11), 0.0), (dict(K=3, beams=32, azimuths=1024, seed=11), 1.5),
                            (dict(K=16), 0.0)):
        c = synth.registration_case(**kw)
        reg = Registrar(device=0, max_keyframes=16)
        for k, (scan, P) in enumerate(zip(c["scans"], c["kf_poses"])):
            reg.put_keyframe(k, FlatTree(scan), T=P)
        reg.set_moving(FlatTree(c["query"]).leaf_means())
        X0 = np.array(c["T_guess"], dtype=np.float64)
        X0[0, 3] += guess_shift
        for iters in (1, 2, 10, 15):
            reg.set_memo(True)
            a = reg.register(X0, iters=iters)
            ta = reg.register_trace()
            reg.set_memo(False)
            b = reg.register(X0, iters=iters)
            tb = reg.register_trace()
            for k in ("X", "H", "b"):
                assert bits_equal(a[k], b[k]), (kw, iters, k)
            assert (a["matched"] == b["matched"]).all() and a["n_matched"] == b["n_matched"]
            assert bits_equal(ta, tb)
        # restart from the converged pose: nothing moves, every walk of rounds >= 1 is skipped
        reg.set_memo(True)
        a = reg.register(b["X"], iters=5)
        reg.set_memo(False)
        b2 = reg.register(b["X"], iters=5)
        assert bits_equal(a["X"], b2["X"]) and bits_equal(a["H"], b2["H"])


def test_round_timeline_stamps(case, host_path):
    """madicp_debug_cta_stamps: every CTA's items start before they end, the tile goes out after them, CTA 0 has folded
    after the last tile went out and hands the pose out after that; the next round starts after the pose is out.  The
    instrumented launches return the same registration as the plain ones."""
    reg, _ = host_path
    iters = 6
    plain = reg.register(case["T_guess"], iters=iters)
    reg.debug_timing(True, fetch=False)
    try:
        out = reg.register(case["T_guess"], iters=iters)
        d = reg.debug_timing(True)
        start, end, pub = (reg.debug_cta_stamps(p, iters) for p in (1, 2, 3))
        trace = reg.debug_cta_stamps(4, iters)[:, :16]
    finally:
        reg.debug_timing(False, fetch=False)
    for k in ("X", "H", "b"):
        assert bits_equal(out[k], plain[k]), k
    assert d.shape[0] == iters and start.shape == end.shape == pub.shape and start.shape[0] == iters
    assert (start > 0).all() and (start <= end).all() and (end <= pub).all()
    folded, handed = d[:, 6], d[:, 7]
    assert (folded >= pub.max(axis=1) - 1024).all()  # (the timer ticks every few hundred ns)
    assert (handed >= folded).all()
    assert (start[1:].min(axis=1) >= handed[:-1] - 1024).all()
    assert (trace[:, 13] > trace[:, 0]).all() and (trace[:, 14] >= 1).all()

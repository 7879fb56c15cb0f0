"""The reference-named pybind surface (pyvector / pymadtree / pymadicp).  The GPU tests read like the
reference's own demos (apps/utils/tools/nn_search.py, mad_registration.py) with the imports swapped."""
import os

import numpy as np
import pytest

from mad_icp_b200 import synth
from util import POSE_M, POSE_RAD, pose_error

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_vector_eigen3d_semantics(built):
    from mad_icp_b200.pybind.pyvector import VectorEigen3d
    a = np.random.RandomState(0).rand(7, 3)
    v = VectorEigen3d(a)
    assert len(v) == 7 and bool(v) and (np.asarray(v) == a).all()
    assert np.asarray(v).strides == (24, 8)  # reference buffer protocol: eigen_stl_bindings.h:73-80
    assert (v[2] == a[2]).all() and (v[-1] == a[-1]).all()
    assert len(list(v)) == 7 and not VectorEigen3d()
    with pytest.raises(RuntimeError):  # py::cast_error on a wrong shape (eigen_stl_bindings.h:48-50)
        VectorEigen3d(np.zeros((4, 2)))
    v.append(np.array([1.0, 2.0, 3.0]))
    assert len(v) == 8 and "std::vector<Eigen::Vector3d> with 8 elements" in repr(v)
    from mad_icp_b200.pybind import pymadicp, pymadtree
    assert pymadicp.VectorEigen3d is VectorEigen3d and pymadtree.VectorEigen3d is VectorEigen3d


def test_signatures_match_reference(built):
    """Names and defaults of pymadicp.cpp:36-52 / pymadtree.cpp:36-48."""
    from mad_icp_b200.pybind import pymadicp, pymadtree
    d = pymadicp.MADicp.compute.__doc__
    for frag in ("T:", "icp_iterations: ", "= 15", "rho_ker: ", "= 0.1", "b_ratio: ", "= 0.02", "print_stats: "):
        assert frag in d, frag
    d = pymadicp.MADicp.setQueryCloud.__doc__
    assert "b_max: " in d and "= 0.2" in d and "b_min: " in d and "= 0.1" in d
    d = pymadtree.MADtree.build.__doc__
    assert "b_max: " in d and "1e-05" in d and "max_parallel_level: " in d and "= 2" in d
    for name in ("search", "searchCloud", "searchCloudDist"):
        assert hasattr(pymadtree.MADtree, name)


def test_no_gpu_fails_loudly(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mad_icp_b200.pybind.pymadicp import MADicp
    m = MADicp(num_threads=2)
    cloud = np.random.RandomState(1).rand(50, 3)
    m.setReferenceCloud(cloud)
    m.setQueryCloud(cloud)
    with pytest.raises(RuntimeError, match="CUDA"):
        m.compute(np.eye(4))


@pytest.mark.gpu
def test_mad_registration_demo(oracle):
    """apps/utils/tools/mad_registration.py:48-69 with the imports swapped."""
    from mad_icp_b200.pybind.pymadicp import MADicp
    from mad_icp_b200.pybind.pyvector import VectorEigen3d
    g = np.load(os.path.join(GOLD, "four_walls_registration.npz"))
    np.random.seed(42)
    ref_cloud = synth.four_walls(points_per_wall=1000)
    query_cloud = ref_cloud.copy()
    T_guess = np.eye(4)
    T_guess[:3, :3] = synth.euler_xyz(0.1, 0.1, 0.1)
    T_guess[:3, 3] = np.random.rand(3)
    assert (T_guess == g["T_guess"]).all()
    madicp = MADicp(num_threads=os.cpu_count())
    madicp.setReferenceCloud(VectorEigen3d(ref_cloud))
    madicp.setQueryCloud(VectorEigen3d(query_cloud))
    T_est = madicp.compute(T_guess, icp_iterations=15)
    assert T_est.shape == (4, 4) and (T_est[3] == [0, 0, 0, 1]).all()
    ang, dt = pose_error(T_est, np.eye(4))
    assert ang < 1e-6 and dt < 1e-6            # ground truth of the demo is the identity
    ang, dt = pose_error(T_est, g["X"])
    assert ang < POSE_RAD and dt < POSE_M      # and it is the oracle's answer
    # one iteration at a time, as the demo's visual loop does (mad_registration.py:86-88)
    T = T_guess.copy()
    for _ in range(15):
        T = madicp.compute(T, icp_iterations=1)
    ang, dt = pose_error(T, g["X"])
    assert ang < POSE_RAD and dt < POSE_M


@pytest.mark.gpu
def test_nn_search_demo(oracle):
    """apps/utils/tools/nn_search.py:36-61: total matching error of the cloud against itself == 0."""
    from mad_icp_b200.pybind.pymadtree import MADtree
    from mad_icp_b200.pybind.pyvector import VectorEigen3d
    np.random.seed(42)
    cloud = synth.four_walls()
    tree = MADtree()
    tree.build(VectorEigen3d(cloud))
    ref_point, ref_normal = tree.search(cloud[0, :])
    assert np.linalg.norm(ref_point - cloud[0]) == 0.0 and ref_normal.shape == (3,)
    ref_cloud = tree.searchCloud(VectorEigen3d(cloud[:2000]))
    tot = 0.0
    for (p, n), q in zip(ref_cloud, cloud[:2000]):
        tot += np.linalg.norm(p - q)
    assert tot == 0.0 and len(ref_cloud) == 2000
    P, N, D = tree.searchCloudArrays(cloud)
    assert np.linalg.norm(P - cloud, axis=1).sum() == 0.0 and (D == 0).all()
    d3 = tree.searchCloudDist(cloud[:10] + 0.01)
    ot = oracle.OracleTree(cloud, b_max=1e-5)
    means, normals, _, _ = ot.leaves()
    oi = ot.search(cloud[:10] + 0.01)
    for (p, n, d), i, q in zip(d3, oi, cloud[:10] + 0.01):
        assert (p == means[i]).all() and (n == normals[i]).all()
        assert abs(d - np.linalg.norm(q - means[i])) <= 4e-16 * d


def _sequence(n, beams=16, azimuths=512):
    scene = synth.StreetScene(seed=7)
    for i in range(n):
        base = synth.pose_xyyaw(0.8 * i, 1.0 + 0.02 * i, 0.004 * i)
        yield 0.1 * i, np.ascontiguousarray(synth.lidar_scan(scene, base, beams=beams, azimuths=azimuths, seed=100 + i))


def test_pipeline_surface(built):
    """pypeline.cpp:57-74: constructor arguments and methods."""
    from mad_icp_b200.pybind.pypeline import Pipeline, VectorEigen3d
    from mad_icp_b200.pybind.pyvector import VectorEigen3d as V2
    assert VectorEigen3d is V2
    d = Pipeline.__init__.__doc__
    for a in ("sensor_hz", "deskew", "b_max", "rho_ker", "p_th", "b_min", "b_ratio", "num_keyframes", "num_threads", "realtime"):
        assert a in d
    for mname in ("currentPose", "trajectory", "keyframePose", "isInitialized", "isMapUpdated", "currentID", "keyframeID",
                  "modelLeaves", "currentLeaves", "compute"):
        assert hasattr(Pipeline, mname)


@pytest.mark.gpu
@pytest.mark.parametrize("deskew", [False, True])
def test_pipeline_tracks_like_the_oracle(oracle, deskew):
    """Streaming odometry (BASELINE cfg5 shape, shortened): the GPU Pipeline and the CPU restatement of
    odometry/pipeline.cpp follow the same trajectory, promote the same keyframes."""
    import ctypes as C
    from mad_icp_b200.pybind.pypeline import Pipeline, VectorEigen3d
    L = oracle.lib()
    L.orc_pipeline_create.restype = C.c_void_p
    L.orc_pipeline_create.argtypes = [C.c_double, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                      C.c_int, C.c_int, C.c_int]
    L.orc_pipeline_compute.argtypes = [C.c_void_p, C.c_double, oracle._dp, C.c_int]
    L.orc_pipeline_state.argtypes = [C.c_void_p, oracle._dp]
    L.orc_pipeline_free.argtypes = [C.c_void_p]
    ref = C.c_void_p(L.orc_pipeline_create(10.0, int(deskew), 0.2, 0.1, 0.8, 0.1, 0.02, 4, 4, 0))
    pipe = Pipeline(sensor_hz=10.0, deskew=deskew, b_max=0.2, rho_ker=0.1, p_th=0.8, b_min=0.1, b_ratio=0.02,
                    num_keyframes=4, num_threads=4, realtime=False)
    st = np.zeros(23)
    n = 25 if not deskew else 10
    for i, (stamp, pts) in enumerate(_sequence(n)):
        L.orc_pipeline_compute(ref, stamp, oracle._d(pts), pts.shape[0])
        L.orc_pipeline_state(ref, oracle._d(st))
        pipe.compute(stamp, VectorEigen3d(pts) if i % 2 == 0 else pts)  # a bound vector or a plain N x 3 array
        T = pipe.currentPose()
        ang, dt = pose_error(T, st[:12].reshape(3, 4))
        assert pipe.currentID() == int(st[13]), i
        if not deskew:
            # the scan does not depend on earlier estimates: identical trees, lock-step trajectories and
            # identical keyframe decisions
            assert ang < POSE_RAD and dt < POSE_M, (i, ang, dt)
            assert pipe.isMapUpdated() == bool(st[12]), i
            assert pipe.keyframeID() == int(st[14]) and pipe.numKeyframes() == int(st[15]), i
            assert abs(pipe.inliersRatio() - st[16]) < 2e-3, i
        else:
            # deskewing feeds the previous pose estimates (equal to ~1e-12, not bit-equal) into the CLOUD; the
            # tree build is discontinuous in its input (a point changing side moves a split), so the two
            # runs register slightly different leaf sets and agree at the sensor-noise level, not at 1e-5
            # (the synthetic sweep is instantaneous, so deskewing it by 0.8 m/scan also distorts it: the
            # registration is softer than in the undistorted case)
            assert ang < 5e-3 and dt < 5e-2, (i, ang, dt)
    assert pipe.isInitialized() and len(pipe.trajectory()) == n
    if not deskew:  # the simulated vehicle moves 0.8 m per scan along x
        assert abs(pipe.currentPose()[0, 3] - 0.8 * (n - 1)) < 0.05
    assert len(pipe.currentLeaves()) > 100 and len(pipe.modelLeaves()) > len(pipe.currentLeaves())
    L.orc_pipeline_free(ref)


def test_deskew_is_bit_identical_to_the_cpu_pipeline(oracle):
    """Pipeline::deskew (pipeline.cpp:79-123) is host code on both sides: same sorted order (ties
    included), same chunk poses, same points."""
    import ctypes as C
    from mad_icp_b200.pybind.pypeline import Pipeline
    L = oracle.lib()
    L.orc_pipeline_create.restype = C.c_void_p
    L.orc_pipeline_create.argtypes = [C.c_double, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                      C.c_int, C.c_int, C.c_int]
    L.orc_pipeline_deskew.argtypes = [C.c_void_p, oracle._dp, C.c_int, oracle._dp, oracle._dp]
    L.orc_pipeline_free.argtypes = [C.c_void_p]
    po = C.c_void_p(L.orc_pipeline_create(10.0, 1, 0.2, 0.1, 0.8, 0.1, 0.02, 4, 1, 0))
    Ta, Tb = synth.pose_xyyaw(0.0, 1.0, 0.0), synth.pose_xyyaw(0.8, 1.05, 0.03)
    Tb[2, 3] = 0.02
    _, pts = next(_sequence(1, beams=32, azimuths=1024))
    rs = np.random.RandomState(5)
    dup = pts[rs.randint(0, pts.shape[0], 500)] * np.array([1.0, 1.0, 0.5])  # same azimuth as another point: ties
    clouds = [pts, np.concatenate([pts, dup])[rs.permutation(pts.shape[0] + 500)], pts[:1], pts[:2]]
    for cloud in clouds:
        cloud = np.ascontiguousarray(cloud)
        want = cloud.copy()
        L.orc_pipeline_deskew(po, oracle._d(want), want.shape[0], oracle._d(np.ascontiguousarray(Ta[:3])),
                              oracle._d(np.ascontiguousarray(Tb[:3])))
        for threads in (1, 4):
            got = np.asarray(Pipeline._deskewOnly(cloud, Ta, Tb, 10.0, threads))
            assert got.shape == want.shape and (got == want).all(), threads
    # full-size scan: the threaded merge sort path (no ties), and with a few exact duplicates (fallback path)
    full = np.ascontiguousarray(synth.registration_case(K=1)["scans"][0])
    for cloud in (full, np.concatenate([full, full[:7]])):
        want = cloud.copy()
        L.orc_pipeline_deskew(po, oracle._d(want), want.shape[0], oracle._d(np.ascontiguousarray(Ta[:3])),
                              oracle._d(np.ascontiguousarray(Tb[:3])))
        got = np.asarray(Pipeline._deskewOnly(cloud, Ta, Tb, 10.0, 8))
        assert (got == want).all()
    L.orc_pipeline_free(po)


def test_threaded_sort_leaves_std_sorts_permutation(built):
    """madicp_deskew sorts by azimuth with a threaded restatement of std::sort; equal keys are the rule
    (one firing column), so the order among them must be std::sort's.  Keys with few distinct values, all
    equal, all distinct; sizes around the algorithm's thresholds."""
    from mad_icp_b200 import _capi
    L = _capi.lib()
    for n in (0, 1, 2, 15, 16, 17, 33, 1000, 20000, 131072):
        for distinct in (1, 2, 7, 100, 10 ** 9):
            for threads in (2, 5, 16):
                assert L.madicp_debug_sort_check(n, 11, distinct, threads) == 0, (n, distinct, threads)

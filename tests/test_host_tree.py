"""CPU tests of the product's host flat-tree builder (mad_icp_b200/csrc/flat_tree.cpp) against the
oracle's pointer tree: every node bit-identical, leaves in getLeafs order, record layout."""
import numpy as np
import pytest

from mad_icp_b200 import FlatTree, MadIcpError, synth
from util import bits_equal


def _same_tree(ft, ot):
    a, b = ft.export(), ot.export()
    assert ft.num_nodes == ot.num_nodes and ft.num_leaves == ot.num_leaves
    for k in ("mean", "eivecs", "bbox"):
        assert bits_equal(a[k], b[k]), k
    for k in ("num_points", "left", "right", "leaf_ordinal"):
        assert (a[k] == b[k]).all(), k
    for x, y in zip(ft.leaves(), ot.leaves()):
        assert bits_equal(x, y) if x.dtype.kind == "f" else (x == y).all()


@pytest.mark.parametrize("b_max,ppw", [(0.2, 1000), (1e-5, 2000), (0.05, 3000)])
def test_four_walls_tree_identical(oracle, built, b_max, ppw):
    np.random.seed(42)
    cloud = synth.four_walls(points_per_wall=ppw)
    _same_tree(FlatTree(cloud, b_max=b_max), oracle.OracleTree(cloud, b_max=b_max))


@pytest.mark.parametrize("threads", [2, 5, 16])
def test_threaded_build_is_bit_identical(oracle, built, threads):
    """Subtrees expanded on a thread pool and spliced back: same nodes, same order, same bits."""
    c = synth.registration_case(K=1, beams=32, azimuths=1024, seed=12)
    _same_tree(FlatTree(c["scans"][0], num_threads=threads), oracle.OracleTree(c["scans"][0]))


@pytest.mark.parametrize("threads", [3, 16, 64])
def test_threaded_build_full_size_scan(oracle, built, threads):
    """131 072 points: the top levels are processed with the passes over one node shared between threads
    (chunked extents, closed-form split applied in parallel), the subtrees below on the pool."""
    c = synth.registration_case(K=1, seed=5)
    _same_tree(FlatTree(c["scans"][0], num_threads=threads), oracle.OracleTree(c["scans"][0]))


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_threaded_build_unstructured_clouds(oracle, built, seed):
    """Blobs, a line and heavy duplicates: split sizes far from n/2, leaves high up in the tree."""
    rs = np.random.RandomState(seed)
    cloud = np.concatenate([rs.normal(0, s, (n, 3)) + rs.uniform(-20, 20, 3)
                            for s, n in ((0.01, 9000), (1.0, 12000), (5.0, 7000), (0.2, 3000))]
                           + [np.c_[np.linspace(0, 30, 4000), np.zeros((4000, 2))]]
                           + [np.repeat(rs.uniform(-5, 5, (10, 3)), 300, axis=0)])
    cloud = cloud[rs.permutation(cloud.shape[0])]
    for b_max in (0.2, 1e-5):
        _same_tree(FlatTree(cloud, b_max=b_max, num_threads=7), oracle.OracleTree(cloud, b_max=b_max))


def test_concurrent_builds_from_two_host_threads(oracle, built):
    """The process-wide pool serves one build; a second one running at the same time gets its own."""
    import threading
    c = synth.registration_case(K=2, beams=32, azimuths=1024, seed=21)
    out = [None, None]

    def work(i):
        for _ in range(3):
            out[i] = FlatTree(c["scans"][i], num_threads=4)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for i in range(2):
        _same_tree(out[i], oracle.OracleTree(c["scans"][i]))


def test_lidar_tree_identical_after_transform(oracle, built):
    c = synth.registration_case(K=1, beams=32, azimuths=1024, seed=11)
    ft, ot = FlatTree(c["scans"][0]), oracle.OracleTree(c["scans"][0])
    _same_tree(ft, ot)
    T = c["kf_poses"][0] @ synth.pose_xyyaw(3.0, -1.0, 0.3)
    ft.apply_transform(T)
    ot.apply_transform(T)
    _same_tree(ft, ot)


@pytest.mark.parametrize("n", [1, 2, 3, 7])
def test_tiny_clouds(oracle, built, n):
    rs = np.random.RandomState(n)
    cloud = rs.uniform(-1, 1, (n, 3))
    _same_tree(FlatTree(cloud, b_max=0.05), oracle.OracleTree(cloud, b_max=0.05))


def test_duplicate_points(oracle, built):
    cloud = np.repeat(np.array([[1.0, 2.0, 3.0], [1.5, 2.0, 3.0]]), 50, axis=0)
    _same_tree(FlatTree(cloud, b_max=0.1), oracle.OracleTree(cloud, b_max=0.1))


def test_empty_cloud_rejected(built):
    with pytest.raises(MadIcpError):
        FlatTree(np.zeros((0, 3)))


def test_record_layout(built):
    """Breadth-first, siblings adjacent, leaves carry -1-ordinal, internal dir = split direction."""
    c = synth.registration_case(K=1, beams=16, azimuths=512, seed=2)
    ft = FlatTree(c["scans"][0])
    r, e = ft.records(), ft.export()
    assert r.dtype.itemsize == 64 and r.shape[0] == ft.num_nodes
    leaf = r["link"] < 0
    assert leaf.sum() == ft.num_leaves
    assert sorted((-1 - r["link"][leaf]).tolist()) == list(range(ft.num_leaves))
    links = r["link"][~leaf]
    assert (links >= 1).all() and (np.diff(links) == 2).all() and links[0] == 1  # BFS: children pairs in order
    # follow the records from the root and compare with the DFS export
    def walk(rec_i, node_i):
        if e["left"][node_i] < 0:
            assert r["link"][rec_i] == -1 - e["leaf_ordinal"][node_i]
            assert bits_equal(r["dir"][rec_i], e["eivecs"][node_i][0:3]) and r["bbox0"][rec_i] == e["bbox"][node_i][0]
            return
        assert bits_equal(r["dir"][rec_i], e["eivecs"][node_i][6:9]) and bits_equal(r["mean"][rec_i], e["mean"][node_i])
        walk(r["link"][rec_i], e["left"][node_i])
        walk(r["link"][rec_i] + 1, e["right"][node_i])
    import sys
    sys.setrecursionlimit(10000)
    walk(0, 0)


@pytest.mark.filterwarnings("ignore::DeprecationWarning")  # Python's generic fork-with-threads warning
def test_build_in_a_forked_child(oracle, built):
    """A forked child (e.g. a data-loader worker) inherits the worker pool object without its threads;
    it must notice and start its own."""
    import multiprocessing as mp
    c = synth.registration_case(K=1, beams=32, azimuths=1024, seed=31)
    parent = FlatTree(c["scans"][0], num_threads=4)  # the pool exists in the parent now
    ctx = mp.get_context("fork")
    q = ctx.Queue()

    def child():
        ft = FlatTree(c["scans"][0], num_threads=4)
        q.put((ft.num_nodes, ft.num_leaves))

    p = ctx.Process(target=child)
    p.start()
    p.join(60)
    assert p.exitcode == 0, "child hung or crashed"
    assert q.get(timeout=5) == (parent.num_nodes, parent.num_leaves)

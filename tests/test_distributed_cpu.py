"""world_size-2 `gloo` tests (CPU) of the N>1 host logic: keyframe sharding, handle exchange and the
rank-ordered reduction rule.  The per-shard compute is the ORACLE (the checker standing in for the
device), so this proves the distributed algebra, not the kernels: sharded GN == unsharded GN."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from mad_icp_b200 import distributed as D, synth
    from oracle import oracle as O
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        K = 5
        case = synth.registration_case(K=K, beams=8, azimuths=256, seed=9)
        trees = []
        for s in range(K):
            t = O.OracleTree(case["scans"][s])
            t.apply_transform(case["kf_poses"][s])
            trees.append(t)
        moving = O.OracleTree(case["query"])
        mine = D.shard_slots(K, rank, world)
        assert sorted(sum((D.shard_slots(K, r, world) for r in range(world)), [])) == list(range(K))
        # handle exchange: 64 opaque bytes per rank, everyone ends with the same table in rank order
        table = D.exchange_handles(bytes([rank] * 64))
        assert table == [bytes([r] * 64) for r in range(world)]
        X = case["T_guess"][:3].copy()
        for _ in range(6):
            H, b, m = O.icp_linearize([trees[s] for s in mine], moving, X)
            tot = D.sum_in_rank_order(np.concatenate([H.ravel(), b]))
            mall = D.sum_in_rank_order(m.astype(np.float64)) > 0     # OR of the matched flags
            _, X = O.solve_update(tot[:36].reshape(6, 6), tot[36:], X)
        q.put((rank, X.tobytes(), mall.tobytes(), O.icp_run(trees, moving, case["T_guess"], iters=6)["X"].tobytes(),
               O.icp_run(trees, moving, case["T_guess"], iters=6)["matched"].tobytes()))
    finally:
        dist.destroy_process_group()


def test_sharded_gauss_newton_equals_unsharded(oracle):
    world, port = 2, 29600 + os.getpid() % 300
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    X0, X1 = (np.frombuffer(r[1]).reshape(3, 4) for r in res)
    assert res[0][1] == res[1][1], "ranks must hold bit-identical poses (rank-ordered sum, identical solve)"
    full = np.frombuffer(res[0][3]).reshape(3, 4)
    assert np.abs(X0 - full).max() < 1e-11
    m_sh = np.frombuffer(res[0][2], dtype=np.bool_)
    m_full = np.frombuffer(res[0][4], dtype=np.uint8) > 0
    assert (m_sh == m_full).mean() > 0.999


def test_shard_slots_round_robin():
    from mad_icp_b200.distributed import shard_slots
    assert shard_slots(16, 3, 8) == [3, 11] and shard_slots(16, 0, 1) == list(range(16))
    assert all(len(shard_slots(16, r, 8)) == 2 for r in range(8))

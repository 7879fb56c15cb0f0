"""Multi-GPU validation (run under torchrun, one rank per GPU):
keyframe slot s lives on rank s % world; every GN round all-reduces the 48-value H/b tile inside the
persistent kernel through NVLink peer mailboxes.  Checks: all ranks end with bit-identical X/H/b,
the result matches a single-GPU registration of the full model within tolerance, matched flags are the
OR over ranks, repeated calls stay consistent (epoch / double-buffer logic).
Usage: torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/multi_gpu_check.py [K] [beams] [azimuths]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from mad_icp_b200 import FlatTree, Registrar, synth

NO_TIMING = "--no-timing" in sys.argv
sys.argv = [a for a in sys.argv if a != "--no-timing"]
K = int(sys.argv[1]) if len(sys.argv) > 1 else 16
beams = int(sys.argv[2]) if len(sys.argv) > 2 else 32
az = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
case = synth.registration_case(K=K, beams=beams, azimuths=az, seed=5)
trees = []
for s in range(K):
    ft = FlatTree(case["scans"][s]); ft.apply_transform(case["kf_poses"][s]); trees.append(ft)
means = FlatTree(case["query"]).leaf_means()

reg = Registrar(device=lr, max_keyframes=K)
for s in range(K):
    if s % world == rank:
        reg.put_keyframe(s, trees[s])
reg.set_moving(means)
h = torch.tensor(list(reg.comm_export()), dtype=torch.uint8, device=f"cuda:{lr}")
allh = [torch.empty_like(h) for _ in range(world)]
dist.all_gather(allh, h)
reg.comm_connect(rank, world, [bytes(t.cpu().tolist()) for t in allh])
dist.barrier()

ref = Registrar(device=lr, max_keyframes=K)  # full model on this GPU: single-GPU answer
for s in range(K):
    ref.put_keyframe(s, trees[s])
ref.set_moving(means)
ok = True
for trial, iters in enumerate((10, 1, 15, 10)):
    out = reg.register(case["T_guess"], iters=iters)
    single = ref.register(case["T_guess"], iters=iters)
    buf = torch.tensor(np.concatenate([out["X"].ravel(), out["H"].ravel(), out["b"]]), device=f"cuda:{lr}")
    gathered = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf)
    same = all(torch.equal(gathered[0].view(torch.int64), g.view(torch.int64)) for g in gathered)
    dX = np.abs(out["X"] - single["X"]).max()
    relH = np.abs(out["H"] - single["H"]).max() / np.abs(single["H"]).max()
    m_ok = bool((out["matched"] == single["matched"]).all()) and out["n_matched"] == single["n_matched"]
    if rank == 0:
        print(f"trial {trial} iters={iters}: ranks bit-identical={same}  |X - X_single|max={dX:.2e}  relH={relH:.2e}  matched ok={m_ok} ({out['n_matched']})", flush=True)
    ok = ok and same and dX < 1e-7 and relH < 1e-9 and m_ok
# timing (device events, max over ranks)
st = torch.cuda.Stream(); reg.set_stream(st.cuda_stream); ref.set_stream(st.cuda_stream)
def timed(r, n=50):
    for _ in range(5): r.register_async(case["T_guess"], 10)
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(n): r.register_async(case["T_guess"], 10)
    b.record(st); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / n * 1e3], device=f"cuda:{lr}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
if not NO_TIMING:
  t_shard, t_single = timed(reg), timed(ref)
  if rank == 0:
    print(f"world={world} K={K} L={means.shape[0]}: sharded {t_shard:.1f} us/scan, single-GPU full model {t_single:.1f} us/scan, speed-up {t_single / t_shard:.2f}x")
okt = torch.tensor([int(ok)], device=f"cuda:{lr}")
dist.all_reduce(okt, op=dist.ReduceOp.MIN)  # every rank's own checks must hold
if rank == 0:
    print("MULTI_GPU_CHECK", "PASS" if int(okt.item()) else "FAIL", flush=True)
dist.barrier(); dist.destroy_process_group()

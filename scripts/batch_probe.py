"""Batched device tree builds: python scripts/batch_probe.py [batch sizes ...]  (MADICP_BUILD_TIMING=1 for the per-level
breakdown; under `ncu --metrics gpu__time_duration.sum` the launch list gives the per-kernel split of a forest build)"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from mad_icp_b200 import Registrar, synth
sizes = [int(x) for x in sys.argv[1:]] or [1, 4, 16, 32]
seq = synth.sequence(max(sizes), workers=8)["scans"]
reg = Registrar(device=0, max_keyframes=2)
pinned = [torch.from_numpy(s).pin_memory().numpy() for s in seq]
reps = int(os.environ.get("PROBE_REPS", "3"))
for B in sizes:
    for rep in range(reps):
        t0 = time.perf_counter(); trees = reg.build_trees(pinned[:B]); reg.synchronize(); t1 = time.perf_counter()
        del trees
    print(f"pinned batch of {B}: {1e3 * (t1 - t0):.2f} ms  ({1e3 * (t1 - t0) / B:.2f} ms per scan)", flush=True)

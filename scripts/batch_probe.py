"""Batched device tree builds: python scripts/batch_probe.py  (MADICP_BUILD_TIMING=1 for the breakdown)"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from mad_icp_b200 import Registrar, synth
seq = synth.sequence(16, workers=8)["scans"]
reg = Registrar(device=0, max_keyframes=2)
pinned = [torch.from_numpy(s).pin_memory().numpy() for s in seq]
for name, clouds in (("pageable", seq), ("pinned", pinned)):
    for B in (1, 4, 16):
        for rep in range(3):
            t0 = time.perf_counter(); trees = reg.build_trees(clouds[:B]); reg.synchronize(); t1 = time.perf_counter()
            del trees
        print(f"{name} batch of {B}: {1e3 * (t1 - t0):.2f} ms  ({1e3 * (t1 - t0) / B:.2f} ms per scan)", flush=True)

"""Device MAD-tree build timing (python scripts/build_probe_gpu.py): per-build wall time, per-level breakdown
(MADICP_BUILD_TIMING=1 prints it), host builder beside it."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from mad_icp_b200 import FlatTree, Registrar, synth
c = synth.registration_case(K=1)
cloud = np.ascontiguousarray(c["query"])
reg = Registrar(device=0, max_keyframes=2)
for rep in range(4):
    t0 = time.perf_counter(); dt = reg.build_tree(cloud); reg.synchronize(); t1 = time.perf_counter()
    print(f"device build: {1e3 * (t1 - t0):.2f} ms  nodes={dt.num_nodes} leaves={dt.num_leaves} levels={dt.num_levels}", flush=True)
for thr in (1, 16):
    for rep in range(3):
        t0 = time.perf_counter(); ft = FlatTree(cloud, num_threads=thr); t1 = time.perf_counter()
    print(f"host build ({thr} threads): {1e3 * (t1 - t0):.2f} ms")
f32 = cloud.astype(np.float32)
for rep in range(3):
    t0 = time.perf_counter(); reg.ingest(f32); dt = reg.build_tree(); reg.synchronize(); t1 = time.perf_counter()
print(f"ingest(float32) + device build: {1e3 * (t1 - t0):.2f} ms")

import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from mad_icp_b200 import FlatTree, Registrar, synth
case = synth.registration_case(K=16)
reg = Registrar(device=0, max_keyframes=16)
for s in range(16):
    ft = FlatTree(case["scans"][s]); ft.apply_transform(case["kf_poses"][s]); reg.put_keyframe(s, ft)
reg.set_moving(FlatTree(case["query"]).leaf_means())
st = torch.cuda.Stream(); reg.set_stream(st.cuda_stream)
for shape in ((768, 1), (704, 1), (896, 1), (640, 1), (1024, 1)):
    reg.set_gn_grid(*shape)
    reg.debug_timing(False, fetch=False)
    ts = []
    for _ in range(20):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st); reg.register_async(case["T_guess"], 10); b.record(st); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    reg.debug_timing(True, fetch=False)
    for _ in range(2):
        reg.register_async(case["T_guess"], 10); torch.cuda.synchronize()
    d = reg.debug_timing(True)
    print("shape", shape, f"10-iter warm median {np.median(ts):.1f} us",
          "per-round all_arrived cycles:", d[:, 1].astype(int).tolist())

"""Streaming odometry (BASELINE.json configs[4] shape, shortened): a synthetic KITTI-shape sequence through
the reference-named Pipeline (pypeline) on the GPU and through the CPU restatement of odometry/pipeline.cpp,
num_keyframes=16.  Reports end-to-end scans/s of Pipeline.compute (host MAD-tree build of every scan
included) and the ATE between the two trajectories.  Not a bench.py line (the path metric is registration).
Usage: python scripts/stream_bench.py [n_scans=60] [beams=64] [azimuths=2048]"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mad_icp_b200 import synth
from mad_icp_b200.pybind.pypeline import Pipeline, VectorEigen3d
from oracle import oracle as O

n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
beams = int(sys.argv[2]) if len(sys.argv) > 2 else 64
az = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
scene = synth.StreetScene(seed=7, x_min=-45.0, x_max=60.0 + 0.8 * n)


def make_scan(i):
    base = synth.pose_xyyaw(0.8 * i, 1.0 + 0.3 * np.sin(0.05 * i), 0.02 * np.sin(0.03 * i))
    return np.ascontiguousarray(synth.lidar_scan(scene, base, beams=beams, azimuths=az, seed=100 + i))


if n > 64:  # ray-casting the sequence is the slow part of a long run: spread it over the host cores (before CUDA starts)
    import multiprocessing as mp
    with mp.get_context("fork").Pool(min(32, os.cpu_count() or 1)) as pool:
        scans = pool.map(make_scan, range(n), chunksize=4)
else:
    scans = [make_scan(i) for i in range(n)]
threads = min(16, os.cpu_count() or 1)
L = O.lib()
L.orc_pipeline_create.restype = C.c_void_p
L.orc_pipeline_create.argtypes = [C.c_double, C.c_int] + [C.c_double] * 5 + [C.c_int] * 3
L.orc_pipeline_compute.argtypes = [C.c_void_p, C.c_double, O._dp, C.c_int]
L.orc_pipeline_state.argtypes = [C.c_void_p, O._dp]
args = dict(sensor_hz=10.0, deskew=False, b_max=0.2, rho_ker=0.1, p_th=0.8, b_min=0.1, b_ratio=0.02, num_keyframes=16,
            num_threads=threads, realtime=False)
pipe = Pipeline(**args)
pipe.compute(0.0, VectorEigen3d(scans[0]))  # initialise (builds keyframe 0)
t0 = time.perf_counter()
gpu_traj = []
for i in range(1, n):
    pipe.compute(0.1 * i, VectorEigen3d(scans[i]))
    gpu_traj.append(pipe.currentPose()[:3, 3].copy())
t_gpu = time.perf_counter() - t0
ref = C.c_void_p(L.orc_pipeline_create(10.0, 0, 0.2, 0.1, 0.8, 0.1, 0.02, 16, threads, 0))
st = np.zeros(23)
L.orc_pipeline_compute(ref, 0.0, O._d(scans[0]), scans[0].shape[0])
t0 = time.perf_counter()
cpu_traj = []
for i in range(1, n):
    L.orc_pipeline_compute(ref, 0.1 * i, O._d(scans[i]), scans[i].shape[0])
    L.orc_pipeline_state(ref, O._d(st))
    cpu_traj.append(st[[3, 7, 11]].copy())
t_cpu = time.perf_counter() - t0
g, c = np.array(gpu_traj), np.array(cpu_traj)
ate = float(np.sqrt(((g - c) ** 2).sum(1).mean()))
print(f"scans={n - 1} pts/scan={scans[0].shape[0]} keyframes(end)={pipe.numKeyframes()} | GPU Pipeline {(n - 1) / t_gpu:.1f} scans/s "
      f"({1e3 * t_gpu / (n - 1):.2f} ms/scan, host tree build included) | CPU restatement ({threads} threads) "
      f"{(n - 1) / t_cpu:.2f} scans/s | ATE(GPU vs CPU) = {ate:.2e} m | path length {np.linalg.norm(g[-1] - g[0]):.1f} m")

"""Synthetic inputs for the registration path: the four-walls demo cloud of the
reference's tools (apps/utils/tools/tools_utils.py:3-21) and a 64-beam x 2048-azimuth LiDAR
ray-caster over a procedural street scene.  Pure numpy; seeds are explicit so every array is
reproducible (SHA-256 digests are pinned in tests/golden/).
"""
import numpy as np


def four_walls(points_per_wall=10000, wall_height=2.0, wall_width=4.0, rng=None):
    """Four 4x2 m walls + floor, uniform samples.  Draw order (x, y, z per plane; walls y=0, y=w, x=0,
    x=w, then floor) matches the reference demo so `np.random.seed(42)` reproduces its cloud."""
    rs = np.random if rng is None else rng

    def plane(xr, yr, zr, n):
        x = rs.uniform(xr[0], xr[1], n)
        y = rs.uniform(yr[0], yr[1], n)
        z = rs.uniform(zr[0], zr[1], n)
        return np.column_stack((x, y, z))

    w, h, n = wall_width, wall_height, points_per_wall
    parts = [plane([0, w], [0, 0], [0, h], n), plane([0, w], [w, w], [0, h], n), plane([0, 0], [0, w], [0, h], n),
             plane([w, w], [0, w], [0, h], n), plane([0, w], [0, w], [0, 0], n)]
    return np.vstack(parts)


def euler_xyz(rx, ry, rz):
    """scipy Rotation.from_euler('xyz', ...) (extrinsic x, then y, then z): R = Rz @ Ry @ Rx."""
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def pose_xyyaw(x, y, yaw, z=0.0):
    T = np.eye(4)
    c, s = np.cos(yaw), np.sin(yaw)
    T[:3, :3] = [[c, -s, 0], [s, c, 0], [0, 0, 1]]
    T[:3, 3] = [x, y, z]
    return T


class StreetScene:
    """Ground plane z=0, canyon side walls (y=-8, +12, 6 m high), end walls (x=-45, +60+ext) and a
    deterministic set of axis-aligned boxes (cars, poles, kiosks)."""

    def __init__(self, seed=7, n_boxes=40, x_min=-45.0, x_max=60.0):
        self.x_min, self.x_max = x_min, x_max
        self.y_lo, self.y_hi, self.wall_h = -8.0, 12.0, 6.0
        r = np.random.RandomState(seed)
        boxes = []
        for _ in range(n_boxes):
            cx = r.uniform(x_min + 3, x_max - 3)
            cy = r.uniform(self.y_lo + 1.0, self.y_hi - 1.0)
            if abs(cy - 1.0) < 2.5:  # keep the driving lane (around y=1) free
                cy += 5.0 if cy >= 1.0 else -5.0
            kind = r.randint(3)
            sx, sy, sz = [(4.2, 1.8, 1.5), (0.3, 0.3, 5.0), (2.0, 2.0, 2.6)][kind]
            boxes.append((cx - sx / 2, cx + sx / 2, cy - sy / 2, cy + sy / 2, 0.0, sz))
        self.boxes = np.array(boxes)

    def cast(self, origin, dirs):
        """Nearest hit distance for rays origin + t*dirs (world frame); inf when nothing is hit."""
        o = np.asarray(origin, dtype=np.float64)
        d = dirs
        t_best = np.full(d.shape[0], np.inf)
        with np.errstate(divide="ignore", invalid="ignore"):
            # ground
            t = -o[2] / d[:, 2]
            px, py = o[0] + t * d[:, 0], o[1] + t * d[:, 1]
            ok = (t > 0) & (px >= self.x_min) & (px <= self.x_max) & (py >= self.y_lo) & (py <= self.y_hi)
            t_best = np.where(ok & (t < t_best), t, t_best)
            # side walls
            for yw in (self.y_lo, self.y_hi):
                t = (yw - o[1]) / d[:, 1]
                px, pz = o[0] + t * d[:, 0], o[2] + t * d[:, 2]
                ok = (t > 0) & (px >= self.x_min) & (px <= self.x_max) & (pz >= 0) & (pz <= self.wall_h)
                t_best = np.where(ok & (t < t_best), t, t_best)
            # end walls
            for xw in (self.x_min, self.x_max):
                t = (xw - o[0]) / d[:, 0]
                py, pz = o[1] + t * d[:, 1], o[2] + t * d[:, 2]
                ok = (t > 0) & (py >= self.y_lo) & (py <= self.y_hi) & (pz >= 0) & (pz <= self.wall_h)
                t_best = np.where(ok & (t < t_best), t, t_best)
            # boxes (slab test)
            inv = 1.0 / d
            for b in self.boxes:
                t1 = (b[0::2] - o) * inv
                t2 = (b[1::2] - o) * inv
                tn = np.nanmax(np.minimum(t1, t2), axis=1)
                tf = np.nanmin(np.maximum(t1, t2), axis=1)
                ok = (tf >= tn) & (tn > 0)
                t_best = np.where(ok & (tn < t_best), tn, t_best)
        return t_best


def lidar_scan(scene, T_sensor_to_world, beams=64, azimuths=2048, seed=0, sigma=0.01, sensor_height=1.73, r_min=0.7,
               r_max=120.0, elev_deg=(-24.8, 2.0)):
    """One sweep of a 64-beam spinning LiDAR (KITTI-like geometry) in the SENSOR frame, float64 Nx3.
    `T_sensor_to_world` is the pose of the vehicle base (z=0); the sensor sits `sensor_height` above it."""
    el = np.deg2rad(np.linspace(elev_deg[0], elev_deg[1], beams))
    az = np.linspace(-np.pi, np.pi, azimuths, endpoint=False)
    ce, se = np.cos(el)[:, None], np.sin(el)[:, None]
    d_s = np.stack([ce * np.cos(az)[None, :], ce * np.sin(az)[None, :], np.broadcast_to(se, (beams, azimuths))],
                   axis=-1).reshape(-1, 3)
    T = np.asarray(T_sensor_to_world, dtype=np.float64)
    R, t = T[:3, :3], T[:3, 3] + np.array([0.0, 0.0, sensor_height])
    rng = scene.cast(t, d_s @ R.T)
    rs = np.random.RandomState(seed)
    rng = rng + sigma * rs.standard_normal(rng.shape[0])
    keep = np.isfinite(rng) & (rng >= r_min) & (rng <= r_max)
    return np.ascontiguousarray(d_s[keep] * rng[keep, None])


def sensor_pose(T_base):
    """Base pose -> sensor pose (same rotation, lifted by the sensor height used in lidar_scan)."""
    T = np.array(T_base, dtype=np.float64)
    T[2, 3] += 1.73
    return T


def keyframe_poses(K, step=2.0, dyaw=0.01):
    """cfg3 model: keyframe k at x = step*k, yaw = dyaw*k, lane centre y=1."""
    return [pose_xyyaw(step * k, 1.0, dyaw * k) for k in range(K)]


def registration_case(K=16, beams=64, azimuths=2048, seed=1, scene_seed=7):
    """Synthetic cfg2/cfg3 workload: K keyframe scans + one query scan.
    Returns dict(scans=[K arrays, sensor frame], kf_poses=[K 4x4 sensor->map], query=Nx3 sensor frame,
    T_true=4x4 query sensor->map, T_guess=4x4)."""
    scene = StreetScene(seed=scene_seed)
    base = keyframe_poses(K)
    scans = [lidar_scan(scene, base[k], beams, azimuths, seed=seed + k) for k in range(K)]
    last = base[-1]
    q_base = last @ pose_xyyaw(0.8, 0.0, 0.02)
    query = lidar_scan(scene, q_base, beams, azimuths, seed=seed + 1000)
    # map frame := sensor frame of keyframe 0's vehicle base lifted by sensor height (poses are sensor->map)
    T_true = sensor_pose(q_base)
    T_guess = T_true @ pose_xyyaw(0.3, 0.0, 0.01)
    return dict(scans=scans, kf_poses=[sensor_pose(b) for b in base], query=query, T_true=T_true, T_guess=T_guess)


def sequence_scan(scene, i, beams=64, azimuths=2048, seed=100, step=0.8):
    """Scan i of the cfg5 sequence: `step` metres per scan along a gently curving lane."""
    base = pose_xyyaw(step * i, 1.0 + 0.3 * np.sin(0.05 * i), 0.02 * np.sin(0.03 * i))
    return np.ascontiguousarray(lidar_scan(scene, base, beams=beams, azimuths=azimuths, seed=seed + i))


def sequence(n_scans=60, beams=64, azimuths=2048, seed=100, step=0.8, workers=1):
    """cfg5 streaming workload: n_scans sensor-frame scans through a street scene long enough for the whole path.
    workers > 1 ray-casts the scans in forked processes (call before CUDA is initialised)."""
    scene = StreetScene(seed=7, x_min=-45.0, x_max=60.0 + step * n_scans)
    if workers > 1 and n_scans > 16:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(workers) as pool:
            scans = pool.starmap(sequence_scan, [(scene, i, beams, azimuths, seed, step) for i in range(n_scans)], chunksize=4)
    else:
        scans = [sequence_scan(scene, i, beams, azimuths, seed, step) for i in range(n_scans)]
    return dict(scans=scans, step=step)

"""Seeded synthetic bundle-adjustment windows.

The reference has no scene generator, no recorded inputs and no tests for its optimiser
(SURVEY.md §4), so the windows every test and the benchmark use are made here.  A window is
produced first in the *reference's* input form — robot poses T_world<-robot, world points,
float key points and float depths, exactly what `Optimizer::localOptimize` receives
(corelib/include/Optimizer/Optimizer.h:46-56) — and then marshalled to the C-ABI's flat arrays
with the same conversions `localOptimize` applies (corelib/src/Optimizer/Optimizer.cpp:100-114
for poses, :184-195 for measurements), so the float rounding of the disparity is part of the data.

Camera: 640x480, fx = fy = 420, cx = 320, cy = 240, baseline 0.05 m
(corelib/include/CameraModels/PinholeModel.h, Interface/ROS/src/InterfaceROS.cpp:20).
"""
from __future__ import annotations

import numpy as np

FX = FY = 420.0
CX, CY = 320.0, 240.0
BASELINE = 0.05
WIDTH, HEIGHT = 640, 480

# GeometricCamera::tansformFromImageToRobot_ (corelib/include/CameraModels/GeometricCamera.h:15-19)
R_ROBOT_FROM_IMAGE = np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])

BASE_SEED = 20261018


def rot_z(a):
    c, s = np.cos(a), np.sin(a)
    R = np.zeros(a.shape + (3, 3))
    R[..., 0, 0], R[..., 0, 1], R[..., 1, 0], R[..., 1, 1], R[..., 2, 2] = c, -s, s, c, 1.0
    return R


def small_rot(w):
    """Rodrigues rotation for axis-angle vectors w[..., 3]."""
    th = np.linalg.norm(w, axis=-1)[..., None, None]
    K = np.zeros(w.shape[:-1] + (3, 3))
    K[..., 0, 1], K[..., 0, 2] = -w[..., 2], w[..., 1]
    K[..., 1, 0], K[..., 1, 2] = w[..., 2], -w[..., 0]
    K[..., 2, 0], K[..., 2, 1] = -w[..., 1], w[..., 0]
    th_safe = np.where(th < 1e-12, 1.0, th)
    A = np.where(th < 1e-12, 1.0, np.sin(th_safe) / th_safe)
    B = np.where(th < 1e-12, 0.5, (1 - np.cos(th_safe)) / th_safe**2)
    return np.eye(3) + A * K + B * (K @ K)


def quat_from_R(R):
    """Eigen::Quaterniond(Matrix3d) followed by CameraPose::normalizeRotation
    (OptimizeTypeDefine.h:30-41): returns (x, y, z, w) with w >= 0, unit norm."""
    R = np.asarray(R, dtype=np.float64)
    out = np.empty(R.shape[:-2] + (4,))
    flat_R = R.reshape(-1, 3, 3)
    flat_q = out.reshape(-1, 4)
    for n, m in enumerate(flat_R):
        t = m[0, 0] + m[1, 1] + m[2, 2]
        if t > 0:
            t = np.sqrt(t + 1.0)
            w = 0.5 * t
            t = 0.5 / t
            x, y, z = (m[2, 1] - m[1, 2]) * t, (m[0, 2] - m[2, 0]) * t, (m[1, 0] - m[0, 1]) * t
        else:
            i = 0
            if m[1, 1] > m[0, 0]:
                i = 1
            if m[2, 2] > m[i, i]:
                i = 2
            j, k = (i + 1) % 3, (i + 2) % 3
            t = np.sqrt(m[i, i] - m[j, j] - m[k, k] + 1.0)
            q = [0.0, 0.0, 0.0]
            q[i] = 0.5 * t
            t = 0.5 / t
            w = (m[k, j] - m[j, k]) * t
            q[j] = (m[j, i] + m[i, j]) * t
            q[k] = (m[k, i] + m[i, k]) * t
            x, y, z = q
        v = np.array([x, y, z, w])
        if w < 0:
            v = -v
        flat_q[n] = v / np.sqrt(np.sum(v * v))
    return out


def R_from_quat(q):
    """Eigen::Quaternion::toRotationMatrix for q = (x, y, z, w)."""
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0], R[..., 0, 1], R[..., 0, 2] = 1 - (tyy + tzz), txy - twz, txz + twy
    R[..., 1, 0], R[..., 1, 1], R[..., 1, 2] = txy + twz, 1 - (txx + tzz), tyz - twx
    R[..., 2, 0], R[..., 2, 1], R[..., 2, 2] = txz - twy, tyz + twx, 1 - (txx + tyy)
    return R


def robot_to_camera_state(T_wr):
    """Optimizer.cpp:100-114: T_wc = T_wr * T_rc ; state = T_cw = T_wc^-1 as (t, q)."""
    T_wr = np.asarray(T_wr)
    R_wc = T_wr[..., :3, :3] @ R_ROBOT_FROM_IMAGE
    t_wc = T_wr[..., :3, 3]
    R_cw = np.swapaxes(R_wc, -1, -2)
    t_cw = -(R_cw @ t_wc[..., None])[..., 0]
    return np.concatenate([t_cw, quat_from_R(R_cw)], axis=-1)


def camera_state_to_robot(tq):
    """Optimizer.cpp:320-340: T_wr = (T_cw)^-1 * T_rc^-1."""
    tq = np.asarray(tq)
    R_cw = R_from_quat(tq[..., 3:7])
    R_wc = np.swapaxes(R_cw, -1, -2)
    t_wc = -(R_wc @ tq[..., :3, None])[..., 0]
    T = np.zeros(tq.shape[:-1] + (4, 4))
    T[..., :3, :3] = R_wc @ R_ROBOT_FROM_IMAGE.T
    T[..., :3, 3] = t_wc
    T[..., 3, 3] = 1.0
    return T


def marshal_observations(kpt_uv, depth, stereo=True):
    """Optimizer.cpp:176-196.  kpt_uv float32 [E,2], depth float32 [E] (NaN / <= 0 = no depth).
    Returns obs float64 [E,3] and kind uint8 [E] (0 stereo, 1 mono)."""
    kpt_uv = np.asarray(kpt_uv, dtype=np.float32)
    depth = np.asarray(depth, dtype=np.float32)
    base = np.float64(np.float32(BASELINE)) if stereo else 0.0  # getBaseLine() returns float
    d64 = depth.astype(np.float64)
    ok = np.isfinite(d64) & (d64 > 0.0) & (base > 0.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        disparity = (base * FX / d64).astype(np.float32)      # static_cast<float>(baseLine * K(0,0) / depth)
    ur = (kpt_uv[:, 0] - disparity).astype(np.float32)        # float - float
    obs = np.zeros((kpt_uv.shape[0], 3))
    obs[:, 0] = kpt_uv[:, 0]
    obs[:, 1] = kpt_uv[:, 1]
    obs[:, 2] = np.where(ok, ur.astype(np.float64), 0.0)
    kind = np.where(ok, 0, 1).astype(np.uint8)
    return obs, kind


def bf_value():
    return np.float64(np.float32(BASELINE)) * FX  # es->bf = baseLine * es->fx


def _trajectory(kind, P, rng):
    """Ground-truth robot poses T_world<-robot, [P,4,4]."""
    T = np.zeros((P, 4, 4))
    T[:, 3, 3] = 1.0
    if kind == "line":      # +x at 0.15 m / frame, yaw jitter N(0, 1 deg)
        yaw = rng.normal(0.0, np.deg2rad(1.0), P)
        T[:, :3, :3] = rot_z(yaw)
        T[:, 0, 3] = 0.15 * np.arange(P)
    elif kind == "loop":    # closed circle, 0.15 m / frame, heading along the tangent
        radius = 0.15 * P / (2 * np.pi)
        ang = 2 * np.pi * np.arange(P) / P
        yaw = ang + np.pi / 2 + rng.normal(0.0, np.deg2rad(1.0), P)
        T[:, :3, :3] = rot_z(yaw)
        T[:, 0, 3] = radius * np.cos(ang)
        T[:, 1, 3] = radius * np.sin(ang)
    elif kind == "orbit":   # circle of radius 8 m looking at the origin
        ang = 2 * np.pi * np.arange(P) / P
        yaw = ang + np.pi + rng.normal(0.0, np.deg2rad(1.0), P)
        T[:, :3, :3] = rot_z(yaw)
        T[:, 0, 3] = 8.0 * np.cos(ang)
        T[:, 1, 3] = 8.0 * np.sin(ang)
    else:
        raise ValueError(kind)
    return T


def make_window(n_poses=10, n_points=2000, views=10, *, seed=BASE_SEED, layout="all", trajectory="line",
                mono_frac=0.0, fixed_point_frac=0.0, outlier_frac=0.05, pixel_noise=0.7,
                pose_noise=(0.02, np.deg2rad(0.5)), point_noise=0.05, iterations=10, solver=0,
                trust_region=0, huber_delta=8.0, pixel_variance=1.5, first_id=1, root="second_newest",
                depth_range=(2.0, 10.0), shuffle_edges=False, links=None, odometry_variance=0.00005,
                link_noise=(0.003, np.deg2rad(0.05))):
    """Build one window.  Returns a dict of numpy arrays in C-ABI form plus the reference-form
    inputs under the keys ``ref_*``.

    links: None, or "chain": an odometry constraint (EdgePoseConstraint, Optimizer.cpp:116-150) between every pair of
    consecutive frames, measurement = ground-truth relative camera motion T_c1c2 plus noise.

    layout: "all" (every point seen by every frame; views ignored), "consecutive" (each point
    seen by `views` consecutive frames, wrapping on a loop trajectory), "random" (`views` random
    frames per point).
    """
    rng = np.random.default_rng(seed)
    P, L = int(n_poses), int(n_points)
    T_wr = _trajectory(trajectory, P, rng)
    tq_gt = robot_to_camera_state(T_wr)
    R_cw = R_from_quat(tq_gt[:, 3:7])
    t_cw = tq_gt[:, :3]

    # which frames see which point
    if layout == "all":
        d = P
        view = np.broadcast_to(np.arange(P), (L, P)).copy()
    elif layout == "consecutive":
        d = min(int(views), P)
        if trajectory == "loop":
            start = rng.integers(0, P, L)
            view = (start[:, None] + np.arange(d)[None, :]) % P
        else:
            start = rng.integers(0, P - d + 1, L)
            view = start[:, None] + np.arange(d)[None, :]
        view = np.sort(view, axis=1)
    elif layout == "random":
        d = min(int(views), P)
        view = np.argsort(rng.random((L, P)), axis=1)[:, :d]
        view = np.sort(view, axis=1)
    else:
        raise ValueError(layout)

    # ground-truth points
    if trajectory == "orbit":
        v = rng.normal(size=(L, 3))
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        pts = v * (2.0 * rng.random(L) ** (1 / 3))[:, None]
    else:
        if layout == "all":
            anchor = np.full(L, P // 2)
        elif trajectory == "loop" and layout == "consecutive":
            anchor = (start + d // 2) % P
        else:
            anchor = view[:, d // 2]
        z = rng.uniform(depth_range[0], depth_range[1], L)
        u = rng.uniform(0, WIDTH, L)
        vv = rng.uniform(0, HEIGHT, L)
        pc = np.stack([(u - CX) / FX * z, (vv - CY) / FY * z, z], axis=1)
        Ra = R_cw[anchor]
        pts = (np.swapaxes(Ra, 1, 2) @ (pc - t_cw[anchor])[:, :, None])[:, :, 0]

    # observations (edge list: grouped by point, ascending pose inside a point)
    e_point = np.repeat(np.arange(L, dtype=np.int32), d)
    e_pose = view.reshape(-1).astype(np.int32)
    E = e_point.shape[0]
    pcs = (R_cw[e_pose] @ pts[e_point][:, :, None])[:, :, 0] + t_cw[e_pose]
    zc = pcs[:, 2]
    u = FX * pcs[:, 0] / zc + CX + rng.normal(0, pixel_noise, E)
    v = FY * pcs[:, 1] / zc + CY + rng.normal(0, pixel_noise, E)
    disp = bf_value() / zc + rng.normal(0, pixel_noise, E)
    gross = rng.random(E) < outlier_frac
    sign = np.where(rng.random(E) < 0.5, -1.0, 1.0)
    u = u + np.where(gross, 20.0 * sign, 0.0)
    disp = np.maximum(disp, 0.05)
    depth = (bf_value() / disp).astype(np.float32)
    mono = rng.random(E) < mono_frac
    depth = np.where(mono, np.float32(np.nan), depth).astype(np.float32)
    kpt = np.stack([u, v], axis=1).astype(np.float32)
    obs, kind = marshal_observations(kpt, depth)

    # initial estimate
    T_init = T_wr.copy()
    T_init[:, :3, 3] += rng.normal(0, pose_noise[0], (P, 3))
    T_init[:, :3, :3] = small_rot(rng.normal(0, pose_noise[1], (P, 3))) @ T_wr[:, :3, :3]
    pts_init = pts + rng.normal(0, point_noise, (L, 3))
    point_fixed = (rng.random(L) < fixed_point_frac).astype(np.uint8)

    pose_id = np.arange(first_id, first_id + P, dtype=np.int64)
    if root == "second_newest":      # Estimator.cpp:252
        root_id = int(pose_id[-1]) - 1
    elif root == "first":
        root_id = int(pose_id[0])
    elif root is None:
        root_id = -1
    else:
        root_id = int(root)
    pose_fixed = (pose_id == root_id).astype(np.uint8)

    if shuffle_edges:
        perm = rng.permutation(E)
        e_point, e_pose, obs, kind, kpt, depth = e_point[perm], e_pose[perm], obs[perm], kind[perm], kpt[perm], depth[perm]

    link_kw = {}
    if links == "chain" and P > 1:
        # T_c1c2 = T_c1w * T_c2w^-1 from the ground truth, perturbed
        Tcw = np.zeros((P, 4, 4)); Tcw[:, 3, 3] = 1.0
        Tcw[:, :3, :3] = R_cw; Tcw[:, :3, 3] = t_cw
        lf = np.arange(0, P - 1, dtype=np.int32); lt = lf + 1
        rel = Tcw[lf] @ np.linalg.inv(Tcw[lt])
        rel[:, :3, 3] += rng.normal(0, link_noise[0], (P - 1, 3))
        rel[:, :3, :3] = small_rot(rng.normal(0, link_noise[1], (P - 1, 3))) @ rel[:, :3, :3]
        # what the reference is handed is the robot-frame motion T_r1r2; Optimizer.cpp:133 turns it into
        # T_c1c2 = T_rc^-1 * T_r1r2 * T_rc, which is the measurement of the C ABI
        T_rc = np.eye(4); T_rc[:3, :3] = R_ROBOT_FROM_IMAGE
        ref_T = T_rc @ rel @ np.linalg.inv(T_rc)
        meas = np.linalg.inv(T_rc) @ ref_T @ T_rc
        ltq = np.zeros((P - 1, 7))
        ltq[:, :3] = meas[:, :3, 3]
        ltq[:, 3:7] = quat_from_R(meas[:, :3, :3])
        link_kw = dict(n_links=P - 1, link_from=lf, link_to=lt, link_tq=np.ascontiguousarray(ltq),
                       odometry_variance=float(odometry_variance), ref_link_T=ref_T)
    elif links is not None:
        raise ValueError(links)

    return dict(
        **link_kw,
        n_poses=P, n_points=L, n_edges=E,
        pose_tq=np.ascontiguousarray(robot_to_camera_state(T_init)), pose_id=pose_id, pose_fixed=pose_fixed,
        point_xyz=np.ascontiguousarray(pts_init), point_id=np.arange(L, dtype=np.int64), point_fixed=point_fixed,
        edge_obs=np.ascontiguousarray(obs), edge_pose=np.ascontiguousarray(e_pose),
        edge_point=np.ascontiguousarray(e_point), edge_kind=np.ascontiguousarray(kind),
        fx=FX, fy=FY, cx=CX, cy=CY, bf=float(bf_value()),
        pixel_variance=float(pixel_variance), huber_delta=float(huber_delta),
        iterations=int(iterations), solver=int(solver), trust_region=int(trust_region), flags=0,
        ref_T_wr=T_init, ref_T_wr_gt=T_wr, ref_points_gt=pts, ref_kpt=kpt, ref_depth=depth, ref_root_id=root_id,
        seed=int(seed),
    )


# The five BASELINE.json configurations; seed = BASE_SEED + config index.
def config_c1(seed=BASE_SEED + 1, **kw):
    """C1: 10 key frames, 2 000 landmarks each seen by all 10 -> 20 000 stereo edges."""
    return make_window(10, 2000, layout="all", seed=seed, **kw)


def config_c2(seed=BASE_SEED + 2, **kw):
    """C2: 20 frames, 10 000 landmarks x 10 consecutive frames -> 100 000 edges, 30 % mono."""
    kw.setdefault("mono_frac", 0.3)
    return make_window(20, 10000, views=10, layout="consecutive", seed=seed, **kw)


def config_c3_windows(n_windows, seed=BASE_SEED + 3, **kw):
    """C3: independent C1 windows with different seeds."""
    return [make_window(10, 2000, layout="all", seed=seed + 1000 * (w + 1), **kw) for w in range(n_windows)]


def config_c4(n_poses=2000, n_points=500000, seed=BASE_SEED + 4, **kw):
    """C4: frames on a loop, each landmark seen by 10 consecutive frames (5 M edges at full size)."""
    return make_window(n_poses, n_points, views=10, layout="consecutive", trajectory="loop", seed=seed, **kw)


def config_c5(n_poses=200, n_points=200000, seed=BASE_SEED + 5, **kw):
    """C5: frames orbiting one scene, each landmark seen by 10 random frames (dense reduced system)."""
    return make_window(n_poses, n_points, views=10, layout="random", trajectory="orbit", seed=seed, **kw)


def algorithmic_bytes_per_trial(w):
    """64 E_stereo + 48 E_mono + 72 L + 112 P."""
    e_mono = int(np.count_nonzero(w["edge_kind"]))
    e_st = int(w["n_edges"]) - e_mono
    return 64 * e_st + 48 * e_mono + 72 * int(w["n_points"]) + 112 * int(w["n_poses"])

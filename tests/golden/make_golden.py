"""Golden vectors for the BA oracle.

The reference holds no golden vectors for its optimiser and cannot be built here, so
the oracle cannot be pinned against the reference itself.  What this script pins instead:

  * `dense_lm()` below — an INDEPENDENT numpy restatement of the same algorithm that never forms a Schur
    complement: it assembles the full (6F + 3NL) normal equations and solves them densely, with g2o's
    Levenberg-Marquardt rules and the two-pass protocol of Optimizer.cpp:261-318.  Agreement between it
    and oracle/ba_oracle.cpp (Schur + skyline Cholesky, C++) on the fixtures checks the Hessian blocks,
    the Schur elimination, the back-substitution and the LM control flow against a second derivation.
  * hand-computed known answers for one edge (identity pose, point on the optical axis).

Run `python tests/golden/make_golden.py` to regenerate tests/golden/*.npz (the committed fixtures were
produced by this script; tests/test_oracle.py re-runs both implementations against them).
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from visfs_b200 import synth  # noqa: E402


def quat_R(q):
    return synth.R_from_quat(np.asarray(q))


def edge_terms(w, pose, point, e_idx):
    """errors (n,3), J_point (n,3,3), J_pose (n,3,6) for the edges e_idx (OptimizeTypeDefine.h:121-187)."""
    ep, el = w["edge_pose"][e_idx], w["edge_point"][e_idx]
    R = quat_R(pose[ep, 3:7])
    pc = np.einsum("nij,nj->ni", R, point[el]) + pose[ep, :3]
    x, y, z = pc[:, 0], pc[:, 1], pc[:, 2]
    fx, fy, cx, cy, bf = w["fx"], w["fy"], w["cx"], w["cy"], w["bf"]
    u = x / z * fx + cx
    v = y / z * fy + cy
    ur = u - bf / z
    obs = w["edge_obs"][e_idx]
    mono = w["edge_kind"][e_idx].astype(bool)
    err = np.stack([obs[:, 0] - u, obs[:, 1] - v, np.where(mono, 0.0, obs[:, 2] - ur)], axis=1)
    z2 = z * z
    n = len(e_idx)
    Jl = np.zeros((n, 3, 3))
    for k in range(3):
        Jl[:, 0, k] = -fx * R[:, 0, k] / z + fx * x * R[:, 2, k] / z2
        Jl[:, 1, k] = -fy * R[:, 1, k] / z + fy * y * R[:, 2, k] / z2
        Jl[:, 2, k] = Jl[:, 0, k] - bf * R[:, 2, k] / z2
    Jp = np.zeros((n, 3, 6))
    Jp[:, 0, 0] = -fx / z
    Jp[:, 0, 2] = x / z2 * fx
    Jp[:, 0, 3] = x * y / z2 * fx
    Jp[:, 0, 4] = -(1 + x * x / z2) * fx
    Jp[:, 0, 5] = y / z * fx
    Jp[:, 1, 1] = -fy / z
    Jp[:, 1, 2] = y / z2 * fy
    Jp[:, 1, 3] = (1 + y * y / z2) * fy
    Jp[:, 1, 4] = -x * y / z2 * fy
    Jp[:, 1, 5] = -x / z * fy
    Jp[:, 2, 0] = Jp[:, 0, 0]
    Jp[:, 2, 2] = Jp[:, 0, 2] - bf / z2
    Jp[:, 2, 3] = Jp[:, 0, 3] - bf * y / z2
    Jp[:, 2, 4] = Jp[:, 0, 4] + bf * x / z2
    Jp[:, 2, 5] = Jp[:, 0, 5]
    Jl[mono, 2, :] = 0.0
    Jp[mono, 2, :] = 0.0
    return err, Jl, Jp


def robust(chi2, delta):
    if delta <= 0:
        return chi2.copy(), np.ones_like(chi2)
    rho, wgt = chi2.copy(), np.ones_like(chi2)
    out = chi2 > delta * delta
    s = np.sqrt(chi2[out])
    rho[out] = 2 * s * delta - delta * delta
    wgt[out] = delta / s
    return rho, wgt


def pose_oplus(tq, d):
    out = tq.copy()
    out[:3] += d[:3]
    ax, ay, az, aw = d[3] / 2, d[4] / 2, d[5] / 2, 1.0
    bx, by, bz, bw = tq[3], tq[4], tq[5], tq[6]
    q = np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                  aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz])
    out[3:7] = q / np.sqrt(np.sum(q * q))
    return out


def q_mul(a, b):   # Eigen (Hamilton) product, (x, y, z, w)
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz])


def q_inv(q):
    return np.array([-q[0], -q[1], -q[2], q[3]]) / float(np.sum(q * q))


def link_error(tq1, tq2, m):
    """EdgePoseConstraint::computeError (OptimizeTypeDefine.cpp:35-51), written from the source independently of the oracle."""
    q12 = q_mul(tq1[3:7], q_inv(tq2[3:7]))
    e_t = quat_R(q12) @ (-tq2[:3]) + tq1[:3] - m[:3]
    e_r = 2.0 * q_mul(q_inv(m[3:7]), q12)[:3]
    return np.concatenate([e_t, e_r])


def link_jacobians_numeric(tq1, tq2, m, h=1e-4):
    """d error / d (oplus increment) of both poses by 4th-order central differences through pose_oplus — NOT the reference's
    closed forms: agreement of the full LM with the oracle then also checks that those closed forms are the derivatives."""
    def col(which, a):
        def f(s):
            d = np.zeros(6); d[a] = s
            return link_error(pose_oplus(tq1, d), tq2, m) if which == 0 else link_error(tq1, pose_oplus(tq2, d), m)
        return (-f(2 * h) + 8 * f(h) - 8 * f(-h) + f(-2 * h)) / (12 * h)
    Ji = np.stack([col(0, a) for a in range(6)], axis=1)
    Jj = np.stack([col(1, a) for a in range(6)], axis=1)
    return Ji, Jj


def optimize_pass(w, pose, point, level, max_iter):
    """One g2o optimize(max_iter) with dense normal equations.  Returns the accepted state and statistics."""
    pfix, lfix = w["pose_fixed"].astype(bool), w["point_fixed"].astype(bool)
    ep, el = w["edge_pose"], w["edge_point"]
    active = (level == 0) & ~(pfix[ep] & lfix[el])
    e_idx = np.nonzero(active)[0]
    P, L = len(pose), len(point)
    pact, lact = np.zeros(P, bool), np.zeros(L, bool)
    pact[ep[e_idx]] = True
    lact[el[e_idx]] = True
    K = int(w.get("n_links", 0))
    links = [(int(w["link_from"][k]), int(w["link_to"][k]), np.asarray(w["link_tq"][k], float)) for k in range(K)]
    links = [(i, j, m) for (i, j, m) in links if not (pfix[i] and pfix[j])]          # e->allVerticesFixed()
    for i, j, _ in links:
        pact[i] = pact[j] = True
    om_link = 1.0 / float(w.get("odometry_variance", 1.0))
    phi = np.full(P, -1)
    free_p = np.nonzero(pact & ~pfix)[0]
    phi[free_p] = np.arange(len(free_p))
    F = len(free_p)
    lhi = np.full(L, -1)
    free_l = np.nonzero(lact & ~lfix)[0]
    lhi[free_l] = np.arange(len(free_l))
    NL = len(free_l)
    n = 6 * F + 3 * NL
    pv, delta = w["pixel_variance"], w["huber_delta"]

    def chi_of(ps, pt):
        err, _, _ = edge_terms(w, ps, pt, e_idx)
        rho, _ = robust(np.sum(err * err, axis=1) / pv, delta)
        chi_links = sum(om_link * float(np.sum(link_error(ps[i], ps[j], m) ** 2)) for i, j, m in links)   # no robust kernel
        return float(np.sum(rho)) + chi_links, err

    stats = dict(iterations=0, trials=0, stop=1, F=F, NL=NL, lam=0.0)
    chi_now, _ = chi_of(pose, point)
    stats["last_trial"] = chi_now
    if n == 0:
        stats["stop"] = 3
        stats["chi2"] = chi_now
        return pose, point, stats, active
    lam, ni = 0.0, 2.0
    for it in range(max_iter):
        cur_chi, err = chi_of(pose, point)
        _, Jl, Jp = edge_terms(w, pose, point, e_idx)
        _, wgt = robust(np.sum(err * err, axis=1) / pv, delta)
        H = np.zeros((n, n))
        b = np.zeros(n)
        for k, e in enumerate(e_idx):
            cols, blocks = [], []
            if phi[ep[e]] >= 0:
                cols.append(np.arange(6) + 6 * phi[ep[e]])
                blocks.append(Jp[k])
            if lhi[el[e]] >= 0:
                cols.append(np.arange(3) + 6 * F + 3 * lhi[el[e]])
                blocks.append(Jl[k])
            if not cols:
                continue
            J = np.concatenate(blocks, axis=1)
            c = np.concatenate(cols)
            wo = wgt[k] / pv
            H[np.ix_(c, c)] += wo * J.T @ J
            b[c] -= wo * J.T @ err[k]
        for i, j, m in links:
            e = link_error(pose[i], pose[j], m)
            Ji, Jj = link_jacobians_numeric(pose[i], pose[j], m)
            cols, blocks = [], []
            if phi[i] >= 0:
                cols.append(np.arange(6) + 6 * phi[i]); blocks.append(Ji)
            if phi[j] >= 0:
                cols.append(np.arange(6) + 6 * phi[j]); blocks.append(Jj)
            if cols:
                J = np.concatenate(blocks, axis=1)
                c = np.concatenate(cols)
                H[np.ix_(c, c)] += om_link * J.T @ J
                b[c] -= om_link * J.T @ e
        if w["trust_region"] == 1:
            try:
                x = np.linalg.solve(H, b)
            except np.linalg.LinAlgError:
                stats["stop"] = 4
                stats["iterations"] += 1
                stats["trials"] += 1
                break
            pose = pose.copy()
            point = point.copy()
            for i in free_p:
                pose[i] = pose_oplus(pose[i], x[6 * phi[i]: 6 * phi[i] + 6])
            point[free_l] += x[6 * F:].reshape(-1, 3)
            stats["iterations"] += 1
            stats["trials"] += 1
            continue
        if it == 0:
            lam = 1e-5 * float(np.max(np.abs(np.diag(H))))
            ni = 2.0
        rho_gain, qmax = 0.0, 0
        while True:
            ok = True
            try:
                Lc = np.linalg.cholesky(H + lam * np.eye(n))
                x = np.linalg.solve(Lc.T, np.linalg.solve(Lc, b))
            except np.linalg.LinAlgError:
                ok, x = False, np.zeros(n)
            tp, tl = pose.copy(), point.copy()
            for i in free_p:
                tp[i] = pose_oplus(pose[i], x[6 * phi[i]: 6 * phi[i] + 6])
            tl[free_l] += x[6 * F:].reshape(-1, 3)
            temp_chi, _ = chi_of(tp, tl)
            stats["last_trial"] = temp_chi
            stats["trials"] += 1
            if not ok:
                temp_chi = np.finfo(float).max
            rho_gain = (cur_chi - temp_chi) / (float(np.sum(x * (lam * x + b))) + 1e-3)
            if rho_gain > 0 and np.isfinite(temp_chi):
                alpha = min(1.0 - (2 * rho_gain - 1) ** 3, 2.0 / 3.0)
                lam *= max(1.0 / 3.0, alpha)
                ni = 2.0
                cur_chi = temp_chi
                pose, point = tp, tl
            else:
                lam *= ni
                ni *= 2
                if not np.isfinite(lam):
                    break
            qmax += 1
            if not (rho_gain < 0 and qmax < 10):
                break
        stats["iterations"] += 1
        if qmax == 10 or rho_gain == 0 or not np.isfinite(lam):
            stats["stop"] = 2
            break
    stats["lam"] = lam
    stats["chi2"], _ = chi_of(pose, point)
    return pose, point, stats, active


def dense_lm(w):
    """Optimizer.cpp:261-318 on top of optimize_pass()."""
    pose, point = w["pose_tq"].copy(), w["point_xyz"].copy()
    level = np.zeros(w["n_edges"], np.uint8)
    half = w["iterations"] // 2
    pose, point, s1, active = optimize_pass(w, pose, point, level, half)
    out = dict(status=0, chi2_pass1=s1["chi2"], chi2_final=s1["chi2"], iterations_run=[s1["iterations"], 0],
               trials_run=[s1["trials"], 0], stop_reason=[s1["stop"], 0], n_outliers=0, lambda_final=[s1["lam"], 0.0])
    if not np.isfinite(s1["chi2"]) or s1["chi2"] > 1e12:
        out["status"] = 3
    elif w["huber_delta"] > 0:
        err, _, _ = edge_terms(w, pose, point, np.arange(w["n_edges"]))
        chi2 = np.sum(err * err, axis=1) / w["pixel_variance"]
        level[active & (chi2 > w["huber_delta"])] = 1
        out["n_outliers"] = int(level.sum())
        pose, point, s2, _ = optimize_pass(w, pose, point, level, half)
        out["chi2_final"] = s2["chi2"]
        out["iterations_run"][1], out["trials_run"][1], out["stop_reason"][1] = s2["iterations"], s2["trials"], s2["stop"]
        out["lambda_final"][1] = s2["lam"]
        if s2["last_trial"] > 1e12:
            out["status"] = 4
    out.update(pose_tq=pose, point_xyz=point, edge_level=level)
    return out


LINK_KEYS = ("n_links", "link_from", "link_to", "link_tq", "odometry_variance")
INPUT_KEYS = ("n_poses", "n_points", "n_edges", "pose_tq", "pose_id", "pose_fixed", "point_xyz", "point_id", "point_fixed",
              "edge_obs", "edge_pose", "edge_point", "edge_kind", "fx", "fy", "cx", "cy", "bf", "pixel_variance",
              "huber_delta", "iterations", "solver", "trust_region", "flags")

CASES = {
    "stereo_4x40": dict(n_poses=4, n_points=40, layout="all", seed=9001),
    "mixed_5x60_fixed_points": dict(n_poses=5, n_points=60, layout="consecutive", views=3, seed=9002, mono_frac=0.3,
                                    fixed_point_frac=0.25),
    "rejecting_4x50": dict(n_poses=4, n_points=50, layout="all", seed=9102, pose_noise=(0.3, np.deg2rad(6.0)), point_noise=0.5,
                           iterations=12, depth_range=(1.0, 6.0)),
    "gauss_newton_4x40": dict(n_poses=4, n_points=40, layout="all", seed=9004, trust_region=1, outlier_frac=0.0, point_noise=0.02),
    "no_fixed_pose_3x30": dict(n_poses=3, n_points=30, layout="all", seed=9005, root=None),
    "odometry_links_5x50": dict(n_poses=5, n_points=50, layout="consecutive", views=3, seed=9006, links="chain", mono_frac=0.2),
}


def main(only=None):
    for name, kw in CASES.items():
        if only and name not in only:
            continue
        w = synth.make_window(**kw)
        out = dense_lm(w)
        path = os.path.join(HERE, name + ".npz")
        keys = INPUT_KEYS + (LINK_KEYS if w.get("n_links", 0) else ())
        np.savez_compressed(path, **{"in_" + k: np.asarray(w[k]) for k in keys},
                            **{"out_" + k: np.asarray(v) for k, v in out.items()})
        print(name, "chi2", out["chi2_pass1"], "->", out["chi2_final"], "iters", out["iterations_run"], "trials", out["trials_run"],
              "outliers", out["n_outliers"], "bytes", os.path.getsize(path))


if __name__ == "__main__":
    main(sys.argv[1:])

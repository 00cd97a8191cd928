// ba_oracle.cpp — CPU restatement of VISFS's local bundle adjustment (g2o branch).
//
// TEST INFRASTRUCTURE ONLY.  Nothing under visfs_b200/ may link, import or call this file.
// It is used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// leg as the checker and as the timed CPU baseline ("g2o-equivalent port").
//
// PARITY, WHAT IS PINNED AND WHAT IS NOT.  The reference ships no golden vectors, known-answer tests or fixtures
// for this path, and g2o / Eigen are absent from the build container, so Optimizer.cpp itself
// cannot be built here.  Since round 2 the reference's OWN per-edge / per-vertex arithmetic IS executed here: its
// unmodified OptimizeTypeDefine.{h,cpp} + Math.h compile against a minimal Eigen / g2o-base-class stand-in
// (oracle/ref_stub, oracle/ref_shim.cpp -> oracle/_ref/libvisfs_ref.so), and tests/test_ref_pin.py holds this file to
// it: rows (1)-(3) below and EdgePoseConstraint are REFERENCE-PINNED.  Rows (4)-(5) — g2o's optimiser semantics and the
// driver protocol of Optimizer.cpp — remain a restatement: PARITY UNPINNED for those.  It follows
//   (1) corelib/include/Optimizer/g2o/OptimizeTypeDefine.h:16-191 — CameraPose, VertexPose, EdgeStereo
//   (2) corelib/src/Optimizer/g2o/OptimizeTypeDefine.cpp:7-14     — CameraPose::update
//   (3) utilite/include/Math.h:277-287                            — deltaQ
//   (4) corelib/src/Optimizer/Optimizer.cpp:72-364                — graph build, two passes, guards
//   (5) upstream g2o (RainerKuemmerle/g2o, unpinned in the reference, API level ~ tag 20201223_git):
//       SparseOptimizer::{initializeOptimization,buildIndexMapping,optimize,activeRobustChi2},
//       BlockSolver<6,3>::{buildStructure,buildSystem,setLambda,solve,restoreDiagonal},
//       BaseBinaryEdge::constructQuadraticForm, RobustKernelHuber::robustify,
//       OptimizationAlgorithmLevenberg::{solve,computeLambdaInit,computeScale},
//       OptimizationAlgorithmGaussNewton::solve, LinearSolverPCG::solve — restated from their
//       published source; every such assumption is listed in SURVEY.md Appendix C.
// Each function below names the lines it follows.
//
// Built twice from this one source (oracle/Makefile): liboracle.so (sequential, g2o summation
// order) and liboracle_omp.so (-fopenmp: landmark-parallel, used only for timing).

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <map>
#include <set>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/visfs_ba.h"

namespace {

// Eigen::Quaternion::toRotationMatrix (called from CameraPose::map, OptimizeTypeDefine.h:46)
inline void quatToR(const double *q /* x y z w */, double *R /* row-major 3x3 */) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz;         R[2] = txz + twy;
    R[3] = txy + twz;         R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;         R[7] = tyz + twx;         R[8] = 1.0 - (txx + tyy);
}

struct Intrinsics { double fx, fy, cx, cy, bf; };

// CameraPose::map (OptimizeTypeDefine.h:45-47)
inline void mapPoint(const double *tq, const double *R, const double *pw, double *pc) {
    for (int i = 0; i < 3; ++i)
        pc[i] = (R[3 * i] * pw[0] + R[3 * i + 1] * pw[1] + R[3 * i + 2] * pw[2]) + tq[i];
}

// EdgeStereo::computeError + project (OptimizeTypeDefine.h:121-126, 180-187).
// Mono edges (dead code in the reference, Optimizer.cpp:197-208; defined here per SURVEY.md
// Appendix A): rows 0-1 only, row 2 forced to zero.
inline void edgeError(const double *tq, const double *R, const double *pw, const double *obs, int kind,
                      const Intrinsics &K, double *e) {
    double pc[3];
    mapPoint(tq, R, pw, pc);
    const double invZ = 1.0 / pc[2];
    const double u = pc[0] * invZ * K.fx + K.cx;
    const double v = pc[1] * invZ * K.fy + K.cy;
    const double ur = u - K.bf * invZ;
    e[0] = obs[0] - u;
    e[1] = obs[1] - v;
    e[2] = (kind == VISFS_BA_EDGE_MONO) ? 0.0 : obs[2] - ur;
}

// EdgeStereo::linearizeOplus (OptimizeTypeDefine.h:134-178); operation order kept.
inline void edgeJacobians(const double *tq, const double *R, const double *pw, int kind, const Intrinsics &K,
                          double *Jl /* 3x3 row-major */, double *Jp /* 3x6 row-major */) {
    double pc[3];
    mapPoint(tq, R, pw, pc);
    const double x = pc[0], y = pc[1], z = pc[2], z_2 = z * z;
    const double fx = K.fx, fy = K.fy, bf = K.bf;
    for (int k = 0; k < 3; ++k) {
        Jl[0 * 3 + k] = -fx * R[0 * 3 + k] / z + fx * x * R[2 * 3 + k] / z_2;
        Jl[1 * 3 + k] = -fy * R[1 * 3 + k] / z + fy * y * R[2 * 3 + k] / z_2;
        Jl[2 * 3 + k] = Jl[0 * 3 + k] - bf * R[2 * 3 + k] / z_2;
    }
    Jp[0] = -1. / z * fx;  Jp[1] = 0.;  Jp[2] = x / z_2 * fx;  Jp[3] = x * y / z_2 * fx;
    Jp[4] = -(1. + (x * x / z_2)) * fx;  Jp[5] = y / z * fx;
    Jp[6] = 0.;  Jp[7] = -1. / z * fy;  Jp[8] = y / z_2 * fy;  Jp[9] = (1. + y * y / z_2) * fy;
    Jp[10] = -x * y / z_2 * fy;  Jp[11] = -x / z * fy;
    Jp[12] = Jp[0];  Jp[13] = 0.;  Jp[14] = Jp[2] - bf / z_2;  Jp[15] = Jp[3] - bf * y / z_2;
    Jp[16] = Jp[4] + bf * x / z_2;  Jp[17] = Jp[5];
    if (kind == VISFS_BA_EDGE_MONO) {
        for (int k = 0; k < 3; ++k) Jl[6 + k] = 0.0;
        for (int k = 0; k < 6; ++k) Jp[12 + k] = 0.0;
    }
}

// g2o RobustKernelHuber::robustify; delta <= 0 means "no kernel" (Optimizer.cpp:212)
inline void huber(double e2, double delta, double *rho0, double *rho1) {
    if (delta <= 0.0) { *rho0 = e2; *rho1 = 1.0; return; }
    const double dsqr = delta * delta;
    if (e2 <= dsqr) { *rho0 = e2; *rho1 = 1.0; }
    else {
        const double sqrte = std::sqrt(e2);
        *rho0 = 2 * sqrte * delta - dsqr;
        *rho1 = delta / sqrte;
    }
}

// CameraPose::update (OptimizeTypeDefine.cpp:7-14) with deltaQ (Math.h:277-287):
// t += d[0:3]; dq = (w=1, xyz = d[3:6]/2); q = dq * q; q.normalize()
inline void poseOplus(double *tq, const double *d) {
    tq[0] += d[0]; tq[1] += d[1]; tq[2] += d[2];
    const double ax = d[3] / 2.0, ay = d[4] / 2.0, az = d[5] / 2.0, aw = 1.0;
    const double bx = tq[3], by = tq[4], bz = tq[5], bw = tq[6];
    // Eigen quaternion product a*b
    const double w = aw * bw - ax * bx - ay * by - az * bz;
    const double x = aw * bx + ax * bw + ay * bz - az * by;
    const double y = aw * by + ay * bw + az * bx - ax * bz;
    const double z = aw * bz + az * bw + ax * by - ay * bx;
    const double n = std::sqrt(x * x + y * y + z * z + w * w);
    tq[3] = x / n; tq[4] = y / n; tq[5] = z / n; tq[6] = w / n;
}

inline bool inv3(const double *A, double *Ai) {  // Eigen 3x3 inverse (cofactor / determinant)
    const double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
    const double det = A[0] * c00 + A[1] * c01 + A[2] * c02;
    const double id = 1.0 / det;
    Ai[0] = c00 * id; Ai[1] = (A[2] * A[7] - A[1] * A[8]) * id; Ai[2] = (A[1] * A[5] - A[2] * A[4]) * id;
    Ai[3] = c01 * id; Ai[4] = (A[0] * A[8] - A[2] * A[6]) * id; Ai[5] = (A[2] * A[3] - A[0] * A[5]) * id;
    Ai[6] = c02 * id; Ai[7] = (A[1] * A[6] - A[0] * A[7]) * id; Ai[8] = (A[0] * A[4] - A[1] * A[3]) * id;
    return std::isfinite(id);
}

// ---- EdgePoseConstraint (OptimizeTypeDefine.h:193-225, OptimizeTypeDefine.cpp:35-88): odometry constraint between two
// camera poses.  Quaternions are (x, y, z, w); products are Eigen's (Hamilton).
inline void qmul(const double *a, const double *b, double *o) {
    const double ax = a[0], ay = a[1], az = a[2], aw = a[3], bx = b[0], by = b[1], bz = b[2], bw = b[3];
    o[3] = aw * bw - ax * bx - ay * by - az * bz;
    o[0] = aw * bx + ax * bw + ay * bz - az * by;
    o[1] = aw * by + ay * bw + az * bx - ax * bz;
    o[2] = aw * bz + az * bw + ax * by - ay * bx;
}
inline void qinv(const double *q, double *o) {   // Eigen::Quaternion::inverse: conjugate / squaredNorm
    const double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    o[0] = -q[0] / n2; o[1] = -q[1] / n2; o[2] = -q[2] / n2; o[3] = q[3] / n2;
}
inline void matvec3(const double *R, const double *v, double *o) {
    for (int i = 0; i < 3; ++i) o[i] = R[3 * i] * v[0] + R[3 * i + 1] * v[1] + R[3 * i + 2] * v[2];
}
inline void skew3(const double *v, double *S) {   // Math.h:293-300
    S[0] = 0; S[1] = -v[2]; S[2] = v[1]; S[3] = v[2]; S[4] = 0; S[5] = -v[0]; S[6] = -v[1]; S[7] = v[0]; S[8] = 0;
}
inline void qpositify(const double *q, double *o) {   // Math.h:308-316
    double s = (q[3] < 0) ? -1.0 : 1.0;
    double t[4] = {s * q[0], s * q[1], s * q[2], s * q[3]};
    const double n = std::sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2] + t[3] * t[3]);
    for (int i = 0; i < 4; ++i) o[i] = t[i] / n;
}
// bottom-right 3x3 of QuaternionLeft(q) (Math.h:324-331): w I + skew(v), and of QuaternionRight(q) (:339-346): w I - skew(v);
// the full 4x4 are [[w, -v^T], [v, w I +- skew(v)]]
inline void qleft(const double *q, double *M /* 4x4 row-major */) {
    double p[4]; qpositify(q, p);
    const double x = p[0], y = p[1], z = p[2], w = p[3];
    const double m[16] = {w, -x, -y, -z,   x, w, -z, y,   y, z, w, -x,   z, -y, x, w};
    std::memcpy(M, m, sizeof m);
}
inline void qright(const double *q, double *M) {
    double p[4]; qpositify(q, p);
    const double x = p[0], y = p[1], z = p[2], w = p[3];
    const double m[16] = {w, -x, -y, -z,   x, w, z, -y,   y, -z, w, x,   z, y, -x, w};
    std::memcpy(M, m, sizeof m);
}

// EdgePoseConstraint::computeError (OptimizeTypeDefine.cpp:35-51): tq1 / tq2 = (t, q) of vertex 0 / 1, m = measurement
inline void linkError(const double *tq1, const double *tq2, const double *m, double *e /* 6 */) {
    double q2i[4], q12[4], R12[9], nP2[3] = {-tq2[0], -tq2[1], -tq2[2]}, rp[3];
    qinv(tq2 + 3, q2i);
    qmul(tq1 + 3, q2i, q12);
    quatToR(q12, R12);
    matvec3(R12, nP2, rp);
    for (int i = 0; i < 3; ++i) e[i] = rp[i] + tq1[i] - m[i];
    double mi[4], t[4];
    qinv(m + 3, mi);
    qmul(mi, q12, t);
    e[3] = 2 * t[0]; e[4] = 2 * t[1]; e[5] = 2 * t[2];
}

// EdgePoseConstraint::linearizeOplus, the live "Left update" branch (OptimizeTypeDefine.cpp:53-72): 6x6 row-major each
inline void linkJacobians(const double *tq1, const double *tq2, const double *m, double *Ji, double *Jj) {
    std::fill(Ji, Ji + 36, 0.0); std::fill(Jj, Jj + 36, 0.0);
    double q1i[4], q2i[4], q12[4], R1[9], R2i[9], R12[9], nP2[3] = {-tq2[0], -tq2[1], -tq2[2]};
    qinv(tq1 + 3, q1i); qinv(tq2 + 3, q2i);
    qmul(tq1 + 3, q2i, q12);
    quatToR(tq1 + 3, R1); quatToR(q2i, R2i); quatToR(q12, R12);
    // Xi
    for (int i = 0; i < 3; ++i) Ji[6 * i + i] = 1.0;
    double t1[3], t2[3], S[9];
    matvec3(R2i, nP2, t1); matvec3(R1, t1, t2);          // sQ1 * (sQ2.inverse() * (-sP2))
    skew3(t2, S);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Ji[6 * i + 3 + j] = -S[3 * i + j];
    double q21[4], QL[16], QR[16];
    qmul(tq2 + 3, q1i, q21);                              // sQ2 * sQ1.inverse()
    qleft(q21, QL); qright(m + 3, QR);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double acc = 0;
            for (int k = 0; k < 4; ++k) acc += QL[4 * (i + 1) + k] * QR[4 * k + (j + 1)];
            Ji[6 * (3 + i) + 3 + j] = acc;
        }
    // Xj
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Jj[6 * i + j] = -R12[3 * i + j];
    double Sn[9], R1R2i[9];
    skew3(nP2, Sn);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        double acc = 0;
        for (int k = 0; k < 3; ++k) acc += R1[3 * i + k] * R2i[3 * k + j];
        R1R2i[3 * i + j] = acc;
    }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        double acc = 0;
        for (int k = 0; k < 3; ++k) acc += R1R2i[3 * i + k] * Sn[3 * k + j];
        Jj[6 * i + 3 + j] = acc;
    }
    double mi[4], t[4], QL2[16];
    qinv(m + 3, mi); qmul(mi, q12, t);                    // mQ12.inverse() * sQ1 * sQ2.inverse()
    qleft(t, QL2);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Jj[6 * (3 + i) + 3 + j] = -QL2[4 * (i + 1) + (j + 1)];
}

struct PassStats {
    int iterations = 0, trials = 0, stop = VISFS_BA_STOP_NOT_RUN, F = 0, NL = 0;
    double lambda = 0, chi2 = 0, chi2_last_trial = 0;
    std::vector<int> trials_per_iteration;
};

struct Oracle {
    int P = 0, L = 0, E = 0;
    std::vector<double> pose, point, obs;
    std::vector<uint8_t> pfix, lfix, kind, level;
    std::vector<int> ep, el;
    Intrinsics K{};
    double pv = 1.5, delta = 8.0;
    int solver = 0, trust = 0;
    int threads = 1;
    // odometry links (EdgePoseConstraint, Optimizer.cpp:116-150): inserted into the graph before the visual edges
    int NK = 0;
    std::vector<int> lkFrom, lkTo;
    std::vector<double> lkM;            // 7 per link
    std::vector<uint8_t> lkAct;
    std::vector<double> lkErr;          // 6 per link
    double ov = 0.00005;
    std::map<std::pair<int, int>, std::vector<double>> HppOff;   // pose-pose blocks (hi < hj) -> 6x6 row-major

    // structure
    std::vector<uint8_t> eact;
    std::vector<int> phidx, lhidx, activeEdges;
    std::vector<int> lmFirst;  // CSR by landmark over activeEdges (sorted by point then insertion)
    int F = 0, NL = 0;
    std::vector<int> hplRow, hplCol;
    std::vector<std::pair<int, int>> spat;  // (col, row) sorted
    std::vector<int> skyFirst;              // scalar skyline: first column of each row of L
    std::vector<size_t> skyOff;
    // landmark -> its edges (all levels) for the schur pattern
    std::vector<std::vector<int>> lmEdges;       // active edges of each landmark (insertion order)
    std::vector<int> hidx2point;

    // system
    std::vector<double> R;     // 9 per pose
    std::vector<double> err;   // 3 per edge
    std::vector<double> Hpp, Hll, Hpl, b, x, S, bs, Dinv, coeff;
    std::vector<double> diagBackupP, diagBackupL;

    // CameraPose::normalizeRotation (OptimizeTypeDefine.h:36-41), run by the CameraPose constructors of Optimizer.cpp:109;
    // g2o::SE3Quat::normalizeRotation does the same to an odometry measurement (Optimizer.cpp:140): w >= 0, unit norm.
    static void normalizeRotation(double *q /* x y z w */) {
        if (q[3] < 0) for (int i = 0; i < 4; ++i) q[i] *= -1;
        const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
        for (int i = 0; i < 4; ++i) q[i] /= n;
    }

    void load(const visfs_ba_problem *p) {
        P = p->n_poses; L = p->n_points; E = p->n_edges;
        pose.assign(p->pose_tq, p->pose_tq + 7 * (size_t)P);
        for (int i = 0; i < P; ++i) normalizeRotation(&pose[7 * (size_t)i + 3]);
        point.assign(p->point_xyz, p->point_xyz + 3 * (size_t)L);
        if (p->edge_obs_f32) { obs.resize(3 * (size_t)E); for (size_t i = 0; i < obs.size(); ++i) obs[i] = (double)p->edge_obs_f32[i]; }
        else obs.assign(p->edge_obs, p->edge_obs + 3 * (size_t)E);
        pfix.assign(P, 0); lfix.assign(L, 0); kind.assign(E, 0); level.assign(E, 0);
        if (p->pose_fixed) pfix.assign(p->pose_fixed, p->pose_fixed + P);
        if (p->point_fixed) lfix.assign(p->point_fixed, p->point_fixed + L);
        if (p->edge_kind) kind.assign(p->edge_kind, p->edge_kind + E);
        ep.assign(p->edge_pose, p->edge_pose + E);
        el.assign(p->edge_point, p->edge_point + E);
        K = Intrinsics{p->fx, p->fy, p->cx, p->cy, p->bf};
        pv = p->pixel_variance; delta = p->huber_delta; solver = p->solver; trust = p->trust_region;
        NK = p->n_links;
        if (NK > 0) {
            lkFrom.assign(p->link_from, p->link_from + NK); lkTo.assign(p->link_to, p->link_to + NK);
            lkM.assign(p->link_tq, p->link_tq + 7 * (size_t)NK);
            for (int k = 0; k < NK; ++k) normalizeRotation(&lkM[7 * (size_t)k + 3]);
            ov = p->odometry_variance;
        }
    }

    void refreshR() {
        R.resize(9 * (size_t)P);
        for (int i = 0; i < P; ++i) quatToR(&pose[7 * i + 3], &R[9 * i]);
    }

    // SparseOptimizer::initializeOptimization(level) + buildIndexMapping, then
    // BlockSolver::buildStructure (patterns only).
    void buildStructure() {
        eact.assign(E, 0);
        std::vector<uint8_t> pact(P, 0), lact(L, 0);
        activeEdges.clear();
        for (int e = 0; e < E; ++e) {
            if (level[e] != 0) continue;
            if (pfix[ep[e]] && lfix[el[e]]) continue;  // e->allVerticesFixed()
            eact[e] = 1; pact[ep[e]] = 1; lact[el[e]] = 1;
            activeEdges.push_back(e);
        }
        lkAct.assign(NK, 0);
        for (int k = 0; k < NK; ++k) {
            if (pfix[lkFrom[k]] && pfix[lkTo[k]]) continue;   // e->allVerticesFixed()
            lkAct[k] = 1; pact[lkFrom[k]] = 1; pact[lkTo[k]] = 1;
        }
        phidx.assign(P, -1); lhidx.assign(L, -1);
        F = 0;
        for (int i = 0; i < P; ++i) if (pact[i] && !pfix[i]) phidx[i] = F++;
        NL = 0;
        for (int l = 0; l < L; ++l) if (lact[l] && !lfix[l]) lhidx[l] = F + NL++;
        hidx2point.assign(NL, -1);
        for (int l = 0; l < L; ++l) if (lhidx[l] >= 0) hidx2point[lhidx[l] - F] = l;

        hplRow.assign(E, -1); hplCol.assign(E, -1);
        lmEdges.assign(L, {});
        for (int e : activeEdges) {
            lmEdges[el[e]].push_back(e);
            if (phidx[ep[e]] >= 0 && lhidx[el[e]] >= 0) { hplRow[e] = phidx[ep[e]]; hplCol[e] = lhidx[el[e]] - F; }
        }
        // Schur pattern: diagonal for every free pose, plus for each free active landmark all pose
        // pairs (i1 <= i2) reachable through ANY of its edges (v->edges(): includes level != 0).
        std::set<std::pair<int, int>> pat;
        for (int i = 0; i < F; ++i) pat.insert({i, i});
        for (int k = 0; k < NK; ++k) {   // BlockSolver::buildStructure: H_pp off-diagonal block of a pose-pose edge
            const int hi = phidx[lkFrom[k]], hj = phidx[lkTo[k]];
            if (lkAct[k] && hi >= 0 && hj >= 0) pat.insert({std::max(hi, hj), std::min(hi, hj)});
        }
        std::vector<std::vector<int>> allEdges(L);
        for (int e = 0; e < E; ++e) allEdges[el[e]].push_back(e);
        for (int l = 0; l < L; ++l) {
            if (lhidx[l] < 0) continue;
            for (int e1 : allEdges[l]) {
                const int i1 = phidx[ep[e1]];
                if (i1 < 0) continue;
                for (int e2 : allEdges[l]) {
                    const int i2 = phidx[ep[e2]];
                    if (i2 < 0) continue;
                    if (i1 <= i2) pat.insert({i2, i1});  // stored (col,row)
                }
            }
        }
        spat.assign(pat.begin(), pat.end());
        // scalar skyline of the lower triangle (row r of L starts at skyFirst[r])
        const int n = 6 * F;
        skyFirst.assign(n, 0);
        for (int r = 0; r < n; ++r) skyFirst[r] = r - (r % 6);
        for (auto &cr : spat) {
            const int col = cr.first, row = cr.second;  // row <= col; lower-tri row block = col
            for (int a = 0; a < 6; ++a) skyFirst[6 * col + a] = std::min(skyFirst[6 * col + a], 6 * row);
        }
        skyOff.assign(n + 1, 0);
        for (int r = 0; r < n; ++r) skyOff[r + 1] = skyOff[r] + (size_t)(r - skyFirst[r] + 1);
    }

    // SparseOptimizer::computeActiveErrors + activeRobustChi2
    double computeErrorsAndChi2() {
        refreshR();
        err.resize(3 * (size_t)E);
        const int na = (int)activeEdges.size();
        std::vector<double> rho(na);
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(threads)
#endif
        for (int k = 0; k < na; ++k) {
            const int e = activeEdges[k];
            edgeError(&pose[7 * ep[e]], &R[9 * ep[e]], &point[3 * el[e]], &obs[3 * e], kind[e], K, &err[3 * e]);
            const double *r = &err[3 * e];
            const double c = (r[0] * r[0] + r[1] * r[1] + r[2] * r[2]) / pv;
            double r0, r1;
            huber(c, delta, &r0, &r1);
            rho[k] = r0;
        }
        double chi = 0.0;
        lkErr.assign(6 * (size_t)NK, 0.0);
        for (int k = 0; k < NK; ++k) {   // no robust kernel on these edges: rho = chi2 = e' (I / ov) e
            if (!lkAct[k]) continue;
            double *e = &lkErr[6 * (size_t)k];
            linkError(&pose[7 * lkFrom[k]], &pose[7 * lkTo[k]], &lkM[7 * (size_t)k], e);
            double c = 0;
            for (int i = 0; i < 6; ++i) c += e[i] * e[i];
            chi += c / ov;
        }
        for (int k = 0; k < na; ++k) chi += rho[k];
        return chi;
    }

    double edgeChi2(int e) const {
        const double *r = &err[3 * e];
        return (r[0] * r[0] + r[1] * r[1] + r[2] * r[2]) / pv;
    }

    // BlockSolver::buildSystem: linearizeOplus + constructQuadraticForm per active edge.
    void buildSystem() {
        const int n = 6 * F + 3 * NL;
        Hpp.assign(36 * (size_t)F, 0.0); Hll.assign(9 * (size_t)NL, 0.0);
        Hpl.assign(18 * (size_t)E, 0.0); b.assign(n, 0.0);
        HppOff.clear();
        for (int k = 0; k < NK; ++k) {   // BaseBinaryEdge::constructQuadraticForm with Omega = I / ov, no robust kernel
            if (!lkAct[k]) continue;
            double Ji[36], Jj[36];
            linkJacobians(&pose[7 * lkFrom[k]], &pose[7 * lkTo[k]], &lkM[7 * (size_t)k], Ji, Jj);
            const double *e = &lkErr[6 * (size_t)k];
            const double om = 1.0 / ov;
            const int hi = phidx[lkFrom[k]], hj = phidx[lkTo[k]];
            auto addDiag = [&](int h, const double *J) {
                for (int a = 0; a < 6; ++a) {
                    for (int c = 0; c < 6; ++c) {
                        double sacc = 0;
                        for (int r = 0; r < 6; ++r) sacc += J[6 * r + a] * om * J[6 * r + c];
                        Hpp[36 * (size_t)h + 6 * a + c] += sacc;
                    }
                    double sb = 0;
                    for (int r = 0; r < 6; ++r) sb += J[6 * r + a] * om * e[r];
                    b[6 * h + a] -= sb;
                }
            };
            if (hi >= 0) addDiag(hi, Ji);
            if (hj >= 0) addDiag(hj, Jj);
            if (hi >= 0 && hj >= 0) {
                // block (row = smaller hessian index): Ji' Omega Jj, transposed when the "to" pose comes first
                const bool tr = hj < hi;
                auto &blk = HppOff[{std::min(hi, hj), std::max(hi, hj)}];
                if (blk.empty()) blk.assign(36, 0.0);
                for (int a = 0; a < 6; ++a)
                    for (int c = 0; c < 6; ++c) {
                        double sacc = 0;
                        for (int r = 0; r < 6; ++r) sacc += Ji[6 * r + a] * om * Jj[6 * r + c];
                        if (!tr) blk[6 * a + c] += sacc; else blk[6 * c + a] += sacc;
                    }
            }
        }
        auto edgeWork = [&](int e, double *HppT, double *bT) {
            const int pi = ep[e], li = el[e];
            double Jl[9], Jp[18];
            edgeJacobians(&pose[7 * pi], &R[9 * pi], &point[3 * li], kind[e], K, Jl, Jp);
            const double *r = &err[3 * e];
            const double c = edgeChi2(e);
            double rho0, w;
            huber(c, delta, &rho0, &w);
            const double wo = w * (1.0 / pv);  // rho' * Omega (Omega = I/pv)
            const int hi = phidx[pi], hl = lhidx[li];
            if (hl >= 0) {
                double *A = &Hll[9 * (size_t)(hl - F)];
                double *bl = &b[6 * F + 3 * (hl - F)];
                for (int a = 0; a < 3; ++a) {
                    for (int c2 = 0; c2 < 3; ++c2) {
                        double s = 0;
                        for (int k = 0; k < 3; ++k) s += Jl[3 * k + a] * wo * Jl[3 * k + c2];
                        A[3 * a + c2] += s;
                    }
                    double s = 0;
                    for (int k = 0; k < 3; ++k) s += Jl[3 * k + a] * (-(1.0 / pv) * r[k] * w);
                    bl[a] += s;
                }
            }
            if (hi >= 0) {
                double *A = &HppT[36 * (size_t)hi];
                double *bp = &bT[6 * hi];
                for (int a = 0; a < 6; ++a) {
                    for (int c2 = 0; c2 < 6; ++c2) {
                        double s = 0;
                        for (int k = 0; k < 3; ++k) s += Jp[6 * k + a] * wo * Jp[6 * k + c2];
                        A[6 * a + c2] += s;
                    }
                    double s = 0;
                    for (int k = 0; k < 3; ++k) s += Jp[6 * k + a] * (-(1.0 / pv) * r[k] * w);
                    bp[a] += s;
                }
            }
            if (hi >= 0 && hl >= 0) {
                double *B = &Hpl[18 * (size_t)e];  // 6x3 row-major
                for (int a = 0; a < 6; ++a)
                    for (int c2 = 0; c2 < 3; ++c2) {
                        double s = 0;
                        for (int k = 0; k < 3; ++k) s += Jp[6 * k + a] * wo * Jl[3 * k + c2];
                        B[3 * a + c2] = s;
                    }
            }
        };
#ifdef _OPENMP
        if (threads > 1) {
            std::vector<std::vector<double>> HppT(threads, std::vector<double>(36 * (size_t)F, 0.0));
            std::vector<std::vector<double>> bT(threads, std::vector<double>(6 * (size_t)F, 0.0));
#pragma omp parallel for schedule(static) num_threads(threads)
            for (int l = 0; l < L; ++l) {
                const int t = omp_get_thread_num();
                for (int e : lmEdges[l]) edgeWork(e, HppT[t].data(), bT[t].data());
            }
            for (int t = 0; t < threads; ++t) {
                for (size_t i = 0; i < Hpp.size(); ++i) Hpp[i] += HppT[t][i];
                for (int i = 0; i < 6 * F; ++i) b[i] += bT[t][i];
            }
            return;
        }
#endif
        for (int e : activeEdges) edgeWork(e, Hpp.data(), b.data());
    }

    // OptimizationAlgorithmLevenberg::computeLambdaInit
    double lambdaInit() const {
        double m = 0;
        for (int i = 0; i < F; ++i) for (int j = 0; j < 6; ++j) m = std::max(std::fabs(Hpp[36 * (size_t)i + 7 * j]), m);
        for (int l = 0; l < NL; ++l) for (int j = 0; j < 3; ++j) m = std::max(std::fabs(Hll[9 * (size_t)l + 4 * j]), m);
        return 1e-5 * m;
    }

    inline double &Sat(std::vector<double> &A, int r, int c) { return A[skyOff[r] + (size_t)(c - skyFirst[r])]; }

    // skyline Cholesky (stands in for CSparse / Cholmod / Eigen LDLT — all exact direct solvers);
    // fails on a non-positive pivot like cs_chol.
    bool choleskySolve(std::vector<double> &A, const std::vector<double> &rhs, double *xs) {
        const int n = 6 * F;
        for (int i = 0; i < n; ++i) {
            for (int j = skyFirst[i]; j <= i; ++j) {
                double s = Sat(A, i, j);
                const int k0 = std::max(skyFirst[i], skyFirst[j]);
                const double *li = &A[skyOff[i] + (size_t)(k0 - skyFirst[i])];
                const double *lj = &A[skyOff[j] + (size_t)(k0 - skyFirst[j])];
                for (int k = 0; k < j - k0; ++k) s -= li[k] * lj[k];
                if (j < i) Sat(A, i, j) = s / Sat(A, j, j);
                else {
                    if (!(s > 0.0)) return false;
                    Sat(A, i, i) = std::sqrt(s);
                }
            }
        }
        std::vector<double> y(n);
        for (int i = 0; i < n; ++i) {
            double s = rhs[i];
            for (int k = skyFirst[i]; k < i; ++k) s -= Sat(A, i, k) * y[k];
            y[i] = s / Sat(A, i, i);
        }
        for (int i = n - 1; i >= 0; --i) {
            y[i] /= Sat(A, i, i);
            const double yi = y[i];
            for (int k = skyFirst[i]; k < i; ++k) y[k] -= Sat(A, i, k) * yi;
        }
        for (int i = 0; i < n; ++i) xs[i] = y[i];
        return true;
    }

    // g2o LinearSolverPCG::solve (block-Jacobi preconditioner, x0 = 0, tolerance 1e-6 on r'M^-1 r,
    // absolute-tolerance quirk: from the second call on d0 = 0 when the previous residual > tol).
    double pcgResidual = -1.0;
    bool pcgSolve(const std::vector<double> &A, const std::vector<double> &rhs, double *xs) {
        const int n = 6 * F;
        // dense symmetric mat-vec through the skyline (lower) storage
        auto mult = [&](const std::vector<double> &v, std::vector<double> &out) {
            std::fill(out.begin(), out.end(), 0.0);
            for (int r = 0; r < n; ++r) {
                const double *row = &A[skyOff[r]];
                const int f = skyFirst[r];
                double s = 0;
                for (int c = f; c < r; ++c) { s += row[c - f] * v[c]; out[c] += row[c - f] * v[r]; }
                out[r] += s + row[r - f] * v[r];
            }
        };
        std::vector<double> J(36 * (size_t)F);
        for (int i = 0; i < F; ++i) {
            double blk[36];
            for (int a = 0; a < 6; ++a) for (int c = 0; c < 6; ++c) {
                const int r = 6 * i + std::max(a, c), cc = 6 * i + std::min(a, c);
                blk[6 * a + c] = A[skyOff[r] + (size_t)(cc - skyFirst[r])];
            }
            // 6x6 inverse by Gauss-Jordan with partial pivoting
            double M[6][12];
            for (int a = 0; a < 6; ++a) for (int c = 0; c < 6; ++c) { M[a][c] = blk[6 * a + c]; M[a][6 + c] = (a == c); }
            for (int c = 0; c < 6; ++c) {
                int piv = c;
                for (int a = c + 1; a < 6; ++a) if (std::fabs(M[a][c]) > std::fabs(M[piv][c])) piv = a;
                if (piv != c) for (int k = 0; k < 12; ++k) std::swap(M[c][k], M[piv][k]);
                const double d = M[c][c];
                for (int k = 0; k < 12; ++k) M[c][k] /= d;
                for (int a = 0; a < 6; ++a) if (a != c) {
                    const double f = M[a][c];
                    if (f != 0) for (int k = 0; k < 12; ++k) M[a][k] -= f * M[c][k];
                }
            }
            for (int a = 0; a < 6; ++a) for (int c = 0; c < 6; ++c) J[36 * (size_t)i + 6 * a + c] = M[a][6 + c];
        }
        auto multDiag = [&](const std::vector<double> &v, std::vector<double> &out) {
            for (int i = 0; i < F; ++i)
                for (int a = 0; a < 6; ++a) {
                    double s = 0;
                    for (int c = 0; c < 6; ++c) s += J[36 * (size_t)i + 6 * a + c] * v[6 * i + c];
                    out[6 * i + a] = s;
                }
        };
        auto dot = [&](const std::vector<double> &u, const std::vector<double> &v) {
            double s = 0; for (int i = 0; i < n; ++i) s += u[i] * v[i]; return s;
        };
        std::vector<double> xv(n, 0.0), r(rhs.begin(), rhs.begin() + n), d(n, 0.0), q(n, 0.0), s(n, 0.0);
        multDiag(r, d);
        double dn = dot(r, d);
        double d0 = 1e-6 * dn;
        if (pcgResidual > 0.0 && pcgResidual > 1e-6) d0 = 0;
        const int maxIter = n;
        for (int it = 0; it < maxIter; ++it) {
            if (dn <= d0) break;
            mult(d, q);
            const double a = dn / dot(d, q);
            for (int i = 0; i < n; ++i) { xv[i] += a * d[i]; r[i] -= a * q[i]; }
            multDiag(r, s);
            const double dold = dn;
            dn = dot(r, s);
            const double ba = dn / dold;
            for (int i = 0; i < n; ++i) d[i] = s[i] + ba * d[i];
        }
        pcgResidual = 0.5 * dn;
        for (int i = 0; i < n; ++i) xs[i] = xv[i];
        return true;
    }

    // BlockSolver::setLambda + solve (Schur) + restoreDiagonal for one damping value.
    bool solveDamped(double lambda) {
        const int np = 6 * F, n = np + 3 * NL;
        x.assign(n, 0.0);
        S.assign(skyOff.empty() ? 0 : skyOff[np], 0.0);
        // _Hschur = _Hpp (+ lambda on the diagonal)
        for (int i = 0; i < F; ++i)
            for (int a = 0; a < 6; ++a)
                for (int c = 0; c <= a; ++c) {
                    double v = Hpp[36 * (size_t)i + 6 * a + c];
                    if (a == c) v += lambda;
                    Sat(S, 6 * i + a, 6 * i + c) = v;
                }
        for (auto &kv : HppOff) {   // _Hschur = _Hpp: pose-pose blocks (i < j) live at lower (6j + c, 6i + a)
            const int i = kv.first.first, j = kv.first.second;
            for (int a = 0; a < 6; ++a)
                for (int c = 0; c < 6; ++c) Sat(S, 6 * j + c, 6 * i + a) += kv.second[6 * a + c];
        }
        coeff.assign(np, 0.0);
        Dinv.assign(9 * (size_t)NL, 0.0);
        auto landmarkWork = [&](int hl, std::vector<double> &St, std::vector<double> &ct) {
            const int l = hidx2point[hl];
            double D[9];
            for (int k = 0; k < 9; ++k) D[k] = Hll[9 * (size_t)hl + k];
            D[0] += lambda; D[4] += lambda; D[8] += lambda;
            double *Di = &Dinv[9 * (size_t)hl];
            inv3(D, Di);
            const double *bl = &b[np + 3 * hl];
            double db[3];
            for (int a = 0; a < 3; ++a) db[a] = Di[3 * a] * bl[0] + Di[3 * a + 1] * bl[1] + Di[3 * a + 2] * bl[2];
            const std::vector<int> &le = lmEdges[l];
            for (size_t ia = 0; ia < le.size(); ++ia) {
                const int e1 = le[ia];
                if (hplRow[e1] < 0) continue;
                const int i1 = hplRow[e1];
                const double *Bi = &Hpl[18 * (size_t)e1];
                double BDinv[18];
                for (int a = 0; a < 6; ++a)
                    for (int c = 0; c < 3; ++c)
                        BDinv[3 * a + c] = Bi[3 * a] * Di[c] + Bi[3 * a + 1] * Di[3 + c] + Bi[3 * a + 2] * Di[6 + c];
                for (int a = 0; a < 6; ++a)
                    ct[6 * i1 + a] += Bi[3 * a] * db[0] + Bi[3 * a + 1] * db[1] + Bi[3 * a + 2] * db[2];
                for (size_t ib = 0; ib < le.size(); ++ib) {
                    const int e2 = le[ib];
                    if (hplRow[e2] < 0) continue;
                    const int i2 = hplRow[e2];
                    if (i2 < i1) continue;  // upper block triangle only (lower_bound on row i1)
                    const double *Bj = &Hpl[18 * (size_t)e2];
                    // Hschur(i1,i2) -= BDinv * Bj^T ; stored through the lower-triangle skyline as (i2,i1)^T
                    for (int a = 0; a < 6; ++a)
                        for (int c = 0; c < 6; ++c) {
                            const int r = 6 * i2 + c, cc = 6 * i1 + a;  // element (a,c) of block (i1,i2) == (r,cc) lower
                            if (cc > r) continue;                        // diagonal block: keep lower half only
                            const double v = BDinv[3 * a] * Bj[3 * c] + BDinv[3 * a + 1] * Bj[3 * c + 1] + BDinv[3 * a + 2] * Bj[3 * c + 2];
                            St[skyOff[r] + (size_t)(cc - skyFirst[r])] -= v;
                        }
                }
            }
        };
#ifdef _OPENMP
        if (threads > 1) {
            std::vector<std::vector<double>> St(threads), ct(threads);
#pragma omp parallel num_threads(threads)
            {
                const int t = omp_get_thread_num();
                St[t].assign(S.size(), 0.0); ct[t].assign(np, 0.0);
#pragma omp for schedule(static)
                for (int hl = 0; hl < NL; ++hl) landmarkWork(hl, St[t], ct[t]);
            }
            for (int t = 0; t < threads; ++t) {
                for (size_t i = 0; i < S.size(); ++i) S[i] += St[t][i];
                for (int i = 0; i < np; ++i) coeff[i] += ct[t][i];
            }
        } else
#endif
        for (int hl = 0; hl < NL; ++hl) landmarkWork(hl, S, coeff);

        bs.assign(np, 0.0);
        for (int i = 0; i < np; ++i) bs[i] = b[i] - coeff[i];
        bool ok = true;
        if (np > 0) {
            if (solver == VISFS_BA_SOLVER_PCG) ok = pcgSolve(S, bs, x.data());
            else { std::vector<double> A = S; ok = choleskySolve(A, bs, x.data()); }
        }
        if (!ok) return false;
        // x_l = Dinv (b_l - Hpl^T x_p)
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(threads)
#endif
        for (int hl = 0; hl < NL; ++hl) {
            const int l = hidx2point[hl];
            double cl[3] = {b[np + 3 * hl], b[np + 3 * hl + 1], b[np + 3 * hl + 2]};
            for (int e : lmEdges[l]) {
                if (hplRow[e] < 0) continue;
                const double *B = &Hpl[18 * (size_t)e];
                const double *xp = &x[6 * hplRow[e]];
                for (int c = 0; c < 3; ++c)
                    for (int a = 0; a < 6; ++a) cl[c] -= B[3 * a + c] * xp[a];
            }
            const double *Di = &Dinv[9 * (size_t)hl];
            for (int a = 0; a < 3; ++a) x[np + 3 * hl + a] = Di[3 * a] * cl[0] + Di[3 * a + 1] * cl[1] + Di[3 * a + 2] * cl[2];
        }
        return true;
    }

    // SparseOptimizer::update: each free active vertex oplus its slice
    void applyUpdate() {
        for (int i = 0; i < P; ++i) if (phidx[i] >= 0) poseOplus(&pose[7 * i], &x[6 * phidx[i]]);
        const int np = 6 * F;
        for (int hl = 0; hl < NL; ++hl) {
            const int l = hidx2point[hl];
            for (int a = 0; a < 3; ++a) point[3 * l + a] += x[np + 3 * hl + a];
        }
    }

    // SparseOptimizer::optimize(maxIter) with OptimizationAlgorithmLevenberg / GaussNewton
    PassStats optimize(int maxIter) {
        PassStats st;
        buildStructure();
        st.F = F; st.NL = NL;
        pcgResidual = -1.0;
        st.chi2 = st.chi2_last_trial = computeErrorsAndChi2();
        if (F + NL == 0) { st.stop = VISFS_BA_STOP_EMPTY; return st; }
        double lambda = 0, ni = 2;
        st.stop = VISFS_BA_STOP_ITERATIONS;
        for (int it = 0; it < maxIter; ++it) {
            double currentChi = computeErrorsAndChi2();
            st.chi2_last_trial = currentChi;
            buildSystem();
            if (trust == VISFS_BA_GAUSS_NEWTON) {
                const bool ok = solveDamped(0.0);
                ++st.trials; ++st.iterations; st.trials_per_iteration.push_back(1);
                if (!ok) { st.stop = VISFS_BA_STOP_SOLVER_FAIL; break; }
                applyUpdate();
                continue;
            }
            if (it == 0) { lambda = lambdaInit(); ni = 2; }
            double rho = 0;
            int qmax = 0;
            do {
                std::vector<double> poseBak = pose, pointBak = point;  // push()
                const bool ok2 = solveDamped(lambda);
                applyUpdate();
                double tempChi = computeErrorsAndChi2();
                st.chi2_last_trial = tempChi;
                if (!ok2) tempChi = std::numeric_limits<double>::max();
                rho = currentChi - tempChi;
                double scale = 0;
                for (size_t j = 0; j < x.size(); ++j) scale += x[j] * (lambda * x[j] + b[j]);
                scale += 1e-3;
                rho /= scale;
                if (rho > 0 && std::isfinite(tempChi)) {
                    double alpha = 1. - std::pow((2 * rho - 1), 3);
                    alpha = std::min(alpha, 2. / 3.);
                    const double scaleFactor = std::max(1. / 3., alpha);
                    lambda *= scaleFactor;
                    ni = 2;
                    currentChi = tempChi;
                } else {
                    lambda *= ni;
                    ni *= 2;
                    pose = poseBak; point = pointBak;  // pop()
                    if (!std::isfinite(lambda)) { ++st.trials; break; }  // g2o breaks before qmax++
                }
                ++qmax; ++st.trials;
            } while (rho < 0 && qmax < 10);
            ++st.iterations;
            st.trials_per_iteration.push_back(qmax);
            if (qmax == 10 || rho == 0 || !std::isfinite(lambda)) { st.stop = VISFS_BA_STOP_TERMINATE; break; }
        }
        st.lambda = lambda;
        st.chi2 = computeErrorsAndChi2();  // accepted state (Optimizer.cpp:270-271)
        return st;
    }
};

void writeState(const Oracle &o, visfs_ba_result *r) {
    if (r->pose_tq) std::memcpy(r->pose_tq, o.pose.data(), sizeof(double) * o.pose.size());
    if (r->point_xyz) std::memcpy(r->point_xyz, o.point.data(), sizeof(double) * o.point.size());
    if (r->edge_level) std::memcpy(r->edge_level, o.level.data(), o.level.size());
}

// Optimizer.cpp:261-318: pass 1, guards, culling, pass 2, guard.
int solveTwoPass(Oracle &o, const visfs_ba_problem *p, visfs_ba_result *r, std::vector<int> *trialLog) {
    const bool single = (p->flags & VISFS_BA_FLAG_SINGLE_PASS) != 0;
    const int half = single ? p->iterations : p->iterations / 2;
    o.level.assign(o.E, 0);
    o.buildStructure();
    r->chi2_initial = o.computeErrorsAndChi2();
    PassStats s1 = o.optimize(half);
    r->iterations_run[0] = s1.iterations; r->trials_run[0] = s1.trials; r->stop_reason[0] = s1.stop;
    r->n_free_poses[0] = s1.F; r->n_free_points[0] = s1.NL; r->lambda_final[0] = s1.lambda;
    r->iterations_run[1] = r->trials_run[1] = 0; r->stop_reason[1] = VISFS_BA_STOP_NOT_RUN;
    r->n_free_poses[1] = r->n_free_points[1] = 0; r->lambda_final[1] = 0;
    r->chi2_pass1 = r->chi2_final = s1.chi2; r->chi2_last_trial = s1.chi2; r->n_outliers = 0;
    if (trialLog) *trialLog = s1.trials_per_iteration;
    const double chi2 = s1.chi2;
    if (std::isnan(chi2) || chi2 > 1000000000000.0 || !std::isfinite(chi2)) {
        writeState(o, r);
        return r->status = VISFS_BA_ERR_NUMERIC_PASS1;
    }
    if (p->huber_delta > 0.0 && !single) {
        int n = 0;
        if (!(p->flags & VISFS_BA_FLAG_NO_CULL))
            for (int e = 0; e < o.E; ++e)
                // never-active (all-fixed) edges carry no computed error in g2o: not culled here
                if (o.level[e] == 0 && o.eact[e] && o.edgeChi2(e) > p->huber_delta) { o.level[e] = 1; ++n; }
        r->n_outliers = n;
        PassStats s2 = o.optimize(half);
        r->iterations_run[1] = s2.iterations; r->trials_run[1] = s2.trials; r->stop_reason[1] = s2.stop;
        r->n_free_poses[1] = s2.F; r->n_free_points[1] = s2.NL; r->lambda_final[1] = s2.lambda;
        r->chi2_final = s2.chi2; r->chi2_last_trial = s2.chi2_last_trial;
        if (trialLog) trialLog->insert(trialLog->end(), s2.trials_per_iteration.begin(), s2.trials_per_iteration.end());
        if (s2.chi2_last_trial > 1000000000000.0) {
            writeState(o, r);
            return r->status = VISFS_BA_ERR_NUMERIC_PASS2;
        }
    }
    writeState(o, r);
    return r->status = VISFS_BA_OK;
}

int defaultThreads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // namespace

extern "C" {

int oracle_threads(void) { return defaultThreads(); }

int oracle_solve(const visfs_ba_problem *p, visfs_ba_result *r, int threads) {
    Oracle o;
    o.load(p);
    o.threads = threads > 0 ? threads : defaultThreads();
    return solveTwoPass(o, p, r, nullptr);
}

// same, also returning seconds spent inside the two optimize() passes (graph set-up excluded),
// and the trial count of every LM iteration (for the "same accept/reject sequence" check).
int oracle_solve_timed(const visfs_ba_problem *p, visfs_ba_result *r, int threads, double *seconds,
                       int32_t *trials_per_iteration, int32_t capacity, int32_t *n_iterations) {
    Oracle o;
    o.load(p);
    o.threads = threads > 0 ? threads : defaultThreads();
    std::vector<int> log;
    const auto t0 = std::chrono::steady_clock::now();
    const int st = solveTwoPass(o, p, r, &log);
    const auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    if (n_iterations) *n_iterations = (int32_t)log.size();
    if (trials_per_iteration)
        for (int i = 0; i < (int)log.size() && i < capacity; ++i) trials_per_iteration[i] = log[i];
    return st;
}

int oracle_linearize(const visfs_ba_problem *p, visfs_ba_linearization *out) {
    Oracle o;
    o.load(p);
    o.refreshR();
    for (int e = 0; e < o.E; ++e) {
        double r[3], Jl[9], Jp[18];
        const int pi = o.ep[e], li = o.el[e];
        edgeError(&o.pose[7 * pi], &o.R[9 * pi], &o.point[3 * li], &o.obs[3 * e], o.kind[e], o.K, r);
        edgeJacobians(&o.pose[7 * pi], &o.R[9 * pi], &o.point[3 * li], o.kind[e], o.K, Jl, Jp);
        const double c = (r[0] * r[0] + r[1] * r[1] + r[2] * r[2]) / o.pv;
        double r0, w;
        huber(c, o.delta, &r0, &w);
        if (out->error) std::memcpy(out->error + 3 * (size_t)e, r, sizeof r);
        if (out->chi2) out->chi2[e] = c;
        if (out->rho) out->rho[e] = r0;
        if (out->weight) out->weight[e] = w;
        if (out->J_point) std::memcpy(out->J_point + 9 * (size_t)e, Jl, sizeof Jl);
        if (out->J_pose) std::memcpy(out->J_pose + 18 * (size_t)e, Jp, sizeof Jp);
    }
    return VISFS_BA_OK;
}

int oracle_structure(const visfs_ba_problem *p, visfs_ba_structure *out) {
    Oracle o;
    o.load(p);
    if (out->edge_level) o.level.assign(out->edge_level, out->edge_level + o.E);
    o.buildStructure();
    if (out->pose_hidx) std::copy(o.phidx.begin(), o.phidx.end(), out->pose_hidx);
    if (out->point_hidx) std::copy(o.lhidx.begin(), o.lhidx.end(), out->point_hidx);
    if (out->edge_active) std::copy(o.eact.begin(), o.eact.end(), out->edge_active);
    if (out->hpl_row) std::copy(o.hplRow.begin(), o.hplRow.end(), out->hpl_row);
    if (out->hpl_col) std::copy(o.hplCol.begin(), o.hplCol.end(), out->hpl_col);
    out->n_schur_blocks = (int32_t)o.spat.size();
    for (int k = 0; k < (int)o.spat.size() && k < out->schur_capacity; ++k) {
        if (out->schur_cols) out->schur_cols[k] = o.spat[k].first;
        if (out->schur_rows) out->schur_rows[k] = o.spat[k].second;
    }
    out->n_free_poses = o.F; out->n_free_points = o.NL;
    out->n_active_edges = (int32_t)o.activeEdges.size();
    int nh = 0;
    for (int e = 0; e < o.E; ++e) nh += o.hplRow[e] >= 0;
    out->n_hpl_blocks = nh;
    return VISFS_BA_OK;
}

// EdgePoseConstraint::computeError / linearizeOplus of every link at the input state (parity hook)
int oracle_link_linearize(const visfs_ba_problem *p, double *err /* [K][6] */, double *Ji /* [K][6][6] */, double *Jj) {
    Oracle o;
    o.load(p);
    for (int k = 0; k < o.NK; ++k) {
        const double *a = &o.pose[7 * o.lkFrom[k]], *b = &o.pose[7 * o.lkTo[k]], *m = &o.lkM[7 * (size_t)k];
        if (err) linkError(a, b, m, err + 6 * (size_t)k);
        double J1[36], J2[36];
        linkJacobians(a, b, m, J1, J2);
        if (Ji) std::memcpy(Ji + 36 * (size_t)k, J1, sizeof J1);
        if (Jj) std::memcpy(Jj + 36 * (size_t)k, J2, sizeof J2);
    }
    return VISFS_BA_OK;
}

// CameraPose::update of n poses (parity hook against the reference's own compiled code, oracle/ref_shim.cpp)
int oracle_pose_oplus(int n, const double *tq_in /* [n][7] */, const double *delta /* [n][6] */, double *tq_out) {
    for (int i = 0; i < n; ++i) {
        double tq[7];
        std::memcpy(tq, tq_in + 7 * (size_t)i, sizeof tq);
        poseOplus(tq, delta + 6 * (size_t)i);
        std::memcpy(tq_out + 7 * (size_t)i, tq, sizeof tq);
    }
    return VISFS_BA_OK;
}

// One damped Schur solve at the input state: fills the reduced system (dense, row-major, full
// symmetric n x n with n = 6 * n_free_poses), its right-hand side and the full step x.
// Used to check the GPU's Hessian / Schur kernels block by block.
int oracle_reduced_system(const visfs_ba_problem *p, double lambda, double *S_dense, double *b_s, double *x_full,
                          int32_t *n_out, double *chi2_out, double *lambda_init_out) {
    Oracle o;
    o.load(p);
    o.level.assign(o.E, 0);
    o.buildStructure();
    const double chi = o.computeErrorsAndChi2();
    o.buildSystem();
    if (chi2_out) *chi2_out = chi;
    if (lambda_init_out) *lambda_init_out = o.lambdaInit();
    const bool ok = o.solveDamped(lambda);
    const int n = 6 * o.F;
    if (n_out) *n_out = n;
    if (S_dense)
        for (int r = 0; r < n; ++r)
            for (int c = 0; c < n; ++c) {
                const int rr = std::max(r, c), cc = std::min(r, c);
                S_dense[(size_t)r * n + c] = (cc >= o.skyFirst[rr]) ? o.S[o.skyOff[rr] + (size_t)(cc - o.skyFirst[rr])] : 0.0;
            }
    if (b_s) std::copy(o.bs.begin(), o.bs.end(), b_s);
    if (x_full) std::copy(o.x.begin(), o.x.end(), x_full);
    return ok ? VISFS_BA_OK : VISFS_BA_ERR_NUMERIC_PASS1;
}

}  // extern "C"

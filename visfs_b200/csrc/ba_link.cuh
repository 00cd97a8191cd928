// ba_link.cuh — odometry pose-pose constraints: EdgePoseConstraint of
// corelib/include/Optimizer/g2o/OptimizeTypeDefine.h:193-225 / corelib/src/Optimizer/g2o/OptimizeTypeDefine.cpp:35-88 with
// the quaternion operators of utilite/include/Math.h:293-346, added to the graph at corelib/src/Optimizer/Optimizer.cpp:116-150
// (information = I6 / Optimizer/OdometryCovariance, no robust kernel, never culled).
//
// A window has at most a few dozen links, so they do not touch the edge kernels: k_link_lin evaluates every link once
// per trial (error, both 6x6 Jacobians, the three Hessian blocks and the two gradient pieces) into a 128-double record,
// k_solve adds the records into the reduced camera system it assembles, k_control(_init) add the link chi2.
#pragma once
#include "ba_math.cuh"

namespace visfs {

constexpr int kLinkStride = 128;     // per link: chi2 (1) | b_i (6) | b_j (6) | H_ii (36) | H_jj (36) | H_ij (36)
constexpr int kLkChi = 0, kLkBi = 1, kLkBj = 7, kLkHii = 13, kLkHjj = 49, kLkHij = 85;

__device__ __forceinline__ void q_mul(const double *a, const double *b, double *o) {   // Eigen (Hamilton) product, (x, y, z, w)
    const double ax = a[0], ay = a[1], az = a[2], aw = a[3], bx = b[0], by = b[1], bz = b[2], bw = b[3];
    o[3] = aw * bw - ax * bx - ay * by - az * bz;
    o[0] = aw * bx + ax * bw + ay * bz - az * by;
    o[1] = aw * by + ay * bw + az * bx - ax * bz;
    o[2] = aw * bz + az * bw + ax * by - ay * bx;
}
__device__ __forceinline__ void q_inv(const double *q, double *o) {   // Eigen::Quaternion::inverse: conjugate / squaredNorm
    const double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    o[0] = -q[0] / n2; o[1] = -q[1] / n2; o[2] = -q[2] / n2; o[3] = q[3] / n2;
}
__device__ __forceinline__ void q_positify(const double *q, double *o) {   // Math.h:308-316
    const double s = (q[3] < 0.0) ? -1.0 : 1.0;
    const double t0 = s * q[0], t1 = s * q[1], t2 = s * q[2], t3 = s * q[3];
    const double n = sqrt(t0 * t0 + t1 * t1 + t2 * t2 + t3 * t3);
    o[0] = t0 / n; o[1] = t1 / n; o[2] = t2 / n; o[3] = t3 / n;
}
__device__ __forceinline__ void mat3_vec(const double *R, const double *v, double *o) {
#pragma unroll
    for (int i = 0; i < 3; ++i) o[i] = R[3 * i] * v[0] + R[3 * i + 1] * v[1] + R[3 * i + 2] * v[2];
}

// EdgePoseConstraint::computeError (OptimizeTypeDefine.cpp:35-51); tq = t(3) q(4) of vertex 0 / vertex 1, m = measurement
__device__ __forceinline__ void link_error(const double *tq1, const double *tq2, const double *m, double *e) {
    double q2i[4], q12[4], R12[9], rp[3];
    const double nP2[3] = {-tq2[0], -tq2[1], -tq2[2]};
    q_inv(tq2 + 3, q2i);
    q_mul(tq1 + 3, q2i, q12);
    quat_to_R(q12, R12);
    mat3_vec(R12, nP2, rp);
#pragma unroll
    for (int i = 0; i < 3; ++i) e[i] = rp[i] + tq1[i] - m[i];
    double mi[4], t[4];
    q_inv(m + 3, mi);
    q_mul(mi, q12, t);
    e[3] = 2.0 * t[0]; e[4] = 2.0 * t[1]; e[5] = 2.0 * t[2];
}

// EdgePoseConstraint::linearizeOplus, the live "Left update" branch (OptimizeTypeDefine.cpp:53-72): 6x6 row-major each
__device__ void link_jacobians(const double *tq1, const double *tq2, const double *m, double *Ji, double *Jj) {
    for (int i = 0; i < 36; ++i) { Ji[i] = 0.0; Jj[i] = 0.0; }
    double q1i[4], q2i[4], q12[4], R1[9], R2i[9], R12[9];
    const double nP2[3] = {-tq2[0], -tq2[1], -tq2[2]};
    q_inv(tq1 + 3, q1i); q_inv(tq2 + 3, q2i);
    q_mul(tq1 + 3, q2i, q12);
    quat_to_R(tq1 + 3, R1); quat_to_R(q2i, R2i); quat_to_R(q12, R12);
    // vertex 0
    for (int i = 0; i < 3; ++i) Ji[6 * i + i] = 1.0;
    double t1[3], t2[3];
    mat3_vec(R2i, nP2, t1); mat3_vec(R1, t1, t2);         // sQ1 * (sQ2.inverse() * (-sP2))
    // -skewSymmetric(t2)
    Ji[0 * 6 + 4] = t2[2];  Ji[0 * 6 + 5] = -t2[1];
    Ji[1 * 6 + 3] = -t2[2]; Ji[1 * 6 + 5] = t2[0];
    Ji[2 * 6 + 3] = t2[1];  Ji[2 * 6 + 4] = -t2[0];
    {
        double q21[4], pl[4], pr[4];
        q_mul(tq2 + 3, q1i, q21);                          // sQ2 * sQ1.inverse()
        q_positify(q21, pl); q_positify(m + 3, pr);
        // rows 1..3 of QuaternionLeft(pl) = [v | w I + skew(v)], columns 1..3 of QuaternionRight(pr) = [-v^T ; w I - skew(v)]
        const double L[3][4] = {{pl[0], pl[3], -pl[2], pl[1]}, {pl[1], pl[2], pl[3], -pl[0]}, {pl[2], -pl[1], pl[0], pl[3]}};
        const double Rr[4][3] = {{-pr[0], -pr[1], -pr[2]}, {pr[3], pr[2], -pr[1]}, {-pr[2], pr[3], pr[0]}, {pr[1], -pr[0], pr[3]}};
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double acc = 0.0;
                for (int k = 0; k < 4; ++k) acc += L[i][k] * Rr[k][j];
                Ji[6 * (3 + i) + 3 + j] = acc;
            }
    }
    // vertex 1
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Jj[6 * i + j] = -R12[3 * i + j];
    {
        const double Sn[9] = {0.0, -nP2[2], nP2[1], nP2[2], 0.0, -nP2[0], -nP2[1], nP2[0], 0.0};   // skewSymmetric(-sP2)
        double R1R2i[9];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) R1R2i[3 * i + j] = R1[3 * i] * R2i[j] + R1[3 * i + 1] * R2i[3 + j] + R1[3 * i + 2] * R2i[6 + j];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                Jj[6 * i + 3 + j] = R1R2i[3 * i] * Sn[j] + R1R2i[3 * i + 1] * Sn[3 + j] + R1R2i[3 * i + 2] * Sn[6 + j];
        double mi[4], t[4], pl[4];
        q_inv(m + 3, mi); q_mul(mi, q12, t);               // mQ12.inverse() * sQ1 * sQ2.inverse()
        q_positify(t, pl);
        // -(w I + skew(v))
        Jj[6 * 3 + 3] = -pl[3]; Jj[6 * 3 + 4] = pl[2];  Jj[6 * 3 + 5] = -pl[1];
        Jj[6 * 4 + 3] = -pl[2]; Jj[6 * 4 + 4] = -pl[3]; Jj[6 * 4 + 5] = pl[0];
        Jj[6 * 5 + 3] = pl[1];  Jj[6 * 5 + 4] = -pl[0]; Jj[6 * 5 + 5] = -pl[3];
    }
}

// every link of the batch at the accepted state: record for k_solve / k_control_init
__global__ void k_link_lin(Batch B) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= B.tot_link) return;
    const int w = B.link_win[k];
    const WinDesc &wd = B.win[w];
    const LMState &st = B.st[w];
    if (st.done) return;
    double *rec = B.link_lin + (size_t)k * kLinkStride;
    const int pf = wd.pose_off + B.link_from[k], pt = wd.pose_off + B.link_to[k];
    if ((B.pose_flags[pf] & kFixed) && (B.pose_flags[pt] & kFixed)) {   // e->allVerticesFixed(): never active
        for (int i = 0; i < kLkHij + 36; ++i) rec[i] = 0.0;
        return;
    }
    const double *ps = B.pose + (size_t)st.cur * B.tot_pose * kPoseStride;
    const double *a = ps + (size_t)pf * kPoseStride, *b = ps + (size_t)pt * kPoseStride, *m = B.link_m + 7 * (size_t)k;
    double e[6], Ji[36], Jj[36];
    link_error(a, b, m, e);
    link_jacobians(a, b, m, Ji, Jj);
    const double om = wd.inv_ov;
    double chi = 0.0;
    for (int i = 0; i < 6; ++i) chi += e[i] * e[i];
    rec[kLkChi] = chi * om;
    for (int x = 0; x < 6; ++x) {
        double si = 0.0, sj = 0.0;
        for (int r = 0; r < 6; ++r) { si += Ji[6 * r + x] * om * e[r]; sj += Jj[6 * r + x] * om * e[r]; }
        rec[kLkBi + x] = -si;
        rec[kLkBj + x] = -sj;
        for (int c = 0; c < 6; ++c) {
            double hii = 0.0, hjj = 0.0, hij = 0.0;
            for (int r = 0; r < 6; ++r) {
                hii += Ji[6 * r + x] * om * Ji[6 * r + c];
                hjj += Jj[6 * r + x] * om * Jj[6 * r + c];
                hij += Ji[6 * r + x] * om * Jj[6 * r + c];
            }
            rec[kLkHii + 6 * x + c] = hii;
            rec[kLkHjj + 6 * x + c] = hjj;
            rec[kLkHij + 6 * x + c] = hij;
        }
    }
}

// chi2 of the links of one window at the state buffer `buf` (block-wide, deterministic); 0 when the window has none
__device__ __forceinline__ double link_chi2_block(const Batch &B, const WinDesc &wd, int buf, double *scratch) {
    double acc = 0.0;
    const double *ps = B.pose + (size_t)buf * B.tot_pose * kPoseStride;
    for (int k = threadIdx.x; k < wd.n_link; k += blockDim.x) {
        const int g = wd.link_off + k;
        const int pf = wd.pose_off + B.link_from[g], pt = wd.pose_off + B.link_to[g];
        if ((B.pose_flags[pf] & kFixed) && (B.pose_flags[pt] & kFixed)) continue;
        double e[6];
        link_error(ps + (size_t)pf * kPoseStride, ps + (size_t)pt * kPoseStride, B.link_m + 7 * (size_t)g, e);
        double c = 0.0;
        for (int i = 0; i < 6; ++i) c += e[i] * e[i];
        acc += c * wd.inv_ov;
    }
    return block_sum(acc, scratch);
}


// parity hooks (visfs_ba_debug_pose_oplus / visfs_ba_debug_link_linearize): the device functions of the product path on
// caller-supplied operands, checked against the reference's own compiled code (tests/test_gpu_ref_pin.py)
__global__ void k_debug_pose_oplus(int n, const double *tq_in, const double *delta, double *tq_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double src[7], rec[kPoseStride];
    for (int k = 0; k < 7; ++k) src[k] = tq_in[7 * (size_t)i + k];
    pose_oplus(src, delta + 6 * (size_t)i, rec);
    for (int k = 0; k < 7; ++k) tq_out[7 * (size_t)i + k] = rec[k];
}

__global__ void k_debug_link(int n, const double *from_tq, const double *to_tq, const double *meas_tq, double *err, double *Ji, double *Jj) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double e[6], A[36], Bm[36];
    link_error(from_tq + 7 * (size_t)k, to_tq + 7 * (size_t)k, meas_tq + 7 * (size_t)k, e);
    link_jacobians(from_tq + 7 * (size_t)k, to_tq + 7 * (size_t)k, meas_tq + 7 * (size_t)k, A, Bm);
    for (int i = 0; i < 6; ++i) err[6 * (size_t)k + i] = e[i];
    for (int i = 0; i < 36; ++i) { Ji[36 * (size_t)k + i] = A[i]; Jj[36 * (size_t)k + i] = Bm[i]; }
}

}  // namespace visfs

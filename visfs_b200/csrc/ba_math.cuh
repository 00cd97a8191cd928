// ba_math.cuh — per-edge arithmetic of the stereo / mono reprojection edge in FP64.
// Formulas: corelib/include/Optimizer/g2o/OptimizeTypeDefine.h:121-187 (EdgeStereo),
// :45-47 (CameraPose::map), corelib/src/Optimizer/g2o/OptimizeTypeDefine.cpp:7-14 (update),
// utilite/include/Math.h:277-287 (deltaQ); Huber as g2o::RobustKernelHuber::robustify.
#pragma once
#include "ba_types.cuh"

namespace visfs {

struct Intr { double fx, fy, cx, cy, bf, inv_pv, delta; };

__device__ __forceinline__ Intr load_intr(const WinDesc &w) {
    return Intr{w.fx, w.fy, w.cx, w.cy, w.bf, w.inv_pv, w.delta};
}

// Eigen::Quaternion::toRotationMatrix, q = (x, y, z, w)
__host__ __device__ __forceinline__ void quat_to_R(const double *q, double *R) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz;         R[2] = txz + twy;
    R[3] = txy + twz;         R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;         R[7] = tyz + twx;         R[8] = 1.0 - (txx + tyy);
}

// pose record: t(3) q(4) R(9).  CameraPose::update: t += d[0:3]; q = normalize(dq * q), dq = (d[3:6]/2, 1)
__device__ __forceinline__ void pose_oplus(const double *src, const double *d, double *dst) {
    dst[0] = src[0] + d[0]; dst[1] = src[1] + d[1]; dst[2] = src[2] + d[2];
    const double ax = d[3] / 2.0, ay = d[4] / 2.0, az = d[5] / 2.0;
    const double bx = src[3], by = src[4], bz = src[5], bw = src[6];
    const double w = bw - ax * bx - ay * by - az * bz;
    const double x = bx + ax * bw + ay * bz - az * by;
    const double y = by + ay * bw + az * bx - ax * bz;
    const double z = bz + az * bw + ax * by - ay * bx;
    const double n = sqrt(x * x + y * y + z * z + w * w);
    dst[3] = x / n; dst[4] = y / n; dst[5] = z / n; dst[6] = w / n;
    quat_to_R(dst + 3, dst + 7);
}

__device__ __forceinline__ void huber(double e2, double delta, double &rho, double &w) {
    rho = e2; w = 1.0;
    if (delta > 0.0) {
        const double dsqr = delta * delta;
        if (e2 > dsqr) {
            const double sqrte = sqrt(e2);
            rho = 2.0 * sqrte * delta - dsqr;
            w = delta / sqrte;
        }
    }
}

// residual only (computeError + project); mono: row 2 = 0
__device__ __forceinline__ void edge_residual(const double *ps, double px, double py, double pz, double ou, double ov,
                                              double our, bool mono, const Intr &K, double &r0, double &r1, double &r2) {
    const double *R = ps + 7;
    const double x = (R[0] * px + R[1] * py + R[2] * pz) + ps[0];
    const double y = (R[3] * px + R[4] * py + R[5] * pz) + ps[1];
    const double z = (R[6] * px + R[7] * py + R[8] * pz) + ps[2];
    const double iz = 1.0 / z;
    const double u = x * iz * K.fx + K.cx;
    const double v = y * iz * K.fy + K.cy;
    r0 = ou - u;
    r1 = ov - v;
    r2 = mono ? 0.0 : our - (u - K.bf * iz);
}

struct EdgeLin {
    double r[3];
    double Jl[9];    // 3x3 row-major  d e / d point
    double Jp[18];   // 3x6 row-major  d e / d (t, theta)
    double chi2, rho, w;
};

// residual + both Jacobians + Huber weight
__device__ __forceinline__ void edge_linearize(const double *ps, double px, double py, double pz, double ou, double ov,
                                               double our, bool mono, const Intr &K, EdgeLin &o) {
    const double *R = ps + 7;
    const double x = (R[0] * px + R[1] * py + R[2] * pz) + ps[0];
    const double y = (R[3] * px + R[4] * py + R[5] * pz) + ps[1];
    const double z = (R[6] * px + R[7] * py + R[8] * pz) + ps[2];
    const double iz = 1.0 / z, iz2 = iz * iz;
    const double fx = K.fx, fy = K.fy, bf = K.bf;
    const double u = x * iz * fx + K.cx;
    const double v = y * iz * fy + K.cy;
    o.r[0] = ou - u;
    o.r[1] = ov - v;
    o.r[2] = mono ? 0.0 : our - (u - bf * iz);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double j0 = -fx * R[k] * iz + fx * x * R[6 + k] * iz2;
        const double j1 = -fy * R[3 + k] * iz + fy * y * R[6 + k] * iz2;
        o.Jl[k] = j0;
        o.Jl[3 + k] = j1;
        o.Jl[6 + k] = mono ? 0.0 : j0 - bf * R[6 + k] * iz2;
    }
    o.Jp[0] = -iz * fx;  o.Jp[1] = 0.0;  o.Jp[2] = x * iz2 * fx;  o.Jp[3] = x * y * iz2 * fx;
    o.Jp[4] = -(1.0 + x * x * iz2) * fx;  o.Jp[5] = y * iz * fx;
    o.Jp[6] = 0.0;  o.Jp[7] = -iz * fy;  o.Jp[8] = y * iz2 * fy;  o.Jp[9] = (1.0 + y * y * iz2) * fy;
    o.Jp[10] = -x * y * iz2 * fy;  o.Jp[11] = -x * iz * fy;
    if (mono) {
#pragma unroll
        for (int k = 0; k < 6; ++k) o.Jp[12 + k] = 0.0;
    } else {
        o.Jp[12] = o.Jp[0];  o.Jp[13] = 0.0;  o.Jp[14] = o.Jp[2] - bf * iz2;  o.Jp[15] = o.Jp[3] - bf * y * iz2;
        o.Jp[16] = o.Jp[4] + bf * x * iz2;  o.Jp[17] = o.Jp[5];
    }
    o.chi2 = (o.r[0] * o.r[0] + o.r[1] * o.r[1] + o.r[2] * o.r[2]) * K.inv_pv;
    huber(o.chi2, K.delta, o.rho, o.w);
}

// inverse of a symmetric 3x3 given as (a00 a01 a02 a11 a12 a22); result in the same packing
__device__ __forceinline__ void inv_sym3(const double *A, double *Ai) {
    const double c00 = A[3] * A[5] - A[4] * A[4];
    const double c01 = A[4] * A[2] - A[1] * A[5];
    const double c02 = A[1] * A[4] - A[3] * A[2];
    const double det = A[0] * c00 + A[1] * c01 + A[2] * c02;
    const double id = 1.0 / det;
    Ai[0] = c00 * id; Ai[1] = c01 * id; Ai[2] = c02 * id;
    Ai[3] = (A[0] * A[5] - A[2] * A[2]) * id;
    Ai[4] = (A[1] * A[2] - A[0] * A[4]) * id;
    Ai[5] = (A[0] * A[3] - A[1] * A[1]) * id;
}

// y = Asym * x with the packing above
__device__ __forceinline__ void sym3_mul(const double *A, const double *x, double *y) {
    y[0] = A[0] * x[0] + A[1] * x[1] + A[2] * x[2];
    y[1] = A[1] * x[0] + A[3] * x[1] + A[4] * x[2];
    y[2] = A[2] * x[0] + A[4] * x[1] + A[5] * x[2];
}

// deterministic CTA-wide sum: shuffle tree inside each warp, then warps added in index order
__device__ __forceinline__ double block_sum(double v, double *scratch /* >= 32 doubles */) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    double s = 0.0;
    const int nw = (blockDim.x + 31) >> 5;
    for (int k = 0; k < nw; ++k) s += scratch[k];
    return s;
}

__device__ __forceinline__ double block_max(double v, double *scratch) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    double s = 0.0;
    const int nw = (blockDim.x + 31) >> 5;
    for (int k = 0; k < nw; ++k) s = fmax(s, scratch[k]);
    return s;
}

}  // namespace visfs

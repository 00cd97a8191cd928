// ba_update.cuh — k_update: landmark back-substitution x_l = Dinv (b_l - sum_e W_e^T x_p), point oplus into the
// trial buffer, robust chi2 of the trial state, and the landmark part of g2o's computeScale().
// Same chunk / tile table as the build kernel; the linearisation is recomputed instead of being stored
// (32 B/edge re-read instead of 288 B/edge written and read back).  One edge per thread; the next tile's edge
// records and points are prefetched into registers while the current tile is computed; per-landmark sums are
// formed by (landmark, entry) owner threads in edge order.
#pragma once
#include "ba_math.cuh"

namespace visfs {

// residual, d e / d point, Huber weight and Jp * x_p for one edge, without materialising the 3x6 pose Jacobian
__device__ __forceinline__ void edge_linearize_jx(const double *ps, double px, double py, double pz, double ou, double ov,
                                                  double our, bool mono, const Intr &K, const double *xp /* 6 or null */,
                                                  double *r, double *Jl, double *jx, double &w) {
    const double *R = ps + 7;
    const double x = (R[0] * px + R[1] * py + R[2] * pz) + ps[0];
    const double y = (R[3] * px + R[4] * py + R[5] * pz) + ps[1];
    const double z = (R[6] * px + R[7] * py + R[8] * pz) + ps[2];
    const double iz = 1.0 / z, iz2 = iz * iz;
    const double fx = K.fx, fy = K.fy, bf = K.bf;
    const double u = x * iz * fx + K.cx;
    const double v = y * iz * fy + K.cy;
    r[0] = ou - u;
    r[1] = ov - v;
    r[2] = mono ? 0.0 : our - (u - bf * iz);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double j0 = -fx * R[k] * iz + fx * x * R[6 + k] * iz2;
        const double j1 = -fy * R[3 + k] * iz + fy * y * R[6 + k] * iz2;
        Jl[k] = j0;
        Jl[3 + k] = j1;
        Jl[6 + k] = mono ? 0.0 : j0 - bf * R[6 + k] * iz2;
    }
    if (xp) {
        const double a0 = -iz * fx, a2 = x * iz2 * fx, a3 = x * y * iz2 * fx, a4 = -(1.0 + x * x * iz2) * fx, a5 = y * iz * fx;
        const double b1 = -iz * fy, b2 = y * iz2 * fy, b3 = (1.0 + y * y * iz2) * fy, b4 = -x * y * iz2 * fy, b5 = -x * iz * fy;
        jx[0] = fma(a0, xp[0], fma(a2, xp[2], fma(a3, xp[3], fma(a4, xp[4], a5 * xp[5]))));
        jx[1] = fma(b1, xp[1], fma(b2, xp[2], fma(b3, xp[3], fma(b4, xp[4], b5 * xp[5]))));
        jx[2] = mono ? 0.0
                     : fma(a0, xp[0], fma(a2 - bf * iz2, xp[2], fma(a3 - bf * y * iz2, xp[3], fma(a4 + bf * x * iz2, xp[4], a5 * xp[5]))));
    } else {
        jx[0] = jx[1] = jx[2] = 0.0;
    }
    double rho;
    huber((r[0] * r[0] + r[1] * r[1] + r[2] * r[2]) * K.inv_pv, K.delta, rho, w);
}

struct UpdateSmem {
    double pose[kMaxSmallPoses * kPoseStride];
    double poseT[kMaxSmallPoses * kPoseStride];
    double xp[kMaxSmallPoses * 6];
    double H[kTileEdges * 12];       // per edge: H_ll (6) b_l (3) W^T x_p (3)
    double lm[kTileLm * 12];         // per landmark sums of the same
    double newp[kTileLm * 3];
    double red[32];
    int hidx[kMaxSmallPoses];
    int lmoff[kTileLm + 1];
};

struct UpdRec {
    double ou, ov, our, px, py, pz;
    int pw, gl;
    uint8_t lf, pf;
};

__device__ __forceinline__ void upd_load_l1(const Batch &B, const WinDesc &wd, const Tile &T, int tid, UpdRec &r) {
    if (tid < T.ne) {
        const int e = T.e0 + tid;
        r.pw = B.edge_pose[e];
        r.gl = wd.point_off + B.edge_point[e];
        r.ou = B.obs_u[e]; r.ov = B.obs_v[e]; r.our = B.obs_r[e];
    }
}
__device__ __forceinline__ void upd_load_l2(const Batch &B, const WinDesc &wd, const Tile &T, int tid, const double *gpoint, UpdRec &r) {
    if (tid < T.ne) {
        r.lf = B.lm_flags[r.gl];
        r.pf = B.pose_flags[wd.pose_off + (r.pw & kPoseMask)];
        r.px = gpoint[3 * (size_t)r.gl]; r.py = gpoint[3 * (size_t)r.gl + 1]; r.pz = gpoint[3 * (size_t)r.gl + 2];
    }
}

constexpr int kUpdThreads = kTileEdges;   // 160: one edge per thread, 5 warps

__global__ void __launch_bounds__(kUpdThreads, 4) k_update(Batch B) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    UpdateSmem &sm = *reinterpret_cast<UpdateSmem *>(smem_raw);
    const int tid = threadIdx.x;
    const Chunk ck = B.chunks[blockIdx.x];
    const WinDesc &wd = B.win[ck.win];
    const LMState &st = B.st[ck.win];
    if (st.done) return;
    const int cur = st.cur;
    const double lambda = (wd.trust == 0) ? (B.dbg_lambda >= 0.0 ? B.dbg_lambda : st.lambda) : 0.0;
    const Intr K = load_intr(wd);
    const int pose_off = wd.pose_off, n_pose = wd.n_pose;
    const double *__restrict__ gpose = B.pose + ((size_t)cur * B.tot_pose + pose_off) * kPoseStride;
    const double *__restrict__ gposeT = B.pose + ((size_t)(1 - cur) * B.tot_pose + pose_off) * kPoseStride;
    const double *__restrict__ gpoint = B.point + (size_t)cur * B.tot_point * 3;
    double *__restrict__ gpointT = B.point + (size_t)(1 - cur) * B.tot_point * 3;
    const int tile0 = B.chunk_tile_off[blockIdx.x], ntiles = B.chunk_tile_off[blockIdx.x + 1] - tile0;
    const Tile *__restrict__ tiles = B.tiles + tile0;
    for (int i = tid; i < n_pose * kPoseStride; i += kUpdThreads) { sm.pose[i] = gpose[i]; sm.poseT[i] = gposeT[i]; }
    for (int i = tid; i < n_pose; i += kUpdThreads) sm.hidx[i] = B.pose_hidx[pose_off + i];
    for (int i = tid; i < st.F * 6; i += kUpdThreads) sm.xp[i] = B.xp[(size_t)pose_off * 6 + i];
    double chi_acc = 0.0, scale_acc = 0.0;
    UpdRec rec, nxt;
    Tile T, Tn;
    if (ntiles > 0) {
        T = tiles[0];
        upd_load_l1(B, wd, T, tid, rec);
        upd_load_l2(B, wd, T, tid, gpoint, rec);
    }
    __syncthreads();

    for (int t = 0; t < ntiles; ++t) {
        const bool more = t + 1 < ntiles;
        if (more) { Tn = tiles[t + 1]; upd_load_l1(B, wd, Tn, tid, nxt); }
        const int ne = T.ne, ntl = T.ntl, lt = T.lt;
        if (tid <= ntl) sm.lmoff[tid] = min(B.lm_edge_off[lt + tid] - T.e0, kTileEdges);
        bool act = false, mono = false;
        int tl = 0, p = 0;
        if (tid < ne) {
            p = rec.pw & kPoseMask;
            mono = (rec.pw & kMonoBit) != 0;
            tl = rec.gl - lt;
            act = !(rec.pw & kCulledBit) && !((rec.lf & kFixed) && (rec.pf & kFixed));
            double *hl = sm.H + tid * 12;
            if (act && (rec.lf & kInHessian)) {
                double r[3], J[9], v[3], w;
                const int hi = sm.hidx[p];
                edge_linearize_jx(sm.pose + p * kPoseStride, rec.px, rec.py, rec.pz, rec.ou, rec.ov, rec.our, mono, K,
                                  hi >= 0 ? sm.xp + hi * 6 : nullptr, r, J, v, w);
                const double wo = w * K.inv_pv;
                hl[0] = wo * fma(J[0], J[0], fma(J[3], J[3], J[6] * J[6]));
                hl[1] = wo * fma(J[0], J[1], fma(J[3], J[4], J[6] * J[7]));
                hl[2] = wo * fma(J[0], J[2], fma(J[3], J[5], J[6] * J[8]));
                hl[3] = wo * fma(J[1], J[1], fma(J[4], J[4], J[7] * J[7]));
                hl[4] = wo * fma(J[1], J[2], fma(J[4], J[5], J[7] * J[8]));
                hl[5] = wo * fma(J[2], J[2], fma(J[5], J[5], J[8] * J[8]));
                hl[6] = -wo * fma(J[0], r[0], fma(J[3], r[1], J[6] * r[2]));
                hl[7] = -wo * fma(J[1], r[0], fma(J[4], r[1], J[7] * r[2]));
                hl[8] = -wo * fma(J[2], r[0], fma(J[5], r[1], J[8] * r[2]));
                // W^T x_p = wo * Jl^T (Jp x_p)
                const double v0 = wo * v[0], v1 = wo * v[1], v2 = wo * v[2];
                hl[9] = fma(J[0], v0, fma(J[3], v1, J[6] * v2));
                hl[10] = fma(J[1], v0, fma(J[4], v1, J[7] * v2));
                hl[11] = fma(J[2], v0, fma(J[5], v1, J[8] * v2));
            } else {
#pragma unroll
                for (int q = 0; q < 12; ++q) hl[q] = 0.0;
            }
        }
        if (more) upd_load_l2(B, wd, Tn, tid, gpoint, nxt);
        __syncthreads();
        for (int task = tid; task < ntl * 12; task += kUpdThreads) {
            const int l = task / 12, q = task - l * 12;
            double s = 0.0;
            for (int e = sm.lmoff[l]; e < sm.lmoff[l + 1]; ++e) s += sm.H[e * 12 + q];
            sm.lm[task] = s;
        }
        __syncthreads();
        if (tid < ntl) {
            const int gl = lt + tid;
            const double px = gpoint[3 * (size_t)gl], py = gpoint[3 * (size_t)gl + 1], pz = gpoint[3 * (size_t)gl + 2];
            double np0 = px, np1 = py, np2 = pz;
            if (B.lm_flags[gl] & kInHessian) {
                const double *ls = sm.lm + tid * 12;
                double A[6] = {ls[0] + lambda, ls[1], ls[2], ls[3] + lambda, ls[4], ls[5] + lambda};
                const double bl[3] = {ls[6], ls[7], ls[8]};
                double c[3] = {bl[0] - ls[9], bl[1] - ls[10], bl[2] - ls[11]};
                double Di[6], xl[3];
                inv_sym3(A, Di);
                sym3_mul(Di, c, xl);
                np0 = px + xl[0]; np1 = py + xl[1]; np2 = pz + xl[2];
                gpointT[3 * (size_t)gl] = np0; gpointT[3 * (size_t)gl + 1] = np1; gpointT[3 * (size_t)gl + 2] = np2;
                scale_acc += xl[0] * (lambda * xl[0] + bl[0]) + xl[1] * (lambda * xl[1] + bl[1]) + xl[2] * (lambda * xl[2] + bl[2]);
            }
            sm.newp[tid * 3] = np0; sm.newp[tid * 3 + 1] = np1; sm.newp[tid * 3 + 2] = np2;
        }
        __syncthreads();
        if (tid < ne && act) {
            double r0, r1, r2;
            edge_residual(sm.poseT + p * kPoseStride, sm.newp[tl * 3], sm.newp[tl * 3 + 1], sm.newp[tl * 3 + 2], rec.ou, rec.ov,
                          rec.our, mono, K, r0, r1, r2);
            double rho, wgt;
            huber((r0 * r0 + r1 * r1 + r2 * r2) * K.inv_pv, K.delta, rho, wgt);
            chi_acc += rho;
        }
        if (more) { T = Tn; rec = nxt; }
        // (the next iteration's writes to sm.H / sm.lmoff come after every thread has passed the barrier above;
        //  sm.newp is rewritten only after the next two barriers)
    }
    const double chi = block_sum(chi_acc, sm.red);
    const double sc = block_sum(scale_acc, sm.red);
    if (tid == 0) { B.part2[2 * (size_t)blockIdx.x] = chi; B.part2[2 * (size_t)blockIdx.x + 1] = sc; }
}

}  // namespace visfs

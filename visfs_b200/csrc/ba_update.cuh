// ba_update.cuh — k_update: landmark back-substitution x_l = Dinv (b_l - sum_e W_e^T x_p), point oplus into the
// trial buffer, robust chi2 of the trial state, and the landmark part of g2o's computeScale().
// Same chunk / tile table as the build kernel; the linearisation is recomputed instead of being stored
// (32 B/edge re-read instead of 288 B/edge written and read back).  One edge per lane; every warp walks its own
// warp tiles (whole landmarks, <= 32 edges) without block-level barriers; per-landmark sums are formed by
// (landmark, entry) owner lanes in edge order.
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
    const double fxiz = fx * iz, fyiz = fy * iz;
    const double bx = fxiz * (x * iz), by = fyiz * (y * iz), bb = bf * iz2;   // fx x / z^2, fy y / z^2, bf / z^2
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double j0 = fma(bx, R[6 + k], -fxiz * R[k]);
        const double j1 = fma(by, R[6 + k], -fyiz * R[3 + k]);
        Jl[k] = j0;
        Jl[3 + k] = j1;
        Jl[6 + k] = mono ? 0.0 : fma(-bb, R[6 + k], j0);
    }
    if (xp) {
        const double a0 = -fxiz, a2 = bx, a3 = bx * y, a4 = -fma(bx, x, fx), a5 = fxiz * y;
        const double b1 = -fyiz, b2 = by, b3 = fma(by, y, fy), b4 = -by * x, b5 = -fyiz * x;
        jx[0] = fma(a0, xp[0], fma(a2, xp[2], fma(a3, xp[3], fma(a4, xp[4], a5 * xp[5]))));
        jx[1] = fma(b1, xp[1], fma(b2, xp[2], fma(b3, xp[3], fma(b4, xp[4], b5 * xp[5]))));
        // row 2 = row 0 + bf / z^2 * (0, 0, -1, -y, x, 0)
        jx[2] = mono ? 0.0 : fma(bb, fma(x, xp[4], -fma(y, xp[3], xp[2])), jx[0]);
    } else {
        jx[0] = jx[1] = jx[2] = 0.0;
    }
    double rho;
    huber((r[0] * r[0] + r[1] * r[1] + r[2] * r[2]) * K.inv_pv, K.delta, rho, w);
}

// Warp tiles: consecutive landmarks of one chunk packed so that their edges fit one warp (<= 32 edges, <= 8 landmarks);
// a landmark with more than 32 edges is a tile of its own and is walked in rounds.  Built once per upload
// (k_count_wtiles / k_fill_wtiles).  Warps never wait for each other: the per-landmark sums go through the warp's own
// shared-memory slab (summed in edge order, like g2o) and only __syncwarp orders them.
constexpr int kWtLm = 8;
constexpr int kUpdWarps = 8;
constexpr int kUpdThreads = kUpdWarps * 32;
constexpr int kHs = 13;                  // padded row length of the per-edge slab (12 values)

struct UpdWarp {
    double H[32 * kHs];                  // per edge: H_ll (6) b_l (3) W^T x_p (3)
    double lm[kWtLm * 12];               // per landmark sums of the same
    int lmoff[kWtLm + 1];
    int pad;
};
struct UpdateSmem {
    double pose[kMaxSmallPoses * kPoseSm];
    double poseT[kMaxSmallPoses * kPoseSm];
    double xp[kMaxSmallPoses * 6];
    UpdWarp w[kUpdWarps];
    double red[32];
    int hidx[kMaxSmallPoses];
    unsigned char pflag[kMaxSmallPoses];
};

template <bool WRITE>
__device__ int walk_wtiles(const int *__restrict__ off, int lm0, int lm1, Tile *out) {
    int n = 0;
    for (int lt = lm0; lt < lm1;) {
        const int e0 = off[lt];
        int l1 = lt;
        while (l1 < lm1 && l1 - lt < kWtLm && off[l1 + 1] - e0 <= 32) ++l1;
        if (l1 == lt) l1 = lt + 1;       // one landmark with more than 32 edges
        if (off[l1] - e0 > 0) {          // landmarks without edges have nothing to update
            if (WRITE) out[n] = Tile{lt, l1 - lt, e0, off[l1] - e0};
            ++n;
        }
        lt = l1;
    }
    return n;
}
__global__ void k_count_wtiles(Batch B, int *ntiles) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > B.n_chunks) return;
    ntiles[c] = (c < B.n_chunks) ? walk_wtiles<false>(B.lm_edge_off, B.chunks[c].lm0, B.chunks[c].lm1, nullptr) : 0;
}
__global__ void k_fill_wtiles(Batch B, const int *tile_off, Tile *tiles) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= B.n_chunks) return;
    walk_wtiles<true>(B.lm_edge_off, B.chunks[c].lm0, B.chunks[c].lm1, tiles + tile_off[c]);
}

// the 12 per-edge terms of the landmark back-substitution
__device__ __forceinline__ void upd_edge_terms(const double *ps, double px, double py, double pz, double ou, double ov, double our,
                                               bool mono, const Intr &K, const double *xp, double *hl) {
    double r[3], J[9], v[3], w;
    edge_linearize_jx(ps, px, py, pz, ou, ov, our, mono, K, xp, r, J, v, w);
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
    const double v0 = wo * v[0], v1 = wo * v[1], v2 = wo * v[2];   // W^T x_p = wo * Jl^T (Jp x_p)
    hl[9] = fma(J[0], v0, fma(J[3], v1, J[6] * v2));
    hl[10] = fma(J[1], v0, fma(J[4], v1, J[7] * v2));
    hl[11] = fma(J[2], v0, fma(J[5], v1, J[8] * v2));
}

// only the last 3 of those terms (W^T x_p): H_ll and b_l come from k_build_ws's per-landmark sums (Batch::lm_sum)
__device__ __forceinline__ void upd_edge_wx(const double *ps, double px, double py, double pz, double ou, double ov, double our,
                                            bool mono, const Intr &K, const double *xp, double *hw) {
    double r[3], J[9], v[3], w;
    edge_linearize_jx(ps, px, py, pz, ou, ov, our, mono, K, xp, r, J, v, w);
    const double wo = w * K.inv_pv;
    const double v0 = wo * v[0], v1 = wo * v[1], v2 = wo * v[2];   // W^T x_p = wo * Jl^T (Jp x_p)
    hw[0] = fma(J[0], v0, fma(J[3], v1, J[6] * v2));
    hw[1] = fma(J[1], v0, fma(J[4], v1, J[7] * v2));
    hw[2] = fma(J[2], v0, fma(J[5], v1, J[8] * v2));
}

// x_l = (H_ll + lambda I)^-1 (b_l - sum W^T x_p) from the 12 landmark sums; returns the landmark's part of computeScale()
__device__ __forceinline__ double upd_point_step(const double *ls, double lambda, double *xl) {
    const double A[6] = {ls[0] + lambda, ls[1], ls[2], ls[3] + lambda, ls[4], ls[5] + lambda};
    const double c[3] = {ls[6] - ls[9], ls[7] - ls[10], ls[8] - ls[11]};
    double Di[6];
    inv_sym3(A, Di);
    sym3_mul(Di, c, xl);
    return xl[0] * (lambda * xl[0] + ls[6]) + xl[1] * (lambda * xl[1] + ls[7]) + xl[2] * (lambda * xl[2] + ls[8]);
}

__global__ void __launch_bounds__(kUpdThreads, 2) k_update(Batch B) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    UpdateSmem &sm = *reinterpret_cast<UpdateSmem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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
    const int tile0 = B.chunk_wtile_off[blockIdx.x], ntiles = B.chunk_wtile_off[blockIdx.x + 1] - tile0;
    const Tile *__restrict__ tiles = B.wtiles + tile0;
    for (int i = tid; i < n_pose * kPoseStride; i += kUpdThreads) {
        const int j = (i >> 4) * kPoseSm + (i & 15);
        sm.pose[j] = gpose[i]; sm.poseT[j] = gposeT[i];
    }
    for (int i = tid; i < n_pose; i += kUpdThreads) { sm.hidx[i] = B.pose_hidx[pose_off + i]; sm.pflag[i] = B.pose_flags[pose_off + i]; }
    for (int i = tid; i < st.F * 6; i += kUpdThreads) sm.xp[i] = B.xp[(size_t)pose_off * 6 + i];
    __syncthreads();
    UpdWarp &W = sm.w[warp];
    double chi_acc = 0.0, scale_acc = 0.0;

    Tile Tnext{0, 0, 0, 0};
    if (warp < ntiles) Tnext = tiles[warp];
    for (int t = warp; t < ntiles; t += kUpdWarps) {
        const Tile T = Tnext;
        if (t + kUpdWarps < ntiles) Tnext = tiles[t + kUpdWarps];   // descriptor of the next tile: off the critical path
        if (T.ne <= 32) {
            // ---- several whole landmarks, one edge per lane.  Everything is addressed from the tile descriptor alone
            //      (landmarks of a tile are consecutive), so all global loads of the tile are issued at once:
            //      lane l holds CSR offset and flags of landmark l, lanes 0..3 ntl - 1 hold the tile's coordinates
            const int off_l = (lane <= T.ntl) ? B.lm_edge_off[T.lt + lane] - T.e0 : 0x7fff;
            const int lf_l = (lane < T.ntl) ? (int)B.lm_flags[T.lt + lane] : 0;
            const double pt = (lane < 3 * T.ntl) ? gpoint[3 * (size_t)T.lt + lane] : 0.0;
            int pw = 0;
            double ou = 0, ov = 0, our = 0;
            if (lane < T.ne) {
                const int e = T.e0 + lane;
                pw = B.edge_pose[e];
                ou = B.obs_u[e]; ov = B.obs_v[e]; our = B.obs_r[e];
            }
            if (lane <= T.ntl) W.lmoff[lane] = off_l;
            int tl = 0;
#pragma unroll
            for (int l = 1; l < kWtLm; ++l) {
                const int v = __shfl_sync(0xffffffffu, off_l, l);
                tl += (l < T.ntl && v <= lane) ? 1 : 0;
            }
            const int lf = __shfl_sync(0xffffffffu, lf_l, tl);
            const double px = __shfl_sync(0xffffffffu, pt, 3 * tl), py = __shfl_sync(0xffffffffu, pt, 3 * tl + 1),
                         pz = __shfl_sync(0xffffffffu, pt, 3 * tl + 2);
            bool act = false, mono = false, lmfree = false;
            int p = 0;
            if (B.lm_sum) {
                // H_ll / b_l of the tile's landmarks as k_build_ws summed them (9 per landmark, contiguous: one coalesced load,
                // issued with the other loads of the tile); only W^T x_p is formed per edge and summed per landmark here
                const double *gs = B.lm_sum + 9 * (size_t)T.lt;
                double s0 = (lane < 9 * T.ntl) ? gs[lane] : 0.0, s1 = (lane + 32 < 9 * T.ntl) ? gs[lane + 32] : 0.0,
                       s2 = (lane + 64 < 9 * T.ntl) ? gs[lane + 64] : 0.0;
                double hw[3] = {0.0, 0.0, 0.0};
                if (lane < T.ne) {
                    p = pw & kPoseMask;
                    mono = (pw & kMonoBit) != 0;
                    lmfree = (lf & kInHessian) != 0;
                    act = !(pw & kCulledBit) && !((lf & kFixed) && (sm.pflag[p] & kFixed));
                    const int hi = sm.hidx[p];
                    if (act && lmfree && hi >= 0) upd_edge_wx(sm.pose + p * kPoseSm, px, py, pz, ou, ov, our, mono, K, sm.xp + hi * 6, hw);
                    W.H[lane * kHs] = hw[0]; W.H[lane * kHs + 1] = hw[1]; W.H[lane * kHs + 2] = hw[2];
                }
                {   // scatter the loaded sums: global entry g = 9 l + q  ->  W.lm[12 l + q]
                    int g = lane;
                    if (g < 9 * T.ntl) W.lm[(g / 9) * 12 + g % 9] = s0;
                    g += 32;
                    if (g < 9 * T.ntl) W.lm[(g / 9) * 12 + g % 9] = s1;
                    g += 32;
                    if (g < 9 * T.ntl) W.lm[(g / 9) * 12 + g % 9] = s2;
                }
                __syncwarp();
                if (lane < T.ntl * 3) {
                    const int l = lane / 3, q = lane - l * 3;
                    double s = 0.0;
                    for (int e = W.lmoff[l]; e < W.lmoff[l + 1]; ++e) s += W.H[e * kHs + q];
                    W.lm[l * 12 + 9 + q] = s;
                }
                __syncwarp();
            } else {
            double hl[12];
#pragma unroll
            for (int q = 0; q < 12; ++q) hl[q] = 0.0;
            if (lane < T.ne) {
                p = pw & kPoseMask;
                mono = (pw & kMonoBit) != 0;
                lmfree = (lf & kInHessian) != 0;
                act = !(pw & kCulledBit) && !((lf & kFixed) && (sm.pflag[p] & kFixed));
                if (act && lmfree) {
                    const int hi = sm.hidx[p];
                    upd_edge_terms(sm.pose + p * kPoseSm, px, py, pz, ou, ov, our, mono, K, hi >= 0 ? sm.xp + hi * 6 : nullptr, hl);
                }
#pragma unroll
                for (int q = 0; q < 12; ++q) W.H[lane * kHs + q] = hl[q];
            }
            __syncwarp();
            for (int task = lane; task < T.ntl * 12; task += 32) {
                const int l = task / 12, q = task - l * 12;
                double s = 0.0;
                for (int e = W.lmoff[l]; e < W.lmoff[l + 1]; ++e) s += W.H[e * kHs + q];
                W.lm[task] = s;
            }
            __syncwarp();
            }
            if (lane < T.ne) {
                double np0 = px, np1 = py, np2 = pz;
                if (lmfree) {
                    double xl[3];
                    const double sc = upd_point_step(W.lm + tl * 12, lambda, xl);
                    np0 = px + xl[0]; np1 = py + xl[1]; np2 = pz + xl[2];
                    if (lane == W.lmoff[tl]) {      // first edge of the landmark: owner of the point
                        const size_t gl = (size_t)T.lt + tl;
                        gpointT[3 * gl] = np0; gpointT[3 * gl + 1] = np1; gpointT[3 * gl + 2] = np2;
                        scale_acc += sc;
                    }
                }
                if (act) {
                    double r0, r1, r2, rho, wgt;
                    edge_residual(sm.poseT + p * kPoseSm, np0, np1, np2, ou, ov, our, mono, K, r0, r1, r2);
                    huber((r0 * r0 + r1 * r1 + r2 * r2) * K.inv_pv, K.delta, rho, wgt);
                    chi_acc += rho;
                }
            }
            __syncwarp();
        } else {
            // ---- one landmark with more than 32 edges: rounds of 32, sums by shuffle
            const int gl = T.lt;
            const uint8_t lf = B.lm_flags[gl];
            const bool lmfree = (lf & kInHessian) != 0;
            const double px = gpoint[3 * (size_t)gl], py = gpoint[3 * (size_t)gl + 1], pz = gpoint[3 * (size_t)gl + 2];
            double ls[12];
#pragma unroll
            for (int q = 0; q < 12; ++q) ls[q] = 0.0;
            if (lmfree) {
                for (int b0 = 0; b0 < T.ne; b0 += 32) {
                    double hl[12];
#pragma unroll
                    for (int q = 0; q < 12; ++q) hl[q] = 0.0;
                    if (b0 + lane < T.ne) {
                        const int e = T.e0 + b0 + lane;
                        const int pw = B.edge_pose[e];
                        const int p = pw & kPoseMask;
                        if (!(pw & kCulledBit) && !((lf & kFixed) && (B.pose_flags[pose_off + p] & kFixed))) {
                            const int hi = sm.hidx[p];
                            upd_edge_terms(sm.pose + p * kPoseSm, px, py, pz, B.obs_u[e], B.obs_v[e], B.obs_r[e], (pw & kMonoBit) != 0, K,
                                           hi >= 0 ? sm.xp + hi * 6 : nullptr, hl);
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 12; ++q) {
                        double v = hl[q];
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                        ls[q] += v;
                    }
                }
            }
            double np0 = px, np1 = py, np2 = pz;
            if (lmfree) {
                double xl[3];
                const double sc = upd_point_step(ls, lambda, xl);
                np0 = px + xl[0]; np1 = py + xl[1]; np2 = pz + xl[2];
                if (lane == 0) {
                    gpointT[3 * (size_t)gl] = np0; gpointT[3 * (size_t)gl + 1] = np1; gpointT[3 * (size_t)gl + 2] = np2;
                    scale_acc += sc;
                }
            }
            for (int b0 = 0; b0 < T.ne; b0 += 32) {
                if (b0 + lane < T.ne) {
                    const int e = T.e0 + b0 + lane;
                    const int pw = B.edge_pose[e];
                    const int p = pw & kPoseMask;
                    if (!(pw & kCulledBit) && !((lf & kFixed) && (B.pose_flags[pose_off + p] & kFixed))) {
                        double r0, r1, r2, rho, wgt;
                        edge_residual(sm.poseT + p * kPoseSm, np0, np1, np2, B.obs_u[e], B.obs_v[e], B.obs_r[e], (pw & kMonoBit) != 0,
                                      K, r0, r1, r2);
                        huber((r0 * r0 + r1 * r1 + r2 * r2) * K.inv_pv, K.delta, rho, wgt);
                        chi_acc += rho;
                    }
                }
            }
        }
    }
    const double chi = block_sum(chi_acc, sm.red);
    const double sc = block_sum(scale_acc, sm.red);
    // The CTA that finishes a window's trial last runs the LM controller (one launch less per trial).  Nobody reads the LM
    // state of the window any more at that point: every other CTA of the window has taken its ticket at its very end.
    __shared__ int s_last;
    if (tid == 0) {
        B.part2[2 * (size_t)blockIdx.x] = chi; B.part2[2 * (size_t)blockIdx.x + 1] = sc;
        __threadfence();
        s_last = atomicAdd(&B.ctl_count[ck.win], 1) == wd.n_chunks - 1;
    }
    __syncthreads();
    if (s_last && warp == 0) {
        __threadfence();
        if (lane == 0) B.ctl_count[ck.win] = 0;
        control_step(B, ck.win, lane);
    }
}

// ------------------------------------------------------------------------------------------------
// k_init: start of a pass — robust chi2 of the accepted state (computeActiveErrors + activeRobustChi2) and the diagonal
// of H for OptimizationAlgorithmLevenberg::computeLambdaInit (max over poses AND landmarks).  Same warp tiles as k_update.
// Per-pose diagonal sums stay deterministic without atomics: every warp owns a private [pose][6] table; inside a tile the
// landmarks are added one after the other (the poses of one landmark are distinct, so the lanes of a round never collide).
// Output per chunk (layout k_control_init reads): [F x 6 diag(H_pp) | chi2 | max |diag H_ll|].
// ------------------------------------------------------------------------------------------------
struct InitSmem {
    double pose[kMaxSmallPoses * kPoseSm];
    double pd[kUpdWarps][kMaxSmallPoses * 6];
    double H[kUpdWarps][32 * 4];
    double red[32];
    int lmoff[kUpdWarps][kWtLm + 1];
    int hidx[kMaxSmallPoses];
    unsigned char pflag[kMaxSmallPoses];
};

__global__ void __launch_bounds__(kUpdThreads, 2) k_init(Batch B) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    InitSmem &sm = *reinterpret_cast<InitSmem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const Chunk ck = B.chunks[blockIdx.x];
    const WinDesc &wd = B.win[ck.win];
    const LMState &st = B.st[ck.win];
    if (st.done) return;
    const int cur = st.cur, F = st.F;
    const Intr K = load_intr(wd);
    const int pose_off = wd.pose_off, n_pose = wd.n_pose;
    const double *__restrict__ gpose = B.pose + ((size_t)cur * B.tot_pose + pose_off) * kPoseStride;
    const double *__restrict__ gpoint = B.point + (size_t)cur * B.tot_point * 3;
    const int tile0 = B.chunk_wtile_off[blockIdx.x], ntiles = B.chunk_wtile_off[blockIdx.x + 1] - tile0;
    const Tile *__restrict__ tiles = B.wtiles + tile0;
    for (int i = tid; i < n_pose * kPoseStride; i += kUpdThreads) sm.pose[(i >> 4) * kPoseSm + (i & 15)] = gpose[i];
    for (int i = tid; i < n_pose; i += kUpdThreads) { sm.hidx[i] = B.pose_hidx[pose_off + i]; sm.pflag[i] = B.pose_flags[pose_off + i]; }
    for (int i = tid; i < kUpdWarps * kMaxSmallPoses * 6; i += kUpdThreads) (&sm.pd[0][0])[i] = 0.0;
    __syncthreads();
    double *pd = sm.pd[warp], *Hs = sm.H[warp];
    int *lmoff = sm.lmoff[warp];
    double chi_acc = 0.0, maxd = 0.0;

    Tile Tnext{0, 0, 0, 0};
    if (warp < ntiles) Tnext = tiles[warp];
    for (int t = warp; t < ntiles; t += kUpdWarps) {
        const Tile T = Tnext;
        if (t + kUpdWarps < ntiles) Tnext = tiles[t + kUpdWarps];
        const int lf_l = (lane < T.ntl) ? (int)B.lm_flags[T.lt + lane] : 0;
        const double pt = (lane < 3 * T.ntl) ? gpoint[3 * (size_t)T.lt + lane] : 0.0;
        double hacc[3] = {0.0, 0.0, 0.0};     // long landmarks (> 32 edges): H_ll diagonal over the rounds
        for (int b0 = 0; b0 < T.ne; b0 += 32) {
            // (tiles of several landmarks have one round; a landmark with more than 32 edges is a tile of its own)
            const int off_l = (lane <= T.ntl) ? B.lm_edge_off[T.lt + lane] - T.e0 - b0 : 0x7fff;
            int pw = 0;
            double ou = 0, ov = 0, our = 0;
            const bool in = b0 + lane < T.ne;
            if (in) {
                const int e = T.e0 + b0 + lane;
                pw = B.edge_pose[e];
                ou = B.obs_u[e]; ov = B.obs_v[e]; our = B.obs_r[e];
            }
            if (lane <= T.ntl) lmoff[lane] = max(min(off_l, 32), 0);
            int tl = 0;
#pragma unroll
            for (int l = 1; l < kWtLm; ++l) {
                const int v = __shfl_sync(0xffffffffu, off_l, l);
                tl += (l < T.ntl && v <= lane) ? 1 : 0;
            }
            const int lf = __shfl_sync(0xffffffffu, lf_l, tl);
            const double px = __shfl_sync(0xffffffffu, pt, 3 * tl), py = __shfl_sync(0xffffffffu, pt, 3 * tl + 1),
                         pz = __shfl_sync(0xffffffffu, pt, 3 * tl + 2);
            bool act = false;
            int hi = -1;
            double dv[6] = {0, 0, 0, 0, 0, 0}, h3[3] = {0.0, 0.0, 0.0};
            if (in) {
                const int p = pw & kPoseMask;
                act = !(pw & kCulledBit) && !((lf & kFixed) && (sm.pflag[p] & kFixed));
                if (act) {
                    EdgeLin lin;
                    edge_linearize(sm.pose + p * kPoseSm, px, py, pz, ou, ov, our, (pw & kMonoBit) != 0, K, lin);
                    chi_acc += lin.rho;
                    const double wo = lin.w * K.inv_pv;
                    hi = sm.hidx[p];
                    if (lf & kInHessian) {
                        const double *J = lin.Jl;
                        h3[0] = wo * (J[0] * J[0] + J[3] * J[3] + J[6] * J[6]);
                        h3[1] = wo * (J[1] * J[1] + J[4] * J[4] + J[7] * J[7]);
                        h3[2] = wo * (J[2] * J[2] + J[5] * J[5] + J[8] * J[8]);
                    }
#pragma unroll
                    for (int a = 0; a < 6; ++a)
                        dv[a] = wo * (lin.Jp[a] * lin.Jp[a] + lin.Jp[6 + a] * lin.Jp[6 + a] + lin.Jp[12 + a] * lin.Jp[12 + a]);
                }
            }
            Hs[lane * 4] = h3[0]; Hs[lane * 4 + 1] = h3[1]; Hs[lane * 4 + 2] = h3[2];
            const int lf_task = __shfl_sync(0xffffffffu, lf_l, min(lane / 3, kWtLm - 1));   // flags of the landmark lane / 3 owns below
            __syncwarp();
            // diag(H_ll) per landmark in edge order
            if (lane < T.ntl * 3) {
                const int l = lane / 3, q = lane - l * 3;
                double sacc = 0.0;
                for (int e = lmoff[l]; e < lmoff[l + 1]; ++e) sacc += Hs[e * 4 + q];
                if (T.ne > 32) { hacc[q] += sacc; sacc = hacc[q]; }
                if (lf_task & kInHessian) maxd = fmax(maxd, fabs(sacc));
            }
            // diag(H_pp) per pose: one landmark per round
            for (int l = 0; l < T.ntl; ++l) {
                if (in && tl == l && act && hi >= 0) {
#pragma unroll
                    for (int a = 0; a < 6; ++a) pd[hi * 6 + a] += dv[a];
                }
                __syncwarp();
            }
        }
    }
    __syncthreads();
    double *part = B.part + wd.part_off + (size_t)(blockIdx.x - wd.chunk_off) * wd.part_stride;
    for (int task = tid; task < F * 6; task += kUpdThreads) {
        double sacc = 0.0;
#pragma unroll
        for (int w = 0; w < kUpdWarps; ++w) sacc += sm.pd[w][task];
        part[task] = sacc;
    }
    const double chi = block_sum(chi_acc, sm.red);
    const double md = block_max(maxd, sm.red);
    if (tid == 0) { part[F * 6] = chi; part[F * 6 + 1] = md; }
}

}  // namespace visfs

// ba_build_ws.cuh — warp-specialised build kernel (linearise + Hessian blocks + Schur partials), F <= 22.
//
// Same arithmetic and the same tile walk as k_build<MODE_BUILD> (ba_kernels.cuh), restructured for the SM:
//   * 5 PRODUCER warps (160 threads, one edge each): stage A (residual, Jacobians, Huber, per-landmark
//     H_ll / b_l, damped 3x3 inverse) and stage B (W = H_pl block, Yn = -W Dinv, H_pp_e, b_p_e, g) into one
//     of two shared-memory stages.
//   * 8 CONSUMER warps (256 threads): stage C.  Every block (i <= j) of the reduced camera system has a fixed
//     owner thread that keeps its 36 entries in registers for the whole chunk and adds Yn_i W_j^T with
//     three chained DFMAs per entry; per-pose sums (H_pp, g, b_p) have fixed owners as well.
//   The two roles overlap on different tiles through full / empty named barriers (bar.arrive / bar.sync),
//   so neither the 72 accumulator registers nor the linearisation temporaries are ever live in the same
//   thread (v1 spilled 900 B / thread to L2), and the FP64 pipe is fed by both roles at once.
//   * Thread-block clusters: when a window is split over several chunks, the CTAs of a cluster add their
//     partial systems through distributed shared memory in rank order and write ONE partial per cluster,
//     so k_solve reads 8x fewer partials.  No atomics; every sum has a fixed order.
#pragma once
#include <cooperative_groups.h>
#include "ba_kernels.cuh"

namespace visfs {
namespace ws {

namespace cg = cooperative_groups;

constexpr int kEdgeThreads = kTileEdges;      // 160
constexpr int kPairThreads = 224;              // 7 warps: 12 warps per CTA = 3 per SM sub-partition (16 K registers each)
constexpr int kThreadsWs = kEdgeThreads + kPairThreads;   // 416
constexpr int kMaxPosesWs = 20;               // F (F + 1) / 2 <= 210 block owners

enum { BAR_PROD = 1, BAR_CONS = 2, BAR_FULL = 3, BAR_EMPTY = 5 };

// (whole warps take every barrier; __syncwarp reconverges the warp after the divergent per-edge code)
__device__ __forceinline__ void bar_sync(int id, int n) { __syncwarp(); asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { __syncwarp(); asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

struct Stage {
    double W[kTileEdges * 18];
    double Yn[kTileEdges * 18];
    double H[kTileEdges * kHStride];           // H_pp_e (21, upper) | g (6) | b_p (6); stage A scratch before that
    double lm[kTileLm * 12];                   // Dinv(6) db(3) bl(3)
    short slot[kTileLm * kMaxSmallPoses];
    int ntl, next_lt, pad0, pad1;
};

struct Smem {
    double pose[kMaxSmallPoses * kPoseStride];
    Stage st[2];
    double pacc[kMaxSmallPoses * kHStride];
    int hidx[kMaxSmallPoses];
    int lmoff[kTileLm + 1];
};
constexpr int kStageDoubles = (int)(sizeof(Stage) / sizeof(double));

// number of doubles of one partial system in the "all pairs" layout
__host__ __device__ __forceinline__ int part_len(int F) { return F * (F + 1) / 2 * 36 + F * kHStride; }

__global__ void __maxnreg__(168) k_build_ws(Batch B, int cluster_size) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    const int tid = threadIdx.x;
    const Chunk ck = B.chunks[blockIdx.x];
    const WinDesc &wd = B.win[ck.win];
    const LMState &st = B.st[ck.win];
    if (st.done) return;   // uniform over the cluster: all its chunks belong to one window
    const int cur = st.cur;
    const int F = st.F;
    const double lambda = (wd.trust == 0) ? (B.dbg_lambda >= 0.0 ? B.dbg_lambda : st.lambda) : 0.0;
    const Intr K = load_intr(wd);
    const int pose_off = wd.pose_off, n_pose = wd.n_pose;
    const double *__restrict__ gpose = B.pose + ((size_t)cur * B.tot_pose + pose_off) * kPoseStride;
    const double *__restrict__ gpoint = B.point + (size_t)cur * B.tot_point * 3;

    for (int i = tid; i < n_pose * kPoseStride; i += kThreadsWs) sm.pose[i] = gpose[i];
    for (int i = tid; i < n_pose; i += kThreadsWs) sm.hidx[i] = B.pose_hidx[pose_off + i];
    for (int i = tid; i < kMaxSmallPoses * kHStride; i += kThreadsWs) sm.pacc[i] = 0.0;
    __syncthreads();

    const int npairs = F * (F + 1) / 2;
    int G = 1;
    if (npairs > 0) {
        G = kPairThreads / npairs;
        if (G > kTileLm) G = kTileLm;
        const int cap = 1 + kStageDoubles / (npairs * 36);
        if (G > cap) G = cap;
        if (G < 1) G = 1;
    }
    const bool producer = tid < kEdgeThreads;
    const int ctid = tid - kEdgeThreads;
    const int grp = (!producer && npairs > 0) ? ctid / npairs : 0;
    const int pt = (!producer && npairs > 0) ? ctid % npairs : 0;
    double *scratch = reinterpret_cast<double *>(&sm.st[0]);   // epilogue aliases: group partials ...
    double *vec = reinterpret_cast<double *>(&sm.st[1]);       // ... and this CTA's partial system

    if (producer) {
        // ======================================================================= producers: stages A and B
        int t = 0, cnt_prev = 0;
        for (int lt = ck.lm0; lt < ck.lm1; ++t) {
            Stage &S = sm.st[t & 1];
            if (t >= 2) bar_sync(BAR_EMPTY + (t & 1), kThreadsWs);
            const int e0 = B.lm_edge_off[lt];
            const int lmax = min(lt + kTileLm, ck.lm1);
            int l1;
            {   // tile end: try the previous tile's landmark count first (uniform degree), else binary search
                const int g = min(lt + max(cnt_prev, 1), lmax);
                const int og = B.lm_edge_off[g] - e0;
                const int og1 = (g < lmax) ? B.lm_edge_off[g + 1] - e0 : (kTileEdges + 1);
                l1 = (og <= kTileEdges && og1 > kTileEdges) ? g : tile_end(B.lm_edge_off, lt, lmax, e0);
            }
            const int ne = min(B.lm_edge_off[l1] - e0, kTileEdges);
            const int ntl = l1 - lt;
            cnt_prev = ntl;
            for (int i = tid; i < ntl * kMaxSmallPoses; i += kEdgeThreads) S.slot[i] = -1;
            if (tid <= ntl) sm.lmoff[tid] = min(B.lm_edge_off[lt + tid] - e0, kTileEdges);

            EdgeLin lin;
            bool act = false, lmfree = false;
            int tl = 0, p = 0;
            if (tid < ne) {
                const int e = e0 + tid;
                const int pw = B.edge_pose[e];
                p = pw & kPoseMask;
                const int gl = wd.point_off + B.edge_point[e];
                tl = gl - lt;
                const uint8_t lf = B.lm_flags[gl];
                const uint8_t pf = B.pose_flags[pose_off + p];
                act = !(pw & kCulledBit) && !((lf & kFixed) && (pf & kFixed));
                lmfree = (lf & kInHessian) != 0;
                double *hl = S.H + tid * kHStride;
                if (act) {
                    const double px = gpoint[3 * (size_t)gl], py = gpoint[3 * (size_t)gl + 1], pz = gpoint[3 * (size_t)gl + 2];
                    edge_linearize(sm.pose + p * kPoseStride, px, py, pz, B.obs_u[e], B.obs_v[e], B.obs_r[e],
                                   (pw & kMonoBit) != 0, K, lin);
                }
                if (act && lmfree) {
                    const double wo = lin.w * K.inv_pv;
                    const double *J = lin.Jl;
                    hl[0] = wo * fma(J[0], J[0], fma(J[3], J[3], J[6] * J[6]));
                    hl[1] = wo * fma(J[0], J[1], fma(J[3], J[4], J[6] * J[7]));
                    hl[2] = wo * fma(J[0], J[2], fma(J[3], J[5], J[6] * J[8]));
                    hl[3] = wo * fma(J[1], J[1], fma(J[4], J[4], J[7] * J[7]));
                    hl[4] = wo * fma(J[1], J[2], fma(J[4], J[5], J[7] * J[8]));
                    hl[5] = wo * fma(J[2], J[2], fma(J[5], J[5], J[8] * J[8]));
                    hl[6] = -wo * fma(J[0], lin.r[0], fma(J[3], lin.r[1], J[6] * lin.r[2]));
                    hl[7] = -wo * fma(J[1], lin.r[0], fma(J[4], lin.r[1], J[7] * lin.r[2]));
                    hl[8] = -wo * fma(J[2], lin.r[0], fma(J[5], lin.r[1], J[8] * lin.r[2]));
                } else {
#pragma unroll
                    for (int q = 0; q < 9; ++q) hl[q] = 0.0;
                }
            }
            bar_sync(BAR_PROD, kEdgeThreads);
            if (tid < ntl) {
                double A[6] = {0, 0, 0, 0, 0, 0}, bl[3] = {0, 0, 0};
                for (int s = sm.lmoff[tid]; s < sm.lmoff[tid + 1]; ++s) {
                    const double *hl = S.H + s * kHStride;
#pragma unroll
                    for (int q = 0; q < 6; ++q) A[q] += hl[q];
                    bl[0] += hl[6]; bl[1] += hl[7]; bl[2] += hl[8];
                }
                double *o = S.lm + tid * 12;
                if (B.lm_flags[lt + tid] & kInHessian) {
                    A[0] += lambda; A[3] += lambda; A[5] += lambda;
                    inv_sym3(A, o);
                    sym3_mul(o, bl, o + 6);
                    o[9] = bl[0]; o[10] = bl[1]; o[11] = bl[2];
                } else {
#pragma unroll
                    for (int q = 0; q < 12; ++q) o[q] = 0.0;
                }
            }
            bar_sync(BAR_PROD, kEdgeThreads);
            if (tid < ne && act) {
                const int hi = sm.hidx[p];
                if (hi >= 0) {
                    const double wo = lin.w * K.inv_pv;
                    double *hs = S.H + tid * kHStride;
                    double *ws = S.W + tid * 18, *ys = S.Yn + tid * 18;
                    const double *lm = S.lm + tl * 12;
                    double Wm[18];
                    if (lmfree) {
                        double Aj[9];
#pragma unroll
                        for (int q = 0; q < 9; ++q) Aj[q] = wo * lin.Jl[q];
#pragma unroll
                        for (int a = 0; a < 6; ++a)
#pragma unroll
                            for (int c = 0; c < 3; ++c)
                                Wm[a * 3 + c] = fma(lin.Jp[a], Aj[c], fma(lin.Jp[6 + a], Aj[3 + c], lin.Jp[12 + a] * Aj[6 + c]));
#pragma unroll
                        for (int a = 0; a < 6; ++a) {
                            ys[a * 3 + 0] = -fma(Wm[a * 3], lm[0], fma(Wm[a * 3 + 1], lm[1], Wm[a * 3 + 2] * lm[2]));
                            ys[a * 3 + 1] = -fma(Wm[a * 3], lm[1], fma(Wm[a * 3 + 1], lm[3], Wm[a * 3 + 2] * lm[4]));
                            ys[a * 3 + 2] = -fma(Wm[a * 3], lm[2], fma(Wm[a * 3 + 1], lm[4], Wm[a * 3 + 2] * lm[5]));
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < 18; ++q) { Wm[q] = 0.0; ys[q] = 0.0; }
                    }
#pragma unroll
                    for (int q = 0; q < 18; ++q) ws[q] = Wm[q];
                    const double wr0 = wo * lin.r[0], wr1 = wo * lin.r[1], wr2 = wo * lin.r[2];
#pragma unroll
                    for (int a = 0; a < 6; ++a) {
                        const double bp = -fma(lin.Jp[a], wr0, fma(lin.Jp[6 + a], wr1, lin.Jp[12 + a] * wr2));
                        hs[27 + a] = bp;
                        hs[21 + a] = bp - fma(Wm[a * 3], lm[6], fma(Wm[a * 3 + 1], lm[7], Wm[a * 3 + 2] * lm[8]));
#pragma unroll
                        for (int c = a; c < 6; ++c)
                            hs[hd_index(a, c)] = wo * fma(lin.Jp[a], lin.Jp[c], fma(lin.Jp[6 + a], lin.Jp[6 + c], lin.Jp[12 + a] * lin.Jp[12 + c]));
                    }
                    S.slot[tl * kMaxSmallPoses + hi] = (short)tid;
                }
            }
            if (tid == 0) { S.ntl = ntl; S.next_lt = l1; }
            __threadfence_block();
            bar_arrive(BAR_FULL + (t & 1), kThreadsWs);
            lt = l1;
        }
        for (int tt = max(0, t - 2); tt < t; ++tt) bar_sync(BAR_EMPTY + (tt & 1), kThreadsWs);   // join the consumers
    } else {
        // ======================================================================= consumers: stage C
        double acc[36];   // lives only in the consumer branch: never competes with the producers' registers
#pragma unroll
        for (int q = 0; q < 36; ++q) acc[q] = 0.0;
        int pi = -1, pj = -1;
        if (grp < G && pt < npairs) {
            int i = 0, base = 0;
            while (base + (F - i) <= pt) { base += F - i; ++i; }
            pi = i; pj = i + (pt - base);
        }
        int t = 0;
        for (int lt = ck.lm0; lt < ck.lm1; ++t) {
            const Stage &S = sm.st[t & 1];
            bar_sync(BAR_FULL + (t & 1), kThreadsWs);
            const int ntl = S.ntl;
            for (int task = ctid; task < F * kHStride; task += kPairThreads) {
                const int i = task / kHStride, k = task - i * kHStride;
                double s = 0.0;
                for (int q = 0; q < ntl; ++q) {
                    const int sl = S.slot[q * kMaxSmallPoses + i];
                    if (sl >= 0) s += S.H[sl * kHStride + k];
                }
                sm.pacc[task] += s;
            }
            if (pi >= 0) {
                for (int q = grp; q < ntl; q += G) {
                    const int si = S.slot[q * kMaxSmallPoses + pi];
                    const int sj = S.slot[q * kMaxSmallPoses + pj];
                    if (si < 0 || sj < 0) continue;
                    double Wj[18];
                    const double2 *wp = reinterpret_cast<const double2 *>(S.W + sj * 18);
#pragma unroll
                    for (int u = 0; u < 9; ++u) { const double2 v = wp[u]; Wj[2 * u] = v.x; Wj[2 * u + 1] = v.y; }
                    const double *yp = S.Yn + si * 18;
#pragma unroll
                    for (int a = 0; a < 6; ++a) {
                        const double y0 = yp[a * 3], y1 = yp[a * 3 + 1], y2 = yp[a * 3 + 2];
#pragma unroll
                        for (int c = 0; c < 6; ++c)
                            acc[a * 6 + c] = fma(y0, Wj[c * 3], fma(y1, Wj[c * 3 + 1], fma(y2, Wj[c * 3 + 2], acc[a * 6 + c])));
                    }
                }
            }
            const int l1 = S.next_lt;
            bar_arrive(BAR_EMPTY + (t & 1), kThreadsWs);
            lt = l1;
        }
        // groups -> one block per owner (consumer-only barriers; the stages are no longer written)
        bar_sync(BAR_CONS, kPairThreads);
        if (G > 1 && grp > 0 && grp < G && pt < npairs) {
            double *dst = scratch + ((size_t)(grp - 1) * npairs + pt) * 36;
#pragma unroll
            for (int q = 0; q < 36; ++q) dst[q] = acc[q];
        }
        bar_sync(BAR_CONS, kPairThreads);
        if (grp == 0 && pt < npairs) {
            for (int g2 = 1; g2 < G; ++g2) {
                const double *src = scratch + ((size_t)(g2 - 1) * npairs + pt) * 36;
#pragma unroll
                for (int q = 0; q < 36; ++q) acc[q] += src[q];
            }
#pragma unroll
            for (int q = 0; q < 36; ++q) vec[pt * 36 + q] = acc[q];
        }
    }
    __syncthreads();

    // ---- epilogue: one vector per CTA -> (cluster sum through DSMEM) -> global partial
    const int offd = npairs * 36;
    const int NP = offd + F * kHStride;
    for (int task = tid; task < F * kHStride; task += kThreadsWs) vec[offd + task] = sm.pacc[task];
    __syncthreads();
    double *part = B.part + wd.part_off + (size_t)((blockIdx.x - wd.chunk_off) / cluster_size) * wd.part_stride;
    if (cluster_size == 1) {
        for (int idx = tid; idx < NP; idx += kThreadsWs) part[idx] = vec[idx];
    } else {
        cg::cluster_group cluster = cg::this_cluster();
        cluster.sync();
        const int rank = (int)cluster.block_rank();
        for (int idx = rank * kThreadsWs + tid; idx < NP; idx += kThreadsWs * cluster_size) {
            double s = 0.0;
            for (int r = 0; r < cluster_size; ++r) s += cluster.map_shared_rank(vec, r)[idx];
            part[idx] = s;
        }
        cluster.sync();
    }
}

}  // namespace ws
}  // namespace visfs

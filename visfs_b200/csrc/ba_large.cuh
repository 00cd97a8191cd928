// ba_large.cuh — the large-window path: windows with more free poses than the shared-memory path holds
// (P > kMaxSmallPoses) and the rank-local part of a landmark-partitioned global BA (BASELINE configs C4 / C5).
//
// The reduced camera system S lives in HBM / L2 as a BLOCK SKYLINE (lower triangle, 6x6 blocks): row r keeps
// the blocks sky_first[r] .. r.  The envelope is built on the device from the edge list once per pass
// (g2o BlockSolver::buildStructure walks the same landmark -> pose pairs) and contains the whole fill-in of
// the Cholesky factor, so the factorisation runs in place.  For a partitioned problem every rank builds the
// envelope of its own landmarks, the ranks agree on it with one integer MIN allreduce, and the per-trial
// partial systems are summed with ONE ncclAllReduce over the contiguous reduce buffer
//      red = [ skyline blocks (n_sky x 36) | reduced rhs g (6F) | raw b_p (6F) ]
//
//   k_build_large   one warp per landmark, one lane per edge: linearise, H_ll / b_l by shuffle reduction,
//                   damped 3x3 inverse, W / Yn staged in the warp's shared-memory slice, then the lanes sweep the
//                   (pose pair, entry) items of Yn_a W_b^T and add them into the skyline with red.global.add.f64
//                   (consecutive lanes -> consecutive addresses of one block).  H_pp, g and b_p likewise.
//   k_solve_large   one CTA: right-looking block-skyline Cholesky (the forward substitution rides along as an
//                   extra right-hand-side row), row-oriented back-substitution, CameraPose::update.
//   k_update_large  one warp per landmark: back-substitution of the landmark, point oplus, chi2 of the trial.
#pragma once
#include <cooperative_groups.h>
#include "ba_kernels.cuh"

namespace visfs {
namespace lg {

namespace cg = cooperative_groups;

constexpr int kWarpsL = 8;
constexpr int kThreadsL = kWarpsL * 32;
constexpr int kMaxDegL = kMaxDegLarge;    // landmark degree limit of the large path (one lane per edge)
constexpr int kSolveThreadsL = 512;

struct WarpStage {
    double W[kMaxDegL * 18];
    double Yn[kMaxDegL * 18];
    long long base[kMaxDegL];            // sky_off[h] - sky_first[h] of the edge's pose row (block units)
    int hi[kMaxDegL];                    // hessian index of the edge's pose, -1: not part of the Schur products
};
struct BuildSmemL { WarpStage w[kWarpsL]; double red[32]; };

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// element (row c, col r) of the lower block (hb, ha), ha <= hb
__device__ __forceinline__ size_t sky_index(long long base_b, int ha, int c, int r) {
    return (size_t)(base_b + ha) * 36 + (size_t)(c * 6 + r);
}

// ------------------------------------------------------------------------------------------------
// structure: envelope of the reduced camera system
// ------------------------------------------------------------------------------------------------
__global__ void k_sky_init(Batch B) {
    const int F = B.st[0].F;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < F; i += gridDim.x * blockDim.x) B.sky_first[i] = i;
}

// per landmark in the Hessian: every pose (hessian index) it touches through ANY edge (g2o walks v->edges(),
// level-1 edges included) reaches back to the smallest such index
__global__ void k_sky_first(Batch B) {
    const WinDesc &wd = B.win[0];
    if (B.st[0].status != 0) return;
    for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < wd.n_point; l += gridDim.x * blockDim.x) {
        if (!(B.lm_flags[l] & kInHessian)) continue;
        const int e0 = B.lm_edge_off[l], e1 = B.lm_edge_off[l + 1];
        int mn = 0x7fffffff;
        for (int e = e0; e < e1; ++e) {
            const int hi = B.pose_hidx[B.edge_pose[e] & kPoseMask];
            if (hi >= 0) mn = min(mn, hi);
        }
        if (mn == 0x7fffffff) continue;
        for (int e = e0; e < e1; ++e) {
            const int hi = B.pose_hidx[B.edge_pose[e] & kPoseMask];
            if (hi > mn && B.sky_first[hi] > mn) atomicMin(&B.sky_first[hi], mn);
        }
    }
}

// one CTA: row offsets (exclusive scan of the row lengths), column counts of the envelope, n_sky -> info[0]
__global__ void k_sky_layout(Batch B, int *col_cnt, long long *info) {
    __shared__ long long s_carry;
    __shared__ long long s_w[32];
    const int F = B.st[0].F;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int i0 = 0; i0 < F; i0 += blockDim.x) {
        const int i = i0 + tid;
        const long long len = (i < F) ? (long long)(i - B.sky_first[i] + 1) : 0;
        long long v = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const long long t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
        if (lane == 31) s_w[wid] = v;
        __syncthreads();
        long long pre = s_carry;
        for (int k = 0; k < wid; ++k) pre += s_w[k];
        if (i < F) B.sky_off[i] = pre + v - len;
        __syncthreads();
        if (tid == 0) { long long t = 0; for (int k = 0; k < nw; ++k) t += s_w[k]; s_carry += t; }
        __syncthreads();
    }
    if (tid == 0) { B.sky_off[F] = s_carry; info[0] = s_carry; info[1] = F; }
    for (int i = tid; i <= F; i += blockDim.x) col_cnt[i] = 0;
}

// column structure of the envelope: rows r > k with sky_first[r] <= k  (count, then fill after a scan)
__global__ void k_col_count(Batch B, int *col_cnt) {
    const int F = B.st[0].F;
    for (int r = blockIdx.x; r < F; r += gridDim.x)
        for (int k = B.sky_first[r] + threadIdx.x; k < r; k += blockDim.x) atomicAdd(&col_cnt[k], 1);
}
__global__ void k_col_scan(Batch B, const int *col_cnt, long long *info) {   // one CTA; info[2] = widest front
    __shared__ int s_carry;
    __shared__ int s_max;
    if (threadIdx.x == 0) s_max = 0;
    __shared__ int s_w[32];
    const int F = B.st[0].F;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int i0 = 0; i0 < F; i0 += blockDim.x) {
        const int i = i0 + tid;
        const int len = (i < F) ? col_cnt[i] : 0;
        if (len > 0) atomicMax(&s_max, len);
        int v = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
        if (lane == 31) s_w[wid] = v;
        __syncthreads();
        int pre = s_carry;
        for (int k = 0; k < wid; ++k) pre += s_w[k];
        if (i < F) B.col_ptr[i] = pre + v - len;
        __syncthreads();
        if (tid == 0) { int t = 0; for (int k = 0; k < nw; ++k) t += s_w[k]; s_carry += t; }
        __syncthreads();
    }
    if (tid == 0) { B.col_ptr[F] = s_carry; info[2] = s_max; }
}
// rows of column k in ascending order: row r lands at position (number of rows r' < r with first[r'] <= k < r').
// One thread per (r, k) pair would need a rank; instead every column is filled by one warp scanning the candidate
// rows k+1 .. F-1 in order (ballot compaction) — F^2 / 32 warp steps in total, once per pass.
__global__ void k_col_fill(Batch B) {
    const int F = B.st[0].F;
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (int k = gw; k < F; k += nw) {
        int pos = B.col_ptr[k];
        const int end = B.col_ptr[k + 1];
        for (int r0 = k + 1; r0 < F && pos < end; r0 += 32) {
            const int r = r0 + lane;
            const bool in = (r < F) && (B.sky_first[r] <= k);
            const unsigned m = __ballot_sync(0xffffffffu, in);
            if (in) B.col_rows[pos + __popc(m & ((1u << lane) - 1u))] = r;
            pos += __popc(m);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_build_large<INIT>
//   INIT : robust chi2 of the accepted state, diag(H_pp) into hdiag (atomics), max |diag H_ll|
//   BUILD: the whole damped Schur system into the reduce buffer
// ------------------------------------------------------------------------------------------------
template <bool INIT>
__global__ void __launch_bounds__(kThreadsL) k_build_large(Batch B) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BuildSmemL &sm = *reinterpret_cast<BuildSmemL *>(smem_raw);
    const WinDesc &wd = B.win[0];
    const LMState &st = B.st[0];
    if (st.done) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cur = st.cur;
    const double lambda = (!INIT && wd.trust == 0) ? (B.dbg_lambda >= 0.0 ? B.dbg_lambda : st.lambda) : 0.0;
    const Intr K = load_intr(wd);
    const double *__restrict__ gpose = B.pose + (size_t)cur * B.tot_pose * kPoseStride;
    const double *__restrict__ gpoint = B.point + (size_t)cur * B.tot_point * 3;
    double *__restrict__ sky = B.red;
    double *__restrict__ gvec = B.red + B.red_g_off;
    double *__restrict__ bpvec = B.red + B.red_bp_off;
    WarpStage &S = sm.w[warp];
    double chi_acc = 0.0, maxd = 0.0;

    for (int l = blockIdx.x * kWarpsL + warp; l < wd.n_point; l += gridDim.x * kWarpsL) {
        const int e0 = B.lm_edge_off[l];
        const int d = min(B.lm_edge_off[l + 1] - e0, kMaxDegL);
        if (d <= 0) continue;
        const uint8_t lf = B.lm_flags[l];
        const bool lmfree = (lf & kInHessian) != 0;
        bool act = false;
        int hi = -1;
        EdgeLin lin;
        if (lane < d) {
            const int e = e0 + lane;
            const int pw = B.edge_pose[e];
            const int p = pw & kPoseMask;
            const uint8_t pf = B.pose_flags[p];
            act = !(pw & kCulledBit) && !((lf & kFixed) && (pf & kFixed));
            if (act) {
                hi = B.pose_hidx[p];
                edge_linearize(gpose + (size_t)p * kPoseStride, gpoint[3 * (size_t)l], gpoint[3 * (size_t)l + 1],
                               gpoint[3 * (size_t)l + 2], B.obs_u[e], B.obs_v[e], B.obs_r[e], (pw & kMonoBit) != 0, K, lin);
                if (INIT) chi_acc += lin.rho;
            }
        }
        const double wo = act ? lin.w * K.inv_pv : 0.0;
        double hl[9];
        if (act && lmfree) {
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
#pragma unroll
        for (int q = 0; q < (INIT ? 6 : 9); ++q) hl[q] = warp_sum(hl[q]);

        if (INIT) {
            if (lmfree) maxd = fmax(maxd, fmax(fabs(hl[0]), fmax(fabs(hl[3]), fabs(hl[5]))));
            if (act && hi >= 0) {
#pragma unroll
                for (int a = 0; a < 6; ++a)
                    atomicAdd(&B.hdiag[6 * (size_t)hi + a],
                              wo * fma(lin.Jp[a], lin.Jp[a], fma(lin.Jp[6 + a], lin.Jp[6 + a], lin.Jp[12 + a] * lin.Jp[12 + a])));
            }
            continue;
        }

        // damped inverse of the landmark block (every lane, redundantly) and Dinv b_l
        double Di[6], db[3] = {0.0, 0.0, 0.0};
        if (lmfree) {
            const double A[6] = {hl[0] + lambda, hl[1], hl[2], hl[3] + lambda, hl[4], hl[5] + lambda};
            const double bl[3] = {hl[6], hl[7], hl[8]};
            inv_sym3(A, Di);
            sym3_mul(Di, bl, db);
        }
        int myhi = -1;
        if (act && hi >= 0) {
            const long long base = B.sky_off[hi] - B.sky_first[hi];
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
                double *ws = S.W + lane * 18, *ys = S.Yn + lane * 18;
#pragma unroll
                for (int a = 0; a < 6; ++a) {
                    ws[a * 3] = Wm[a * 3]; ws[a * 3 + 1] = Wm[a * 3 + 1]; ws[a * 3 + 2] = Wm[a * 3 + 2];
                    ys[a * 3 + 0] = -fma(Wm[a * 3], Di[0], fma(Wm[a * 3 + 1], Di[1], Wm[a * 3 + 2] * Di[2]));
                    ys[a * 3 + 1] = -fma(Wm[a * 3], Di[1], fma(Wm[a * 3 + 1], Di[3], Wm[a * 3 + 2] * Di[4]));
                    ys[a * 3 + 2] = -fma(Wm[a * 3], Di[2], fma(Wm[a * 3 + 1], Di[4], Wm[a * 3 + 2] * Di[5]));
                }
                S.base[lane] = base;
                myhi = hi;
            } else {
#pragma unroll
                for (int q = 0; q < 18; ++q) Wm[q] = 0.0;
            }
            // per-pose sums: H_pp_e into the diagonal block (lower half), g = b_p_e - W Dinv b_l, raw b_p_e
            const double wr0 = wo * lin.r[0], wr1 = wo * lin.r[1], wr2 = wo * lin.r[2];
            double *dblk = sky + (size_t)(base + hi) * 36;
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                const double bp = -fma(lin.Jp[a], wr0, fma(lin.Jp[6 + a], wr1, lin.Jp[12 + a] * wr2));
                atomicAdd(&bpvec[6 * (size_t)hi + a], bp);
                atomicAdd(&gvec[6 * (size_t)hi + a], bp - fma(Wm[a * 3], db[0], fma(Wm[a * 3 + 1], db[1], Wm[a * 3 + 2] * db[2])));
#pragma unroll
                for (int c = a; c < 6; ++c)
                    atomicAdd(&dblk[c * 6 + a],
                              wo * fma(lin.Jp[a], lin.Jp[c], fma(lin.Jp[6 + a], lin.Jp[6 + c], lin.Jp[12 + a] * lin.Jp[12 + c])));
            }
        }
        if (lane < d) S.hi[lane] = myhi;
        __syncwarp();
        if (lmfree) {
            // Schur products: for a <= b (ascending pose => ascending hessian index) block (h_a, h_b) += Yn_a W_b^T,
            // stored as the lower block (h_b, h_a): element (r, c) of the product sits at row c, column r
            for (int a = 0; a < d; ++a) {
                const int ha = S.hi[a];
                if (ha < 0) continue;
                const double *ya = S.Yn + a * 18;
                const int items = (d - a) * 36;
                for (int it = lane; it < items; it += 32) {
                    const int bo = it / 36, q = it - bo * 36;
                    const int b = a + bo;
                    if (S.hi[b] < 0) continue;
                    const int r = q / 6, c = q - r * 6;
                    if (bo == 0 && c < r) continue;            // diagonal block: lower half only
                    const double *wb = S.W + b * 18 + c * 3;
                    const double v = fma(ya[r * 3], wb[0], fma(ya[r * 3 + 1], wb[1], ya[r * 3 + 2] * wb[2]));
                    atomicAdd(&sky[sky_index(S.base[b], ha, c, r)], v);
                }
            }
        }
        __syncwarp();
    }
    if (INIT) {
        const double chi = block_sum(chi_acc, sm.red);
        const double md = block_max(maxd, sm.red);
        if (tid == 0) {
            B.part2[2 * (size_t)blockIdx.x] = chi;
            B.part2[2 * (size_t)blockIdx.x + 1] = md;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_build_large_run: the same build for landmarks of degree <= 10, with the atomics taken out of the inner loop.
// The landmarks are walked in the order of their first pose (lm_order, sorted on the device once per pass), every warp
// takes a contiguous piece of that order, and consecutive landmarks that touch the SAME poses form a run: the Schur
// products of a run are accumulated in registers (a lane owns up to two pose pairs, 2 x 36 accumulators), the per-pose
// sums in the warp's shared-memory slab, and only the end of a run goes to the skyline with red.global.add.f64.  In a map
// whose landmarks are seen by consecutive frames a run is a few hundred landmarks long (C4: ~250).
// ------------------------------------------------------------------------------------------------
constexpr int kRunDeg = 10;                          // landmark degree limit of this kernel: d (d + 1) / 2 <= 64 pose pairs
struct RunWarp {
    double W[kRunDeg * 18];
    double Yn[kRunDeg * 18];
    double PA[kRunDeg * kHStride];                   // per edge slot: H_pp (21, upper) | g (6) | b_p (6), summed over the run
    long long base[kRunDeg];
    int hi[kRunDeg];
    int pad[2];
};
struct RunSmem { RunWarp w[kWarpsL]; };

__global__ void __launch_bounds__(kThreadsL, 1) k_build_large_run(Batch B, const int4 *__restrict__ lm_rec, int n_lm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RunSmem &sm = *reinterpret_cast<RunSmem *>(smem_raw);
    const WinDesc &wd = B.win[0];
    const LMState &st = B.st[0];
    if (st.done) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cur = st.cur;
    const double lambda = (wd.trust == 0) ? (B.dbg_lambda >= 0.0 ? B.dbg_lambda : st.lambda) : 0.0;
    const Intr K = load_intr(wd);
    const double *__restrict__ gpose = B.pose + (size_t)cur * B.tot_pose * kPoseStride;
    const double *__restrict__ gpoint = B.point + (size_t)cur * B.tot_point * 3;
    double *__restrict__ sky = B.red;
    double *__restrict__ gvec = B.red + B.red_g_off;
    double *__restrict__ bpvec = B.red + B.red_bp_off;
    RunWarp &S = sm.w[warp];

    // this warp's piece of the order
    const int nwarp = gridDim.x * kWarpsL, gw = blockIdx.x * kWarpsL + warp;
    const int per = (n_lm + nwarp - 1) / nwarp;   // n_lm records: the whole sorted map, or what the band chunks left over
    const int i0 = min(gw * per, n_lm), i1 = min(i0 + per, n_lm);

    double acc0[36], acc1[36];
#pragma unroll
    for (int q = 0; q < 36; ++q) { acc0[q] = 0.0; acc1[q] = 0.0; }
    int run_d = -1, run_sig = -3, run_free = 0, run_len = 0;   // run_sig: this lane's entry of the signature
    int pa0 = -1, pb0 = -1, pa1 = -1, pb1 = -1;                // edge slots (a <= b) of the two pairs this lane owns

    auto flush = [&]() {
        if (run_len > 0) {
            if (run_free) {
                // pairs: product block (h_a, h_b) -> lower block (h_b, h_a), element (r, c) at row c, column r
                if (pa0 >= 0 && S.hi[pa0] >= 0 && S.hi[pb0] >= 0) {
                    double *blk = sky + (size_t)(S.base[pb0] + S.hi[pa0]) * 36;
#pragma unroll
                    for (int r = 0; r < 6; ++r)
#pragma unroll
                        for (int c = 0; c < 6; ++c) if (pa0 != pb0 || c >= r) atomicAdd(&blk[c * 6 + r], acc0[r * 6 + c]);
                }
                if (pa1 >= 0 && S.hi[pa1] >= 0 && S.hi[pb1] >= 0) {
                    double *blk = sky + (size_t)(S.base[pb1] + S.hi[pa1]) * 36;
#pragma unroll
                    for (int r = 0; r < 6; ++r)
#pragma unroll
                        for (int c = 0; c < 6; ++c) if (pa1 != pb1 || c >= r) atomicAdd(&blk[c * 6 + r], acc1[r * 6 + c]);
                }
            }
            // per-pose sums of this lane's edge slot
            if (lane < run_d && run_sig >= 0) {
                const double *pa = S.PA + lane * kHStride;
                double *dblk = sky + (size_t)(S.base[lane] + run_sig) * 36;
#pragma unroll
                for (int a = 0; a < 6; ++a) {
                    atomicAdd(&gvec[6 * (size_t)run_sig + a], pa[21 + a]);
                    atomicAdd(&bpvec[6 * (size_t)run_sig + a], pa[27 + a]);
#pragma unroll
                    for (int c = a; c < 6; ++c) atomicAdd(&dblk[c * 6 + a], pa[hd_index(a, c)]);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 36; ++q) { acc0[q] = 0.0; acc1[q] = 0.0; }
        run_len = 0;
        __syncwarp();
    };

    // software pipeline: the sorted records are read 32 at a time (lane j holds landmark blk0 + j), the edge records and the
    // point of landmark idx + 1 are requested while landmark idx is computed — their addresses come from the record block,
    // not from a chain of dependent loads
    int4 rblk = make_int4(0, 0, 0, 0);
    int blk0 = i0 - 32;
    int n_pw = 0;
    double n_ou = 0.0, n_ov = 0.0, n_our = 0.0, n_pt = 0.0;   // next landmark: edge record of this lane, point coordinate (lanes 0..2)
    auto request = [&](int idx) {
        if (idx >= i1) return;
        if (idx >= blk0 + 32) { blk0 = idx; rblk = (idx + lane < i1) ? lm_rec[idx + lane] : make_int4(0, 0, 0, 0); }
        const int j = idx - blk0;
        const int l = __shfl_sync(0xffffffffu, rblk.x, j), e0 = __shfl_sync(0xffffffffu, rblk.y, j);
        const int d = __shfl_sync(0xffffffffu, rblk.z, j);
        if (lane < d) {
            const int e = e0 + lane;
            n_pw = B.edge_pose[e];
            n_ou = B.obs_u[e]; n_ov = B.obs_v[e]; n_our = B.obs_r[e];
        }
        if (lane < 3) n_pt = gpoint[3 * (size_t)l + lane];
    };
    request(i0);
    for (int idx = i0; idx < i1; ++idx) {
        const int j = idx - blk0;
        const int l = __shfl_sync(0xffffffffu, rblk.x, j);
        const int d = __shfl_sync(0xffffffffu, rblk.z, j);
        const uint8_t lf = (uint8_t)__shfl_sync(0xffffffffu, rblk.w, j);
        (void)l;
        const int pw = n_pw;
        const double ou = n_ou, ov = n_ov, our = n_our;
        const double px = __shfl_sync(0xffffffffu, n_pt, 0), py = __shfl_sync(0xffffffffu, n_pt, 1), pz = __shfl_sync(0xffffffffu, n_pt, 2);
        request(idx + 1);
        if (d <= 0) continue;
        const bool lmfree = (lf & kInHessian) != 0;
        bool act = false;
        int hi = -1;
        EdgeLin lin;
        if (lane < d) {
            const int p = pw & kPoseMask;
            act = !(pw & kCulledBit) && !((lf & kFixed) && (B.pose_flags[p] & kFixed));
            if (act) {
                hi = B.pose_hidx[p];
                edge_linearize(gpose + (size_t)p * kPoseStride, px, py, pz, ou, ov, our, (pw & kMonoBit) != 0, K, lin);
            }
        }
        const int sig = (lane < d) ? ((act && hi >= 0) ? hi : -1) : -2;
        // same poses as the running run?
        const bool same = (d == run_d) && (lmfree == (run_free != 0)) && __all_sync(0xffffffffu, sig == run_sig);
        if (!same) {
            flush();
            run_d = d; run_sig = sig; run_free = lmfree ? 1 : 0;
            if (lane < d) {
                S.hi[lane] = sig;
                S.base[lane] = (sig >= 0) ? (B.sky_off[sig] - B.sky_first[sig]) : 0;
#pragma unroll
                for (int q = 0; q < kHStride; ++q) S.PA[lane * kHStride + q] = 0.0;
            }
            // pair p (row-major over a <= b < d) -> (a, b) for p = lane and lane + 32
            const int np = d * (d + 1) / 2;
            pa0 = pb0 = pa1 = pb1 = -1;
            {
                int p = lane, a = 0;
                if (p < np) { while (p >= d - a) { p -= d - a; ++a; } pa0 = a; pb0 = a + p; }
                p = lane + 32; a = 0;
                if (p < np) { while (p >= d - a) { p -= d - a; ++a; } pa1 = a; pb1 = a + p; }
            }
            __syncwarp();
        }
        run_len += 1;

        const double wo = act ? lin.w * K.inv_pv : 0.0;
        double hl[9];
        if (act && lmfree) {
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
#pragma unroll
        for (int q = 0; q < 9; ++q) hl[q] = warp_sum(hl[q]);
        double Di[6], db[3] = {0.0, 0.0, 0.0};
        if (lmfree) {
            const double A[6] = {hl[0] + lambda, hl[1], hl[2], hl[3] + lambda, hl[4], hl[5] + lambda};
            const double bl[3] = {hl[6], hl[7], hl[8]};
            inv_sym3(A, Di);
            sym3_mul(Di, bl, db);
        }
        if (sig >= 0) {
            double Wm[18];
            double *ws = S.W + lane * 18, *ys = S.Yn + lane * 18;
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
                    ws[a * 3] = Wm[a * 3]; ws[a * 3 + 1] = Wm[a * 3 + 1]; ws[a * 3 + 2] = Wm[a * 3 + 2];
                    ys[a * 3 + 0] = -fma(Wm[a * 3], Di[0], fma(Wm[a * 3 + 1], Di[1], Wm[a * 3 + 2] * Di[2]));
                    ys[a * 3 + 1] = -fma(Wm[a * 3], Di[1], fma(Wm[a * 3 + 1], Di[3], Wm[a * 3 + 2] * Di[4]));
                    ys[a * 3 + 2] = -fma(Wm[a * 3], Di[2], fma(Wm[a * 3 + 1], Di[4], Wm[a * 3 + 2] * Di[5]));
                }
            } else {
#pragma unroll
                for (int q = 0; q < 18; ++q) Wm[q] = 0.0;
            }
            const double wr0 = wo * lin.r[0], wr1 = wo * lin.r[1], wr2 = wo * lin.r[2];
            double *pa = S.PA + lane * kHStride;
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                const double bp = -fma(lin.Jp[a], wr0, fma(lin.Jp[6 + a], wr1, lin.Jp[12 + a] * wr2));
                pa[27 + a] += bp;
                pa[21 + a] += bp - fma(Wm[a * 3], db[0], fma(Wm[a * 3 + 1], db[1], Wm[a * 3 + 2] * db[2]));
#pragma unroll
                for (int c = a; c < 6; ++c)
                    pa[hd_index(a, c)] += wo * fma(lin.Jp[a], lin.Jp[c], fma(lin.Jp[6 + a], lin.Jp[6 + c], lin.Jp[12 + a] * lin.Jp[12 + c]));
            }
        }
        __syncwarp();
        if (lmfree) {
            if (pa0 >= 0 && S.hi[pa0] >= 0 && S.hi[pb0] >= 0) {
                const double *ya = S.Yn + pa0 * 18, *wb = S.W + pb0 * 18;
                double Wj[18];
#pragma unroll
                for (int q = 0; q < 18; ++q) Wj[q] = wb[q];
#pragma unroll
                for (int r = 0; r < 6; ++r) {
                    const double y0 = ya[r * 3], y1 = ya[r * 3 + 1], y2 = ya[r * 3 + 2];
#pragma unroll
                    for (int c = 0; c < 6; ++c) acc0[r * 6 + c] = fma(y0, Wj[c * 3], fma(y1, Wj[c * 3 + 1], fma(y2, Wj[c * 3 + 2], acc0[r * 6 + c])));
                }
            }
            if (pa1 >= 0 && S.hi[pa1] >= 0 && S.hi[pb1] >= 0) {
                const double *ya = S.Yn + pa1 * 18, *wb = S.W + pb1 * 18;
                double Wj[18];
#pragma unroll
                for (int q = 0; q < 18; ++q) Wj[q] = wb[q];
#pragma unroll
                for (int r = 0; r < 6; ++r) {
                    const double y0 = ya[r * 3], y1 = ya[r * 3 + 1], y2 = ya[r * 3 + 2];
#pragma unroll
                    for (int c = 0; c < 6; ++c) acc1[r * 6 + c] = fma(y0, Wj[c * 3], fma(y1, Wj[c * 3 + 1], fma(y2, Wj[c * 3 + 2], acc1[r * 6 + c])));
                }
            }
        }
        __syncwarp();
    }
    flush();
}

// first pose (smallest hessian index) of every landmark in the Hessian, the sort key of k_build_large_run's order
__global__ void k_lm_first(Batch B, int *key, int *idx) {
    const WinDesc &wd = B.win[0];
    for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < wd.n_point; l += gridDim.x * blockDim.x) {
        int mn = 0x7fffffff, mx = -1;
        for (int e = B.lm_edge_off[l]; e < B.lm_edge_off[l + 1]; ++e) {
            const int hi = B.pose_hidx[B.edge_pose[e] & kPoseMask];
            if (hi >= 0) { mn = min(mn, hi); mx = max(mx, hi); }
        }
        // landmarks that span more poses than a band chunk holds (loop closures) sort behind all others, again by first pose:
        // together with their neighbours they would push every chunk they land in over the pose limit (ba_band.cuh)
        key[l] = (mx - mn >= 19 && mn < (1 << 24)) ? mn + (1 << 24) : mn;
        idx[l] = l;
    }
}

// sorted landmark records for k_build_large_run: rec[idx] = (landmark, first edge, degree, flags) so that the kernel reads
// them contiguously, and the number of runs (landmarks whose pose list differs from their predecessor's in the order)
__global__ void k_run_prep(Batch B, const int *__restrict__ order, int4 *rec, int *n_runs) {
    const WinDesc &wd = B.win[0];
    int cnt = 0;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < wd.n_point; idx += gridDim.x * blockDim.x) {
        const int l = order[idx];
        const int e0 = B.lm_edge_off[l], d = B.lm_edge_off[l + 1] - e0;
        rec[idx] = make_int4(l, e0, d, (int)B.lm_flags[l]);
        bool boundary = (idx == 0);
        if (!boundary) {
            const int lp = order[idx - 1];
            const int p0 = B.lm_edge_off[lp], dp = B.lm_edge_off[lp + 1] - p0;
            boundary = (dp != d) || ((B.lm_flags[lp] ^ B.lm_flags[l]) & kInHessian);
            for (int k = 0; k < d && !boundary; ++k)
                boundary = ((B.edge_pose[e0 + k] ^ B.edge_pose[p0 + k]) & (kPoseMask | kCulledBit)) != 0;
        }
        cnt += boundary ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(n_runs, cnt);
}

// sum / max of the per-CTA partials in CTA order -> out[0], out[1]   (one CTA)
__global__ void k_fold_part2(Batch B, int n, int second_is_max, double *out) {
    __shared__ double red[32];
    double a = 0.0, b = 0.0;
    if (B.st[0].done) return;
    // fixed order: thread t owns the partials t, t + blockDim, ...
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        a += B.part2[2 * (size_t)i];
        b = second_is_max ? fmax(b, B.part2[2 * (size_t)i + 1]) : b + B.part2[2 * (size_t)i + 1];
    }
    const double sa = block_sum(a, red);
    const double sb = second_is_max ? block_max(b, red) : block_sum(b, red);
    if (threadIdx.x == 0) { out[0] = sa; out[1] = sb; }
}

// lambda init + pass bookkeeping (k_control_init of the small path): scal = {chi2, max |diag H_ll|}
__global__ void k_control_init_large(Batch B, const double *scal) {
    const WinDesc &wd = B.win[0];
    LMState &st = B.st[0];
    if (st.done) return;
    const int F = st.F;
    double md = 0.0;
    for (int i = threadIdx.x; i < 6 * F; i += blockDim.x) md = fmax(md, fabs(B.hdiag[i]));
    __shared__ double red[32];
    md = block_max(md, red);
    if (threadIdx.x == 0) {
        md = fmax(md, scal[1]);
        const double chi = scal[0];
        st.cur_chi = chi;
        st.chi_last_trial = chi;
        if (st.pass == 0) st.chi_initial = chi;
        st.lambda = 1e-5 * md;
        st.ni = 2.0;
        st.iter = 0; st.qmax = 0;
        st.pcg_residual = -1.0;
        if (st.err != 0) finish_pass(st, VISFS_BA_STOP_NOT_RUN, B.n_running);
        else if (st.F + st.NL == 0) finish_pass(st, VISFS_BA_STOP_EMPTY, B.n_running);
        else if (wd.max_iter <= 0) finish_pass(st, VISFS_BA_STOP_ITERATIONS, B.n_running);
    }
}

// ------------------------------------------------------------------------------------------------
// k_solve_large: in-place block-skyline Cholesky of the damped reduced system + solve + pose update
// ------------------------------------------------------------------------------------------------
struct Chol6 {
    double L10, L20, L21, L30, L31, L32, L40, L41, L42, L43, L50, L51, L52, L53, L54;
    double L00, L11, L22, L33, L44, L55;
    double i0, i1, i2, i3, i4, i5;
    bool ok;
};

// Cholesky of a 6x6 block given by its lower half d[c*6 + a] (row c, column a), `lam` added to the diagonal
__device__ __forceinline__ void chol6(const double *d, double lam, Chol6 &f) {
    f.ok = true;
    double v;
#define VISFS_PIV(expr, inv, diag) v = (expr); if (!(v > 0.0)) { f.ok = false; v = 1.0; } inv = rsqrt(v); diag = v * inv;
    VISFS_PIV(d[0] + lam, f.i0, f.L00)
    f.L10 = d[6] * f.i0; f.L20 = d[12] * f.i0; f.L30 = d[18] * f.i0; f.L40 = d[24] * f.i0; f.L50 = d[30] * f.i0;
    VISFS_PIV(fma(-f.L10, f.L10, d[7] + lam), f.i1, f.L11)
    f.L21 = fma(-f.L20, f.L10, d[13]) * f.i1; f.L31 = fma(-f.L30, f.L10, d[19]) * f.i1;
    f.L41 = fma(-f.L40, f.L10, d[25]) * f.i1; f.L51 = fma(-f.L50, f.L10, d[31]) * f.i1;
    VISFS_PIV(fma(-f.L21, f.L21, fma(-f.L20, f.L20, d[14] + lam)), f.i2, f.L22)
    f.L32 = fma(-f.L31, f.L21, fma(-f.L30, f.L20, d[20])) * f.i2;
    f.L42 = fma(-f.L41, f.L21, fma(-f.L40, f.L20, d[26])) * f.i2;
    f.L52 = fma(-f.L51, f.L21, fma(-f.L50, f.L20, d[32])) * f.i2;
    VISFS_PIV(fma(-f.L32, f.L32, fma(-f.L31, f.L31, fma(-f.L30, f.L30, d[21] + lam))), f.i3, f.L33)
    f.L43 = fma(-f.L42, f.L32, fma(-f.L41, f.L31, fma(-f.L40, f.L30, d[27]))) * f.i3;
    f.L53 = fma(-f.L52, f.L32, fma(-f.L51, f.L31, fma(-f.L50, f.L30, d[33]))) * f.i3;
    VISFS_PIV(fma(-f.L43, f.L43, fma(-f.L42, f.L42, fma(-f.L41, f.L41, fma(-f.L40, f.L40, d[28] + lam)))), f.i4, f.L44)
    f.L54 = fma(-f.L53, f.L43, fma(-f.L52, f.L42, fma(-f.L51, f.L41, fma(-f.L50, f.L40, d[34])))) * f.i4;
    VISFS_PIV(fma(-f.L54, f.L54, fma(-f.L53, f.L53, fma(-f.L52, f.L52, fma(-f.L51, f.L51, fma(-f.L50, f.L50, d[35] + lam))))),
              f.i5, f.L55)
#undef VISFS_PIV
}

// x L^T = v  (one row against the factor of the diagonal block), in place
__device__ __forceinline__ void row_solve6(const Chol6 &f, double *x) {
    const double x0 = x[0] * f.i0;
    const double x1 = fma(-x0, f.L10, x[1]) * f.i1;
    const double x2 = fma(-x1, f.L21, fma(-x0, f.L20, x[2])) * f.i2;
    const double x3 = fma(-x2, f.L32, fma(-x1, f.L31, fma(-x0, f.L30, x[3]))) * f.i3;
    const double x4 = fma(-x3, f.L43, fma(-x2, f.L42, fma(-x1, f.L41, fma(-x0, f.L40, x[4])))) * f.i4;
    const double x5 = fma(-x4, f.L54, fma(-x3, f.L53, fma(-x2, f.L52, fma(-x1, f.L51, fma(-x0, f.L50, x[5]))))) * f.i5;
    x[0] = x0; x[1] = x1; x[2] = x2; x[3] = x3; x[4] = x4; x[5] = x5;
}

// COOP = false: one CTA (banded systems: the column steps are short and __syncthreads is the cheap barrier)
// COOP = true : cooperative launch over the whole GPU, grid.sync() between the panel and the trailing update of a column
//               (wide fronts / dense systems: the trailing update of one column is thousands of 6x6 products);
//               CTA 0 alone runs the back-substitution and the epilogue.
template <bool COOP>
__global__ void __launch_bounds__(kSolveThreadsL) k_solve_large(Batch B, int *flag) {
    const WinDesc &wd = B.win[0];
    LMState &st = B.st[0];
    if (st.done) return;
    const int tid = threadIdx.x;
    const int gtid = COOP ? blockIdx.x * kSolveThreadsL + tid : tid;
    const int gsize = COOP ? gridDim.x * kSolveThreadsL : kSolveThreadsL;
    const int F = st.F, n = 6 * F;
    const double lambda = (wd.trust == 0) ? (B.dbg_lambda >= 0.0 ? B.dbg_lambda : st.lambda) : 0.0;
    __shared__ double s_red[32];
    __shared__ double s_x[6];
    if (n == 0) {
        if (gtid == 0) { st.ok = 1; st.scale_p = 0.0; }
        return;
    }
    double *__restrict__ sky = B.red;
    double *__restrict__ y = B.red + B.red_g_off;          // rhs -> y -> x, in place
    const double *__restrict__ braw = B.red + B.red_bp_off;
    const int *__restrict__ first = B.sky_first;
    const long long *__restrict__ off = B.sky_off;
    cg::grid_group grid = cg::this_grid();
    auto barrier = [&]() { if (COOP) grid.sync(); else __syncthreads(); };

    // ---- factorisation, one block column per step
    for (int k = 0; k < F; ++k) {
        const int c0 = B.col_ptr[k], m = B.col_ptr[k + 1] - c0;
        const int *rows = B.col_rows + c0;
        double *dk = sky + (size_t)(off[k] + (k - first[k])) * 36;
        Chol6 f;
        const int ntask = 6 * m + 1;
        bool owner = false;                  // the thread that forward-substitutes the rhs also stores the diagonal factor
        if (gtid < ntask) chol6(dk, lambda, f);   // every task owner factors the diagonal block redundantly
        for (int t = gtid; t < ntask; t += gsize) {
            if (t < 6 * m) {                 // panel row: block L_rk, row a
                const int r = rows[t / 6], a = t - (t / 6) * 6;
                double *x = sky + (size_t)(off[r] + (k - first[r])) * 36 + a * 6;
                double xv[6] = {x[0], x[1], x[2], x[3], x[4], x[5]};
                row_solve6(f, xv);
#pragma unroll
                for (int q = 0; q < 6; ++q) x[q] = xv[q];
            } else {                         // forward substitution of the rhs: y_k = L_kk^-1 b_k
                double xv[6] = {y[6 * k], y[6 * k + 1], y[6 * k + 2], y[6 * k + 3], y[6 * k + 4], y[6 * k + 5]};
                row_solve6(f, xv);           // (row vector times L^-T) == L^-1 applied to the column
#pragma unroll
                for (int q = 0; q < 6; ++q) y[6 * k + q] = xv[q];
                if (!f.ok) *flag = 1;     // cleared by the host before the launch
                owner = true;
            }
        }
        barrier();
        if (owner) {   // nobody reads the diagonal block again before the back-substitution
            dk[0] = f.L00;
            dk[6] = f.L10; dk[7] = f.L11;
            dk[12] = f.L20; dk[13] = f.L21; dk[14] = f.L22;
            dk[18] = f.L30; dk[19] = f.L31; dk[20] = f.L32; dk[21] = f.L33;
            dk[24] = f.L40; dk[25] = f.L41; dk[26] = f.L42; dk[27] = f.L43; dk[28] = f.L44;
            dk[30] = f.L50; dk[31] = f.L51; dk[32] = f.L52; dk[33] = f.L53; dk[34] = f.L54; dk[35] = f.L55;
        }
        if (m > 0) {
            const double y0 = y[6 * k], y1 = y[6 * k + 1], y2 = y[6 * k + 2], y3 = y[6 * k + 3], y4 = y[6 * k + 4], y5 = y[6 * k + 5];
            // rhs update: b_r -= L_rk y_k, one thread per (row block, row)
            for (int t = gtid; t < 6 * m; t += gsize) {
                const int r = rows[t / 6], a = t - (t / 6) * 6;
                const double *x = sky + (size_t)(off[r] + (k - first[r])) * 36 + a * 6;
                y[6 * r + a] -= fma(x[0], y0, fma(x[1], y1, fma(x[2], y2, fma(x[3], y3, fma(x[4], y4, x[5] * y5)))));
            }
            // trailing update: A_rc -= L_rk L_ck^T for r >= c in the column structure (all inside the envelope);
            // one thread per 6x6 row (6 entries): L_rk row a against the six rows of L_ck
            const int pairs = m * (m + 1) / 2;
            for (int item = gtid; item < pairs * 6; item += gsize) {
                const int pr = item / 6, a = item - pr * 6;
                int ri = (int)((sqrt(8.0 * (double)pr + 1.0) - 1.0) * 0.5);   // pr -> (ri >= ci)
                while ((ri + 1) * (ri + 2) / 2 <= pr) ++ri;
                while (ri * (ri + 1) / 2 > pr) --ri;
                const int ci = pr - ri * (ri + 1) / 2;
                const int r = rows[ri], cc = rows[ci];
                const double *lr = sky + (size_t)(off[r] + (k - first[r])) * 36 + a * 6;
                const double *lc = sky + (size_t)(off[cc] + (k - first[cc])) * 36;
                double *dst = sky + (size_t)(off[r] + (cc - first[r])) * 36 + a * 6;
                const double l0 = lr[0], l1 = lr[1], l2 = lr[2], l3 = lr[3], l4 = lr[4], l5 = lr[5];
                double d[6];
#pragma unroll
                for (int c = 0; c < 6; ++c) d[c] = dst[c];
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    const double *w = lc + c * 6;
                    d[c] -= fma(l0, w[0], fma(l1, w[1], fma(l2, w[2], fma(l3, w[3], fma(l4, w[4], l5 * w[5])))));
                }
                const int cmax = (ri == ci) ? a : 5;    // diagonal block: lower half only
#pragma unroll
                for (int c = 0; c < 6; ++c) if (c <= cmax) dst[c] = d[c];
            }
        }
        barrier();
    }
    if (COOP && blockIdx.x != 0) return;

    // ---- back-substitution L^T x = y, row-oriented: x_r = L_rr^-T y_r, then y_c -= L_rc^T x_r for the row's blocks
    for (int r = F - 1; r >= 0; --r) {
        const double *dr = sky + (size_t)(off[r] + (r - first[r])) * 36;
        if (tid == 0) {
            double x5 = y[6 * r + 5] / dr[35];
            double x4 = (y[6 * r + 4] - dr[34] * x5) / dr[28];
            double x3 = (y[6 * r + 3] - dr[33] * x5 - dr[27] * x4) / dr[21];
            double x2 = (y[6 * r + 2] - dr[32] * x5 - dr[26] * x4 - dr[20] * x3) / dr[14];
            double x1 = (y[6 * r + 1] - dr[31] * x5 - dr[25] * x4 - dr[19] * x3 - dr[13] * x2) / dr[7];
            double x0 = (y[6 * r] - dr[30] * x5 - dr[24] * x4 - dr[18] * x3 - dr[12] * x2 - dr[6] * x1) / dr[0];
            y[6 * r] = x0; y[6 * r + 1] = x1; y[6 * r + 2] = x2; y[6 * r + 3] = x3; y[6 * r + 4] = x4; y[6 * r + 5] = x5;
            s_x[0] = x0; s_x[1] = x1; s_x[2] = x2; s_x[3] = x3; s_x[4] = x4; s_x[5] = x5;
        }
        __syncthreads();
        const int f0 = first[r], len = r - f0;
        const double *rowblk = sky + (size_t)off[r] * 36;
        for (int t = tid; t < 6 * len; t += kSolveThreadsL) {
            const int cb = t / 6, a = t - cb * 6;        // column block f0 + cb, column a inside it
            const double *blk = rowblk + (size_t)cb * 36;
            double s = 0.0;
#pragma unroll
            for (int u = 0; u < 6; ++u) s = fma(blk[u * 6 + a], s_x[u], s);
            y[6 * (f0 + cb) + a] -= s;
        }
        __syncthreads();
    }

    // ---- solution checks, pose step, trial poses, pose part of g2o's computeScale
    double bad = 0.0;
    for (int i = tid; i < n; i += kSolveThreadsL) if (!isfinite(y[i])) bad = 1.0;
    const double anybad = block_sum(bad, s_red);
    const bool ok = (*flag == 0) && (anybad == 0.0);
    __syncthreads();
    double sc = 0.0;
    for (int i = tid; i < n; i += kSolveThreadsL) {
        const double x = ok ? y[i] : 0.0;
        B.xp[i] = x;
        sc += x * (lambda * x + braw[i]);
    }
    const double scale = block_sum(sc, s_red);
    const int cur = st.cur;
    const double *src = B.pose + (size_t)cur * B.tot_pose * kPoseStride;
    double *dst = B.pose + (size_t)(1 - cur) * B.tot_pose * kPoseStride;
    for (int p = tid; p < wd.n_pose; p += kSolveThreadsL) {
        const int hi = B.pose_hidx[p];
        if (hi >= 0) {
            double dlt[6];
            for (int a = 0; a < 6; ++a) dlt[a] = ok ? y[6 * hi + a] : 0.0;
            pose_oplus(src + (size_t)p * kPoseStride, dlt, dst + (size_t)p * kPoseStride);
        }
    }
    if (tid == 0) { st.ok = ok ? 1 : 0; st.scale_p = scale; }
}

// ------------------------------------------------------------------------------------------------
// k_solve_front: the same factorisation for NARROW fronts (banded systems with a few long rows: a trajectory with loop
// closures).  The active front — pivot row k plus the rows of its column structure — lives in shared memory as a
// kFrontSlots x kFrontSlots matrix of 6x6 blocks; a row occupies one slot from the column where it enters the envelope
// until one column after it has been the pivot.  The host plans the whole schedule from sky_first once per pass
// (FrontPlan): slots, the rows entering at every column and the list of skyline blocks to bring in for them.  Per column
// the panel and the trailing update run entirely in shared memory; everything the NEXT column needs from L2 (entering
// blocks, slots, right-hand side) is loaded into registers at the start of the column and parked in shared memory at its
// end, so no global-memory latency sits on the critical path.  Finished columns of L go back to the skyline for the
// back-substitution, which keeps x in shared memory and prefetches one row ahead.
// ------------------------------------------------------------------------------------------------
constexpr int kFrontSlots = 24;
constexpr int kFrontMaxF = 2560;          // y (6F doubles) + row tables reuse the front's shared memory in the back-substitution
constexpr int kFrontBS = 37;             // doubles per 6x6 block in shared memory: odd, so different blocks start in different banks
constexpr int kFrontPitch = kFrontSlots + 1;   // blocks per row of the slot matrix: odd pitch * odd block size spreads a block column over the banks
constexpr int kFrontLoaders = 128;        // threads (4 warps) that only stream the next column's data in; the rest compute

struct FrontPlan {                        // device pointers of the host-built schedule
    const unsigned char *pslot;           // [F]  slot of row r
    const int *ent_ptr;                   // [F + 1] rows entering at column k ...
    const int *ent_row, *ent_base;        // ... their index and sky_off - sky_first (blocks)
    const int *ld_ptr;                    // [F + 1] skyline blocks to bring in before column k ...
    const int *ld_src;                    // ... block index in the skyline
    const unsigned short *ld_dst;         // ... slot_hi * kFrontPitch + slot_lo
    const unsigned char *fr_slot;         // slots of the column structure, aligned with col_rows
    const int *row_len, *row_off;         // [F] r - sky_first[r], sky_off[r]
};

struct FrontSmem {
    double blk[kFrontSlots * kFrontPitch * kFrontBS];
    double yf[kFrontSlots * 6];
    double red[32];
    double sx[6];
    int col_ptr[kFrontMaxF + 2];
    int ent_ptr[kFrontMaxF + 2];
    int ld_ptr[kFrontMaxF + 2];
    int s_base[kFrontSlots];
    int rslot[2][kFrontSlots];
    int fail;
    unsigned char slot[kFrontMaxF];
};

__global__ void __launch_bounds__(kSolveThreadsL) k_solve_front(Batch B, FrontPlan P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FrontSmem &sm = *reinterpret_cast<FrontSmem *>(smem_raw);
    const WinDesc &wd = B.win[0];
    LMState &st = B.st[0];
    if (st.done) return;
    const int tid = threadIdx.x;
    const int F = st.F, n = 6 * F;
    const double lambda = (wd.trust == 0) ? (B.dbg_lambda >= 0.0 ? B.dbg_lambda : st.lambda) : 0.0;
    if (n == 0) {
        if (tid == 0) { st.ok = 1; st.scale_p = 0.0; }
        return;
    }
    constexpr int SL = kFrontPitch, BS = kFrontBS;
    constexpr int NC = kSolveThreadsL - kFrontLoaders;   // compute threads
    const bool loader = tid >= NC;
    const int ltid = tid - NC;
    double *__restrict__ sky = B.red;
    double *__restrict__ y = B.red + B.red_g_off;
    const double *__restrict__ braw = B.red + B.red_bp_off;
    long long tclk[6];
    tclk[0] = clock64();
    for (int i = tid; i <= F; i += kSolveThreadsL) { sm.col_ptr[i] = B.col_ptr[i]; sm.ent_ptr[i] = P.ent_ptr[i]; sm.ld_ptr[i] = P.ld_ptr[i]; }
    for (int i = tid; i < F; i += kSolveThreadsL) sm.slot[i] = P.pslot[i];
    if (tid == 0) { sm.fail = 0; sm.col_ptr[F + 1] = B.col_ptr[F]; sm.ent_ptr[F + 1] = P.ent_ptr[F]; sm.ld_ptr[F + 1] = P.ld_ptr[F]; }
    __syncthreads();
    // what column 0 needs, without prefetch
    {
        const int m0 = sm.col_ptr[1], ne0 = sm.ent_ptr[1], nl0 = sm.ld_ptr[1] * 36;
        if (tid < m0) sm.rslot[0][tid] = P.fr_slot[tid];
        for (int i = tid; i < ne0; i += kSolveThreadsL) sm.s_base[sm.slot[P.ent_row[i]]] = P.ent_base[i];
        for (int i = tid; i < ne0 * 6; i += kSolveThreadsL) { const int r = P.ent_row[i / 6]; sm.yf[(int)sm.slot[r] * 6 + i % 6] = y[6 * r + i % 6]; }
        for (int i = tid; i < nl0; i += kSolveThreadsL) {
            const int b = i / 36, ent = i - b * 36;
            sm.blk[(int)P.ld_dst[b] * BS + ent] = sky[(size_t)P.ld_src[b] * 36 + ent];
        }
    }
    __syncthreads();
    tclk[1] = clock64();

    long long phA = 0, phB = 0, phL = 0;
    for (int k = 0; k < F; ++k) {
        const long long c_top = clock64();
        const int m = sm.col_ptr[k + 1] - sm.col_ptr[k];
        const int *rslot = sm.rslot[k & 1];
        const int sk = sm.slot[k];
        if (loader) {
            // ---- loader warps: everything column k + 1 needs from L2 goes straight into shared memory (its new rows use
            //      slots outside the current front, so nothing the compute warps touch in this column is overwritten)
            const int c1 = sm.col_ptr[k + 1], m1 = sm.col_ptr[k + 2] - c1;
            const int e1 = sm.ent_ptr[k + 1], ne1 = sm.ent_ptr[k + 2] - e1;
            const int l1 = sm.ld_ptr[k + 1], nl1 = (sm.ld_ptr[k + 2] - l1) * 36;
            if (ltid < m1) sm.rslot[(k + 1) & 1][ltid] = P.fr_slot[c1 + ltid];
            for (int i = ltid; i < ne1; i += kFrontLoaders) sm.s_base[sm.slot[P.ent_row[e1 + i]]] = P.ent_base[e1 + i];
            for (int i = ltid; i < ne1 * 6; i += kFrontLoaders) {
                const int r = P.ent_row[e1 + i / 6];
                sm.yf[(int)sm.slot[r] * 6 + i % 6] = y[6 * r + i % 6];
            }
            // one 6-entry row of a block per thread and step: one index load, then three 16-byte loads
            const int nrow = (sm.ld_ptr[k + 2] - l1) * 6;
            for (int i = ltid; i < nrow; i += kFrontLoaders) {
                const int b = i / 6, a = i - b * 6;
                const double2 *src = reinterpret_cast<const double2 *>(sky + (size_t)P.ld_src[l1 + b] * 36 + a * 6);
                const double2 v0 = src[0], v1 = src[1], v2 = src[2];
                double *d = sm.blk + (int)P.ld_dst[l1 + b] * BS + a * 6;
                d[0] = v0.x; d[1] = v0.y; d[2] = v1.x; d[3] = v1.y; d[4] = v2.x; d[5] = v2.y;
            }
            (void)nl1;
            phL += clock64() - c_top;
            __syncthreads();   // end of the column (the barrier after the panel is the compute warps' own)
            continue;
        }
        // ---- panel: L_rk = A_rk L_kk^-T for the rows of the column structure, y_k = L_kk^-1 b_k
        Chol6 f;
        const int ntask = 6 * m + 1;
        bool owner = false;
        if (tid < ntask) chol6(sm.blk + (sk * SL + sk) * BS, lambda, f);
        for (int t = tid; t < ntask; t += NC) {
            if (t < 6 * m) {
                const int ri = t / 6, a = t - ri * 6;
                const int rs = rslot[ri];
                double *x = sm.blk + (rs * SL + sk) * BS + a * 6;
                double xv[6] = {x[0], x[1], x[2], x[3], x[4], x[5]};
                row_solve6(f, xv);
                double *g = sky + (size_t)(sm.s_base[rs] + k) * 36 + a * 6;
#ifdef VISFS_FRONT_DEBUG
                if (sm.s_base[rs] + k < 0 || (long long)(sm.s_base[rs] + k) * 36 >= B.red_g_off) { printf("panel oob k %d rs %d base %d\n", k, rs, sm.s_base[rs]); continue; }
#endif
#pragma unroll
                for (int q = 0; q < 6; ++q) { x[q] = xv[q]; g[q] = xv[q]; }
            } else {
                double xv[6];
#pragma unroll
                for (int q = 0; q < 6; ++q) xv[q] = sm.yf[sk * 6 + q];
                row_solve6(f, xv);
#pragma unroll
                for (int q = 0; q < 6; ++q) { y[6 * k + q] = xv[q]; sm.sx[q] = xv[q]; }
                if (!f.ok) sm.fail = 1;
                owner = true;
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NC) : "memory");   // compute warps only: the loaders keep streaming
        const long long c_mid = clock64();
        phA += c_mid - c_top;
        if (owner) {   // factor of the diagonal block for the back-substitution: strict lower part and RECIPROCAL pivots
            double *dk = sky + (size_t)(sm.s_base[sk] + k) * 36;
            dk[0] = f.i0;
            dk[6] = f.L10; dk[7] = f.i1;
            dk[12] = f.L20; dk[13] = f.L21; dk[14] = f.i2;
            dk[18] = f.L30; dk[19] = f.L31; dk[20] = f.L32; dk[21] = f.i3;
            dk[24] = f.L40; dk[25] = f.L41; dk[26] = f.L42; dk[27] = f.L43; dk[28] = f.i4;
            dk[30] = f.L50; dk[31] = f.L51; dk[32] = f.L52; dk[33] = f.L53; dk[34] = f.L54; dk[35] = f.i5;
        }
        // ---- rhs and trailing update of the front, all in shared memory: one thread per half block (3 x 6 entries).  Every
        //      operand is loaded before the first store, so the loads pipeline; consecutive lanes share the column block,
        //      whose rows are then broadcast loads
        const int pairs = m * (m + 1) / 2;
        for (int item = tid; item < pairs * 2 + 6 * m; item += NC) {
            if (item < pairs * 2) {
                const int pr = item >> 1, a0 = (item & 1) * 3;
                int ci = 0, base = 0;                       // pr -> (ci <= ri), column-major
                while (base + (m - ci) <= pr) { base += m - ci; ++ci; }
                const int ri = ci + (pr - base);
                const double *lr = sm.blk + (rslot[ri] * SL + sk) * BS + a0 * 6;
                const double *lc = sm.blk + (rslot[ci] * SL + sk) * BS;
                double *dst = sm.blk + (rslot[ri] * SL + rslot[ci]) * BS + a0 * 6;
                double L[18], acc[18];
#pragma unroll
                for (int q = 0; q < 18; ++q) { L[q] = lr[q]; acc[q] = dst[q]; }
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    const double w0 = lc[c * 6], w1 = lc[c * 6 + 1], w2 = lc[c * 6 + 2], w3 = lc[c * 6 + 3], w4 = lc[c * 6 + 4], w5 = lc[c * 6 + 5];
#pragma unroll
                    for (int a = 0; a < 3; ++a)
                        acc[a * 6 + c] -= fma(L[a * 6], w0, fma(L[a * 6 + 1], w1, fma(L[a * 6 + 2], w2, fma(L[a * 6 + 3], w3, fma(L[a * 6 + 4], w4, L[a * 6 + 5] * w5)))));
                }
#pragma unroll
                for (int q = 0; q < 18; ++q) dst[q] = acc[q];
            } else {
                const int t = item - pairs * 2;
                const int ri = t / 6, a = t - ri * 6;
                const double *x = sm.blk + (rslot[ri] * SL + sk) * BS + a * 6;
                sm.yf[rslot[ri] * 6 + a] -= fma(x[0], sm.sx[0], fma(x[1], sm.sx[1], fma(x[2], sm.sx[2], fma(x[3], sm.sx[3], fma(x[4], sm.sx[4], x[5] * sm.sx[5])))));
            }
        }
        __syncthreads();
        phB += clock64() - c_mid;
    }
#ifdef VISFS_FRONT_DEBUG
    if (tid == 0 || tid == NC) printf("front tid %d: panel+barrier %lld  trailing+barrier %lld  loader busy %lld (cycles, %d columns)\n", tid, phA, phB, phL, F);
#endif

    // ---- back-substitution L^T x = y: x in shared memory (the front is no longer needed), row r - 1 prefetched while
    //      row r is applied
    tclk[2] = clock64();
    double *ys = sm.blk;
    int *row_len = reinterpret_cast<int *>(sm.blk + 6 * kFrontMaxF);
    int *row_off = row_len + kFrontMaxF;
    for (int i = tid; i < n; i += kSolveThreadsL) ys[i] = y[i];
    for (int i = tid; i < F; i += kSolveThreadsL) { row_len[i] = P.row_len[i]; row_off[i] = P.row_off[i]; }
    __syncthreads();
    double pf_d[21], pf_b[6];
#pragma unroll
    for (int q = 0; q < 21; ++q) pf_d[q] = 0.0;
    auto prefetch_row = [&](int r) {
        const int len = row_len[r];
        const double *rowblk = sky + (size_t)row_off[r] * 36;
        if (tid == 0) {
            const double *dr = rowblk + (size_t)len * 36;
            int q = 0;
#pragma unroll
            for (int a = 0; a < 6; ++a)
#pragma unroll
                for (int c = 0; c <= a; ++c) pf_d[q++] = dr[a * 6 + c];
        }
        if (tid < 6 * len) {
            const int cb = tid / 6, a = tid - cb * 6;
            const double *blk = rowblk + (size_t)cb * 36;
#pragma unroll
            for (int u = 0; u < 6; ++u) pf_b[u] = blk[u * 6 + a];
        }
    };
#ifdef VISFS_FRONT_DEBUG
    if (tid == 0) printf("front: factor done F %d\n", F);
#endif
    prefetch_row(F - 1);
    for (int r = F - 1; r >= 0; --r) {
        double d[21], bq[6];
#pragma unroll
        for (int q = 0; q < 21; ++q) d[q] = pf_d[q];
#pragma unroll
        for (int q = 0; q < 6; ++q) bq[q] = pf_b[q];
        if (r > 0) prefetch_row(r - 1);
        if (tid == 0) {   // d: packed lower triangle, row a starts at a (a + 1) / 2; diagonal entries are reciprocals
            const double x5 = ys[6 * r + 5] * d[20];
            const double x4 = (ys[6 * r + 4] - d[19] * x5) * d[14];
            const double x3 = (ys[6 * r + 3] - d[18] * x5 - d[13] * x4) * d[9];
            const double x2 = (ys[6 * r + 2] - d[17] * x5 - d[12] * x4 - d[8] * x3) * d[5];
            const double x1 = (ys[6 * r + 1] - d[16] * x5 - d[11] * x4 - d[7] * x3 - d[4] * x2) * d[2];
            const double x0 = (ys[6 * r] - d[15] * x5 - d[10] * x4 - d[6] * x3 - d[3] * x2 - d[1] * x1) * d[0];
            ys[6 * r] = x0; ys[6 * r + 1] = x1; ys[6 * r + 2] = x2; ys[6 * r + 3] = x3; ys[6 * r + 4] = x4; ys[6 * r + 5] = x5;
            sm.sx[0] = x0; sm.sx[1] = x1; sm.sx[2] = x2; sm.sx[3] = x3; sm.sx[4] = x4; sm.sx[5] = x5;
        }
        __syncthreads();
        const int len = row_len[r], f0 = r - len;
        if (tid < 6 * len) {
            const int cb = tid / 6, a = tid - cb * 6;
            double s = 0.0;
#pragma unroll
            for (int u = 0; u < 6; ++u) s = fma(bq[u], sm.sx[u], s);
            ys[6 * (f0 + cb) + a] -= s;
        }
        if (6 * len > kSolveThreadsL) {   // long rows (loop closures): the rest straight from L2
            const double *rowblk = sky + (size_t)row_off[r] * 36;
            for (int t = tid + kSolveThreadsL; t < 6 * len; t += kSolveThreadsL) {
                const int cb = t / 6, a = t - cb * 6;
                const double *blk = rowblk + (size_t)cb * 36;
                double s = 0.0;
#pragma unroll
                for (int u = 0; u < 6; ++u) s = fma(blk[u * 6 + a], sm.sx[u], s);
                ys[6 * (f0 + cb) + a] -= s;
            }
        }
        __syncthreads();
    }

    tclk[3] = clock64();
#ifdef VISFS_FRONT_DEBUG
    if (tid == 0) printf("front: backsub done\n");
#endif
    // ---- solution checks, pose step, trial poses, pose part of g2o's computeScale
    double bad = 0.0;
    for (int i = tid; i < n; i += kSolveThreadsL) if (!isfinite(ys[i])) bad = 1.0;
    const double anybad = block_sum(bad, sm.red);
    const bool ok = (sm.fail == 0) && (anybad == 0.0);
    __syncthreads();
    double sc = 0.0;
    for (int i = tid; i < n; i += kSolveThreadsL) {
        const double x = ok ? ys[i] : 0.0;
        B.xp[i] = x;
        sc += x * (lambda * x + braw[i]);
    }
    const double scale = block_sum(sc, sm.red);
    const int cur = st.cur;
    const double *src = B.pose + (size_t)cur * B.tot_pose * kPoseStride;
    double *dst = B.pose + (size_t)(1 - cur) * B.tot_pose * kPoseStride;
    for (int p = tid; p < wd.n_pose; p += kSolveThreadsL) {
        const int hi = B.pose_hidx[p];
        if (hi >= 0) {
            double dlt[6];
            for (int a = 0; a < 6; ++a) dlt[a] = ok ? ys[6 * hi + a] : 0.0;
            pose_oplus(src + (size_t)p * kPoseStride, dlt, dst + (size_t)p * kPoseStride);
        }
    }
    if (tid == 0) {
        st.ok = ok ? 1 : 0; st.scale_p = scale;
        tclk[4] = tclk[5] = clock64();
        for (int q = 0; q < 6; ++q) st.t_solve[q] = tclk[q] - tclk[0];
    }
}

// ------------------------------------------------------------------------------------------------
// odometry links on the block skyline (ba_link.cuh): envelope, Hessian / gradient pieces, chi2
// ------------------------------------------------------------------------------------------------
__global__ void k_sky_links(Batch B) {   // a link couples its two poses: the later row reaches back to the earlier one
    if (B.st[0].status != 0) return;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < B.tot_link; k += gridDim.x * blockDim.x) {
        const int hi = B.pose_hidx[B.link_from[k]], hj = B.pose_hidx[B.link_to[k]];
        if (hi >= 0 && hj >= 0 && hi != hj) atomicMin(&B.sky_first[max(hi, hj)], min(hi, hj));
    }
}

// records of k_link_lin -> skyline blocks, reduced rhs and raw b_p (one thread per link; called on ONE rank of a
// partitioned run, before the all-reduce)
__global__ void k_link_add_large(Batch B) {
    if (B.st[0].done) return;
    double *sky = B.red, *gvec = B.red + B.red_g_off, *bpvec = B.red + B.red_bp_off;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < B.tot_link; k += gridDim.x * blockDim.x) {
        const double *rec = B.link_lin + (size_t)k * kLinkStride;
        const int hi = B.pose_hidx[B.link_from[k]], hj = B.pose_hidx[B.link_to[k]];
        for (int side = 0; side < 2; ++side) {
            const int h = side ? hj : hi;
            if (h < 0) continue;
            const double *H = rec + (side ? kLkHjj : kLkHii), *b = rec + (side ? kLkBj : kLkBi);
            double *dblk = sky + (size_t)(B.sky_off[h] + (h - B.sky_first[h])) * 36;
            for (int a = 0; a < 6; ++a) {
                atomicAdd(&gvec[6 * (size_t)h + a], b[a]);
                atomicAdd(&bpvec[6 * (size_t)h + a], b[a]);
                for (int c = 0; c <= a; ++c) atomicAdd(&dblk[a * 6 + c], H[a * 6 + c]);   // lower half
            }
        }
        if (hi >= 0 && hj >= 0) {
            // H_ij = J_i' Omega J_j (row pose i, column pose j); lower block (max, min): rows of the later pose
            const int hr = max(hi, hj), hc = min(hi, hj);
            double *blk = sky + (size_t)(B.sky_off[hr] + (hc - B.sky_first[hr])) * 36;
            for (int a = 0; a < 6; ++a)
                for (int c = 0; c < 6; ++c) {
                    const double v = rec[kLkHij + a * 6 + c];
                    if (hj > hi) atomicAdd(&blk[c * 6 + a], v); else atomicAdd(&blk[a * 6 + c], v);
                }
        }
    }
}

// start of a pass: link part of diag(H_pp) and of chi2 (one CTA; ONE rank of a partitioned run, before the all-reduce)
__global__ void k_link_init_large(Batch B, double *scal) {
    if (B.st[0].done) return;
    __shared__ double red[32];
    double chi = 0.0;
    for (int k = threadIdx.x; k < B.tot_link; k += blockDim.x) {
        const double *rec = B.link_lin + (size_t)k * kLinkStride;
        chi += rec[kLkChi];
        const int hi = B.pose_hidx[B.link_from[k]], hj = B.pose_hidx[B.link_to[k]];
        for (int a = 0; a < 6; ++a) {
            if (hi >= 0) atomicAdd(&B.hdiag[6 * (size_t)hi + a], rec[kLkHii + 7 * a]);
            if (hj >= 0) atomicAdd(&B.hdiag[6 * (size_t)hj + a], rec[kLkHjj + 7 * a]);
        }
    }
    const double t = block_sum(chi, red);
    if (threadIdx.x == 0) scal[0] += t;
}

// chi2 of the links at the trial poses (after a large-path solve kernel has written them); one CTA
__global__ void k_link_chi_large(Batch B) {
    __shared__ double red[32];
    LMState &st = B.st[0];
    if (st.done) return;
    const double c = link_chi2_block(B, B.win[0], 1 - st.cur, red);
    if (threadIdx.x == 0) st.link_chi_trial = c;
}

// ------------------------------------------------------------------------------------------------
// k_solve_pcg: Optimizer/Solver = 2 on the block skyline — g2o LinearSolverPCG (block-Jacobi preconditioner = inverse 6x6
// diagonal blocks, x0 = 0, at most n iterations, the tolerance quirk of k_solve).  Cooperative launch: one warp per block
// row for S d (the lower part is the row's own skyline storage, the upper part the transposed blocks of its column
// structure), dot products as per-CTA partials added in CTA order by every thread that needs them (deterministic),
// three grid.sync() per iteration.  work = [r | d | q | s | x] (5n) + Minv (36F) + dot partials (2 x gridDim).
// ------------------------------------------------------------------------------------------------
constexpr int kPcgThreads = 256;

__device__ __forceinline__ double grid_dot(cg::grid_group &grid, const double *u, const double *v, int n, double *partial, double *red) {
    double acc = 0.0;
    for (int i = blockIdx.x * kPcgThreads + threadIdx.x; i < n; i += gridDim.x * kPcgThreads) acc += u[i] * v[i];
    const double t = block_sum(acc, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
    grid.sync();
    double s = 0.0;
    for (int b = 0; b < (int)gridDim.x; ++b) s += partial[b];   // same order on every thread
    return s;
}

__global__ void __launch_bounds__(kPcgThreads) k_solve_pcg(Batch B, double *work) {
    const WinDesc &wd = B.win[0];
    LMState &st = B.st[0];
    if (st.done) return;
    cg::grid_group grid = cg::this_grid();
    __shared__ double s_red[32];
    const int tid = threadIdx.x, lane = tid & 31;
    const int gwarp = (blockIdx.x * kPcgThreads + tid) >> 5, nwarp = (gridDim.x * kPcgThreads) >> 5;
    const int gtid = blockIdx.x * kPcgThreads + tid, gsize = gridDim.x * kPcgThreads;
    const int F = st.F, n = 6 * F;
    const double lambda = (wd.trust == 0) ? (B.dbg_lambda >= 0.0 ? B.dbg_lambda : st.lambda) : 0.0;
    if (n == 0) {
        if (gtid == 0) { st.ok = 1; st.scale_p = 0.0; }
        return;
    }
    const double *__restrict__ sky = B.red;
    const double *__restrict__ rhs = B.red + B.red_g_off;
    const double *__restrict__ braw = B.red + B.red_bp_off;
    const int *__restrict__ first = B.sky_first;
    const long long *__restrict__ off = B.sky_off;
    double *r = work, *d = work + n, *q = work + 2 * (size_t)n, *sv = work + 3 * (size_t)n, *xv = work + 4 * (size_t)n;
    double *Minv = work + 5 * (size_t)n;
    double *part0 = Minv + 36 * (size_t)F, *part1 = part0 + gridDim.x;

    // block-Jacobi preconditioner: inverse of the damped diagonal blocks (Gauss-Jordan with partial pivoting)
    for (int i = gtid; i < F; i += gsize) {
        const double *blk = sky + (size_t)(off[i] + (i - first[i])) * 36;
        double M[6][12];
        for (int a = 0; a < 6; ++a)
            for (int c = 0; c < 6; ++c) {
                M[a][c] = blk[max(a, c) * 6 + min(a, c)] + (a == c ? lambda : 0.0);
                M[a][6 + c] = (a == c) ? 1.0 : 0.0;
            }
        for (int c = 0; c < 6; ++c) {
            int piv = c;
            for (int a = c + 1; a < 6; ++a) if (fabs(M[a][c]) > fabs(M[piv][c])) piv = a;
            if (piv != c) for (int k = 0; k < 12; ++k) { const double t = M[c][k]; M[c][k] = M[piv][k]; M[piv][k] = t; }
            const double dd = M[c][c];
            for (int k = 0; k < 12; ++k) M[c][k] /= dd;
            for (int a = 0; a < 6; ++a) if (a != c) {
                const double f = M[a][c];
                for (int k = 0; k < 12; ++k) M[a][k] -= f * M[c][k];
            }
        }
        for (int a = 0; a < 6; ++a) for (int c = 0; c < 6; ++c) Minv[36 * (size_t)i + 6 * a + c] = M[a][6 + c];
    }
    for (int i = gtid; i < n; i += gsize) { r[i] = rhs[i]; xv[i] = 0.0; }
    grid.sync();
    auto precond = [&](const double *in, double *out) {
        for (int i = gtid; i < n; i += gsize) {
            const int blk = i / 6, a = i - 6 * blk;
            double acc = 0.0;
            for (int c = 0; c < 6; ++c) acc += Minv[36 * (size_t)blk + 6 * a + c] * in[6 * blk + c];
            out[i] = acc;
        }
    };
    precond(r, d);
    grid.sync();
    double dn = grid_dot(grid, r, d, n, part0, s_red);
    double d0 = 1e-6 * dn;
    const double prev_res = st.pcg_residual;
    if (prev_res > 0.0 && prev_res > 1e-6) d0 = 0.0;
    for (int it = 0; it < n; ++it) {
        if (dn <= d0) break;                       // dn is identical on every thread
        // q = S d: one warp per block row
        for (int row = gwarp; row < F; row += nwarp) {
            double acc[6] = {0, 0, 0, 0, 0, 0};
            const int f0 = first[row], len = row - f0;
            const double *rowblk = sky + (size_t)off[row] * 36;
            for (int ci = lane; ci <= len; ci += 32) {
                const double *blk = rowblk + (size_t)ci * 36;
                const double *x = d + 6 * (size_t)(f0 + ci);
                if (ci < len) {
#pragma unroll
                    for (int a = 0; a < 6; ++a)
#pragma unroll
                        for (int u = 0; u < 6; ++u) acc[a] = fma(blk[a * 6 + u], x[u], acc[a]);
                } else {                           // diagonal block: lower half stored, damping added here
#pragma unroll
                    for (int a = 0; a < 6; ++a)
#pragma unroll
                        for (int u = 0; u < 6; ++u)
                            acc[a] = fma(blk[max(a, u) * 6 + min(a, u)] + (a == u ? lambda : 0.0), x[u], acc[a]);
                }
            }
            const int c0 = B.col_ptr[row], m = B.col_ptr[row + 1] - c0;
            for (int j = lane; j < m; j += 32) {
                const int r2 = B.col_rows[c0 + j];
                const double *blk = sky + (size_t)(off[r2] + (row - first[r2])) * 36;
                const double *x = d + 6 * (size_t)r2;
#pragma unroll
                for (int a = 0; a < 6; ++a)
#pragma unroll
                    for (int u = 0; u < 6; ++u) acc[a] = fma(blk[u * 6 + a], x[u], acc[a]);
            }
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                double v = acc[a];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) q[6 * (size_t)row + a] = v;
            }
        }
        grid.sync();
        const double alpha = dn / grid_dot(grid, d, q, n, part1, s_red);
        for (int i = gtid; i < n; i += gsize) { xv[i] += alpha * d[i]; r[i] -= alpha * q[i]; }
        grid.sync();
        precond(r, sv);
        grid.sync();
        const double dold = dn;
        dn = grid_dot(grid, r, sv, n, part0, s_red);
        const double beta = dn / dold;
        for (int i = gtid; i < n; i += gsize) d[i] = sv[i] + beta * d[i];
        grid.sync();
    }
    if (gtid == 0) st.pcg_residual = 0.5 * dn;
    if (blockIdx.x != 0) return;

    // ---- solution checks, pose step, trial poses, pose part of g2o's computeScale (CTA 0)
    double bad = 0.0;
    for (int i = tid; i < n; i += kPcgThreads) if (!isfinite(xv[i])) bad = 1.0;
    const double anybad = block_sum(bad, s_red);
    const bool ok = anybad == 0.0;
    __syncthreads();
    double sc = 0.0;
    for (int i = tid; i < n; i += kPcgThreads) {
        const double x = ok ? xv[i] : 0.0;
        B.xp[i] = x;
        sc += x * (lambda * x + braw[i]);
    }
    const double scale = block_sum(sc, s_red);
    const int cur = st.cur;
    const double *src = B.pose + (size_t)cur * B.tot_pose * kPoseStride;
    double *dst = B.pose + (size_t)(1 - cur) * B.tot_pose * kPoseStride;
    for (int p = tid; p < wd.n_pose; p += kPcgThreads) {
        const int hi = B.pose_hidx[p];
        if (hi >= 0) {
            double dlt[6];
            for (int a = 0; a < 6; ++a) dlt[a] = ok ? xv[6 * hi + a] : 0.0;
            pose_oplus(src + (size_t)p * kPoseStride, dlt, dst + (size_t)p * kPoseStride);
        }
    }
    if (tid == 0) { st.ok = ok ? 1 : 0; st.scale_p = scale; }
}

// ------------------------------------------------------------------------------------------------
// k_update_large: landmark back-substitution, point oplus, chi2 of the trial state
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreadsL) k_update_large(Batch B, int skip_band) {
    // Several whole landmarks per warp step, one edge per lane (the warp tiles of the small-window k_update, formed on
    // the fly): the warp takes the next landmarks of its block whose edges fit 32 lanes (<= 8 landmarks).  Per-landmark
    // sums go through the warp's shared-memory slab in edge order.
    __shared__ double red[32];
    __shared__ double s_H[kWarpsL][32 * kHs];
    __shared__ double s_lm[kWarpsL][kWtLm * 12];
    __shared__ int s_off[kWarpsL][kWtLm + 1];
    const WinDesc &wd = B.win[0];
    const LMState &st = B.st[0];
    if (st.done) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cur = st.cur;
    const double lambda = (wd.trust == 0) ? (B.dbg_lambda >= 0.0 ? B.dbg_lambda : st.lambda) : 0.0;
    const Intr K = load_intr(wd);
    const double *__restrict__ gpose = B.pose + (size_t)cur * B.tot_pose * kPoseStride;
    const double *__restrict__ gposeT = B.pose + (size_t)(1 - cur) * B.tot_pose * kPoseStride;
    const double *__restrict__ gpoint = B.point + (size_t)cur * B.tot_point * 3;
    double *__restrict__ gpointT = B.point + (size_t)(1 - cur) * B.tot_point * 3;
    double *Hs = s_H[warp], *Ls = s_lm[warp];
    int *lmoff = s_off[warp];
    double chi_acc = 0.0, scale_acc = 0.0;

    // contiguous block of landmarks per warp
    const int nwarp = gridDim.x * kWarpsL, gw = blockIdx.x * kWarpsL + warp;
    const int per = (wd.n_point + nwarp - 1) / nwarp;
    const int l_end = min((gw + 1) * per, wd.n_point);
    for (int lt = min(gw * per, wd.n_point); lt < l_end;) {
        // landmarks lt .. lt + ntl - 1: as many as fit 32 edges (at least one; kMaxDegL bounds a single landmark)
        const int navail = min(kWtLm, l_end - lt);
        const int off_abs = (lane <= navail) ? B.lm_edge_off[lt + lane] : 0x7fffffff;
        const int e0 = __shfl_sync(0xffffffffu, off_abs, 0);
        const int off_l = (lane <= navail) ? off_abs - e0 : 0x7fff;
        int ntl = 1;
#pragma unroll
        for (int l = 2; l <= kWtLm; ++l) {
            const int v = __shfl_sync(0xffffffffu, off_l, l);
            ntl += (l <= navail && v <= 32) ? 1 : 0;
        }
        const int ne = min(__shfl_sync(0xffffffffu, off_l, ntl), 32);
        const int lf_l = (lane < ntl) ? (int)B.lm_flags[lt + lane] : 0;
        const double pt = (lane < 3 * ntl) ? gpoint[3 * (size_t)lt + lane] : 0.0;
        int pw = 0;
        double ou = 0, ov = 0, our = 0;
        if (lane < ne) {
            const int e = e0 + lane;
            pw = B.edge_pose[e];
            ou = B.obs_u[e]; ov = B.obs_v[e]; our = B.obs_r[e];
        }
        if (lane <= ntl) lmoff[lane] = min(off_l, 32);
        int tl = 0;
#pragma unroll
        for (int l = 1; l < kWtLm; ++l) {
            const int v = __shfl_sync(0xffffffffu, off_l, l);
            tl += (l < ntl && v <= lane) ? 1 : 0;
        }
        const int lf = __shfl_sync(0xffffffffu, lf_l, tl);
        const double px = __shfl_sync(0xffffffffu, pt, 3 * tl), py = __shfl_sync(0xffffffffu, pt, 3 * tl + 1),
                     pz = __shfl_sync(0xffffffffu, pt, 3 * tl + 2);
        bool act = false, mono = false, lmfree = false;
        int p = 0;
        double hl[12];
#pragma unroll
        for (int q = 0; q < 12; ++q) hl[q] = 0.0;
        if (lane < ne) {
            p = pw & kPoseMask;
            mono = (pw & kMonoBit) != 0;
            const bool skip = skip_band && (lf & 0x40);   // a landmark of the band chunks: bd::k_update_band owns it (ba_band.cuh)
            lmfree = (lf & kInHessian) != 0 && !skip;
            act = !(pw & kCulledBit) && !((lf & kFixed) && (B.pose_flags[p] & kFixed)) && !skip;
            if (act && lmfree) {
                const int hi = B.pose_hidx[p];
                upd_edge_terms(gpose + (size_t)p * kPoseStride, px, py, pz, ou, ov, our, mono, K, hi >= 0 ? B.xp + 6 * (size_t)hi : nullptr, hl);
            }
#pragma unroll
            for (int q = 0; q < 12; ++q) Hs[lane * kHs + q] = hl[q];
        }
        __syncwarp();
        for (int task = lane; task < ntl * 12; task += 32) {
            const int l = task / 12, q = task - l * 12;
            double sacc = 0.0;
            for (int e = lmoff[l]; e < lmoff[l + 1]; ++e) sacc += Hs[e * kHs + q];
            Ls[task] = sacc;
        }
        __syncwarp();
        if (lane < ne) {
            double np0 = px, np1 = py, np2 = pz;
            if (lmfree) {
                double xl[3];
                const double sc = upd_point_step(Ls + tl * 12, lambda, xl);
                np0 = px + xl[0]; np1 = py + xl[1]; np2 = pz + xl[2];
                if (lane == lmoff[tl]) {      // first edge of the landmark: owner of the point
                    const size_t gl = (size_t)lt + tl;
                    gpointT[3 * gl] = np0; gpointT[3 * gl + 1] = np1; gpointT[3 * gl + 2] = np2;
                    scale_acc += sc;
                }
            }
            if (act) {
                double r0, r1, r2, rho, wgt;
                edge_residual(gposeT + (size_t)p * kPoseStride, np0, np1, np2, ou, ov, our, mono, K, r0, r1, r2);
                huber((r0 * r0 + r1 * r1 + r2 * r2) * K.inv_pv, K.delta, rho, wgt);
                chi_acc += rho;
            }
        }
        __syncwarp();
        lt += ntl;
    }
    const double chi = block_sum(chi_acc, red);
    const double sc = block_sum(scale_acc, red);
    if (tid == 0) { B.part2[2 * (size_t)blockIdx.x] = chi; B.part2[2 * (size_t)blockIdx.x + 1] = sc; }
}

// small helpers for the partitioned (multi-GPU) bookkeeping
__global__ void k_get_counts(Batch B, int *out) { if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = B.st[0].NL; out[1] = B.st[0].err; } }
__global__ void k_set_counts(Batch B, const int *in) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        LMState &st = B.st[0];
        st.NL = in[0];
        st.nNL[st.pass] = in[0];
        if (in[1] != 0 && st.err == 0) st.err = in[1];
    }
}

// dense copy of the damped reduced system for the parity hook: out[n*n] row-major symmetric, rhs[n]
__global__ void k_sky_to_dense(Batch B, double lambda, double *out, double *rhs) {
    const int F = B.st[0].F, n = 6 * F;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < (size_t)n * n; idx += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(idx / n), c = (int)(idx - (size_t)r * n);
        const int rr = max(r, c), cc = min(r, c);
        const int rb = rr / 6, cb = cc / 6;
        double v = 0.0;
        if (cb >= B.sky_first[rb]) v = B.red[(size_t)(B.sky_off[rb] + (cb - B.sky_first[rb])) * 36 + (rr - 6 * rb) * 6 + (cc - 6 * cb)];
        if (r == c) v += lambda;
        out[idx] = v;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) rhs[i] = B.red[B.red_g_off + i];
}

// Schur block pattern of the large path as (col,row)-sorted keys: emitted per landmark, sorted and made unique by the host
__global__ void k_pattern_count(Batch B, int *cnt) {
    const WinDesc &wd = B.win[0];
    for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < wd.n_point; l += gridDim.x * blockDim.x) {
        int h = 0;
        if (B.lm_flags[l] & kInHessian)
            for (int e = B.lm_edge_off[l]; e < B.lm_edge_off[l + 1]; ++e) h += B.pose_hidx[B.edge_pose[e] & kPoseMask] >= 0;
        cnt[l] = h * (h + 1) / 2;
    }
}
__global__ void k_pattern_emit(Batch B, const int *offs, unsigned long long *keys, int key_base) {
    const WinDesc &wd = B.win[0];
    const int F = B.st[0].F;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < F; i += gridDim.x * blockDim.x)
        keys[i] = (unsigned long long)i * (unsigned long long)F + (unsigned long long)i;
    for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < wd.n_point; l += gridDim.x * blockDim.x) {
        if (!(B.lm_flags[l] & kInHessian)) continue;
        int o = key_base + offs[l];
        const int e0 = B.lm_edge_off[l], e1 = B.lm_edge_off[l + 1];
        for (int ea = e0; ea < e1; ++ea) {
            const int ha = B.pose_hidx[B.edge_pose[ea] & kPoseMask];
            if (ha < 0) continue;
            for (int eb = ea; eb < e1; ++eb) {
                const int hb = B.pose_hidx[B.edge_pose[eb] & kPoseMask];
                if (hb < 0) continue;
                keys[o++] = (unsigned long long)hb * (unsigned long long)F + (unsigned long long)ha;   // (col, row)
            }
        }
    }
}

}  // namespace lg
}  // namespace visfs

// ba_build_ds.cuh — build kernel for windows of <= 10 poses (BASELINE C1 / C3, the reference's own 6-frame window): the
// same arithmetic as k_build_ws (linearise + Hessian blocks + Schur partials; g2o's linearizeOplus / constructQuadraticForm /
// BlockSolver::buildSystem + the Schur part of BlockSolver::solve), with the Schur products on the FP64 TENSOR pipe.
//
// What changed against k_build_ws, and why (profiles/r2_ncu_summary.md §1): there every 6x6 block of the reduced system had
// an owner thread that fetched Yn_i and W_j (288 B of shared memory) per landmark for 108 DFMAs — 45 % of the kernel's
// shared-memory wavefronts and most of its issue slots; shared memory (0.60 wavefronts per cycle) and the FP64 pipe (39 %)
// were co-limiting with three warps per scheduler.  Here a tile's W and Yn blocks are stored LANDMARK-STACKED, as the two
// K-major matrices
//        Wt[k][r], Yt[k][r]      k = 3 * (landmark of the tile) + c   (<= 48),   r = 6 * (hessian index of the pose) + a   (< 64)
// so that the tile's contribution  S += Yn_stack * W_stack^T  is a plain (64 x K) x (K x 64) product: 8 x 8 output tiles,
// K in steps of 4, mma.sync.aligned.m8n8k4.f64 (DMMA.8x8x4), accumulators in registers for the whole chunk, fragments
// read straight from the stage with conflict-free 64-bit loads (pitch 68 = 4 mod 16), one A fragment shared by a whole row
// of tiles.  A (pose, landmark) pair without an edge is a zero block, so any visibility pattern is handled.
// Per 12-landmark tile a consumer warp issues ~80 DMMA + ~160 LDS for stage C instead of ~750 instructions per thread.
//
//   8 PRODUCER warps in two groups of 128 threads (one edge each): stage A as before, W to its stacked position.  With the
//     consumers this light the kernel is bound by the latency of the producers' dependent chains, so TWO tiles are produced
//     at once: group g fills stage g with the tiles g, g + 2, ...
//   4 CONSUMER warps, one per scheduler: stage B per edge (damped 3x3 inverse, Yn = -W Dinv into Yt, g), per-pose sums, then
//     the DMMA product over the lower triangle of 8 x 8 tiles (dealt round-robin to the warps), then the stage's W / Yn
//     columns are cleared for its next use.
// Summation order is fixed (deterministic reruns); it differs from k_build_ws in the last bits only (S parity gate 1e-10).
#pragma once
#include "ba_build_ws.cuh"
#include "ba_dense.cuh"

namespace visfs {
namespace ds {

constexpr int kLm = 12;                       // landmarks per tile
constexpr int kEdges = 128;                   // edges per tile = threads of one producer group
constexpr int kGroups = 2;                    // producer groups: group g fills stage g with the tiles t = g, g + 2, ...
constexpr int kProd = kGroups * kEdges, kCons = 128, kThreadsDs = kProd + kCons;   // 8 + 4 warps at 168 registers (3 per scheduler)
constexpr int kMaxPosesDs = 10;               // poses of a window (free + fixed): 6 F <= 60 rows
constexpr int kRows = 64;                     // 8 row tiles
constexpr int kLDR = 68;                      // pitch of the K-major stage matrices (4 mod 16: the fragment loads hit 16 banks)
constexpr int kK = 3 * kLm;                   // 36
constexpr int kMaxTilesPerWarp = 9;           // 36 lower 8 x 8 tiles over 4 consumer warps

enum { BAR_PROD0 = 1, BAR_CONS = 3, BAR_FULL0 = 4, BAR_EMPTY0 = 6 };   // (+ g for the group's own barriers)

struct Stage {
    double Wt[kK * kLDR];
    double Yt[kK * kLDR];
    double H[kEdges * kHStride];              // per edge: H_pp_e (21, upper) | g (6) | b_p (6)
    double hl[kEdges * 9];                    // per edge: H_ll (6) b_l (3) terms, summed per landmark in edge order
    double lm[kLm * 12];
    short slot[kLm * kMaxSmallPoses];
    short emeta[kEdges];                      // -1 = contributes nothing, else tile-local landmark | landmark-in-Hessian << 6 | hessian index << 7
    int lmoff[kLm + 1];
    int ntl, cnt, need_clear;                 // cnt: edges of the tile with a free pose; need_clear: some (pose, landmark) block has no edge
};

struct Smem {
    double pose[kMaxPosesDs * kPoseSm];
    Stage st[kGroups];
    double pacc[kGroups][kMaxPosesDs * kHStride];   // per-pose sums (H_pp, g, b_p) of each producer group's tiles
    int hidx[kMaxSmallPoses];
};
static_assert(sizeof(Smem) <= 232448, "k_build_ds: shared memory over the 227 KB a CTA may use");
static_assert(kGroups * sizeof(Stage) >= sizeof(double) * (kRows * (kRows + 1) + kRows * kRows + 64), "the epilogue aliases S and the partial system onto the stages");

__global__ void __maxnreg__(168) k_build_ds(Batch B, int cluster_size) {
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
    const int tile0 = B.chunk_tile_off[blockIdx.x], ntiles = B.chunk_tile_off[blockIdx.x + 1] - tile0;
    const Tile *__restrict__ tiles = B.tiles + tile0;

    for (int i = tid; i < n_pose * kPoseStride; i += kThreadsDs) sm.pose[(i >> 4) * kPoseSm + (i & 15)] = gpose[i];
    for (int i = tid; i < n_pose; i += kThreadsDs) sm.hidx[i] = B.pose_hidx[pose_off + i];
    for (int i = tid; i < kGroups * kMaxPosesDs * kHStride; i += kThreadsDs) (&sm.pacc[0][0])[i] = 0.0;
    if (tid < kGroups) sm.st[tid].cnt = 0;
    for (int s = 0; s < kGroups; ++s)
        for (int i = tid; i < kK * kLDR; i += kThreadsDs) { sm.st[s].Wt[i] = 0.0; sm.st[s].Yt[i] = 0.0; }
    __syncthreads();

    const int npairs = F * (F + 1) / 2;
    const int nrt = (6 * F + 7) >> 3;              // row tiles of the reduced system
    const bool producer = tid < kProd;
    double *Sfull = reinterpret_cast<double *>(&sm.st[0]);     // epilogue aliases: the full reduced system [64][65] ...
    double *vec = Sfull + kRows * (kRows + 1) + 7;             // ... and this CTA's partial system in the layout k_solve reads

    if (producer) {
        // ======================================================================= producers: stage A, group grp on stage grp
        const int grp = tid / kEdges, gt = tid - grp * kEdges;
        Stage &S = sm.st[grp];
        ws::EdgeRec rec, nxt;
        Tile T, Tn;
        if (grp < ntiles) {
            T = tiles[grp];
            ws::load_edge_l1(B, wd, T, gt, rec);
            ws::load_edge_l2(B, wd, T, gt, gpoint, rec);
        }
        int mine = 0;
        for (int t = grp; t < ntiles; t += kGroups, ++mine) {
            const bool more = t + kGroups < ntiles;
            if (more) { Tn = tiles[t + kGroups]; ws::load_edge_l1(B, wd, Tn, gt, nxt); }   // prefetch, level 1
            if (mine >= 1) ws::bar_sync(BAR_EMPTY0 + grp, kEdges + kCons);                 // the consumers are done with this stage
            const int ne = T.ne, ntl = T.ntl, lt = T.lt;
            for (int i = gt; i < ntl * kMaxSmallPoses; i += kEdges) S.slot[i] = -1;
            if (gt <= ntl) S.lmoff[gt] = min(B.lm_edge_off[lt + gt] - T.e0, kEdges);

            short meta = -1;
            int slot_idx = -1;
            if (gt < ne) {
                const int p = rec.pw & kPoseMask;
                const int tl = rec.gl - lt;
                const bool act = !(rec.pw & kCulledBit) && !((rec.lf & kFixed) && (rec.pf & kFixed));
                const bool lmfree = (rec.lf & kInHessian) != 0;
                double *hl = S.hl + gt * 9;
                EdgeLin lin;
                if (act) edge_linearize(sm.pose + p * kPoseSm, rec.px, rec.py, rec.pz, rec.ou, rec.ov, rec.our,
                                        (rec.pw & kMonoBit) != 0, K, lin);
                const double wo = act ? lin.w * K.inv_pv : 0.0;
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
                const int hi = sm.hidx[p];
                if (hi >= 0 && !act) {   // a culled edge of a free pose: its blocks are written as zeros (a block that is written
                                         // by every tile needs no clearing between tiles)
                    double *wt = S.Wt + (3 * tl) * kLDR + 6 * hi, *yt = S.Yt + (3 * tl) * kLDR + 6 * hi;
#pragma unroll
                    for (int a = 0; a < 6; ++a) {
                        wt[a] = 0.0; wt[kLDR + a] = 0.0; wt[2 * kLDR + a] = 0.0;
                        yt[a] = 0.0; yt[kLDR + a] = 0.0; yt[2 * kLDR + a] = 0.0;
                    }
                    atomicAdd(&S.cnt, 1);
                } else if (hi >= 0) {
                    double *hs = S.H + gt * kHStride;
                    double *wt = S.Wt + (3 * tl) * kLDR + 6 * hi;
                    if (lmfree) {   // W = J_p^T Omega_w J_l to its stacked position
                        double Aj[9];
#pragma unroll
                        for (int q = 0; q < 9; ++q) Aj[q] = wo * lin.Jl[q];
#pragma unroll
                        for (int a = 0; a < 6; ++a)
#pragma unroll
                            for (int c = 0; c < 3; ++c)
                                wt[c * kLDR + a] = fma(lin.Jp[a], Aj[c], fma(lin.Jp[6 + a], Aj[3 + c], lin.Jp[12 + a] * Aj[6 + c]));
                    } else {
#pragma unroll
                        for (int a = 0; a < 6; ++a) { wt[a] = 0.0; wt[kLDR + a] = 0.0; wt[2 * kLDR + a] = 0.0; }
                    }
                    const double wr0 = wo * lin.r[0], wr1 = wo * lin.r[1], wr2 = wo * lin.r[2];
#pragma unroll
                    for (int a = 0; a < 6; ++a) {
                        hs[27 + a] = -fma(lin.Jp[a], wr0, fma(lin.Jp[6 + a], wr1, lin.Jp[12 + a] * wr2));
#pragma unroll
                        for (int c = a; c < 6; ++c)
                            hs[hd_index(a, c)] = wo * fma(lin.Jp[a], lin.Jp[c], fma(lin.Jp[6 + a], lin.Jp[6 + c], lin.Jp[12 + a] * lin.Jp[12 + c]));
                    }
                    slot_idx = tl * kMaxSmallPoses + hi;
                    meta = (short)(tl | (lmfree ? 64 : 0) | (hi << 7));
                    atomicAdd(&S.cnt, 1);
                }
            }
            S.emeta[gt] = meta;
            if (more) ws::load_edge_l2(B, wd, Tn, gt, gpoint, nxt);                    // prefetch, level 2
            ws::bar_sync(BAR_PROD0 + grp, kEdges);
            if (slot_idx >= 0) S.slot[slot_idx] = (short)gt;   // (after the barrier: other threads cleared the table above)
            // per-landmark H_ll (6) / b_l (3): one owner thread per (landmark, entry), edges added in edge order
            for (int task = gt; task < ntl * 9; task += kEdges) {
                const int l = task / 9, q = task - l * 9;
                double s = 0.0;
                for (int e = S.lmoff[l]; e < S.lmoff[l + 1]; ++e) s += S.hl[e * 9 + q];
                S.lm[l * 12 + q] = s;
            }
            // the columns between 3 ntl and the next multiple of 4 are read by the tensor-pipe product: keep them zero
            for (int i = gt; i < (((3 * ntl + 3) & ~3) - 3 * ntl) * kLDR; i += kEdges) { S.Wt[3 * ntl * kLDR + i] = 0.0; S.Yt[3 * ntl * kLDR + i] = 0.0; }
            ws::bar_sync(BAR_PROD0 + grp, kEdges);
            // stage B of this thread's edge: damped inverse of its landmark block (redundantly per edge), Yn = -W Dinv to its
            // stacked position, g = b_p_e - W Dinv b_l
            if (meta >= 0) {
                const int tl = meta & 63, hi = meta >> 7;
                double *hs = S.H + gt * kHStride;
                double *yt = S.Yt + (3 * tl) * kLDR + 6 * hi;
                if (meta & 64) {
                    const double *ls = S.lm + tl * 12;
                    double A[6] = {ls[0] + lambda, ls[1], ls[2], ls[3] + lambda, ls[4], ls[5] + lambda};
                    const double bl[3] = {ls[6], ls[7], ls[8]};
                    double Di[6], db[3];
                    inv_sym3(A, Di);
                    sym3_mul(Di, bl, db);
                    const double *wt = S.Wt + (3 * tl) * kLDR + 6 * hi;
#pragma unroll
                    for (int a = 0; a < 6; ++a) {
                        const double w0 = wt[a], w1 = wt[kLDR + a], w2 = wt[2 * kLDR + a];
                        yt[a] = -fma(w0, Di[0], fma(w1, Di[1], w2 * Di[2]));
                        yt[kLDR + a] = -fma(w0, Di[1], fma(w1, Di[3], w2 * Di[4]));
                        yt[2 * kLDR + a] = -fma(w0, Di[2], fma(w1, Di[4], w2 * Di[5]));
                        hs[21 + a] = hs[27 + a] - fma(w0, db[0], fma(w1, db[1], w2 * db[2]));
                    }
                } else {
#pragma unroll
                    for (int a = 0; a < 6; ++a) { yt[a] = 0.0; yt[kLDR + a] = 0.0; yt[2 * kLDR + a] = 0.0; hs[21 + a] = hs[27 + a]; }
                }
            }
            if (gt == 0) { S.ntl = ntl; S.need_clear = (S.cnt != ntl * F) ? 1 : 0; S.cnt = 0; }
            ws::bar_sync(BAR_PROD0 + grp, kEdges);
            // per-pose sums of the tile (H_pp, g, b_p): fixed owner per (pose, entry), landmarks in order
            for (int task = gt; task < F * kHStride; task += kEdges) {
                const int i = task / kHStride, k = task - i * kHStride;
                double s = 0.0;
                for (int q = 0; q < ntl; ++q) {
                    const int sl = S.slot[q * kMaxSmallPoses + i];
                    if (sl >= 0) s += S.H[sl * kHStride + k];
                }
                sm.pacc[grp][task] += s;
            }
            __threadfence_block();
            ws::bar_arrive(BAR_FULL0 + grp, kEdges + kCons);
            if (more) { T = Tn; rec = nxt; }
        }
        if (mine >= 1) ws::bar_sync(BAR_EMPTY0 + grp, kEdges + kCons);   // the consumers' last arrival on this stage
    } else {
        // ======================================================================= consumers: stage B, per-pose sums, stage C (DMMA)
        const int ctid = tid - kProd;
        const int warp = ctid >> 5, lane = ctid & 31, gl = lane >> 2, tl4 = lane & 3;
        // the lower 8 x 8 tiles of the reduced system, dealt round-robin to the four warps: tile q of this warp is
        // (R, C) = the (warp + 4 q)-th pair with C <= R < nrt, kept as shared-memory offsets (8 R, 8 C; -1: none)
        const int n_lower = nrt * (nrt + 1) / 2;
        int offR[kMaxTilesPerWarp], offC[kMaxTilesPerWarp];
#pragma unroll
        for (int q = 0; q < kMaxTilesPerWarp; ++q) {
            const int idx = warp + 4 * q;
            int R = 0;
            while ((R + 1) * (R + 2) / 2 <= idx) ++R;
            offR[q] = (idx < n_lower) ? 8 * R : -1;
            offC[q] = 8 * (idx - R * (R + 1) / 2);
        }
        double acc[kMaxTilesPerWarp][2];
#pragma unroll
        for (int q = 0; q < kMaxTilesPerWarp; ++q) { acc[q][0] = 0.0; acc[q][1] = 0.0; }
        for (int t = 0; t < ntiles; ++t) {
            const int grp = t & 1;
            Stage &S = sm.st[grp];
            ws::bar_sync(BAR_FULL0 + grp, kEdges + kCons);
            const int ntl = S.ntl;
#ifndef VISFS_DS_SKIP_C
            {   // stage C: S += Yn_stack * W_stack^T over the K = 3 ntl columns of this tile (padded to a multiple of 4 with zeros);
                // the fragments of the next k-step are requested before the tensor-pipe instructions of the current one
                const int kend = (3 * ntl + 3) & ~3;
                const double *yb = S.Yt + tl4 * kLDR + gl;
                const double *wb = S.Wt + tl4 * kLDR + gl;
                double a[kMaxTilesPerWarp], b[kMaxTilesPerWarp], an[kMaxTilesPerWarp], bn[kMaxTilesPerWarp];
#pragma unroll
                for (int q = 0; q < kMaxTilesPerWarp; ++q) {
                    a[q] = (offR[q] >= 0) ? yb[offR[q]] : 0.0;
                    b[q] = (offR[q] >= 0) ? wb[offC[q]] : 0.0;
                }
                for (int kk = 0; kk < kend; kk += 4) {
                    const bool more_k = kk + 4 < kend;
#pragma unroll
                    for (int q = 0; q < kMaxTilesPerWarp; ++q) {
                        an[q] = (more_k && offR[q] >= 0) ? yb[(kk + 4) * kLDR + offR[q]] : 0.0;
                        bn[q] = (more_k && offR[q] >= 0) ? wb[(kk + 4) * kLDR + offC[q]] : 0.0;
                    }
#pragma unroll
                    for (int q = 0; q < kMaxTilesPerWarp; ++q) if (offR[q] >= 0) dn::dmma884(acc[q][0], acc[q][1], a[q], b[q]);
#pragma unroll
                    for (int q = 0; q < kMaxTilesPerWarp; ++q) { a[q] = an[q]; b[q] = bn[q]; }
                }
            }
#endif
            if (S.need_clear) {   // some (pose, landmark) pair of this tile had no edge: its block must not survive into the next tile
                ws::bar_sync(BAR_CONS, kCons);
                for (int i = ctid; i < 3 * ntl * kLDR; i += kCons) { S.Wt[i] = 0.0; S.Yt[i] = 0.0; }
            }
            __threadfence_block();
            ws::bar_arrive(BAR_EMPTY0 + grp, kEdges + kCons);
        }
        // the tiles into the full reduced system (lower tiles; lane l holds C[l >> 2][2 (l & 3) .. + 1]); the producers may still
        // be at their last barrier, but they touch the stages no more
        ws::bar_sync(BAR_CONS, kCons);
#pragma unroll
        for (int q = 0; q < kMaxTilesPerWarp; ++q)
            if (offR[q] >= 0) { double *p = Sfull + (offR[q] + gl) * (kRows + 1) + offC[q] + 2 * tl4; p[0] = acc[q][0]; p[1] = acc[q][1]; }
    }
    __syncthreads();

    // ---- epilogue: the layout k_solve reads (blocks (i <= j) x 36, then per pose H_pp g b_p) -> (cluster sum) -> global partial
    const int offd = npairs * 36;
    const int NP = offd + F * kHStride;
    for (int idx = tid; idx < offd; idx += kThreadsDs) {
        const int pt = idx / 36, q = idx - pt * 36, a = q / 6, c = q - a * 6;
        int i = 0, base = 0;
        while (base + (F - i) <= pt) { base += F - i; ++i; }
        const int j = i + (pt - base);
        const int r = 6 * i + a, cc = 6 * j + c;          // entry (a, c) of block (i, j), i <= j: S(r, cc); tiles hold rows >= columns
        vec[idx] = ((r >> 3) >= (cc >> 3)) ? Sfull[r * (kRows + 1) + cc] : Sfull[cc * (kRows + 1) + r];
    }
    for (int task = tid; task < F * kHStride; task += kThreadsDs) vec[offd + task] = sm.pacc[0][task] + sm.pacc[1][task];
    __syncthreads();
    double *part = B.part + wd.part_off + (size_t)((blockIdx.x - wd.chunk_off) / cluster_size) * wd.part_stride;
    if (cluster_size == 1) {
        for (int idx = tid; idx < NP; idx += kThreadsDs) part[idx] = vec[idx];
    } else {
        ws::cg::cluster_group cluster = ws::cg::this_cluster();
        cluster.sync();
        const int rank = (int)cluster.block_rank();
        for (int idx = rank * kThreadsDs + tid; idx < NP; idx += kThreadsDs * cluster_size) {
            double s = 0.0;
            for (int r = 0; r < cluster_size; ++r) s += cluster.map_shared_rank(vec, r)[idx];
            part[idx] = s;
        }
        cluster.sync();
    }
}

}  // namespace ds
}  // namespace visfs

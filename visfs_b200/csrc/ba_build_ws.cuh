// ba_build_ws.cuh — warp-specialised build kernel (linearise + Hessian blocks + Schur partials), F <= 20.
//
// Same arithmetic as k_build<MODE_BUILD> (ba_kernels.cuh), restructured for the SM:
//   * 6 PRODUCER warps (192 threads, one edge each): stage A (residual, Jacobians, Huber, per-edge H_ll / b_l
//     terms, W = H_pl block, H_pp_e, b_p_e) and the per-landmark sums in edge order by (landmark, entry) owner
//     threads, into one of two shared-memory stages.  The edge records and points of the NEXT tile are
//     prefetched into registers while the current tile is computed; tile bounds come from a table built once
//     at upload (k_fill_tiles).
//   * 6 CONSUMER warps (192 threads): stage B, one edge per thread (damped 3x3 inverse, Yn = -W Dinv, g), then
//     stage C: every block (i <= j) of the reduced camera system has a fixed owner thread that keeps its 36
//     entries in registers for the whole chunk and adds Yn_i W_j^T with three chained DFMAs per entry;
//     per-pose sums (H_pp, g, b_p) have fixed owners as well.
//   The two roles overlap on different tiles through full / empty named barriers (bar.arrive / bar.sync),
//   so neither the 72 accumulator registers nor the linearisation temporaries are ever live in the same
//   thread (v1 spilled 900 B / thread to L2), and the FP64 pipe is fed by both roles at once.
//   12 warps = 3 per SM sub-partition (16 K registers each) at 168 registers per thread.
//   * Thread-block clusters: when ONE window is split over many chunks, the CTAs of a cluster add their
//     partial systems through distributed shared memory in rank order and write one partial per cluster,
//     so k_solve reads fewer partials.  No atomics; every sum has a fixed order.
#pragma once
#include <cooperative_groups.h>
#include "ba_kernels.cuh"

namespace visfs {
namespace ws {

namespace cg = cooperative_groups;

constexpr int kEdgeThreads = kTileEdges;      // 192
constexpr int kPairThreads = 192;
constexpr int kThreadsWs = kEdgeThreads + kPairThreads;   // 384
constexpr int kMaxFreeWs = 19;                // free poses F: F (F + 1) / 2 <= 190 block owners
constexpr int kMaxPosesWs = 20;               // poses of a window (free + fixed)

enum { BAR_PROD = 1, BAR_CONS = 2, BAR_FULL = 3, BAR_EMPTY = 5 };

// (whole warps take every barrier; __syncwarp reconverges the warp after the divergent per-edge code)
__device__ __forceinline__ void bar_sync(int id, int n) { __syncwarp(); asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { __syncwarp(); asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

struct Stage {
    double W[kTileEdges * 18];
    double Yn[kTileEdges * 18];
    double H[kTileEdges * kHStride];           // H_pp_e (21, upper) | g (6) | b_p (6); stage A scratch before that
    double lm[kTileLm * 12];                   // per landmark: H_ll (6) b_l (3), summed in edge order
    short slot[kTileLm * kMaxSmallPoses];
    short emeta[kTileEdges];                   // per edge: -1 = contributes nothing, else tile-local landmark | landmark-in-Hessian << 6
    int ntl, pad0, pad1, pad2;
};

struct Smem {
    double pose[kMaxPosesWs * kPoseSm];
    Stage st[2];
    double pacc[kMaxFreeWs * kHStride];
    int hidx[kMaxSmallPoses];
    int lmoff[kTileLm + 1];
    short slot_pose[kTileEdges];               // regular chunks: hessian index of the pose of edge slot t (-1: fixed pose / no edge)
};
constexpr int kStageDoubles = (int)(sizeof(Stage) / sizeof(double));
static_assert(sizeof(Smem) <= 232448, "k_build_ws: shared memory over the 227 KB a CTA may use");

// number of doubles of one partial system in the "all pairs" layout
__host__ __device__ __forceinline__ int part_len(int F) { return F * (F + 1) / 2 * 36 + F * kHStride; }

// ---- band chunks (ba_band.cuh): the same kernel on pieces of a LARGE window.  The landmarks are taken in the order of
// their first pose; a band chunk is a contiguous piece of that order whose landmarks touch <= kBandPoses poses in all, so
// inside the chunk the poses are numbered 0 .. n_pose - 1 and everything above applies.  Its partial system goes to the
// chunk's own slot of `part` and k_band_gather adds the slots into the block skyline in a fixed order (no atomics).
constexpr int kBandPoses = 19;                 // poses of a band chunk, free or fixed (<= kMaxFreeWs block owners per row)
constexpr int kBandPartStride = 7472;          // doubles per chunk slot >= part_len(19) = 7467
struct BandChunk {
    int lm0, lm1;                              // landmark range in the sorted order
    int n_pose, F;                             // poses the chunk touches / the free ones among them; n_pose < 0: not a band chunk
    int regular, pad;                          // 1: all landmarks of the chunk are seen by the same poses in the same order
};
struct Band {
    const BandChunk *chunk;
    const int *chunk_pose;                     // [n_chunk][kBandPoses] global pose index, ascending
    const int *sorted_off;                     // [L + 1] edge offsets of the landmarks in the sorted order
    // per-pass copies of the edge records in the sorted order, so that the kernel's prefetch has no dependent address:
    const int *s_pw;                           // [E] pose word (mono / culled bits) with the CHUNK-LOCAL pose index
    const int *s_gl;                           // [E] landmark
    const int *s_sl;                           // [E] sorted landmark index
    const double *s_ou, *s_ov, *s_our;         // [E] observation
    const int4 *lm_rec;                        // [L] sorted landmark -> (landmark, first edge, degree, flags)
    const Tile *tiles;
    const int *chunk_tile_off;
    const int *order;                          // [n_chunk] launch order: chunks by decreasing cost (tiles x block pairs)
    double *part;                              // [n_chunk][kBandPartStride]
    int n_chunk;
};

// what a producer thread holds for its edge of one tile
struct EdgeRec {
    double ou, ov, our, px, py, pz;
    int pw, gl, sl;                            // sl: sorted landmark index (band chunks only)
    uint8_t lf, pf;
};

template <bool BAND>
__device__ __forceinline__ void load_edge_l1_t(const Batch &B, const WinDesc &wd, const Band &bd, const Tile &T, int tid, EdgeRec &r) {
    if (BAND) {
        if (tid < T.ne) {
            const int k = T.e0 + tid;
            r.sl = bd.s_sl[k];
            r.pw = bd.s_pw[k];
            r.gl = bd.s_gl[k];
            r.ou = bd.s_ou[k]; r.ov = bd.s_ov[k]; r.our = bd.s_our[k];
        }
        return;
    }
    if (tid < T.ne) {
        const int e = T.e0 + tid;
        r.pw = B.edge_pose[e];
        r.gl = wd.point_off + B.edge_point[e];
        r.ou = B.obs_u[e]; r.ov = B.obs_v[e]; r.our = B.obs_r[e];
    }
}
template <bool BAND>
__device__ __forceinline__ void load_edge_l2_t(const Batch &B, const WinDesc &wd, const int *__restrict__ cpose, const Tile &T, int tid,
                                               const double *gpoint, EdgeRec &r) {
    if (tid < T.ne) {
        r.lf = B.lm_flags[r.gl];
        r.pf = B.pose_flags[wd.pose_off + (BAND ? cpose[r.pw & kPoseMask] : (r.pw & kPoseMask))];
        r.px = gpoint[3 * (size_t)r.gl]; r.py = gpoint[3 * (size_t)r.gl + 1]; r.pz = gpoint[3 * (size_t)r.gl + 2];
    }
}
__device__ __forceinline__ void load_edge_l1(const Batch &B, const WinDesc &wd, const Tile &T, int tid, EdgeRec &r) {
    load_edge_l1_t<false>(B, wd, Band{}, T, tid, r);
}
__device__ __forceinline__ void load_edge_l2(const Batch &B, const WinDesc &wd, const Tile &T, int tid, const double *gpoint, EdgeRec &r) {
    load_edge_l2_t<false>(B, wd, nullptr, T, tid, gpoint, r);
}

// stage B of one edge: damped inverse of its landmark block (redundantly per edge), Yn = -W Dinv over the scratch stage A
// left in the Yn row, g = b_p_e - W Dinv b_l.  Needs the landmark sums of the tile (S.lm).
__device__ __forceinline__ void stage_b_edge(Stage &S, int e, double lambda) {
    const int meta = S.emeta[e];
    if (meta < 0) return;
    double *ys = S.Yn + e * 18;
    double *hs = S.H + e * kHStride;
    if (meta & 64) {
        const double *ls = S.lm + (meta & 63) * 12;
        double A[6] = {ls[0] + lambda, ls[1], ls[2], ls[3] + lambda, ls[4], ls[5] + lambda};
        const double bl[3] = {ls[6], ls[7], ls[8]};
        double Di[6], db[3];
        inv_sym3(A, Di);
        sym3_mul(Di, bl, db);
        const double *wsrc = S.W + e * 18;
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            const double w0 = wsrc[a * 3], w1 = wsrc[a * 3 + 1], w2 = wsrc[a * 3 + 2];
            ys[a * 3 + 0] = -fma(w0, Di[0], fma(w1, Di[1], w2 * Di[2]));
            ys[a * 3 + 1] = -fma(w0, Di[1], fma(w1, Di[3], w2 * Di[4]));
            ys[a * 3 + 2] = -fma(w0, Di[2], fma(w1, Di[4], w2 * Di[5]));
            hs[21 + a] = hs[27 + a] - fma(w0, db[0], fma(w1, db[1], w2 * db[2]));
        }
    } else {
#pragma unroll
        for (int q = 0; q < 18; ++q) ys[q] = 0.0;
#pragma unroll
        for (int a = 0; a < 6; ++a) hs[21 + a] = hs[27 + a];
    }
}

// Regular chunks (k_chunk_regular): edge slot e belongs to the same pose in every tile, so the per-pose sums need no slot table.
// The producers ADD H_pp_e (21) and b_p_e (6) into the H row of the slot (it persists over the tiles of its stage) and leave this
// tile's b_p_e in entries 21..26; stage B reads it there and adds g = b_p_e - W Dinv b_l into the consumer thread's registers.
constexpr int kRegScratch = 5760;              // doubles of stage 0 the group partials may use; the g accumulators follow (192 x 6)
constexpr int kRegMaxFree = 16;                // the CTA's partial system must stay inside the W / Yn part of stage 1: F <= 16
__device__ __forceinline__ void stage_b_edge_reg(Stage &S, int e, double lambda, double *gacc) {
    const int meta = S.emeta[e];
    if (meta < 0) return;
    double *ys = S.Yn + e * 18;
    const double *hs = S.H + e * kHStride;
    if (meta & 64) {
        const double *ls = S.lm + (meta & 63) * 12;
        double A[6] = {ls[0] + lambda, ls[1], ls[2], ls[3] + lambda, ls[4], ls[5] + lambda};
        const double bl[3] = {ls[6], ls[7], ls[8]};
        double Di[6], db[3];
        inv_sym3(A, Di);
        sym3_mul(Di, bl, db);
        const double *wsrc = S.W + e * 18;
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            const double w0 = wsrc[a * 3], w1 = wsrc[a * 3 + 1], w2 = wsrc[a * 3 + 2];
            ys[a * 3 + 0] = -fma(w0, Di[0], fma(w1, Di[1], w2 * Di[2]));
            ys[a * 3 + 1] = -fma(w0, Di[1], fma(w1, Di[3], w2 * Di[4]));
            ys[a * 3 + 2] = -fma(w0, Di[2], fma(w1, Di[4], w2 * Di[5]));
            gacc[a] += hs[21 + a] - fma(w0, db[0], fma(w1, db[1], w2 * db[2]));
        }
    } else {
#pragma unroll
        for (int q = 0; q < 18; ++q) ys[q] = 0.0;
#pragma unroll
        for (int a = 0; a < 6; ++a) gacc[a] += hs[21 + a];
    }
}

// Of every 4 edges of a tile, how many get their stage B from the producers (after one more producer barrier) instead of
// the consumers.  Measured on C3 x 512 (ms per step): 0 -> 22.7, 1 -> 23.7, 2 -> 23.7; the old all-producer design 23.1.
#ifndef VISFS_WS_PROD_SHARE
#define VISFS_WS_PROD_SHARE 0
#endif
constexpr int kProdShare = VISFS_WS_PROD_SHARE;

template <bool BAND>
__device__ __forceinline__ void build_ws_body(const Batch &B, int cluster_size, const Band &bd) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    const int tid = threadIdx.x;
    const int cidx = BAND ? bd.order[blockIdx.x] : (int)blockIdx.x;   // band chunks are launched longest first
    const int win = BAND ? 0 : B.chunks[cidx].win;
    const WinDesc &wd = B.win[win];
    const LMState &st = B.st[win];
    if (st.done) return;   // uniform over the cluster: all its chunks belong to one window
    BandChunk bc{0, 0, 0, 0, 0, 0};
    if (BAND) { bc = bd.chunk[cidx]; if (bc.n_pose < 0) return; }
    const int *__restrict__ cpose = BAND ? bd.chunk_pose + (size_t)cidx * kBandPoses : nullptr;
    const int cur = st.cur;
    const int F = BAND ? bc.F : st.F;
    const double lambda = (wd.trust == 0) ? (B.dbg_lambda >= 0.0 ? B.dbg_lambda : st.lambda) : 0.0;
    const Intr K = load_intr(wd);
    const int pose_off = wd.pose_off, n_pose = BAND ? bc.n_pose : wd.n_pose;
    const double *__restrict__ gpose = B.pose + ((size_t)cur * B.tot_pose + pose_off) * kPoseStride;
    const double *__restrict__ gpoint = B.point + (size_t)cur * B.tot_point * 3;
    const int *__restrict__ tile_off = BAND ? bd.chunk_tile_off : B.chunk_tile_off;
    const int *__restrict__ lm_off = BAND ? bd.sorted_off : B.lm_edge_off;
    const int tile0 = tile_off[cidx], ntiles = tile_off[cidx + 1] - tile0;
    const Tile *__restrict__ tiles = (BAND ? bd.tiles : B.tiles) + tile0;

    if (BAND) {   // the chunk's poses, numbered 0 .. n_pose - 1 in ascending pose order; the free ones 0 .. F - 1 in that order
        for (int i = tid; i < n_pose * kPoseStride; i += kThreadsWs) sm.pose[(i >> 4) * kPoseSm + (i & 15)] = gpose[(size_t)cpose[i >> 4] * kPoseStride + (i & 15)];
        if (tid < 32) {   // free poses of the chunk numbered by a ballot over their global hessian indices
            const bool fr = tid < n_pose && B.pose_hidx[pose_off + cpose[tid]] >= 0;
            const unsigned m = __ballot_sync(0xffffffffu, fr);
            if (tid < n_pose) sm.hidx[tid] = fr ? __popc(m & ((1u << tid) - 1u)) : -1;
        }
    } else {
        for (int i = tid; i < n_pose * kPoseStride; i += kThreadsWs) sm.pose[(i >> 4) * kPoseSm + (i & 15)] = gpose[i];
        for (int i = tid; i < n_pose; i += kThreadsWs) sm.hidx[i] = B.pose_hidx[pose_off + i];
    }
    for (int i = tid; i < kMaxFreeWs * kHStride; i += kThreadsWs) sm.pacc[i] = 0.0;
    // (from 4 tiles up: clearing the accumulators and reducing them by pose costs about as much as two tiles' worth of slot-table sums)
    const bool reg = kProdShare == 0 && ntiles >= 4 && F <= kRegMaxFree &&
                     (BAND ? bc.regular != 0 : (B.chunk_regular != nullptr && B.chunk_regular[blockIdx.x] != 0));
    if (reg)
        for (int i = tid; i < kTileEdges * kHStride; i += kThreadsWs) { sm.st[0].H[i] = 0.0; sm.st[1].H[i] = 0.0; }
    __syncthreads();

    const int npairs = F * (F + 1) / 2;
    int G = 1;
    if (npairs > 0) {
        G = kPairThreads / npairs;
        if (G > kTileLm) G = kTileLm;
        const int cap = 1 + (reg ? kRegScratch : kStageDoubles) / (npairs * 36);
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
        EdgeRec rec, nxt;
        Tile T, Tn;
        if (ntiles > 0) {
            T = tiles[0];
            load_edge_l1_t<BAND>(B, wd, bd, T, tid, rec);
            load_edge_l2_t<BAND>(B, wd, cpose, T, tid, gpoint, rec);
        }
        for (int t = 0; t < ntiles; ++t) {
            Stage &S = sm.st[t & 1];
            const bool more = t + 1 < ntiles;
            if (more) { Tn = tiles[t + 1]; load_edge_l1_t<BAND>(B, wd, bd, Tn, tid, nxt); }   // prefetch, level 1
            if (t >= 2) bar_sync(BAR_EMPTY + (t & 1), kThreadsWs);
            const int ne = T.ne, ntl = T.ntl, lt = T.lt;
            for (int i = tid; i < ntl * kMaxSmallPoses; i += kEdgeThreads) S.slot[i] = -1;
            if (tid <= ntl) sm.lmoff[tid] = min(lm_off[lt + tid] - T.e0, kTileEdges);

            // stage A, everything that needs this edge only: residual, Jacobians, Huber weight; the H_ll / b_l terms go to
            // the (still unused) Yn rows as scratch, W, H_pp_e and b_p_e to their final place.  Nothing of the
            // linearisation stays live across the barrier; the landmark-level half (damped inverse, Yn, g) is done by the
            // consumers, who would otherwise wait for this stage.
            short meta = -1;
            int slot_idx = -1;
            if (tid < ne) {
                const int p = rec.pw & kPoseMask;
                const int tl = (BAND ? rec.sl : rec.gl) - lt;
                const bool act = !(rec.pw & kCulledBit) && !((rec.lf & kFixed) && (rec.pf & kFixed));
                const bool lmfree = (rec.lf & kInHessian) != 0;
                double *hl = S.Yn + tid * 18;
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
                const int hi = act ? sm.hidx[p] : -1;
                if (hi >= 0) {
                    double *hs = S.H + tid * kHStride;
                    double *ws = S.W + tid * 18;
                    if (lmfree) {
                        double Aj[9];
#pragma unroll
                        for (int q = 0; q < 9; ++q) Aj[q] = wo * lin.Jl[q];
#pragma unroll
                        for (int a = 0; a < 6; ++a)
#pragma unroll
                            for (int c = 0; c < 3; ++c)
                                ws[a * 3 + c] = fma(lin.Jp[a], Aj[c], fma(lin.Jp[6 + a], Aj[3 + c], lin.Jp[12 + a] * Aj[6 + c]));
                    } else {
#pragma unroll
                        for (int q = 0; q < 18; ++q) ws[q] = 0.0;
                    }
                    const double wr0 = wo * lin.r[0], wr1 = wo * lin.r[1], wr2 = wo * lin.r[2];
                    if (reg) {
#pragma unroll
                        for (int a = 0; a < 6; ++a) {
                            const double bp = -fma(lin.Jp[a], wr0, fma(lin.Jp[6 + a], wr1, lin.Jp[12 + a] * wr2));
                            hs[21 + a] = bp;
                            hs[27 + a] += bp;
#pragma unroll
                            for (int c = a; c < 6; ++c)
                                hs[hd_index(a, c)] += wo * fma(lin.Jp[a], lin.Jp[c], fma(lin.Jp[6 + a], lin.Jp[6 + c], lin.Jp[12 + a] * lin.Jp[12 + c]));
                        }
                    } else {
#pragma unroll
                        for (int a = 0; a < 6; ++a) {
                            hs[27 + a] = -fma(lin.Jp[a], wr0, fma(lin.Jp[6 + a], wr1, lin.Jp[12 + a] * wr2));
#pragma unroll
                            for (int c = a; c < 6; ++c)
                                hs[hd_index(a, c)] = wo * fma(lin.Jp[a], lin.Jp[c], fma(lin.Jp[6 + a], lin.Jp[6 + c], lin.Jp[12 + a] * lin.Jp[12 + c]));
                        }
                    }
                    slot_idx = tl * kMaxSmallPoses + hi;
                    meta = (short)(tl | (lmfree ? 64 : 0));
                }
                if (reg && t == 0) sm.slot_pose[tid] = (short)sm.hidx[p];   // structure, not activity: the same in every tile
            } else if (reg && t == 0) {
                sm.slot_pose[tid] = (short)-1;
            }
            S.emeta[tid] = meta;
            if (more) load_edge_l2_t<BAND>(B, wd, cpose, Tn, tid, gpoint, nxt);     // prefetch, level 2
            bar_sync(BAR_PROD, kEdgeThreads);
            if (slot_idx >= 0) S.slot[slot_idx] = (short)tid;   // (after the barrier: other threads cleared the table above)
            // per-landmark H_ll (6) / b_l (3): one owner thread per (landmark, entry), edges added in edge order
            for (int task = tid; task < ntl * 9; task += kEdgeThreads) {
                const int l = task / 9, q = task - l * 9;
                double s = 0.0;
                for (int e = sm.lmoff[l]; e < sm.lmoff[l + 1]; ++e) s += S.Yn[e * 18 + q];
                S.lm[l * 12 + q] = s;
                if (!BAND && B.lm_sum) B.lm_sum[(size_t)(lt + l) * 9 + q] = s;   // k_update needs the same sums: it reads them instead of forming them again
            }
            if (kProdShare > 0) {
                bar_sync(BAR_PROD, kEdgeThreads);
                if ((tid & 3) < kProdShare) stage_b_edge(S, tid, lambda);
            }
            if (tid == 0) S.ntl = ntl;
            __threadfence_block();
            bar_arrive(BAR_FULL + (t & 1), kThreadsWs);
            if (more) { T = Tn; rec = nxt; }
        }
        for (int tt = max(0, ntiles - 2); tt < ntiles; ++tt) bar_sync(BAR_EMPTY + (tt & 1), kThreadsWs);   // join the consumers
    } else {
        // ======================================================================= consumers: stage C
        double acc[36];   // lives only in the consumer branch: never competes with the producers' registers
#pragma unroll
        for (int q = 0; q < 36; ++q) acc[q] = 0.0;
        double gacc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};   // regular chunks: g of this thread's edge slot, summed over the tiles
        int pi = -1, pj = -1;
        if (grp < G && pt < npairs) {
            int i = 0, base = 0;
            while (base + (F - i) <= pt) { base += F - i; ++i; }
            pi = i; pj = i + (pt - base);
        }
        for (int t = 0; t < ntiles; ++t) {
            Stage &S = sm.st[t & 1];
            bar_sync(BAR_FULL + (t & 1), kThreadsWs);
            const int ntl = S.ntl;
            if (reg) stage_b_edge_reg(S, ctid, lambda, gacc);
            else if ((ctid & 3) >= kProdShare) stage_b_edge(S, ctid, lambda);
            bar_sync(BAR_CONS, kPairThreads);
            if (!reg)
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
            bar_arrive(BAR_EMPTY + (t & 1), kThreadsWs);
        }
        // groups -> one block per owner (consumer-only barriers; the stages are no longer written)
        bar_sync(BAR_CONS, kPairThreads);
        if (reg) {
#pragma unroll
            for (int a = 0; a < 6; ++a) scratch[kRegScratch + ctid * 6 + a] = gacc[a];
        }
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
    if (reg) {
        // per-pose sums from the slot accumulators: both stages, slots in order; g from the consumers' registers
        for (int task = tid; task < F * kHStride; task += kThreadsWs) {
            const int i = task / kHStride, k = task - i * kHStride;
            double s = 0.0;
            if (k >= 21 && k < 27) {
                for (int t = 0; t < kTileEdges; ++t) if (sm.slot_pose[t] == i) s += scratch[kRegScratch + t * 6 + (k - 21)];
            } else {
                for (int t = 0; t < kTileEdges; ++t) if (sm.slot_pose[t] == i) s += sm.st[0].H[t * kHStride + k] + sm.st[1].H[t * kHStride + k];
            }
            sm.pacc[task] = s;
        }
        __syncthreads();
    }
    for (int task = tid; task < F * kHStride; task += kThreadsWs) vec[offd + task] = sm.pacc[task];
    __syncthreads();
    double *part = BAND ? bd.part + (size_t)cidx * kBandPartStride
                        : B.part + wd.part_off + (size_t)((blockIdx.x - wd.chunk_off) / cluster_size) * wd.part_stride;
    if (BAND || cluster_size == 1) {
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

__global__ void __maxnreg__(168) k_build_ws(Batch B, int cluster_size) { build_ws_body<false>(B, cluster_size, Band{}); }

// the same kernel on the band chunks of a large window (ba_band.cuh); one CTA per chunk, no clusters
__global__ void __maxnreg__(168) k_build_band(Batch B, Band bd) { build_ws_body<true>(B, 1, bd); }

}  // namespace ws
}  // namespace visfs

// ba_mf.cuh — multifrontal Cholesky of a BANDED reduced camera system (BASELINE config C4: 2 000 key frames on a loop, every
// landmark seen by 10 consecutive frames => S is block-banded with half-width 9 plus a few long "loop closure" rows).
//
// Why: a band Cholesky is a chain of F dependent block columns; round 1 ran it on ONE SM (lg::k_solve_front, 3.3 us per
// block column = 6.7 ms per solve, 65 % of C4 on one GPU and 91 % on eight, where it is replicated on every rank).  Nested
// dissection of the chain breaks the dependency: the key frames are cut into ~F / 16 leaves of <= 9 poses separated by
// separators of w poses (w = band half-width); all leaves are eliminated at once (one CTA each), then all lowest
// separators, ... up a binary tree of depth log2(F / 16); the long rows ("arrows") go last.  Every tree node is a small
// DENSE front [eliminated | boundary] x [eliminated] handled by the building blocks of ba_dense.cuh: panel factorisation
// in shared memory, the Schur update of the boundary corner on the FP64 tensor pipe (DMMA), the inverse of the diagonal
// block as a by-product.  A parent adds the corners of its children into its own front (extend-add, fixed child order:
// deterministic).  Back-substitution runs down the tree the same way.  C4: 9 levels instead of 1 999 steps.
//
// The block skyline the build kernels (and the NCCL all-reduce of a partitioned run) produce is only READ here.
// Replaces, like the other direct solvers, g2o's linear solver call of BlockSolver::solve (Optimizer.cpp:76-91).
#pragma once
#include "ba_dense.cuh"

namespace visfs {
namespace mf {

constexpr int kMaxBand = 9;             // separator width (blocks) the fronts are sized for: one panel of dn::kNB columns
constexpr int kMaxArrow = 9;            // long rows carried in every boundary
constexpr int kMaxFrontBlocks = 2 * kMaxBand + 2 * kMaxBand + kMaxArrow;   // eliminated (<= 18 in a leaf) + boundary blocks
constexpr int kMaxRows = 6 * kMaxFrontBlocks + 1 + dn::kNB;   // rows of a panel matrix: front rows + rhs row + unit rows

struct Prob {
    long long d_off;        // doubles: offset of the front D_P [LD][LD] in the front buffer
    long long linv_off;     // doubles: L_kk^-T of this front's panels, [panels][kNB][kNB]
    int ne, nb;             // eliminated / boundary 6 x 6 blocks
    int LD;                 // leading dimension (= rows) of D_P
    int idx_off;            // into idx[]: hessian indices of the local blocks, eliminated first, then boundary
    int child_off, n_child; // into child[]: problem ids of the children (ascending)
    int map_off;            // into pmap[]: for boundary block b of THIS front its local block index in the parent's front
};

struct Plan {
    const Prob *prob;
    const int *idx, *child, *pmap;
    double *fronts, *linvt, *x;
    int *flag;
    long long *prof;        // optional (VISFS_BA_DENSE_PROF): globaltimer stamps of CTA 0 of every level, 8 per level
};

// Which columns does a long row ("arrow") really couple with?  The envelope only knows its first column.  touch[a][h] = 1
// when arrow a and hessian index h share a landmark (any edge, like g2o's buildStructure) or an odometry link; the host
// uses it to keep arrows out of the fronts they cannot reach.
__global__ void k_arrow_touch(Batch B, const int *__restrict__ arrow_of /*[F] arrow index or -1*/, int *touch, int F) {
    const WinDesc &wd = B.win[0];
    if (B.st[0].status != 0) return;
    const int t0 = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    for (int l = t0; l < wd.n_point; l += stride) {
        if (!(B.lm_flags[l] & kInHessian)) continue;
        const int e0 = B.lm_edge_off[l], e1 = B.lm_edge_off[l + 1];
        for (int e = e0; e < e1; ++e) {
            const int hi = B.pose_hidx[B.edge_pose[e] & kPoseMask];
            const int a = hi >= 0 ? arrow_of[hi] : -1;
            if (a < 0) continue;
            for (int e2 = e0; e2 < e1; ++e2) {
                const int hj = B.pose_hidx[B.edge_pose[e2] & kPoseMask];
                if (hj >= 0) touch[(size_t)a * F + hj] = 1;
            }
        }
    }
    for (int k = t0; k < wd.n_link; k += stride) {
        const int hi = B.pose_hidx[B.link_from[wd.link_off + k]], hj = B.pose_hidx[B.link_to[wd.link_off + k]];
        if (hi < 0 || hj < 0) continue;
        if (arrow_of[hi] >= 0) touch[(size_t)arrow_of[hi] * F + hj] = 1;
        if (arrow_of[hj] >= 0) touch[(size_t)arrow_of[hj] * F + hi] = 1;
    }
}

// One CTA per front of the level: zero, assemble (entries of S whose earlier-eliminated end is eliminated here + the
// children's corners), eliminate the front's own columns panel by panel, leave the updated boundary corner for the parent.
__global__ void __launch_bounds__(dn::kThreadsD) k_mf_factor(Batch B, Plan P, int first_prob, int level) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *T = reinterpret_cast<double *>(smem_raw);
    double *S6 = T + (size_t)kMaxRows * dn::kTP;
    const WinDesc &wd = B.win[0];
    const LMState &st = B.st[0];
    if (st.done) return;
    const int tid = threadIdx.x, warp = tid >> 5;
    const Prob pb = P.prob[first_prob + blockIdx.x];
    const int ne = pb.ne, nbd = pb.nb, LD = pb.LD;
    const int n_e = 6 * ne, n_t = 6 * (ne + nbd);
    double *__restrict__ D = P.fronts + pb.d_off;
    const int *__restrict__ idx = P.idx + pb.idx_off;
    const double lambda = (wd.trust == 0) ? (B.dbg_lambda >= 0.0 ? B.dbg_lambda : st.lambda) : 0.0;
    long long *pr = (P.prof && blockIdx.x == 0 && tid == 0) ? P.prof + 8 * level : nullptr;
    if (pr) pr[0] = dn::gtime();

    // ---- assembly by GATHER: every lower entry of the front (and its rhs row) is written once, as the sum of the entry of S
    //      (when its column is eliminated here) and of the children's corner entries (fixed child order: deterministic); all
    //      loads of an item are independent.  Index tables first, in shared memory.
    __shared__ int s_idx[kMaxFrontBlocks], s_first[kMaxFrontBlocks], s_inv[2][kMaxFrontBlocks];
    __shared__ long long s_off[kMaxFrontBlocks];
    __shared__ unsigned short s_pair[kMaxFrontBlocks * (kMaxFrontBlocks + 1) / 2];   // lower block pairs (i << 8 | j), j <= i
    const int nloc = ne + nbd;
    const int nch = min(pb.n_child, 2);
    if (tid < nloc) {
        const int r = idx[tid];
        s_idx[tid] = r; s_first[tid] = B.sky_first[r]; s_off[tid] = B.sky_off[r];
        s_inv[0][tid] = -1; s_inv[1][tid] = -1;
    }
    for (int t = tid; t < nloc * (nloc + 1) / 2; t += dn::kThreadsD) {
        int i = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
        while ((i + 1) * (i + 2) / 2 <= t) ++i;
        while (i * (i + 1) / 2 > t) --i;
        s_pair[t] = (unsigned short)((i << 8) | (t - i * (i + 1) / 2));
    }
    __syncthreads();
    Prob chp[2];
    for (int k = 0; k < nch; ++k) {
        chp[k] = P.prob[P.child[pb.child_off + k]];
        if (tid < chp[k].nb) s_inv[k][P.pmap[chp[k].map_off + tid]] = tid;      // parent local block -> boundary block of child k
    }
    __syncthreads();
    if (pr) pr[1] = dn::gtime();
    {
        const double *__restrict__ sky = B.red;
        const double *__restrict__ g = B.red + B.red_g_off;
        const int n_items = nloc * (nloc + 1) / 2 * 6;
#pragma unroll 2
        for (int item = tid; item < n_items + nloc * 6; item += dn::kThreadsD) {
            if (item >= n_items) {                              // rhs row: entry 6 j + c
                const int t = item - n_items, j = t / 6, c = t - j * 6;
                double v = (j < ne) ? g[6 * s_idx[j] + c] : 0.0;
                for (int k = 0; k < nch; ++k) {
                    const int bj = s_inv[k][j];
                    if (bj >= 0) v += __ldcg(P.fronts + chp[k].d_off + (size_t)(6 * (chp[k].ne + chp[k].nb)) * chp[k].LD + 6 * chp[k].ne + 6 * bj + c);
                }
                D[(size_t)n_t * LD + t] = v;
                continue;
            }
            const int pr_ = item / 6, a = item - pr_ * 6;
            const int i = s_pair[pr_] >> 8, j = s_pair[pr_] & 255;
            double v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            if (j < ne) {                                       // entry of S: block (i, j), row a
                const int ri = s_idx[i], cj = s_idx[j];
                const bool lower = ri >= cj;                    // stored with the larger hessian index as its row
                const int hb = lower ? i : j, ha_idx = lower ? cj : ri;
                if (ha_idx >= s_first[hb]) {
                    const double *blk = sky + (size_t)(s_off[hb] + (ha_idx - s_first[hb])) * 36;
                    if (lower) {
                        const double2 *p2 = reinterpret_cast<const double2 *>(blk + a * 6);
                        const double2 v0 = p2[0], v1 = p2[1], v2 = p2[2];
                        v[0] = v0.x; v[1] = v0.y; v[2] = v1.x; v[3] = v1.y; v[4] = v2.x; v[5] = v2.y;
                    } else {
#pragma unroll
                        for (int c = 0; c < 6; ++c) v[c] = blk[c * 6 + a];
                    }
                    if (i == j) v[a] += lambda;
                }
            }
            for (int k = 0; k < nch; ++k) {
                const int bi = s_inv[k][i], bj = s_inv[k][j];
                if (bi < 0 || bj < 0) continue;
                const double *Dc = P.fronts + chp[k].d_off;
                const int ce = 6 * chp[k].ne, cLD = chp[k].LD;
                if (bi >= bj) {                                 // the child holds this block as (bi, bj): row a is contiguous
                    const double2 *p2 = reinterpret_cast<const double2 *>(Dc + (size_t)(ce + 6 * bi + a) * cLD + ce + 6 * bj);
                    const double2 v0 = __ldcg(p2), v1 = __ldcg(p2 + 1), v2 = __ldcg(p2 + 2);
                    v[0] += v0.x; v[1] += v0.y; v[2] += v1.x; v[3] += v1.y; v[4] += v2.x; v[5] += v2.y;
                } else {                                        // ... as (bj, bi): transposed
#pragma unroll
                    for (int c = 0; c < 6; ++c) v[c] += __ldcg(Dc + (size_t)(ce + 6 * bj + c) * cLD + ce + 6 * bi + a);
                }
            }
            double *dst = D + (size_t)(6 * i + a) * LD + 6 * j;
#pragma unroll
            for (int c = 0; c < 6; ++c) if (i != j || c <= a) dst[c] = v[c];
        }
    }
    if (pr) pr[2] = dn::gtime();
    __threadfence();
    __syncthreads();
    if (pr) pr[3] = dn::gtime();
    // ---- eliminate the front's own columns
    for (int k0 = 0; k0 < n_e; k0 += dn::kNB) {
        const int nbp = min(dn::kNB, n_e - k0), base = k0 + nbp, m = n_t + 1 - base;
        dn::panel_factor(D, LD, k0, nbp, base, m, 0, 1, 0, nbp, P.linvt + pb.linv_off + (size_t)(k0 / dn::kNB) * dn::kNB * dn::kNB, P.flag, true, T, S6);
        __threadfence();
        __syncthreads();
        if (pr && k0 == 0) pr[4] = dn::gtime();
        if (m > 1) dn::trailing_update_smem(D, LD, T, nbp, base, m, warp, dn::kThreadsD / 32);
        __threadfence();
        __syncthreads();
        if (pr && k0 == 0) pr[5] = dn::gtime();
    }
    if (pr) pr[6] = dn::gtime();
}

// One CTA per front of the level, parents before children: x of the boundary is known, solve for the eliminated part
__global__ void __launch_bounds__(dn::kThreadsD) k_mf_back(Batch B, Plan P, int first_prob) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *Ls = reinterpret_cast<double *>(smem_raw);           // [kNB][kTP]
    double *ys = Ls + dn::kNB * dn::kTP, *xs = ys + dn::kNB, *xb = xs + dn::kNB;   // xb [6 * boundary blocks]
    if (B.st[0].done) return;
    const int tid = threadIdx.x;
    const Prob pb = P.prob[first_prob + blockIdx.x];
    const int ne = pb.ne, nbd = pb.nb, LD = pb.LD;
    const int n_e = 6 * ne, n_t = 6 * (ne + nbd);
    double *__restrict__ D = P.fronts + pb.d_off;
    const int *__restrict__ idx = P.idx + pb.idx_off;
    double *__restrict__ yrow = D + (size_t)n_t * LD;
    for (int t = tid; t < 6 * nbd; t += dn::kThreadsD) xb[t] = dn::ld_l2(P.x + 6 * idx[ne + t / 6] + t % 6);
    __syncthreads();
    if (nbd > 0) {   // y_c -= sum over the boundary rows of L[r][c] x_r: four lanes per column, each a quarter of the rows
        const int part = tid & 3, nrow = 6 * nbd;
        for (int c0 = 0; c0 < n_e; c0 += dn::kThreadsD / 4) {
            const int c = c0 + (tid >> 2);
            double sacc = 0.0;
            if (c < n_e) {
                const double *col = D + (size_t)n_e * LD + c;
                int r = part;
                for (; r + 28 < nrow; r += 32) {
                    double v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) v[u] = dn::ld_l2(col + (size_t)(r + 4 * u) * LD);
#pragma unroll
                    for (int u = 0; u < 8; ++u) sacc = fma(v[u], xb[r + 4 * u], sacc);
                }
                for (; r < nrow; r += 4) sacc = fma(dn::ld_l2(col + (size_t)r * LD), xb[r], sacc);
            }
            sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
            sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
            if (c < n_e && part == 0) yrow[c] -= sacc;
        }
    }
    __threadfence();
    __syncthreads();
    const int npan = (n_e + dn::kNB - 1) / dn::kNB;
    for (int p = npan - 1; p >= 0; --p) {
        const int k0 = p * dn::kNB, nbp = min(dn::kNB, n_e - k0);
        dn::back_panel_x(P.linvt + pb.linv_off + (size_t)p * dn::kNB * dn::kNB, yrow, k0, nbp, Ls, ys, xs);
        if (tid < nbp) P.x[6 * idx[(k0 + tid) / 6] + (k0 + tid) % 6] = xs[tid];
        dn::back_update_cols(D, LD, yrow, k0, nbp, xs, tid, dn::kThreadsD, k0);
        __threadfence();
        __syncthreads();
    }
}

}  // namespace mf
}  // namespace visfs

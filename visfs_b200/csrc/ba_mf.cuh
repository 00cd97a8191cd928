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
};

// One CTA per front of the level: zero, assemble (entries of S whose earlier-eliminated end is eliminated here + the
// children's corners), eliminate the front's own columns panel by panel, leave the updated boundary corner for the parent.
__global__ void __launch_bounds__(dn::kThreadsD) k_mf_factor(Batch B, Plan P, int first_prob) {
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

    // index tables of the front in shared memory (the assembly below is bound by dependent global loads otherwise)
    __shared__ int s_idx[kMaxFrontBlocks], s_first[kMaxFrontBlocks], s_map[kMaxFrontBlocks];
    __shared__ long long s_off[kMaxFrontBlocks];
    const int nloc = ne + nbd;
    if (tid < nloc) {
        const int r = idx[tid];
        s_idx[tid] = r; s_first[tid] = B.sky_first[r]; s_off[tid] = B.sky_off[r];
    }
    // ---- zero the part of the front the tiles can touch
    {
        double2 *z = reinterpret_cast<double2 *>(D);
        const int cnt = (n_t + 32) * LD / 2;                    // LD is a multiple of 8
        for (int i = tid; i < cnt; i += dn::kThreadsD) z[i] = make_double2(0.0, 0.0);
    }
    __threadfence();
    __syncthreads();
    // ---- entries of S: block (local i, local j) for j eliminated here, i >= j; one item = one 6-entry row of a block
    {
        const double *__restrict__ sky = B.red;
        for (int item = tid; item < nloc * ne * 6; item += dn::kThreadsD) {
            const int i = item / (ne * 6), rem = item - i * (ne * 6), j = rem / 6, a = rem - j * 6;
            if (i < j) continue;
            const int ri = s_idx[i], cj = s_idx[j];
            const bool lower = ri >= cj;                        // the block is stored with the larger hessian index as its row
            const int hb = lower ? i : j, ha_idx = lower ? cj : ri;
            if (ha_idx < s_first[hb]) continue;                 // outside the envelope: structurally zero
            const double *blk = sky + (size_t)(s_off[hb] + (ha_idx - s_first[hb])) * 36;
            double v[6];
            if (lower) {
                const double2 *p2 = reinterpret_cast<const double2 *>(blk + a * 6);
                const double2 v0 = p2[0], v1 = p2[1], v2 = p2[2];
                v[0] = v0.x; v[1] = v0.y; v[2] = v1.x; v[3] = v1.y; v[4] = v2.x; v[5] = v2.y;
            } else {
#pragma unroll
                for (int c = 0; c < 6; ++c) v[c] = blk[c * 6 + a];
            }
            double *dst = D + (size_t)(6 * i + a) * LD + 6 * j;
#pragma unroll
            for (int c = 0; c < 6; ++c)
                if (i != j || c <= a) dst[c] = v[c] + ((i == j && a == c) ? lambda : 0.0);
        }
        const double *g = B.red + B.red_g_off;
        for (int t = tid; t < n_e; t += dn::kThreadsD) D[(size_t)n_t * LD + t] = g[6 * s_idx[t / 6] + t % 6];
    }
    __syncthreads();
    // ---- extend-add: the children's boundary corners (and the boundary part of their rhs rows), child by child;
    //      one item = one 6-entry row of a corner block, every entry of the parent has one writer per child
    for (int k = 0; k < pb.n_child; ++k) {
        const Prob ch = P.prob[P.child[pb.child_off + k]];
        const double *__restrict__ Dc = P.fronts + ch.d_off;
        const int ce = 6 * ch.ne, ct = 6 * (ch.ne + ch.nb), cLD = ch.LD;
        if (tid < ch.nb) s_map[tid] = P.pmap[ch.map_off + tid];
        __syncthreads();
        for (int item = tid; item < (ch.nb * 6 + 1) * ch.nb; item += dn::kThreadsD) {
            const int row = item / ch.nb, bj = item - row * ch.nb;      // row: 6 bi + a, or ch.nb * 6 = the child's rhs row
            const int bi = row / 6, a = row - bi * 6;
            const bool rhs = bi == ch.nb;
            if (!rhs && bj > bi) continue;
            const double2 *p2 = reinterpret_cast<const double2 *>(Dc + (size_t)(rhs ? ct : ce + row) * cLD + ce + 6 * bj);
            const double2 v0 = __ldcg(p2), v1 = __ldcg(p2 + 1), v2 = __ldcg(p2 + 2);
            const double v[6] = {v0.x, v0.y, v1.x, v1.y, v2.x, v2.y};
            const int pj = s_map[bj];
            if (rhs) {
                double *dst = D + (size_t)n_t * LD + 6 * pj;
#pragma unroll
                for (int c = 0; c < 6; ++c) dst[c] += v[c];
                continue;
            }
            const int pi = s_map[bi];
            if (pi >= pj) {
                double *dst = D + (size_t)(6 * pi + a) * LD + 6 * pj;
#pragma unroll
                for (int c = 0; c < 6; ++c) if (bi != bj || c <= a) dst[c] += v[c];
            } else {
#pragma unroll
                for (int c = 0; c < 6; ++c) D[(size_t)(6 * pj + c) * LD + 6 * pi + a] += v[c];
            }
        }
        __syncthreads();
    }
    __threadfence();
    __syncthreads();
    // ---- eliminate the front's own columns
    for (int k0 = 0; k0 < n_e; k0 += dn::kNB) {
        const int nbp = min(dn::kNB, n_e - k0), base = k0 + nbp, m = n_t + 1 - base;
        dn::panel_factor(D, LD, k0, nbp, base, m, 0, 1, 0, nbp, P.linvt + pb.linv_off + (size_t)(k0 / dn::kNB) * dn::kNB * dn::kNB, P.flag, true, T, S6);
        __threadfence();
        __syncthreads();
        if (m > 1) dn::trailing_update(D, LD, k0, nbp, base, m, warp, dn::kThreadsD / 32);
        __threadfence();
        __syncthreads();
    }
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
    if (nbd > 0) dn::back_update_cols(D, LD, yrow, n_e, 6 * nbd, xb, tid, dn::kThreadsD, n_e);
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

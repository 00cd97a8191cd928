// ba_dense.cuh — dense FP64 Cholesky of the reduced camera system on the tensor pipe (north_star item 3: "a dense FP64
// Cholesky on tensor cores only when the window makes it a dense contraction"; BASELINE config C5: 200 key frames that all
// share landmarks, S = 1194 x 1194 dense).  Replaces, for such windows, what g2o's direct linear solvers do after
// BlockSolver::solve has formed S (selected at corelib/src/Optimizer/Optimizer.cpp:76-91).
//
// Layout: one dense row-major matrix D [NP][LD] in HBM / L2 (C5: 12.5 MB), rows 0..n-1 = lower triangle of the damped S,
// row n = the reduced right-hand side (the forward substitution rides along as one more row), everything else zero padding
// (so that no tile needs an edge guard).  The block skyline the build kernels produced is left untouched.
//
//   k_dense_fill     skyline blocks + lambda -> D
//   k_dense_chol     cooperative launch, one CTA per SM: blocked right-looking Cholesky, panels of kNB = 54 columns
//        phase A     every CTA factors the 54 x 54 diagonal block in shared memory (left-looking over 6 x 6 register blocks)
//                    together with ITS rows of the panel (row r below the block belongs to CTA (r - base) mod G): the
//                    triangular solve of a row is done by the thread that owns the row, nothing is broadcast.  CTA c < 54
//                    also carries the unit row e_c, which comes out as row c of L_kk^-T
//        grid.sync
//        phase B     trailing update D[r][c] -= P_r . P_c on 32 x 32 tiles, one warp per tile, K = 54 in fourteen
//                    mma.sync.aligned.m8n8k4.f64 steps (DMMA), fragments double-buffered straight from L2, results sent
//                    with red.global.add.f64 (one writer per entry and panel)
//        grid.sync
//        back-substitution, right-looking and spread over the grid: x_panel = L_kk^-T y_panel (a 54 x 54 product, every
//                    CTA for itself), then every column of y left of the panel is updated by its owner thread; one
//                    grid.sync per panel
//   k_dense_back     solution checks, CameraPose::update of the trial poses, pose part of g2o's computeScale — the epilogue
//                    of the other solvers
#pragma once
#include <cooperative_groups.h>
#include "ba_large.cuh"

namespace visfs {
namespace dn {

namespace cg = cooperative_groups;

constexpr int kNB = 54;                 // panel width: nine 6 x 6 pose blocks (= the separator width of the multifrontal solver for a 10-view band)
constexpr int kThreadsD = 256;
constexpr int kTP = kNB + 1;            // shared-memory pitch of the panel matrix (odd: rows land in different banks)
constexpr int kMaxOwn = 200;            // rows of a panel one CTA may own (row owner = one thread)
constexpr int kBackThreads = 1024;

struct DenseMat {
    double *D;       // [NP][LD]
    int n, LD, NP;
    int *flag;       // set to 1 on a non-positive pivot (cleared by the host before the launch)
    double *Linvt;   // [panels][kNB][kNB] L_kk^-T of every diagonal block (a by-product of the panel solve)
    double *x;       // [n] solution
    long long *prof; // optional (VISFS_BA_DENSE_PROF): globaltimer stamps of CTA 0 at the phase boundaries, 5 per panel
};

__device__ __forceinline__ long long gtime() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

__host__ __device__ inline int dense_pad(int n) { return ((n + 1 + 31) / 32) * 32 + 32; }

// skyline (+ lambda on the diagonal) -> dense lower triangle, rhs -> row n.  D has been zeroed by the caller.
__global__ void k_dense_fill(Batch B, DenseMat M) {
    const WinDesc &wd = B.win[0];
    const LMState &st = B.st[0];
    if (st.done) return;
    const int F = st.F, n = 6 * F;
    const double lambda = (wd.trust == 0) ? (B.dbg_lambda >= 0.0 ? B.dbg_lambda : st.lambda) : 0.0;
    const double *__restrict__ sky = B.red;
    for (int r = blockIdx.x; r < F; r += gridDim.x) {
        const int f0 = B.sky_first[r], len = r - f0 + 1;
        const double *src = sky + (size_t)B.sky_off[r] * 36;
        for (int idx = threadIdx.x; idx < len * 36; idx += blockDim.x) {
            const int cb = idx / 36, e = idx - cb * 36, a = e / 6, c = e - a * 6;   // block (r, f0 + cb), entry (row a, col c)
            const int row = 6 * r + a, col = 6 * (f0 + cb) + c;
            if (col > row) continue;
            M.D[(size_t)row * M.LD + col] = src[idx] + (row == col ? lambda : 0.0);
        }
    }
    const double *g = B.red + B.red_g_off;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) M.D[(size_t)n * M.LD + i] = g[i];
}

// Cholesky of the 6 x 6 block whose lower half sits at d[c * ld + a] (row c, column a <= c)
__device__ __forceinline__ void chol6_ld(const double *d, int ld, lg::Chol6 &f) {
    f.ok = true;
    double v;
#define VISFS_PIV(expr, inv, diag) v = (expr); if (!(v > 0.0)) { f.ok = false; v = 1.0; } inv = rsqrt(v); diag = v * inv;
    const double *d1 = d + ld, *d2 = d + 2 * ld, *d3 = d + 3 * ld, *d4 = d + 4 * ld, *d5 = d + 5 * ld;
    VISFS_PIV(d[0], f.i0, f.L00)
    f.L10 = d1[0] * f.i0; f.L20 = d2[0] * f.i0; f.L30 = d3[0] * f.i0; f.L40 = d4[0] * f.i0; f.L50 = d5[0] * f.i0;
    VISFS_PIV(fma(-f.L10, f.L10, d1[1]), f.i1, f.L11)
    f.L21 = fma(-f.L20, f.L10, d2[1]) * f.i1; f.L31 = fma(-f.L30, f.L10, d3[1]) * f.i1;
    f.L41 = fma(-f.L40, f.L10, d4[1]) * f.i1; f.L51 = fma(-f.L50, f.L10, d5[1]) * f.i1;
    VISFS_PIV(fma(-f.L21, f.L21, fma(-f.L20, f.L20, d2[2])), f.i2, f.L22)
    f.L32 = fma(-f.L31, f.L21, fma(-f.L30, f.L20, d3[2])) * f.i2;
    f.L42 = fma(-f.L41, f.L21, fma(-f.L40, f.L20, d4[2])) * f.i2;
    f.L52 = fma(-f.L51, f.L21, fma(-f.L50, f.L20, d5[2])) * f.i2;
    VISFS_PIV(fma(-f.L32, f.L32, fma(-f.L31, f.L31, fma(-f.L30, f.L30, d3[3]))), f.i3, f.L33)
    f.L43 = fma(-f.L42, f.L32, fma(-f.L41, f.L31, fma(-f.L40, f.L30, d4[3]))) * f.i3;
    f.L53 = fma(-f.L52, f.L32, fma(-f.L51, f.L31, fma(-f.L50, f.L30, d5[3]))) * f.i3;
    VISFS_PIV(fma(-f.L43, f.L43, fma(-f.L42, f.L42, fma(-f.L41, f.L41, fma(-f.L40, f.L40, d4[4])))), f.i4, f.L44)
    f.L54 = fma(-f.L53, f.L43, fma(-f.L52, f.L42, fma(-f.L51, f.L41, fma(-f.L50, f.L40, d5[4])))) * f.i4;
    VISFS_PIV(fma(-f.L54, f.L54, fma(-f.L53, f.L53, fma(-f.L52, f.L52, fma(-f.L51, f.L51, fma(-f.L50, f.L50, d5[5]))))), f.i5, f.L55)
#undef VISFS_PIV
}

// D (8x8) += A (8x4, row) * B (4x8, col), FP64 tensor pipe.  Lane l holds A[l >> 2][l & 3], B[l & 3][l >> 2] and
// C[l >> 2][2 (l & 3) .. + 1].
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ---------------------------------------------------------------------------------------------------------------------
// Building blocks shared by the dense solver below (one matrix, all CTAs of a cooperative grid) and by the multifrontal
// solver of ba_mf.cuh (many small fronts, one CTA each: SINGLE).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kDiagLoads = (kNB * kNB + kThreadsD - 1) / kThreadsD;

__device__ __forceinline__ double ld_l2(const double *p) { return __ldcg(p); }   // D is written by other CTAs / by red.global: read it at L2

// Phase A of one panel: columns [k0, k0 + nb) of D.  The diagonal block and the CTA's rows of the panel go to shared
// memory T, are factored / solved there (left-looking over 6 x 6 register blocks, row owner = thread, nothing is
// broadcast) and the rows go back to D.  Rows: q-th own row = base + c_id + G q, q < nr.  `nunit` unit rows e_u ride along
// (u = unit0 + 0 .. nunit-1): after the solve they hold rows of L_kk^-T, stored to linvt[u][0..nb).
__device__ __forceinline__ void panel_factor(double *__restrict__ D, int LD, int k0, int nb, int base, int nr, int c_id, int G,
                                             int unit0, int nunit, double *__restrict__ linvt, int *flag, bool report,
                                             double *T, double *S6) {
    const int tid = threadIdx.x;
    const int rows_all = nb + nr + nunit;
    {
        double v[kDiagLoads], w[2];
        const int own_items = nr * nb;
#pragma unroll
        for (int u = 0; u < kDiagLoads; ++u) {
            const int idx = tid + u * kThreadsD, i = idx / nb, j = idx - i * nb;
            v[u] = (idx < nb * nb) ? ld_l2(D + (size_t)(k0 + i) * LD + k0 + j) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {       // (the common case: at most 512 entries of own rows, in the same batch of loads)
            const int idx = tid + u * kThreadsD, q = idx / nb, j = idx - q * nb;
            w[u] = (idx < own_items) ? ld_l2(D + (size_t)(base + c_id + G * q) * LD + k0 + j) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < kDiagLoads; ++u) {
            const int idx = tid + u * kThreadsD, i = idx / nb, j = idx - i * nb;
            if (idx < nb * nb && j <= i) T[i * kTP + j] = v[u];
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int idx = tid + u * kThreadsD, q = idx / nb, j = idx - q * nb;
            if (idx < own_items) T[(nb + q) * kTP + j] = w[u];
        }
        for (int i0 = 2 * kThreadsD; i0 < own_items; i0 += 8 * kThreadsD) {
            double z[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = i0 + tid + u * kThreadsD, q = idx / nb, j = idx - q * nb;
                z[u] = (idx < own_items) ? ld_l2(D + (size_t)(base + c_id + G * q) * LD + k0 + j) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = i0 + tid + u * kThreadsD, q = idx / nb, j = idx - q * nb;
                if (idx < own_items) T[(nb + q) * kTP + j] = z[u];
            }
        }
        for (int idx = tid; idx < nunit * nb; idx += kThreadsD) {
            const int u = idx / nb, j = idx - u * nb;
            T[(nb + nr + u) * kTP + j] = (j == unit0 + u) ? 1.0 : 0.0;
        }
    }
    __syncthreads();
    // While the row owners run the dependent chain of a block's factorisation the other warps pre-compute, for the NEXT
    // block column, the part of its update that only needs finished columns (possible when the owners fit warps 0-1).
    const bool helpers = rows_all <= 64;
    for (int b0 = 0; b0 < nb; b0 += 6) {
        if (b0 > 0) {
            for (int i = b0 + tid; i < rows_all; i += kThreadsD) {
                double u[6];
                const int q0 = helpers ? b0 - 6 : 0;
#pragma unroll
                for (int c = 0; c < 6; ++c) u[c] = helpers ? S6[i * 6 + c] : 0.0;
                const double *ri = T + i * kTP, *p0 = T + b0 * kTP;
                for (int q = q0; q < b0; ++q) {
                    const double v = ri[q];
#pragma unroll
                    for (int c = 0; c < 6; ++c) u[c] = fma(v, p0[c * kTP + q], u[c]);
                }
                double *x = T + i * kTP + b0;
#pragma unroll
                for (int c = 0; c < 6; ++c) x[c] -= u[c];
            }
            __syncthreads();
        }
        lg::Chol6 f;
        const bool solver = b0 + 6 + tid < rows_all;                   // owner of (at least) one row below the block
        if (solver || tid == 0) chol6_ld(T + b0 * kTP + b0, kTP, f);   // every row owner factors the block redundantly
        if (solver) {
            for (int i = b0 + 6 + tid; i < rows_all; i += kThreadsD) {
                double *x = T + i * kTP + b0;
                double xv[6] = {x[0], x[1], x[2], x[3], x[4], x[5]};
                lg::row_solve6(f, xv);
#pragma unroll
                for (int q = 0; q < 6; ++q) x[q] = xv[q];
            }
        } else if (helpers && tid >= 64 && b0 + 6 < nb && b0 > 0) {
            // helper warps: S6[i][c] = sum_{q < b0} T[i][q] T[b0 + 6 + c][q] for the rows at and below the next block
            const int nrow = rows_all - b0 - 6;
            for (int item = tid - 64; item < nrow * 6; item += kThreadsD - 64) {
                const int c = item / nrow, ir = b0 + 6 + (item - c * nrow);
                const double *ri = T + ir * kTP, *pc = T + (b0 + 6 + c) * kTP;
                double s0 = 0.0, s1 = 0.0;
                int q = 0;
                for (; q + 1 < b0; q += 2) { s0 = fma(ri[q], pc[q], s0); s1 = fma(ri[q + 1], pc[q + 1], s1); }
                if (q < b0) s0 = fma(ri[q], pc[q], s0);
                S6[ir * 6 + c] = s0 + s1;
            }
        } else if (helpers && tid >= 64 && b0 == 0) {
            for (int item = tid - 64; item < rows_all * 6; item += kThreadsD - 64) S6[item] = 0.0;
        }
        __syncthreads();
        if (tid == 0) {   // the factor of the diagonal block is stored only now: nobody reads the original any more
            double *d0 = T + b0 * kTP + b0, *d1 = d0 + kTP, *d2 = d1 + kTP, *d3 = d2 + kTP, *d4 = d3 + kTP, *d5 = d4 + kTP;
            d0[0] = f.L00;
            d1[0] = f.L10; d1[1] = f.L11;
            d2[0] = f.L20; d2[1] = f.L21; d2[2] = f.L22;
            d3[0] = f.L30; d3[1] = f.L31; d3[2] = f.L32; d3[3] = f.L33;
            d4[0] = f.L40; d4[1] = f.L41; d4[2] = f.L42; d4[3] = f.L43; d4[4] = f.L44;
            d5[0] = f.L50; d5[1] = f.L51; d5[2] = f.L52; d5[3] = f.L53; d5[4] = f.L54; d5[5] = f.L55;
            if (!f.ok && report) *flag = 1;
        }
    }
    __syncthreads();
    // own rows of the panel back, and the rows of L_kk^-T (the factor of the diagonal block itself is not needed again:
    // the back-substitution works with its inverse)
    for (int idx = tid; idx < nr * nb; idx += kThreadsD) {
        const int q = idx / nb, j = idx - q * nb;
        D[(size_t)(base + c_id + G * q) * LD + k0 + j] = T[(nb + q) * kTP + j];
    }
    for (int idx = tid; idx < nunit * nb; idx += kThreadsD) {
        const int u = idx / nb, j = idx - u * nb;
        linvt[(size_t)(unit0 + u) * kNB + j] = T[(nb + nr + u) * kTP + j];
    }
}

// Phase B of one panel: D[r][c] -= sum_k P[r][k] P[c][k] over the lower triangle of rows / columns [base, base + m), 32 x 32
// tiles, one warp each; tile t of the sequence t0, t0 + tstep, ... is this warp's.  Fragments double-buffered (the loads of
// the next 16 columns are in flight while the tensor pipe works on the current ones), results sent with red.global.add
// (every entry has exactly one writer per panel: deterministic, and no load to wait for).
__device__ __forceinline__ void trailing_update(double *__restrict__ D, int LD, int k0, int nb, int base, int m, int t0, int tstep) {
    const int lane = threadIdx.x & 31;
    const int nt = (m + 31) >> 5;
    const int ntiles = nt * (nt + 1) / 2;
    const int gl = lane >> 2, tl = lane & 3;
    for (int tile = t0; tile < ntiles; tile += tstep) {
        int ti = (int)((sqrt(8.0 * (double)tile + 1.0) - 1.0) * 0.5);
        while ((ti + 1) * (ti + 2) / 2 <= tile) ++ti;
        while (ti * (ti + 1) / 2 > tile) --ti;
        const int tj = tile - ti * (ti + 1) / 2;
        const int r0 = base + 32 * ti, c0 = base + 32 * tj;
        const double *pa = D + (size_t)(r0 + gl) * LD + k0 + tl;
        const double *pb = D + (size_t)(c0 + gl) * LD + k0 + tl;
        double acc[4][4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
        double fa0[4][4], fb0[4][4], fa1[4][4], fb1[4][4];
#define VISFS_LOAD(fa, fb, kk)                                                                        \
        _Pragma("unroll") for (int s_ = 0; s_ < 4; ++s_) {                                             \
            const bool in_ = (kk) + 4 * s_ + tl < nb; /* (a panel may be narrower than kNB) */         \
            _Pragma("unroll") for (int i_ = 0; i_ < 4; ++i_) {                                         \
                fa[s_][i_] = in_ ? ld_l2(pa + (size_t)(8 * i_) * LD + (kk) + 4 * s_) : 0.0;            \
                fb[s_][i_] = in_ ? ld_l2(pb + (size_t)(8 * i_) * LD + (kk) + 4 * s_) : 0.0;            \
            }                                                                                         \
        }
#define VISFS_MMA(fa, fb)                                                                             \
        _Pragma("unroll") for (int s_ = 0; s_ < 4; ++s_)                                               \
            _Pragma("unroll") for (int i_ = 0; i_ < 4; ++i_)                                           \
                _Pragma("unroll") for (int j_ = 0; j_ < 4; ++j_) dmma884(acc[i_][j_][0], acc[i_][j_][1], fa[s_][i_], fb[s_][j_]);
        VISFS_LOAD(fa0, fb0, 0)
#pragma unroll 1
        for (int kk = 0; kk < nb; kk += 32) {
            const bool more1 = kk + 16 < nb, more2 = kk + 32 < nb;
            if (more1) { VISFS_LOAD(fa1, fb1, kk + 16) }
            VISFS_MMA(fa0, fb0)
            if (more1) {
                if (more2) { VISFS_LOAD(fa0, fb0, kk + 32) }
                VISFS_MMA(fa1, fb1)
            }
        }
#undef VISFS_LOAD
#undef VISFS_MMA
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                double *p = D + (size_t)(r0 + 8 * i + gl) * LD + c0 + 8 * j + 2 * tl;
                asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(-acc[i][j][0]) : "memory");
                asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p + 1), "d"(-acc[i][j][1]) : "memory");
            }
    }
}

// The same update for a front that ONE CTA owns: the solved panel rows are still in shared memory (T rows nb .. nb + m - 1 are
// the rows base .. base + m - 1 of D), so the fragments come from there instead of L2.
__device__ __forceinline__ void trailing_update_smem(double *__restrict__ D, int LD, const double *T, int nb, int base, int m, int t0, int tstep) {
    const int lane = threadIdx.x & 31;
    const int nt = (m + 31) >> 5;
    const int ntiles = nt * (nt + 1) / 2;
    const int gl = lane >> 2, tl = lane & 3;
    for (int tile = t0; tile < ntiles; tile += tstep) {
        int ti = (int)((sqrt(8.0 * (double)tile + 1.0) - 1.0) * 0.5);
        while ((ti + 1) * (ti + 2) / 2 <= tile) ++ti;
        while (ti * (ti + 1) / 2 > tile) --ti;
        const int tj = tile - ti * (ti + 1) / 2;
        const int xr = 32 * ti + gl, xc = 32 * tj + gl;          // rows relative to base
        double acc[4][4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
#pragma unroll 2
        for (int kk = 0; kk < nb; kk += 4) {
            const bool in = kk + tl < nb;
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                a[i] = (in && xr + 8 * i < m) ? T[(nb + xr + 8 * i) * kTP + kk + tl] : 0.0;
                b[i] = (in && xc + 8 * i < m) ? T[(nb + xc + 8 * i) * kTP + kk + tl] : 0.0;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                double *p = D + (size_t)(base + xr + 8 * i) * LD + base + 32 * tj + 8 * j + 2 * tl;
                asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(-acc[i][j][0]) : "memory");
                asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p + 1), "d"(-acc[i][j][1]) : "memory");
            }
    }
}

// Back-substitution step of one panel: x_panel = L_kk^-T y_panel (Ls / ys / xs in shared memory, four lanes per row).
// Leaves xs valid for all threads (ends with a barrier).  `lin` = the panel's L_kk^-T [kNB][kNB] in global memory.
__device__ __forceinline__ void back_panel_x(const double *__restrict__ lin, const double *__restrict__ yrow, int k0, int nb,
                                             double *Ls, double *ys, double *xs) {
    const int tid = threadIdx.x;
    {
        double v[kDiagLoads];
#pragma unroll
        for (int u = 0; u < kDiagLoads; ++u) {
            const int idx = tid + u * kThreadsD;
            v[u] = (idx < nb * kNB) ? ld_l2(lin + idx) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < kDiagLoads; ++u) {
            const int idx = tid + u * kThreadsD, i = idx / kNB, j = idx - i * kNB;
            if (idx < nb * kNB) Ls[i * kTP + j] = v[u];
        }
        if (tid < nb) ys[tid] = ld_l2(yrow + k0 + tid);
    }
    __syncthreads();
    {   // x_i = sum_{j >= i} (L^-T)[i][j] y_j
        const int i = tid >> 2, part = tid & 3;
        double sacc = 0.0;
        if (i < nb) for (int j = i + part; j < nb; j += 4) sacc = fma(Ls[i * kTP + j], ys[j], sacc);
        sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
        sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
        if (i < nb && part == 0) xs[i] = sacc;
    }
    __syncthreads();
}

// y_c -= sum_{r < nrow} D[row0 + r][c] xv[r] for the columns c = c0, c0 + cstep, ... < cend (column c has ONE owner thread)
__device__ __forceinline__ void back_update_cols(double *__restrict__ D, int LD, double *__restrict__ yrow, int row0, int nrow,
                                                 const double *xv, int c0, int cstep, int cend) {
    for (int c = c0; c < cend; c += cstep) {
        const double *col = D + (size_t)row0 * LD + c;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        for (int r = 0; r < nrow; r += 6) {      // nrow is a multiple of 6
            const double v0 = ld_l2(col + (size_t)r * LD), v1 = ld_l2(col + (size_t)(r + 1) * LD), v2 = ld_l2(col + (size_t)(r + 2) * LD);
            const double v3 = ld_l2(col + (size_t)(r + 3) * LD), v4 = ld_l2(col + (size_t)(r + 4) * LD), v5 = ld_l2(col + (size_t)(r + 5) * LD);
            s0 = fma(v0, xv[r], s0); s1 = fma(v1, xv[r + 1], s1); s2 = fma(v2, xv[r + 2], s2);
            s0 = fma(v3, xv[r + 3], s0); s1 = fma(v4, xv[r + 4], s1); s2 = fma(v5, xv[r + 5], s2);
        }
        yrow[c] -= (s0 + s1) + s2;
    }
}

__global__ void __launch_bounds__(kThreadsD) k_dense_chol(Batch B, DenseMat M) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *T = reinterpret_cast<double *>(smem_raw);   // [(kNB + own rows + 1)][kTP]
    if (B.st[0].done) return;
    cg::grid_group grid = cg::this_grid();
    const int tid = threadIdx.x, warp = tid >> 5;
    const int G = gridDim.x, c_id = blockIdx.x;
    const int n = M.n, LD = M.LD;
    double *__restrict__ D = M.D;
    const int own_max = (n + 1 + G - 1) / G + 1;
    double *S6 = T + (size_t)(kNB + own_max) * kTP;      // [rows][6] partial sums of the helper warps

    for (int k0 = 0; k0 < n; k0 += kNB) {
        const int nb = min(kNB, n - k0);
        const int base = k0 + nb;                  // first row below the diagonal block
        const int m = n + 1 - base;                // rows below it (the rhs row n included)
        const int nr = (m > c_id) ? (m - c_id + G - 1) / G : 0;   // this CTA's rows: base + c_id + G q
        long long *pr = (M.prof && c_id == 0 && tid == 0) ? M.prof + 5 * (k0 / kNB) : nullptr;
        if (pr) pr[0] = gtime();
        // CTA c < nb also carries the unit row e_c: after the panel solve it holds row c of L_kk^-T, which turns the
        // back-substitution's triangular solves into plain products
        panel_factor(D, LD, k0, nb, base, nr, c_id, G, c_id, (c_id < nb) ? 1 : 0, M.Linvt + (size_t)(k0 / kNB) * kNB * kNB, M.flag,
                     c_id == 0, T, S6);
        if (pr) pr[1] = gtime();
        __threadfence();
        grid.sync();
        if (pr) pr[2] = gtime();
        if (base >= n) break;                      // last panel: nothing is left to update (the same decision in every CTA)
        // tiles dealt round-robin over the CTAs first (every SM gets work)
        trailing_update(D, LD, k0, nb, base, m, c_id + G * warp, G * (kThreadsD / 32));
        if (pr) pr[3] = gtime();
        __threadfence();
        grid.sync();
        if (pr) pr[4] = gtime();
    }

    // ---- back-substitution L^T x = y (y = row n of D), right-looking and spread over the grid: every CTA forms
    //      x_panel = L_kk^-T y_panel itself, then column c < k0 of y is brought up to date by ITS owner thread (independent
    //      loads of L[panel rows][c]); one grid.sync per panel, no partial sums to exchange
    long long tb = (M.prof && c_id == 0 && tid == 0) ? gtime() : 0;
    double *Ls = T, *ys = T + kNB * kTP, *xs = ys + kNB;
    const int npan = (n + kNB - 1) / kNB;
    double *__restrict__ yrow = D + (size_t)n * LD;
    for (int p = npan - 1; p >= 0; --p) {
        const int k0 = p * kNB, nb = min(kNB, n - k0);
        back_panel_x(M.Linvt + (size_t)p * kNB * kNB, yrow, k0, nb, Ls, ys, xs);
        if (c_id == 0 && tid < nb) M.x[k0 + tid] = xs[tid];
        back_update_cols(D, LD, yrow, k0, nb, xs, c_id * kThreadsD + tid, G * kThreadsD, k0);
        if (p > 0) { __threadfence(); grid.sync(); }
    }
    if (M.prof && c_id == 0 && tid == 0) M.prof[2530] = gtime() - tb;
}

// the epilogue shared with the other direct solvers: solution checks, pose step, trial poses, pose part of computeScale.
// One thread-block cluster of kBackCluster CTAs (a single CTA took 41 us for C4's 2 000 poses, on every rank, every trial): every
// CTA takes a slice, the two reductions go through distributed shared memory and are added in rank order by every CTA, so all of
// them reach the same verdict and the sums do not depend on the schedule.
constexpr int kBackCluster = 8;
__global__ void __launch_bounds__(kBackThreads) k_dense_back(Batch B, DenseMat M) {
    namespace cg = cooperative_groups;
    __shared__ double s_red[32];
    __shared__ double s_part[2];
    cg::cluster_group cluster = cg::this_cluster();
    const int nc = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const WinDesc &wd = B.win[0];
    LMState &st = B.st[0];
    if (st.done) return;   // uniform over the cluster
    const int tid = threadIdx.x, gt = rank * kBackThreads + tid, stride = nc * kBackThreads;
    const int n = M.n;
    const double *__restrict__ x = M.x;
    const double lambda = (wd.trust == 0) ? (B.dbg_lambda >= 0.0 ? B.dbg_lambda : st.lambda) : 0.0;
    if (n == 0) {
        if (gt == 0) { st.ok = 1; st.scale_p = 0.0; }
        return;
    }
    const double *__restrict__ braw = B.red + B.red_bp_off;
    double bad = 0.0, sc = 0.0;
    for (int i = gt; i < n; i += stride) {
        const double xi = x[i];
        if (!isfinite(xi)) bad = 1.0;
        sc += xi * (lambda * xi + braw[i]);   // (only used when every entry is finite and the factorisation succeeded)
    }
    const double my_bad = block_sum(bad, s_red);
    __syncthreads();
    const double my_sc = block_sum(sc, s_red);
    if (tid == 0) { s_part[0] = my_bad; s_part[1] = my_sc; }
    cluster.sync();
    double anybad = 0.0, scale = 0.0;
    for (int r = 0; r < nc; ++r) {
        const double *p = cluster.map_shared_rank(s_part, r);
        anybad += p[0]; scale += p[1];
    }
    cluster.sync();   // nobody leaves while its shared memory is still being read
    const bool ok = (*M.flag == 0) && (anybad == 0.0);
    for (int i = gt; i < n; i += stride) B.xp[i] = ok ? x[i] : 0.0;
    const int cur = st.cur;
    const double *src = B.pose + (size_t)cur * B.tot_pose * kPoseStride;
    double *dst = B.pose + (size_t)(1 - cur) * B.tot_pose * kPoseStride;
    for (int p = gt; p < wd.n_pose; p += stride) {
        const int hi = B.pose_hidx[p];
        if (hi >= 0) {
            double dlt[6];
            for (int a = 0; a < 6; ++a) dlt[a] = ok ? x[6 * hi + a] : 0.0;
            pose_oplus(src + (size_t)p * kPoseStride, dlt, dst + (size_t)p * kPoseStride);
        }
    }
    if (gt == 0) { st.ok = ok ? 1 : 0; st.scale_p = ok ? scale : 0.0; }
}

}  // namespace dn
}  // namespace visfs

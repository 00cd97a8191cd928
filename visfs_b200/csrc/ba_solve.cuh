// ba_solve.cuh — k_solve: one CTA per window.  Adds the partial systems in part order, assembles the damped
// reduced camera system S (packed lower triangle in shared memory), solves it, applies CameraPose::update
// into the trial buffer and leaves sum x_p (lambda x_p + b_p) for k_control.
//   direct (Optimizer/Solver 0, 1, 3): blocked (6x6) Cholesky; a non-positive pivot fails the solve like
//           cs_chol does, which g2o treats as a rejected step
//   PCG    (Optimizer/Solver 2): g2o LinearSolverPCG, block-Jacobi preconditioner
// Two partial layouts (WinDesc::layout):
//   1  blocks (i <= j) x 36, then per pose H_pp (21) g (6) b_p (6)        — k_build_ws
//   0  blocks (i <  j) x 36, then per pose Hd = H_pp - Y W^T (21) g b_p   — k_build<MODE_BUILD, PPT>
#pragma once
#include "ba_math.cuh"
#include "ba_link.cuh"

namespace visfs {

constexpr int kSolveThreads = 256;

__device__ __forceinline__ int tri(int r, int c) { return r * (r + 1) / 2 + c; }  // r >= c

// Single-window latency: k_solve is one CTA and would pull every partial system through one SM (C2: 1.9 MB, 37 us).
// This kernel folds them into the first partial with many CTAs, in part order (deterministic), so k_solve reads one.
__global__ void k_reduce_parts(Batch B) {
    const int w = blockIdx.y;
    const WinDesc &wd = B.win[w];
    if (B.st[w].done || wd.n_parts <= 1) return;
    double *part = B.part + wd.part_off;
    const size_t stride = (size_t)wd.part_stride;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < wd.part_stride; idx += gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int c0 = 0; c0 < wd.n_parts; c0 += 16) {
            double v[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) v[u] = (c0 + u < wd.n_parts) ? part[(size_t)(c0 + u) * stride + idx] : 0.0;
#pragma unroll
            for (int u = 0; u < 16; ++u) s += v[u];
        }
        part[idx] = s;
    }
}

// k_solve: all registers to one CTA (single windows: latency); k_solve2: 128 registers, two CTAs per SM (batches: the kernel is a
// chain of dependent steps per window, a second resident window hides them; C3 x 1024 step 42.85 -> 42.13 ms, but one C1 window
// 0.717 -> 0.743 ms, hence the choice by batch size at launch)
__device__ __forceinline__ void solve_body(const Batch &B) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *S = reinterpret_cast<double *>(smem_raw);
    const int w = blockIdx.x;
    const WinDesc &wd = B.win[w];
    LMState &st = B.st[w];
    if (st.done) return;
    const int tid = threadIdx.x;
    const int F = st.F, n = 6 * F;
    const int ntri = n * (n + 1) / 2;   // row n of the packed triangle (n + 1 entries) holds the reduced rhs: the
    double *bs = S + ntri;               // factorisation's panel solves then perform the forward substitution for free
    double *braw = bs + n + 1;   // [n] raw b_p
    double *dinv = braw + n;     // [n] 1 / L_kk
    double *zb = dinv + n;       // [8] back-substitution scratch
    double *aux = zb + 8;        // PCG vectors: r, d, q, s, x, Minv blocks
    __shared__ int s_ok;
    __shared__ double s_red[32];
    __shared__ unsigned char tabR[kMaxSmallPoses * (kMaxSmallPoses + 1) / 2], tabC[kMaxSmallPoses * (kMaxSmallPoses + 1) / 2];
    const double lambda = (wd.trust == 0) ? (B.dbg_lambda >= 0.0 ? B.dbg_lambda : st.lambda) : 0.0;
    long long tclk[6];
    tclk[0] = clock64();
    if (tid == 0) s_ok = 1;
    if (n == 0) {
        if (tid == 0) { st.ok = 1; st.scale_p = 0.0; }
        return;
    }
    const bool all_pairs = wd.layout == 1;
    const int npairs = all_pairs ? F * (F + 1) / 2 : F * (F - 1) / 2;
    const int offd = npairs * 36;
    const double *part = B.part + wd.part_off;
    const int nparts = B.parts_reduced ? min(1, wd.n_parts) : wd.n_parts;
    const size_t stride = (size_t)wd.part_stride;
    // up to 32 independent loads in flight per entry (one CTA has to pull every partial through its own SM),
    // added in part order
    auto part_sum = [&](int idx) {
        double s = 0.0;
        for (int c0 = 0; c0 < nparts; c0 += 32) {
            double v[32];
#pragma unroll
            for (int u = 0; u < 32; ++u) v[u] = (c0 + u < nparts) ? part[(size_t)(c0 + u) * stride + idx] : 0.0;
#pragma unroll
            for (int u = 0; u < 32; ++u) s += v[u];
        }
        return s;
    };
    for (int i = tid; i < ntri; i += kSolveThreads) S[i] = 0.0;
    for (int t = tid; t < F * (F + 1) / 2; t += kSolveThreads) {   // block-pair table: t -> (rr >= cc)
        int rr = 0;
        while ((rr + 1) * (rr + 2) / 2 <= t) ++rr;
        tabR[t] = (unsigned char)rr;
        tabC[t] = (unsigned char)(t - rr * (rr + 1) / 2);
    }
    __syncthreads();
    tclk[1] = clock64();
    // pass 1: pair blocks (each entry of S written by exactly one thread)
    for (int idx = tid; idx < offd; idx += kSolveThreads) {
        const double v = part_sum(idx);
        const int p = idx / 36, q = idx - p * 36;
        int i = 0, base = 0, j;
        if (all_pairs) {
            while (base + (F - i) <= p) { base += F - i; ++i; }
            j = i + (p - base);
        } else {
            while (base + (F - 1 - i) <= p) { base += F - 1 - i; ++i; }
            j = i + 1 + (p - base);
        }
        const int a = q / 6, cc = q - a * 6;
        if (i < j) S[tri(6 * j + cc, 6 * i + a)] = v;          // block (i,j) entry (a,cc) lives at lower (6j+cc, 6i+a)
        else if (a >= cc) S[tri(6 * i + a, 6 * i + cc)] = v;   // diagonal block: lower half only
    }
    __syncthreads();
    // pass 2: per-pose sums
    for (int t2 = tid; t2 < F * kHStride; t2 += kSolveThreads) {
        const double v = part_sum(offd + t2);
        const int i = t2 / kHStride, k = t2 - i * kHStride;
        if (k < 21) {
            int a = 0, rem = k;
            while (rem >= 6 - a) { rem -= 6 - a; ++a; }
            const int cc = a + rem;
            S[tri(6 * i + cc, 6 * i + a)] += v + (a == cc ? lambda : 0.0);
        } else if (k < 27) {
            bs[6 * i + (k - 21)] = v;
        } else {
            braw[6 * i + (k - 27)] = v;
        }
    }
    __syncthreads();
    if (wd.n_link > 0) {
        // ---- odometry links: H_pp blocks (diagonal and pose-pose) and gradient pieces from the records of k_link_lin.
        //      Every diagonal entry / rhs entry has one owner thread that walks the links in order; the pose-pose blocks
        //      of different links are different entries (atomicAdd only guards duplicate links between one pair of poses)
        const int *hidx = B.pose_hidx + wd.pose_off;
        const double *lin = B.link_lin + (size_t)wd.link_off * kLinkStride;
        const int *lf = B.link_from + wd.link_off, *lt = B.link_to + wd.link_off;
        for (int t2 = tid; t2 < F * 36; t2 += kSolveThreads) {
            const int i = t2 / 36, q = t2 - i * 36, a = q / 6, cc = q - a * 6;
            if (cc > a) continue;
            double sacc = 0.0;
            for (int k = 0; k < wd.n_link; ++k) {
                if (hidx[lf[k]] == i) sacc += lin[(size_t)k * kLinkStride + kLkHii + q];
                if (hidx[lt[k]] == i) sacc += lin[(size_t)k * kLinkStride + kLkHjj + q];
            }
            S[tri(6 * i + a, 6 * i + cc)] += sacc;
        }
        for (int t2 = tid; t2 < F * 6; t2 += kSolveThreads) {
            const int i = t2 / 6, a = t2 - i * 6;
            double sacc = 0.0;
            for (int k = 0; k < wd.n_link; ++k) {
                if (hidx[lf[k]] == i) sacc += lin[(size_t)k * kLinkStride + kLkBi + a];
                if (hidx[lt[k]] == i) sacc += lin[(size_t)k * kLinkStride + kLkBj + a];
            }
            bs[t2] += sacc;
            braw[t2] += sacc;
        }
        for (int t2 = tid; t2 < wd.n_link * 36; t2 += kSolveThreads) {
            const int k = t2 / 36, q = t2 - k * 36, a = q / 6, cc = q - a * 6;
            const int i = hidx[lf[k]], j = hidx[lt[k]];
            if (i < 0 || j < 0) continue;
            const double v = lin[(size_t)k * kLinkStride + kLkHij + q];   // (J_i' Omega J_j)(a, cc): row pose i, column pose j
            if (i < j) atomicAdd(&S[tri(6 * j + cc, 6 * i + a)], v);
            else atomicAdd(&S[tri(6 * i + a, 6 * j + cc)], v);
        }
        __syncthreads();
    }
    if (B.dbg && w == 0) {
        for (int i = tid; i < ntri + n; i += kSolveThreads) B.dbg[i] = S[i];  // bs follows S
        __syncthreads();
    }

    tclk[2] = clock64();
    if (wd.solver != 2) {
        // ---- blocked right-looking Cholesky, 6x6 blocks, on rows 0..n (row n = rhs)
        for (int kb = 0; kb < F; ++kb) {
            const int base = 6 * kb;
            // every row owner factors the diagonal block redundantly in registers (no broadcast, no extra barrier)
            double L00, L10, L11, L20, L21, L22, L30, L31, L32, L33, L40, L41, L42, L43, L44, L50, L51, L52, L53, L54, L55;
            double i0, i1, i2, i3, i4, i5;
            bool okl = true;
            if (tid <= n - base) {
                const double *d0 = S + tri(base, base), *d1 = S + tri(base + 1, base), *d2 = S + tri(base + 2, base);
                const double *d3 = S + tri(base + 3, base), *d4 = S + tri(base + 4, base), *d5 = S + tri(base + 5, base);
                double d;
#define VISFS_PIVOT(dd, inv, diag) d = (dd); if (!(d > 0.0)) { okl = false; d = 1.0; } inv = rsqrt(d); diag = d * inv;
                VISFS_PIVOT(d0[0], i0, L00)
                L10 = d1[0] * i0; L20 = d2[0] * i0; L30 = d3[0] * i0; L40 = d4[0] * i0; L50 = d5[0] * i0;
                VISFS_PIVOT(fma(-L10, L10, d1[1]), i1, L11)
                L21 = fma(-L20, L10, d2[1]) * i1; L31 = fma(-L30, L10, d3[1]) * i1; L41 = fma(-L40, L10, d4[1]) * i1; L51 = fma(-L50, L10, d5[1]) * i1;
                VISFS_PIVOT(fma(-L21, L21, fma(-L20, L20, d2[2])), i2, L22)
                L32 = fma(-L31, L21, fma(-L30, L20, d3[2])) * i2; L42 = fma(-L41, L21, fma(-L40, L20, d4[2])) * i2;
                L52 = fma(-L51, L21, fma(-L50, L20, d5[2])) * i2;
                VISFS_PIVOT(fma(-L32, L32, fma(-L31, L31, fma(-L30, L30, d3[3]))), i3, L33)
                L43 = fma(-L42, L32, fma(-L41, L31, fma(-L40, L30, d4[3]))) * i3;
                L53 = fma(-L52, L32, fma(-L51, L31, fma(-L50, L30, d5[3]))) * i3;
                VISFS_PIVOT(fma(-L43, L43, fma(-L42, L42, fma(-L41, L41, fma(-L40, L40, d4[4])))), i4, L44)
                L54 = fma(-L53, L43, fma(-L52, L42, fma(-L51, L41, fma(-L50, L40, d5[4])))) * i4;
                VISFS_PIVOT(fma(-L54, L54, fma(-L53, L53, fma(-L52, L52, fma(-L51, L51, fma(-L50, L50, d5[5]))))), i5, L55)
#undef VISFS_PIVOT
                if (tid >= 6) {   // panel row r (r == n: the rhs): solve x L_d^T = S[r, base..base+5]
                    double *xr = S + tri(base + tid, base);
                    const double x0 = xr[0] * i0;
                    const double x1 = fma(-x0, L10, xr[1]) * i1;
                    const double x2 = fma(-x1, L21, fma(-x0, L20, xr[2])) * i2;
                    const double x3 = fma(-x2, L32, fma(-x1, L31, fma(-x0, L30, xr[3]))) * i3;
                    const double x4 = fma(-x3, L43, fma(-x2, L42, fma(-x1, L41, fma(-x0, L40, xr[4])))) * i4;
                    const double x5 = fma(-x4, L54, fma(-x3, L53, fma(-x2, L52, fma(-x1, L51, fma(-x0, L50, xr[5]))))) * i5;
                    xr[0] = x0; xr[1] = x1; xr[2] = x2; xr[3] = x3; xr[4] = x4; xr[5] = x5;
                }
            }
            __syncthreads();
            if (tid == 0) {   // the diagonal block's factor is stored only now: nobody reads that block any more
                double *d0 = S + tri(base, base), *d1 = S + tri(base + 1, base), *d2 = S + tri(base + 2, base);
                double *d3 = S + tri(base + 3, base), *d4 = S + tri(base + 4, base), *d5 = S + tri(base + 5, base);
                d0[0] = L00;
                d1[0] = L10; d1[1] = L11;
                d2[0] = L20; d2[1] = L21; d2[2] = L22;
                d3[0] = L30; d3[1] = L31; d3[2] = L32; d3[3] = L33;
                d4[0] = L40; d4[1] = L41; d4[2] = L42; d4[3] = L43; d4[4] = L44;
                d5[0] = L50; d5[1] = L51; d5[2] = L52; d5[3] = L53; d5[4] = L54; d5[5] = L55;
                dinv[base] = i0; dinv[base + 1] = i1; dinv[base + 2] = i2; dinv[base + 3] = i3; dinv[base + 4] = i4; dinv[base + 5] = i5;
                if (!okl) s_ok = 0;
            }
            // trailing update, one thread per half block (3 x 6 entries in registers, all operands loaded before the first
            // store so the shared-memory loads pipeline); the rhs row against every trailing column rides along
            const int m = F - kb - 1;
            const int halves = m * (m + 1);
            const int extra = 6 * m;
            for (int item = tid; item < halves + extra; item += kSolveThreads) {
                if (item < halves) {
                    const int blk = item >> 1, a0 = (item & 1) * 3;
                    const int r0 = 6 * (kb + 1 + tabR[blk]) + a0, c0 = 6 * (kb + 1 + tabC[blk]);
                    double L[18], acc[18];
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        const double *lr = S + tri(r0 + a, base);
                        const double *dst = S + tri(r0 + a, c0);
#pragma unroll
                        for (int k = 0; k < 6; ++k) { L[a * 6 + k] = lr[k]; acc[a * 6 + k] = (c0 + k <= r0 + a) ? dst[k] : 0.0; }
                    }
#pragma unroll
                    for (int c = 0; c < 6; ++c) {
                        const double *lc = S + tri(c0 + c, base);
                        const double w0 = lc[0], w1 = lc[1], w2 = lc[2], w3 = lc[3], w4 = lc[4], w5 = lc[5];
#pragma unroll
                        for (int a = 0; a < 3; ++a)
                            acc[a * 6 + c] -= fma(L[a * 6], w0, fma(L[a * 6 + 1], w1, fma(L[a * 6 + 2], w2, fma(L[a * 6 + 3], w3, fma(L[a * 6 + 4], w4, L[a * 6 + 5] * w5)))));
                    }
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        double *dst = S + tri(r0 + a, c0);
#pragma unroll
                        for (int k = 0; k < 6; ++k) if (c0 + k <= r0 + a) dst[k] = acc[a * 6 + k];   // diagonal block: lower half only
                    }
                } else {
                    const int cc = base + 6 + (item - halves);
                    const double *lr = S + tri(n, base), *lc = S + tri(cc, base);
                    double sacc = 0.0;
#pragma unroll
                    for (int k = 0; k < 6; ++k) sacc = fma(lr[k], lc[k], sacc);
                    S[tri(n, cc)] -= sacc;
                }
            }
            __syncthreads();
        }
        tclk[3] = clock64();
        // bs now holds y = L^-1 b (row n).  Back-substitution L^T x = y, one 6-block per step: six warps form the
        // column sums over the rows below the block, thread 0 solves the 6x6 triangle.
        for (int kb = F - 1; kb >= 0; --kb) {
            const int base = 6 * kb;
            const int wid = tid >> 5, lane = tid & 31;
            if (wid < 6) {
                double sacc = 0.0;
                for (int r = base + 6 + lane; r < n; r += 32) sacc = fma(S[tri(r, base + wid)], bs[r], sacc);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sacc += __shfl_down_sync(0xffffffffu, sacc, o);
                if (lane == 0) zb[wid] = bs[base + wid] - sacc;
            }
            __syncthreads();
            if (tid == 0) {
                const double *d1 = S + tri(base + 1, base), *d2 = S + tri(base + 2, base), *d3 = S + tri(base + 3, base);
                const double *d4 = S + tri(base + 4, base), *d5 = S + tri(base + 5, base);
                const double x5 = zb[5] * dinv[base + 5];
                const double x4 = fma(-d5[4], x5, zb[4]) * dinv[base + 4];
                const double x3 = fma(-d5[3], x5, fma(-d4[3], x4, zb[3])) * dinv[base + 3];
                const double x2 = fma(-d5[2], x5, fma(-d4[2], x4, fma(-d3[2], x3, zb[2]))) * dinv[base + 2];
                const double x1 = fma(-d5[1], x5, fma(-d4[1], x4, fma(-d3[1], x3, fma(-d2[1], x2, zb[1])))) * dinv[base + 1];
                const double x0 = fma(-d5[0], x5, fma(-d4[0], x4, fma(-d3[0], x3, fma(-d2[0], x2, fma(-d1[0], x1, zb[0]))))) * dinv[base];
                bs[base] = x0; bs[base + 1] = x1; bs[base + 2] = x2; bs[base + 3] = x3; bs[base + 4] = x4; bs[base + 5] = x5;
            }
            __syncthreads();
        }
    } else {
        // g2o LinearSolverPCG::solve: block-Jacobi preconditioner (inverse 6x6 diagonal blocks), x0 = 0,
        // at most n iterations; tolerance handling below.
        double *r = aux, *d = aux + n, *q = aux + 2 * n, *s = aux + 3 * n, *xv = aux + 4 * n, *Minv = aux + 5 * n;
        // invert the 6x6 diagonal blocks (one thread per pose, Gauss-Jordan with partial pivoting)
        for (int i = tid; i < F; i += kSolveThreads) {
            double M[6][12];
            for (int a = 0; a < 6; ++a)
                for (int c = 0; c < 6; ++c) {
                    const int rr = 6 * i + max(a, c), cc = 6 * i + min(a, c);
                    M[a][c] = S[tri(rr, cc)];
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
            for (int a = 0; a < 6; ++a) for (int c = 0; c < 6; ++c) Minv[36 * i + 6 * a + c] = M[a][6 + c];
        }
        for (int i = tid; i < n; i += kSolveThreads) { r[i] = bs[i]; xv[i] = 0.0; }
        __syncthreads();
        auto precond = [&](const double *in, double *out) {
            for (int i = tid; i < n; i += kSolveThreads) {
                const int blk = i / 6, a = i - 6 * blk;
                double acc = 0.0;
                for (int c = 0; c < 6; ++c) acc += Minv[36 * blk + 6 * a + c] * in[6 * blk + c];
                out[i] = acc;
            }
            __syncthreads();
        };
        auto dot = [&](const double *u, const double *v) {
            double acc = 0.0;
            for (int i = tid; i < n; i += kSolveThreads) acc += u[i] * v[i];
            const double t = block_sum(acc, s_red);
            __syncthreads();
            return t;
        };
        precond(r, d);
        double dn = dot(r, d);
        // upstream quirk kept: relative tolerance 1e-6 on the first solve after init(); once a solve has ended
        // with residual > 1e-6 the bound becomes 0 and the loop runs all n iterations
        double d0 = 1e-6 * dn;
        const double prev_res = st.pcg_residual;
        if (prev_res > 0.0 && prev_res > 1e-6) d0 = 0.0;
        for (int it = 0; it < n; ++it) {
            if (dn <= d0) break;
            for (int i = tid; i < n; i += kSolveThreads) {
                double acc = 0.0;
                for (int c = 0; c < n; ++c) acc += S[i >= c ? tri(i, c) : tri(c, i)] * d[c];
                q[i] = acc;
            }
            __syncthreads();
            const double a = dn / dot(d, q);
            for (int i = tid; i < n; i += kSolveThreads) { xv[i] += a * d[i]; r[i] -= a * q[i]; }
            __syncthreads();
            precond(r, s);
            const double dold = dn;
            dn = dot(r, s);
            const double ba = dn / dold;
            for (int i = tid; i < n; i += kSolveThreads) d[i] = s[i] + ba * d[i];
            __syncthreads();
        }
        for (int i = tid; i < n; i += kSolveThreads) bs[i] = xv[i];
        if (tid == 0) st.pcg_residual = 0.5 * dn;
        __syncthreads();
    }

    tclk[4] = clock64();
    // solution checks, pose step, trial poses, scale
    double bad = 0.0;
    for (int i = tid; i < n; i += kSolveThreads) if (!isfinite(bs[i])) bad = 1.0;
    const double anybad = block_sum(bad, s_red);
    const bool ok = (s_ok != 0) && (anybad == 0.0);
    __syncthreads();
    double *xp = B.xp + (size_t)wd.pose_off * 6;
    double sc = 0.0;
    for (int i = tid; i < n; i += kSolveThreads) {
        const double x = ok ? bs[i] : 0.0;
        xp[i] = x;
        sc += x * (lambda * x + braw[i]);
    }
    const double scale = block_sum(sc, s_red);
    const int cur = st.cur;
    const double *src = B.pose + ((size_t)cur * B.tot_pose + wd.pose_off) * kPoseStride;
    double *dst = B.pose + ((size_t)(1 - cur) * B.tot_pose + wd.pose_off) * kPoseStride;
    for (int p = tid; p < wd.n_pose; p += kSolveThreads) {
        const int hi = B.pose_hidx[wd.pose_off + p];
        if (hi >= 0) {
            double dlt[6];
            for (int a = 0; a < 6; ++a) dlt[a] = ok ? bs[6 * hi + a] : 0.0;
            pose_oplus(src + p * kPoseStride, dlt, dst + p * kPoseStride);
        }
    }
    double lchi = 0.0;
    if (wd.n_link > 0) {   // chi2 of the links at the trial poses (this CTA has just written them)
        __syncthreads();
        lchi = link_chi2_block(B, wd, 1 - cur, s_red);
    }
    if (tid == 0) {
        st.ok = ok ? 1 : 0; st.scale_p = scale;
        st.link_chi_trial = lchi;
        tclk[5] = clock64();
        if (wd.solver == 2) tclk[3] = tclk[4];
        for (int k = 0; k < 6; ++k) st.t_solve[k] = tclk[k] - tclk[0];
    }
}

__global__ void __launch_bounds__(kSolveThreads) k_solve(Batch B) { solve_body(B); }
__global__ void __launch_bounds__(kSolveThreads, 2) k_solve2(Batch B) { solve_body(B); }

}  // namespace visfs

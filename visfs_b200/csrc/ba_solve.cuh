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

namespace visfs {

constexpr int kSolveThreads = 256;

__device__ __forceinline__ int tri(int r, int c) { return r * (r + 1) / 2 + c; }  // r >= c

__global__ void __launch_bounds__(kSolveThreads) k_solve(Batch B) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *S = reinterpret_cast<double *>(smem_raw);
    const int w = blockIdx.x;
    const WinDesc &wd = B.win[w];
    LMState &st = B.st[w];
    if (st.done) return;
    const int tid = threadIdx.x;
    const int F = st.F, n = 6 * F;
    const int ntri = n * (n + 1) / 2;
    double *bs = S + ntri;       // [n] reduced rhs, overwritten by the solution
    double *braw = bs + n;       // [n] raw b_p
    double *dinv = braw + n;     // [n] 1 / L_kk
    double *aux = dinv + n;      // PCG vectors: r, d, q, s, x, Minv blocks
    __shared__ int s_ok;
    __shared__ double s_red[32];
    __shared__ unsigned char tabR[kMaxSmallPoses * (kMaxSmallPoses + 1) / 2], tabC[kMaxSmallPoses * (kMaxSmallPoses + 1) / 2];
    const double lambda = (wd.trust == 0) ? (B.dbg_lambda >= 0.0 ? B.dbg_lambda : st.lambda) : 0.0;
    if (tid == 0) s_ok = 1;
    if (n == 0) {
        if (tid == 0) { st.ok = 1; st.scale_p = 0.0; }
        return;
    }
    const bool all_pairs = wd.layout == 1;
    const int npairs = all_pairs ? F * (F + 1) / 2 : F * (F - 1) / 2;
    const int offd = npairs * 36;
    const double *part = B.part + wd.part_off;
    const int nparts = wd.n_parts;
    const size_t stride = (size_t)wd.part_stride;
    auto part_sum = [&](int idx) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int c = 0;
        for (; c + 3 < nparts; c += 4) {
            s0 += part[(size_t)c * stride + idx];
            s1 += part[(size_t)(c + 1) * stride + idx];
            s2 += part[(size_t)(c + 2) * stride + idx];
            s3 += part[(size_t)(c + 3) * stride + idx];
        }
        for (; c < nparts; ++c) s0 += part[(size_t)c * stride + idx];
        return (s0 + s1) + (s2 + s3);
    };
    for (int i = tid; i < ntri; i += kSolveThreads) S[i] = 0.0;
    for (int t = tid; t < F * (F + 1) / 2; t += kSolveThreads) {   // block-pair table: t -> (rr >= cc)
        int rr = 0;
        while ((rr + 1) * (rr + 2) / 2 <= t) ++rr;
        tabR[t] = (unsigned char)rr;
        tabC[t] = (unsigned char)(t - rr * (rr + 1) / 2);
    }
    __syncthreads();
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
    if (B.dbg && w == 0) {
        for (int i = tid; i < ntri + n; i += kSolveThreads) B.dbg[i] = S[i];  // bs follows S
        __syncthreads();
    }

    if (wd.solver != 2) {
        // ---- blocked right-looking Cholesky, 6x6 blocks
        for (int kb = 0; kb < F; ++kb) {
            const int base = 6 * kb;
            // every row owner factors the diagonal block redundantly in registers (no broadcast, no extra barrier)
            double L[21], inv[6];
            bool okl = true;
            if (tid < n - base) {
#pragma unroll
                for (int a = 0; a < 6; ++a)
#pragma unroll
                    for (int c = 0; c <= a; ++c) L[a * (a + 1) / 2 + c] = S[tri(base + a, base + c)];
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    double d = L[c * (c + 1) / 2 + c];
#pragma unroll
                    for (int k = 0; k < c; ++k) d = fma(-L[c * (c + 1) / 2 + k], L[c * (c + 1) / 2 + k], d);
                    if (!(d > 0.0)) { okl = false; d = 1.0; }
                    const double r = rsqrt(d);
                    inv[c] = r;
                    L[c * (c + 1) / 2 + c] = d * r;
#pragma unroll
                    for (int a = c + 1; a < 6; ++a) {
                        double v = L[a * (a + 1) / 2 + c];
#pragma unroll
                        for (int k = 0; k < c; ++k) v = fma(-L[a * (a + 1) / 2 + k], L[c * (c + 1) / 2 + k], v);
                        L[a * (a + 1) / 2 + c] = v * r;
                    }
                }
                if (tid >= 6) {   // panel row r: solve x L_d^T = S[r, base..base+5]
                    const int r = base + tid;
                    double x[6];
#pragma unroll
                    for (int c = 0; c < 6; ++c) x[c] = S[tri(r, base + c)];
#pragma unroll
                    for (int c = 0; c < 6; ++c) {
                        double v = x[c];
#pragma unroll
                        for (int k = 0; k < c; ++k) v = fma(-x[k], L[c * (c + 1) / 2 + k], v);
                        x[c] = v * inv[c];
                    }
#pragma unroll
                    for (int c = 0; c < 6; ++c) S[tri(r, base + c)] = x[c];
                }
            }
            __syncthreads();
            if (tid < 6) {   // the diagonal block's factor is stored only now: nobody reads that block any more
#pragma unroll
                for (int a = 0; a < 6; ++a)
                    if (a == tid) {
#pragma unroll
                        for (int c = 0; c <= a; ++c) S[tri(base + a, base + c)] = L[a * (a + 1) / 2 + c];
                        dinv[base + a] = inv[a];
                    }
                if (tid == 0 && !okl) s_ok = 0;
            }
            const int m = F - kb - 1;
            const int items = m * (m + 1) / 2 * 36;
            for (int item = tid; item < items; item += kSolveThreads) {
                const int blk = item / 36, q = item - blk * 36;
                const int a = q / 6, c = q - a * 6;
                const int r = 6 * (kb + 1 + tabR[blk]) + a, cc = 6 * (kb + 1 + tabC[blk]) + c;
                if (cc > r) continue;
                const double *lr = S + tri(r, base), *lc = S + tri(cc, base);
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < 6; ++k) s = fma(lr[k], lc[k], s);
                S[tri(r, cc)] -= s;
            }
            __syncthreads();
        }
        // triangular solves by warp 0 (column-oriented), x overwrites bs
        if (tid < 32) {
            for (int k = 0; k < n; ++k) {
                const double xk = bs[k] * dinv[k];
                __syncwarp();
                if (tid == 0) bs[k] = xk;
                for (int r = k + 1 + tid; r < n; r += 32) bs[r] = fma(-S[tri(r, k)], xk, bs[r]);
                __syncwarp();
            }
            for (int k = n - 1; k >= 0; --k) {
                const double xk = bs[k] * dinv[k];
                __syncwarp();
                if (tid == 0) bs[k] = xk;
                for (int r = tid; r < k; r += 32) bs[r] = fma(-S[tri(k, r)], xk, bs[r]);
                __syncwarp();
            }
        }
        __syncthreads();
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
    if (tid == 0) { st.ok = ok ? 1 : 0; st.scale_p = scale; }
}

}  // namespace visfs

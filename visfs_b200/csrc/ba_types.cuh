// ba_types.cuh — device-side data model of the bundle-adjustment batch.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace visfs {

constexpr int kMaxSmallPoses = 32;   // windows with <= 32 poses take the shared-memory ("small") path
constexpr int kTileEdges = 192;      // edge slots staged per tile (one thread per edge)
constexpr int kTileLm = 32;          // landmarks per tile
constexpr int kMaxDegLarge = 32;     // landmark degree limit of the large-window path (one lane per edge)
constexpr int kThreads = 256;        // CTA size of the build / update kernels
constexpr int kPoseStride = 16;      // doubles per pose record: t(3) q(4: x y z w) R(9 row-major)
constexpr int kPoseSm = 17;          // shared-memory stride of a pose record: odd, so lanes reading different poses hit different banks
constexpr int kHStride = 33;         // per-edge pose-side staging: Hd(21) g(6) b_p(6)

// edge_pose word: bits 0-23 window-local pose index, bit 24 mono, bit 25 culled (level 1)
constexpr int kPoseMask = 0x00FFFFFF;
constexpr int kMonoBit = 1 << 24;
constexpr int kCulledBit = 1 << 25;

// lm_flags / pose_flags bits
constexpr uint8_t kFixed = 1;        // vertex->fixed()
constexpr uint8_t kInHessian = 2;    // free and active in the current pass (has a hessian index)

struct WinDesc {
    int pose_off, n_pose;            // into the concatenated pose arrays
    int point_off, n_point;
    int edge_off, n_edge;
    int chunk_off, n_chunks;         // work items of this window
    long long part_off;              // doubles, into the partial buffer (chunk k at part_off + k * part_stride)
    int part_stride;
    int max_iter;                    // LM iterations per pass
    int solver, trust;
    unsigned flags;
    int large;                       // 1: block-sparse / global-memory path (P > kMaxSmallPoses)
    int layout;                      // partial-system layout written by the build kernel (see ba_solve.cuh)
    int n_parts;                     // partial systems k_solve adds (= n_chunks / cluster size)
    int link_off, n_link;            // odometry links of this window (ba_link.cuh)
    double fx, fy, cx, cy, bf, inv_pv, delta;
    double inv_ov;                   // 1 / Optimizer/OdometryCovariance
};

struct LMState {
    double lambda, ni, cur_chi, trial_chi, rho, scale_p;
    double chi_initial, chi_pass[2], chi_last_trial, lambda_final[2];
    double pcg_residual;             // g2o LinearSolverPCG::_residual (reset by init() at every optimize())
    double link_chi_trial;           // chi2 of the odometry links at the trial state (k_solve)
    int iter, qmax, done, cur;       // cur: which state buffer holds the accepted estimate
    int F, NL, ok, fresh;            // fresh: 1 until the first trial of the pass has run
    int iterations_run[2], trials_run[2], stop[2], nF[2], nNL[2];
    int n_outliers, status, pass, err;
    long long t_solve[6];            // k_solve phase clocks of the last trial (assemble, factor, back-subst, epilogue)
};

struct Chunk {
    int win;
    int lm0, lm1;                    // global landmark range [lm0, lm1)
};

// One tile of a chunk: <= kTileLm landmarks, <= kTileEdges edges (built on the device at upload, k_fill_tiles)
struct Tile {
    int lt, ntl;                     // first global landmark, landmark count
    int e0, ne;                      // first global edge, edge count
};

// All device pointers of one uploaded batch.
struct Batch {
    int n_win, n_chunks;
    int tot_pose, tot_point, tot_edge;
    const WinDesc *win;
    LMState *st;
    const Chunk *chunks;
    double *pose;                    // [2][tot_pose][kPoseStride]
    double *point;                   // [2][tot_point][3]
    uint8_t *pose_flags;             // [tot_pose]
    uint8_t *lm_flags;               // [tot_point]
    int *pose_hidx;                  // [tot_pose]
    int *pose_active;                // [tot_pose] scratch
    int *point_hidx;                 // [tot_point]  (export only)
    const int *lm_edge_off;          // [tot_point + 1] global edge offsets (CSR by landmark)
    const double *obs_u, *obs_v, *obs_r;   // [tot_edge] SoA
    int *edge_pose;                  // [tot_edge] packed word (see above)
    const int *edge_point;           // [tot_edge] window-local landmark index
    const int *edge_orig;            // [tot_edge] caller's edge index (window-local)
    unsigned *covis;                 // [tot_pose] row i: bit j set if S block (i,j), i<=j exists (small path)
    double *part;                    // per-chunk partial sums
    double *part2;                   // [n_chunks][2] chi2 / scale partials of the update kernel
    double *lm_sum;                  // [tot_point][9] H_ll (6) b_l (3) of the current trial, left by k_build_ws for k_update (null: k_update forms them itself)
    double *xp;                      // [tot_pose][6] pose step per hessian index (window-local)
    int *n_running;                  // windows still running in the current pass
    int *ctl_count;                  // [n_win] CTAs of k_update that have finished this trial: the last one runs the LM controller
    const Tile *tiles;               // tile table, chunk c owns tiles [chunk_tile_off[c], chunk_tile_off[c + 1])
    const int *chunk_tile_off;       // [n_chunks + 1]
    const int *chunk_regular;        // [n_chunks] 1: all landmarks of the chunk are seen by the same poses in the same order (or null)
    const Tile *wtiles;              // warp tiles of k_update (<= 32 edges, whole landmarks), same indexing
    const int *chunk_wtile_off;      // [n_chunks + 1]
    // large-window path (ba_large.cuh): block-skyline reduced camera system in the reduce buffer
    int *sky_first;                  // [F] first block column of lower row r
    long long *sky_off;              // [F + 1] row offsets, in blocks
    int *col_ptr;                    // [F + 1] column structure of the envelope (rows r > k with sky_first[r] <= k)
    int *col_rows;
    double *red;                     // [n_sky * 36 | g (6F) | b_p (6F)]  — what a partitioned run all-reduces
    long long red_g_off, red_bp_off;
    double *hdiag;                   // [6F] diag(H_pp) of the INIT pass (lambda init)
    // odometry links (ba_link.cuh)
    int tot_link;
    const int *link_win, *link_from, *link_to;   // [tot_link] window, window-local pose indices
    const double *link_m;            // [tot_link][7] measurement t, q
    double *link_lin;                // [tot_link][kLinkStride] per-trial records of k_link_lin
    int parts_reduced;               // 1: k_reduce_parts has folded all partial systems of a window into its first one
    double *dbg;                     // parity hook: k_solve dumps packed S and b_s of window 0 here (else null)
    double dbg_lambda;               // parity hook: damping override (< 0: keep the LM state's)
};

}  // namespace visfs

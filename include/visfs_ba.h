/*
 * visfs_ba.h — C ABI of the B200-native local bundle adjustment (VISFS drop-in).
 *
 * This is the one boundary between VISFS's C++17 host code and the sm_100a CUDA
 * kernels.  Plain C, POD only, no exceptions cross it, no torch / Eigen types.
 *
 * What each entry point replaces in the reference (paths relative to the VISFS tree):
 *
 *   visfs_ba_create / visfs_ba_destroy
 *       the `g2o::SparseOptimizer optimizer;` + linear-solver + algorithm set-up that
 *       corelib/src/Optimizer/Optimizer.cpp:75-97 performs on every call.  The handle
 *       pools device buffers and one CUDA stream; results never depend on earlier calls.
 *
 *   visfs_ba_solve
 *       everything between graph construction and read-back in
 *       corelib/src/Optimizer/Optimizer.cpp:100-318 for the visual (stereo / mono)
 *       edges: vertex + edge set-up (100-114, 152-223), initializeOptimization +
 *       optimize(iterations/2) (261-265), the chi2 guards (268-280, 315-318), outlier
 *       culling by plain chi2 > delta (283-309) and the second pass (310-311); with n_links > 0
 *       also the odometry pose-pose constraints of lines 116-150 (EdgePoseConstraint).
 *       The arithmetic it reproduces is corelib/include/Optimizer/g2o/OptimizeTypeDefine.h
 *       :16-191 (CameraPose, VertexPose, EdgeStereo) and
 *       corelib/src/Optimizer/g2o/OptimizeTypeDefine.cpp:7-14 (CameraPose::update),
 *       plus g2o's BlockSolver_6_3 / OptimizationAlgorithmLevenberg /
 *       RobustKernelHuber semantics.
 *
 *   visfs_ba_solve_batch
 *       N independent visfs_ba_solve calls (the reference would run them one after
 *       another on the Estimator thread, corelib/src/Estimator.cpp:254); on the GPU all
 *       windows advance through one launch sequence with per-window LM state.
 *
 *   visfs_ba_linearize     (parity / debug)
 *       EdgeStereo::computeError + linearizeOplus (OptimizeTypeDefine.h:121-178) and
 *       g2o::RobustKernelHuber::robustify for every edge at the input state.
 *
 *   visfs_ba_structure     (parity / debug)
 *       g2o SparseOptimizer::initializeOptimization(level) / buildIndexMapping and
 *       BlockSolver::buildStructure as invoked from Optimizer.cpp:262 and :310.
 *
 *   visfs_ba_upload / visfs_ba_run_resident / visfs_ba_download
 *       the same solve split into H2D, device-only LM, D2H so that callers who keep a
 *       window resident and the benchmark can time the device part.
 *
 *   visfs_ba_comm_* / VISFS_BA_FLAG_PARTITIONED
 *       no reference equivalent (the reference is single process).  Global BA with the
 *       landmarks partitioned across ranks; reduced camera system summed with one
 *       ncclAllReduce per LM trial (NCCL is bound with dlopen at visfs_ba_comm_init).
 *
 * Limits
 *   - windows of up to 32 poses: any number per batch, landmarks with up to 192 observations;
 *   - windows with more poses, and partitioned problems: one per call, landmarks with up to 32
 *     observations (block-skyline path); all Optimizer/Solver values and odometry links (n_links > 0) are implemented
 *     on both paths.  In a partitioned run every rank passes ALL links.
 *
 * Conventions
 *   - poses are T_camera<-world, stored t(3) then quaternion x,y,z,w  (CameraPose::toVector,
 *     OptimizeTypeDefine.h:57-67); `pose_id` strictly ascending (std::map order of
 *     Optimizer.cpp:100).  Like the CameraPose constructor (OptimizeTypeDefine.h:36-41) the library forces
 *     w >= 0 and unit norm on every input quaternion, and like g2o::SE3Quat on every odometry measurement
 *     (link_tq): the sign of a quaternion the caller passes never changes a result.
 *   - points are world-frame xyz; `point_id` strictly ascending (Optimizer.cpp:156).
 *   - edges in g2o insertion order: grouped by point, ascending pose inside a point
 *     (Optimizer.cpp:156-169).  Unsorted edge lists are accepted and stably sorted on
 *     the device by (point, pose); all per-edge outputs are in the CALLER's edge order.
 *     At most ONE edge per (point, pose) pair — the reference's std::map<pose id, FeatureBA> per feature
 *     (Optimizer.h:52) cannot hold two; a duplicate is rejected with VISFS_BA_ERR_INVALID.
 *   - edge_obs = (u, v, u_right) as built at Optimizer.cpp:187-188; u_right is ignored
 *     for mono edges (edge_kind == VISFS_BA_EDGE_MONO).
 *   - every function returns a visfs_ba_status; on error visfs_ba_last_error() has text.
 *   - there is no CPU fallback: without a CUDA device every compute call fails with
 *     VISFS_BA_ERR_CUDA.
 */
#ifndef VISFS_BA_H
#define VISFS_BA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VISFS_BA_ABI_VERSION 5

typedef enum visfs_ba_status {
    VISFS_BA_OK = 0,
    VISFS_BA_ERR_INVALID = 1,     /* bad arguments / malformed problem                     */
    VISFS_BA_ERR_CUDA = 2,        /* CUDA or NCCL failure (text in visfs_ba_last_error)    */
    VISFS_BA_ERR_NUMERIC_PASS1 = 3, /* chi2 NaN / inf / > 1e12 after pass 1 (Optimizer.cpp:272-280) */
    VISFS_BA_ERR_NUMERIC_PASS2 = 4, /* chi2 > 1e12 after pass 2 (Optimizer.cpp:315-318)      */
    VISFS_BA_ERR_UNSUPPORTED = 5
} visfs_ba_status;

enum { VISFS_BA_EDGE_STEREO = 0, VISFS_BA_EDGE_MONO = 1 };

/* Optimizer/Solver values (corelib/include/Parameters.h:185) */
enum { VISFS_BA_SOLVER_CSPARSE = 0, VISFS_BA_SOLVER_CHOLMOD = 1, VISFS_BA_SOLVER_PCG = 2, VISFS_BA_SOLVER_EIGEN = 3 };
/* Optimizer/TrustRegion values (Parameters.h:186) */
enum { VISFS_BA_LEVENBERG = 0, VISFS_BA_GAUSS_NEWTON = 1 };

/* why a pass stopped */
enum {
    VISFS_BA_STOP_NOT_RUN = 0,
    VISFS_BA_STOP_ITERATIONS = 1,   /* ran iterations/2 LM iterations                              */
    VISFS_BA_STOP_TERMINATE = 2,    /* g2o "Terminate": 10 failed trials, rho == 0 or lambda not finite */
    VISFS_BA_STOP_EMPTY = 3,        /* no free active vertex (g2o optimize() returns -1)           */
    VISFS_BA_STOP_SOLVER_FAIL = 4   /* Gauss-Newton linear solve failed (g2o "Fail")               */
};

enum {
    VISFS_BA_FLAG_PARTITIONED = 1u << 0, /* problem holds one rank's landmark partition (global BA) */
    VISFS_BA_FLAG_SINGLE_PASS = 1u << 1, /* run pass 1 only, all `iterations` of it (benchmark aid)  */
    VISFS_BA_FLAG_NO_CULL     = 1u << 2  /* keep every edge at level 0 after pass 1                  */
};

typedef struct visfs_ba_handle visfs_ba_handle;

typedef struct visfs_ba_config {
    int32_t abi_version;     /* VISFS_BA_ABI_VERSION                                        */
    int32_t device;          /* CUDA device ordinal                                         */
    int32_t profile_kernels; /* != 0: bracket kernel classes with CUDA events (small cost)   */
    int32_t reserved[5];
} visfs_ba_config;

typedef struct visfs_ba_problem {
    int32_t n_poses, n_points, n_edges;
    uint32_t flags;
    const double  *pose_tq;      /* [P][7]  tx ty tz qx qy qz qw  (T_cw)                    */
    const int64_t *pose_id;      /* [P]     strictly ascending; may be NULL (= 0..P-1)       */
    const uint8_t *pose_fixed;   /* [P]     1 = fixed (Optimizer.cpp:111)                    */
    const double  *point_xyz;    /* [L][3]                                                   */
    const int64_t *point_id;     /* [L]     strictly ascending; may be NULL                  */
    const uint8_t *point_fixed;  /* [L]     1 = fixed (Optimizer.cpp:165)                    */
    const double  *edge_obs;     /* [E][3]  u, v, u_right                                    */
    const int32_t *edge_pose;    /* [E]     index into poses                                 */
    const int32_t *edge_point;   /* [E]     index into points                                */
    const uint8_t *edge_kind;    /* [E]     VISFS_BA_EDGE_*; may be NULL (= all stereo)      */
    double fx, fy, cx, cy, bf;   /* EdgeStereo intrinsics (Optimizer.cpp:191-195)            */
    double pixel_variance;       /* Optimizer/PixelVariance; information = I / variance      */
    double huber_delta;          /* Optimizer/RobustKernelDelta; <= 0 : no kernel, one pass  */
    int32_t iterations;          /* Optimizer/Iterations; each pass runs iterations / 2      */
    int32_t solver;              /* VISFS_BA_SOLVER_*                                        */
    int32_t trust_region;        /* VISFS_BA_LEVENBERG / VISFS_BA_GAUSS_NEWTON               */
    int32_t n_links;             /* odometry pose-pose constraints, 0 = none (Optimizer.cpp:116-150) */
    /* EdgePoseConstraint (OptimizeTypeDefine.h:193-225, OptimizeTypeDefine.cpp:35-88): vertex 0 = from, vertex 1 = to,
     * measurement T_c1c2 = T_rc^-1 * transform * T_rc (Optimizer.cpp:131-140), information = I6 / odometry_variance,
     * no robust kernel, never culled.  Links in std::map order of their index; from != to. */
    const int32_t *link_from;    /* [n_links] index into poses                               */
    const int32_t *link_to;      /* [n_links] index into poses                               */
    const double  *link_tq;      /* [n_links][7] measurement as tx ty tz qx qy qz qw         */
    double odometry_variance;    /* Optimizer/OdometryCovariance (Parameters.h:189)          */
    /* Optional: the observations as floats, used INSTEAD of edge_obs when not NULL (edge_obs may then be NULL).  The
     * reference's observations are floats widened to double (cv::KeyPoint and FeatureBA::depth are float, and u_right
     * = x - disparity is a float subtraction, Optimizer.cpp:187-188), so this form loses nothing and moves 12 instead of
     * 24 bytes per edge over PCIe; the device widens them back. */
    const float   *edge_obs_f32; /* [E][3]  u, v, u_right                                    */
} visfs_ba_problem;

typedef struct visfs_ba_result {
    /* caller-allocated outputs; any of the three may be NULL */
    double  *pose_tq;            /* [P][7]  optimised T_cw                                   */
    double  *point_xyz;          /* [L][3]  optimised points (raw; the 5 m clamp of          */
                                 /*         Optimizer.cpp:350 is applied by the C++ shim)    */
    uint8_t *edge_level;         /* [E]     1 = culled after pass 1 (Optimizer.cpp:285-286)  */
    /* filled by the library */
    int32_t status;              /* visfs_ba_status of this window                           */
    int32_t n_outliers;
    int32_t iterations_run[2];   /* LM iterations executed per pass                          */
    int32_t trials_run[2];       /* LM trials (damped solves) executed per pass              */
    int32_t stop_reason[2];      /* VISFS_BA_STOP_*                                          */
    int32_t n_free_poses[2];     /* free active poses per pass                               */
    int32_t n_free_points[2];    /* free active points per pass                              */
    double  chi2_initial;        /* robust chi2 of the input state                           */
    double  chi2_pass1;          /* robust chi2 of the accepted state after pass 1 (:271)    */
    double  chi2_final;          /* robust chi2 of the accepted state after pass 2           */
    double  chi2_last_trial;     /* robust chi2 of the last evaluated trial (:315 reads this)*/
    double  lambda_final[2];     /* LM damping at the end of each pass                       */
} visfs_ba_result;

/* per-edge linearisation at the INPUT state (all arrays caller-allocated, may be NULL) */
typedef struct visfs_ba_linearization {
    double *error;     /* [E][3]  obs - proj (row 2 = 0 for mono)                            */
    double *chi2;      /* [E]     e' Omega e                                                 */
    double *rho;       /* [E]     robustified chi2 (= chi2 when huber_delta <= 0)            */
    double *weight;    /* [E]     rho' (1 for inliers / no kernel)                           */
    double *J_point;   /* [E][3][3] row-major d e / d point                                  */
    double *J_pose;    /* [E][3][6] row-major d e / d (t, theta)                             */
} visfs_ba_linearization;

/* index maps and block patterns for one optimisation level (bit-exact gate) */
typedef struct visfs_ba_structure {
    const uint8_t *edge_level;   /* in  [E] or NULL: edges with level != 0 are inactive      */
    int32_t *pose_hidx;          /* out [P]  hessian index, -1 = fixed or inactive           */
    int32_t *point_hidx;         /* out [L]  hessian index (poses first), -1 likewise        */
    uint8_t *edge_active;        /* out [E]  1 = in g2o's activeEdges                        */
    int32_t *hpl_row;            /* out [E]  pose block row of the edge's H_pl block, -1 none*/
    int32_t *hpl_col;            /* out [E]  landmark block column (0-based), -1 none        */
    int32_t *schur_rows;         /* out [schur_capacity] block rows, sorted by (col,row)     */
    int32_t *schur_cols;         /* out [schur_capacity]                                     */
    int32_t schur_capacity;      /* in                                                       */
    int32_t n_schur_blocks;      /* out (may exceed capacity: then lists are truncated)      */
    int32_t n_free_poses;        /* out                                                      */
    int32_t n_free_points;       /* out                                                      */
    int32_t n_active_edges;      /* out                                                      */
    int32_t n_hpl_blocks;        /* out                                                      */
} visfs_ba_structure;

/* device-side timing and work counters of the last run (CUDA events on the library's stream) */
typedef struct visfs_ba_timing {
    double total_ms;             /* whole device LM (both passes, all windows)               */
    double build_ms;             /* linearise + Hessian + Schur kernel launches, summed      */
    double solve_ms;             /* reduced-system solve launches                            */
    double update_ms;            /* back-substitution + update + chi2 launches               */
    double other_ms;             /* structure, control, culling                              */
    int64_t build_launches, solve_launches, update_launches, other_launches;
    int64_t lm_iterations;       /* summed over windows and passes                           */
    int64_t lm_trials;           /* summed over windows and passes                           */
    int64_t edge_trials;         /* sum over trials of the active edges they linearised      */
    int64_t alg_bytes_build;     /* algorithmic bytes of all build launches   */
    int64_t alg_bytes_update;    /* algorithmic bytes of all update launches                 */
    int64_t kernel_launches;     /* kernels launched by the last visfs_ba_run_resident       */
    int64_t h2d_bytes;           /* bytes copied host->device by the last upload             */
    int64_t d2h_bytes;           /* bytes copied device->host by the last run + download     */
    int64_t solve_clocks[6];     /* window 0, last trial: SM clocks at k_solve's phase boundaries */
} visfs_ba_timing;

int  visfs_ba_abi_version(void);
int  visfs_ba_create(const visfs_ba_config *cfg, visfs_ba_handle **out);
void visfs_ba_destroy(visfs_ba_handle *h);
const char *visfs_ba_last_error(const visfs_ba_handle *h);

int visfs_ba_solve(visfs_ba_handle *h, const visfs_ba_problem *problem, visfs_ba_result *result);
int visfs_ba_solve_batch(visfs_ba_handle *h, int32_t n, const visfs_ba_problem *problems,
                         visfs_ba_result *results);

int visfs_ba_linearize(visfs_ba_handle *h, const visfs_ba_problem *problem,
                       visfs_ba_linearization *out);
int visfs_ba_structure_build(visfs_ba_handle *h, const visfs_ba_problem *problem,
                             visfs_ba_structure *out);

/* parity / debug: one damped trial at the input state.  Fills the dense reduced camera system
 * (n x n row-major, n = 6 * free poses, damping included), its right-hand side, the pose step and the
 * trial points.  lambda < 0 uses g2o's initial damping 1e-5 * max|diag H|. */
int visfs_ba_debug_trial(visfs_ba_handle *h, const visfs_ba_problem *problem, double lambda, double *S_dense,
                         double *b_s, double *x_pose, double *trial_points, int32_t *n_out, double *chi2_out,
                         double *lambda_out, double *trial_chi2_out);

/* parity / debug: the device functions of the product path on caller-supplied operands, so that tests can hold them to
 * the reference's own compiled code (oracle/_ref).  CameraPose::update + deltaQ (OptimizeTypeDefine.cpp:7-14,
 * Math.h:277-287) of n poses: tq_in [n][7], delta [n][6] -> tq_out [n][7]. */
int visfs_ba_debug_pose_oplus(visfs_ba_handle *h, int32_t n, const double *tq_in, const double *delta, double *tq_out);
/* EdgePoseConstraint::computeError / linearizeOplus (OptimizeTypeDefine.cpp:35-72) of n (from, to, measurement) triples,
 * each [n][7]; the measurement is used as given (the solve path normalises it like g2o::SE3Quat first).
 * err [n][6], J_from / J_to [n][6][6] row-major. */
int visfs_ba_debug_link_linearize(visfs_ba_handle *h, int32_t n, const double *from_tq, const double *to_tq,
                                  const double *meas_tq, double *err, double *J_from, double *J_to);

/* resident-window API: upload once, run the device LM any number of times, download */
int visfs_ba_upload(visfs_ba_handle *h, int32_t n, const visfs_ba_problem *problems);
int visfs_ba_run_resident(visfs_ba_handle *h);
int visfs_ba_download(visfs_ba_handle *h, int32_t n, visfs_ba_result *results);
int visfs_ba_get_timing(const visfs_ba_handle *h, visfs_ba_timing *out);

/* ---------------------------------------------------------------------------------------------------------------
 * Resident local map.  The reference re-marshals a window that changed by ONE frame on every
 * call: LocalMap keeps the signatures and features (corelib/src/LocalMap.cpp), Estimator::process turns them into fresh
 * std::maps and hands them to localOptimize (corelib/src/Estimator.cpp:216-254), then feeds the result back
 * (Estimator.cpp:391-395 -> LocalMap::updateLocalMap).  A visfs_ba_window keeps the same state in HBM and takes the same
 * DELTAS; a solve moves one small table to the device and the poses / outlier list back.
 *
 *   visfs_ba_window_set_points          a feature enters the map or changes (LocalMap::insertSignature "add new feature",
 *                                       LocalMap.cpp:61-80; Feature::setFeaturePose / STABLE, :84-88, :170-183)
 *   visfs_ba_window_insert_frame        LocalMap::insertSignature (LocalMap.cpp:48-131): the signature's pose as T_cw and
 *                                       its observations, as built at Optimizer.cpp:184-195 (u, v, u_right as floats)
 *   visfs_ba_window_remove_frame        LocalMap::removeSignature (LocalMap.cpp:133-168)
 *   visfs_ba_window_remove_points       features_.erase (LocalMap.cpp:158-160)
 *   visfs_ba_window_remove_observations LocalMap::updateLocalMap's outlier handling (LocalMap.cpp:203-226)
 *   visfs_ba_window_set_poses           Signature::setPose (LocalMap.cpp:170-176; Estimator.cpp:393 overrides the newest)
 *   visfs_ba_window_solve               LocalMap::getSignaturePoses + getFeaturePosesAndObservations (LocalMap.cpp:228-236,
 *                                       274-294: features observed more than once, fixed iff STABLE) + localOptimize
 *                                       (Optimizer.cpp:58-364) + the write-back of Optimizer.cpp:320-358 into the resident
 *                                       state (points move only when displaced by less than 5 m)
 * Ids are the reference's signature / feature ids.  One observation per (feature, frame).  Results are identical to
 * visfs_ba_solve on the window the reference would have built (tests/test_gpu_window.py). */
typedef struct visfs_ba_window visfs_ba_window;

typedef struct visfs_ba_window_config {
    int32_t max_frames;          /* frames resident at once (LocalMap/MapSize + 1), <= 32                   */
    int32_t max_points;          /* features resident at once                                                */
    int32_t max_observations;    /* capacity of the observation pool (dead entries are compacted when full)  */
    int32_t reserved0;
    double fx, fy, cx, cy, bf;   /* as visfs_ba_problem                                                      */
    double pixel_variance, huber_delta;
    int32_t iterations, solver, trust_region, reserved1;
} visfs_ba_window_config;

typedef struct visfs_ba_window_result {
    /* caller-allocated outputs, any may be NULL */
    int64_t *frame_id;           /* [max_frames] ids of the frames of the window, ascending                  */
    double  *pose_tq;            /* [max_frames][7] optimised T_cw in the same order                         */
    int64_t *outlier_point_id;   /* [outlier_capacity] culled observations (Optimizer.cpp:283-297), sorted   */
    int64_t *outlier_frame_id;   /* [outlier_capacity]   by (point, frame)                                   */
    int32_t outlier_capacity;
    /* filled by the library */
    int32_t n_frames, n_points, n_edges;   /* the window that was solved                                     */
    int32_t n_outliers;                    /* may exceed outlier_capacity (lists truncated)                  */
    int32_t status;
    int32_t iterations_run[2], trials_run[2], stop_reason[2];
    double chi2_initial, chi2_pass1, chi2_final;
    int64_t h2d_bytes, d2h_bytes;          /* moved by this solve                                            */
} visfs_ba_window_result;

int  visfs_ba_window_create(visfs_ba_handle *h, const visfs_ba_window_config *cfg, visfs_ba_window **out);
void visfs_ba_window_destroy(visfs_ba_window *w);
int  visfs_ba_window_set_points(visfs_ba_window *w, int32_t n, const int64_t *point_id, const double *xyz, const uint8_t *fixed);
int  visfs_ba_window_insert_frame(visfs_ba_window *w, int64_t frame_id, const double *pose_tq, int32_t n_obs,
                                  const int64_t *point_id, const float *obs_uvr /* [n_obs][3] */, const uint8_t *kind /* or NULL */);
int  visfs_ba_window_remove_frame(visfs_ba_window *w, int64_t frame_id);
int  visfs_ba_window_remove_points(visfs_ba_window *w, int32_t n, const int64_t *point_id);
int  visfs_ba_window_remove_observations(visfs_ba_window *w, int32_t n, const int64_t *point_id, const int64_t *frame_id);
int  visfs_ba_window_set_poses(visfs_ba_window *w, int32_t n, const int64_t *frame_id, const double *pose_tq);
/* Odometry links between frames of the map (LocalMap::getSignatureLinks, LocalMap.cpp:238-272 -> Optimizer.cpp:116-150): REPLACES
 * the link set.  link_tq as visfs_ba_problem::link_tq (T_c1c2 as t, q); a link whose frames are not both in the window at the
 * time of a solve takes no part in it, so links may be declared before their frames arrive and need not be withdrawn when a
 * frame leaves.  odometry_variance as visfs_ba_problem::odometry_variance (> 0 when n > 0). */
int  visfs_ba_window_set_links(visfs_ba_window *w, int32_t n, const int64_t *from_frame_id, const int64_t *to_frame_id,
                               const double *link_tq /* [n][7] */, double odometry_variance);
/* root_frame_id: the fixed pose (Estimator.cpp:252: newest id - 1); a value that is not in the window fixes none */
int  visfs_ba_window_solve(visfs_ba_window *w, int64_t root_frame_id, visfs_ba_window_result *result);
int  visfs_ba_window_get_points(visfs_ba_window *w, int32_t n, const int64_t *point_id, double *xyz_out);
int64_t visfs_ba_window_h2d_bytes_total(const visfs_ba_window *w);   /* every byte the window ever sent to the device */

/* multi-GPU global BA: one handle per rank; id from rank 0 is broadcast by the caller */
#define VISFS_BA_COMM_ID_BYTES 128
int visfs_ba_comm_unique_id(void *id_out /* VISFS_BA_COMM_ID_BYTES */);
int visfs_ba_comm_init(visfs_ba_handle *h, int32_t n_ranks, int32_t rank, const void *id);
int visfs_ba_comm_destroy(visfs_ba_handle *h);

/* Page-locked host memory.  Arrays of a visfs_ba_problem / visfs_ba_result that live in page-locked memory (allocated
 * here, or by cudaHostAlloc / cudaHostRegister / torch pin_memory) are moved by DMA straight between the caller's arrays
 * and the device; pageable arrays go through the library's own page-locked staging buffer (one more host copy).  Either
 * way the caller's arrays must not change while a call that reads them is running.  visfs_ba_host_alloc returns NULL on
 * failure; the memory is portable across the devices of the process. */
void *visfs_ba_host_alloc(size_t bytes);
void  visfs_ba_host_free(void *p);

/* FP64 FMA peak probe used by bench.py for the second roofline (returns TFLOP/s) */
int visfs_ba_probe_fp64(visfs_ba_handle *h, double *tflops_out);

#ifdef __cplusplus
}
#endif
#endif /* VISFS_BA_H */

// visfs_ba.cu — C ABI of include/visfs_ba.h: batch upload, on-device structure build, the LM launch
// sequence, download.  No CPU fallback: every compute entry point needs a CUDA device.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <functional>
#include <map>
#include <unordered_map>
#include <chrono>
#include <thread>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "../../include/visfs_ba.h"
#include "ba_kernels.cuh"
#include "ba_build_ws.cuh"
#include "ba_build_ds.cuh"
#include "ba_large.cuh"
#include "ba_band.cuh"
#include "ba_dense.cuh"
#include "ba_mf.cuh"
#include "ba_window.cuh"

#include <dlfcn.h>
#include <nccl.h>   // types and enums only: the library is bound at run time with dlopen (no link-time dependency)

using namespace visfs;

namespace {

thread_local std::string g_create_error;

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        const size_t want = bytes + bytes / 2 + 256;   // (x 1.5: a map that grows frame by frame re-allocates every ~5 frames instead of every 2;
                                                       //  cudaFree + cudaMalloc cost milliseconds in a process with many streams and page-locked buffers)
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    template <typename T> T *as() const { return static_cast<T *>(p); }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }   // locals of the debug / structure entry points are freed on every early return
};

struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        const size_t want = bytes + bytes / 2 + 256;   // (x 1.5: a map that grows frame by frame re-allocates every ~5 frames instead of every 2;
                                                       //  cudaFree + cudaMalloc cost milliseconds in a process with many streams and page-locked buffers)
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    template <typename T> T *as() const { return static_cast<T *>(p); }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    PinBuf() = default;
    PinBuf(const PinBuf &) = delete;
    PinBuf &operator=(const PinBuf &) = delete;
    ~PinBuf() { release(); }
};

enum { EV_BUILD = 0, EV_SOLVE = 1, EV_UPDATE = 2, EV_OTHER = 3, EV_CLASSES = 4 };

}  // namespace

struct visfs_ba_handle {
    int device = 0;
    bool profile = false;
    cudaStream_t stream = nullptr;
    std::string error;
    int sm_count = 148;

    // uploaded batch (host mirrors)
    int n_win = 0, n_chunks = 0, tot_pose = 0, tot_point = 0, tot_edge = 0, max_pose = 0, max_iter = 0, tot_link = 0;
    DevBuf d_link_win, d_link_from, d_link_to, d_link_m, d_link_lin;
    bool resident = false, has_run = false, sorted = true, use_ws = false;
    int max_clusters8 = 0;            // resident clusters of 8 k_build_ws CTAs (cudaOccupancyMaxActiveClusters)
    bool use_ds = false;              // k_build_ds (Schur products on the FP64 tensor pipe): windows of <= 10 poses; opt-in
                                      // (VISFS_BA_USE_DS=1) -- measured 19 % slower per C3 step than k_build_ws, DESIGN.md §5
    int cluster = 1;
    std::vector<WinDesc> win;
    std::vector<Chunk> chunks;
    std::vector<LMState> st_host;
    size_t solve_smem = 0;
    int grid_lm_x = 1, grid_edge_x = 1, reduce_grid = 1;

    // device memory
    DevBuf d_st, d_pose, d_point, d_pose_flags, d_lm_flags, d_pose_hidx, d_pose_active, d_point_hidx,
        d_lm_edge_off, d_obs_u, d_obs_v, d_obs_r, d_edge_pose, d_edge_point, d_edge_orig, d_covis, d_part, d_part2, d_xp,
        d_n_running, d_ctl_count, d_lm_sum, d_chunk_regular, d_tiles, d_tile_off, d_tile_cnt, d_wtiles, d_wtile_off;
    // the caller's arrays, window and chunk descriptors: ONE device buffer with the layout of the pinned staging buffer,
    // filled by ONE H2D copy (a single-window call is latency-bound: ten small copies cost ~50 us)
    DevBuf d_in;
    struct InPtrs {
        double *pose = nullptr, *point = nullptr, *obs = nullptr;
        int *epose = nullptr, *epoint = nullptr;
        uint8_t *pfix = nullptr, *lfix = nullptr, *ekind = nullptr;
        WinDesc *win = nullptr;
        Chunk *chunks = nullptr;
    } in;
    DevBuf d_out_pose, d_out_point, d_out_level, d_tmp, d_tmp2, d_keys, d_keys2, d_perm;
    PinBuf h_stage, h_out, h_small;

    // timing
    std::vector<cudaEvent_t> ev_pool;
    std::vector<std::pair<int, int>> ev_used[EV_CLASSES];  // (start, stop) indices
    size_t ev_next = 0;
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
    visfs_ba_timing timing{};
    int64_t launches = 0, h2d_bytes = 0, d2h_bytes = 0;
    int idx_base = 0;             // sub-handle of a pipelined batch: index of its first window in the caller's batch (messages)
    bool direct_groups = false;   // sub-handle of a pipelined batch whose groups DMA page-locked caller arrays directly
    int direct_h2d = 0;   // pieces of the last upload DMA'd straight from the caller's page-locked arrays

    Batch batch{};

    // visfs_ba_solve_batch pipeline: the batch is cut into groups, every group has its own sub-handle (stream,
    // device buffers, pinned staging) and host thread, so packing, H2D, the LM kernels and D2H of different groups overlap
    std::vector<visfs_ba_handle *> subs;
    bool is_sub = false;

    // large-window path (ba_large.cuh) and the partitioned global BA
    bool large = false, partitioned = false;
    int grid_build_l = 1, grid_update_l = 1;
    long long n_sky = 0;
    int max_front = 0, coop_grid = 0, pcg_grid = 0;
    int max_deg = -1;                 // largest landmark degree of the uploaded batch (-1: unknown, unsorted edge list)
    bool use_run = false;             // k_build_large_run: landmarks walked in the order of their first pose
    DevBuf d_lm_key, d_lm_key2, d_lm_idx, d_lm_order, d_sort_tmp, d_lm_rec;
    // band chunks (ba_band.cuh): the large build through ws::k_build_band + k_band_gather, no atomics
    bool use_band = false, band_rest = false;   // band_rest: some landmarks are outside the band chunks (k_build_large_run takes them)
    int band_n_chunk = 0, band_n_seg = 0;
    DevBuf d_bd_part2, d_bd_scan, d_bd_edge, d_bd_obs, d_bd_chunk, d_bd_tiles, d_bd_ent, d_bd_val, d_bd_part, d_bd_rest;
    int band_n_rest = 0;
    ws::Band band{};
    const int *band_key = nullptr, *band_seg = nullptr;
    const unsigned long long *band_val = nullptr;
    DevBuf d_pcg;
    bool use_front = false;
    bool use_dense = false;           // dense DMMA Cholesky (ba_dense.cuh): wide fronts whose envelope is mostly full
    int dense_grid = 0;               // CTAs of the cooperative k_dense_chol launch (0: cooperative launches unavailable)
    size_t dense_smem = 0;
    DevBuf d_dense, d_dense_prof;
    dn::DenseMat dense{};
    int st_F_hint = 0;                // free poses of the current pass (host copy)
    bool use_mf = false;              // multifrontal nested-dissection Cholesky (ba_mf.cuh): long banded systems
    DevBuf d_mf_meta, d_mf_fronts, d_mf_touch;
    mf::Plan mf_plan{};
    std::vector<int> mf_level_off;    // problems of level l: [mf_level_off[l], mf_level_off[l + 1])
    DevBuf d_plan;
    lg::FrontPlan front_plan{};
    DevBuf d_sky_first, d_sky_off, d_col_ptr, d_col_cnt, d_col_rows, d_red, d_hdiag, d_scal, d_info, d_cnt;
    Batch batch_ctl{};      // same as `batch`, with part2 pointing at the folded (and all-reduced) trial sums
    ncclComm_t comm = nullptr;
    int comm_ranks = 1, comm_rank = 0;
    int64_t allreduce_bytes = 0, allreduce_calls = 0;

    int fail(int status, const std::string &msg) { error = msg; return status; }
    int cuda_fail(cudaError_t e, const char *what) {
        error = std::string(what) + ": " + cudaGetErrorString(e);
        return VISFS_BA_ERR_CUDA;
    }
};

#define CK(call)                                                     \
    do {                                                             \
        cudaError_t e__ = (call);                                    \
        if (e__ != cudaSuccess) return h->cuda_fail(e__, #call);     \
    } while (0)

namespace {

// VISFS_BA_LAUNCH_DEBUG: name the launch that failed (the error would otherwise surface at the end of the pass)
#define LAUNCH_CHECK(what)                                                                                     \
    do {                                                                                                       \
        static const bool dbg__ = getenv("VISFS_BA_LAUNCH_DEBUG") != nullptr;                                  \
        if (dbg__) {                                                                                           \
            cudaError_t e__ = cudaGetLastError();                                                              \
            if (e__ == cudaSuccess) e__ = cudaStreamSynchronize(h->stream);                                    \
            if (e__ != cudaSuccess) return h->cuda_fail(e__, what);                                            \
        }                                                                                                      \
    } while (0)

int ev_begin(visfs_ba_handle *h, int cls) {
    if (!h->profile) return -1;
    if (h->ev_next + 2 > h->ev_pool.size()) {
        const size_t old = h->ev_pool.size();
        h->ev_pool.resize(old + 256);
        for (size_t i = old; i < h->ev_pool.size(); ++i) cudaEventCreate(&h->ev_pool[i]);
    }
    const int a = (int)h->ev_next++;
    const int b = (int)h->ev_next++;
    cudaEventRecord(h->ev_pool[a], h->stream);
    h->ev_used[cls].push_back({a, b});
    return b;
}
void ev_end(visfs_ba_handle *h, int b) {
    if (b >= 0) cudaEventRecord(h->ev_pool[b], h->stream);
}

// page-locked host memory (cudaHostAlloc, cudaHostRegister, visfs_ba_host_alloc): DMA source / target without staging
bool host_is_pinned(const void *p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

int validate(visfs_ba_handle *h, const visfs_ba_problem &p, int idx, bool *sorted, int *max_degree) {
    char buf[160];
    auto bad = [&](const char *m) {
        snprintf(buf, sizeof buf, "problem %d: %s", h->idx_base + idx, m);
        return h->fail(VISFS_BA_ERR_INVALID, buf);
    };
    if (p.n_poses < 0 || p.n_points < 0 || p.n_edges < 0) return bad("negative size");
    if (p.n_poses > kPoseMask) return bad("too many poses");
    if ((p.n_poses && !p.pose_tq) || (p.n_points && !p.point_xyz)) return bad("null pose_tq / point_xyz");
    if (p.n_edges && ((!p.edge_obs && !p.edge_obs_f32) || !p.edge_pose || !p.edge_point)) return bad("null edge arrays");
    if (!(p.pixel_variance > 0.0)) return bad("pixel_variance must be > 0");
    if (p.n_links < 0 || (p.n_links > 0 && (!p.link_from || !p.link_to || !p.link_tq))) return bad("bad odometry link arrays");
    if (p.n_links > 0 && !(p.odometry_variance > 0.0)) return bad("odometry_variance must be > 0");
    for (int k = 0; k < p.n_links; ++k)
        if (p.link_from[k] < 0 || p.link_from[k] >= p.n_poses || p.link_to[k] < 0 || p.link_to[k] >= p.n_poses || p.link_from[k] == p.link_to[k])
            return bad("odometry link index out of range (or from == to)");
    if (p.pose_id)
        for (int i = 1; i < p.n_poses; ++i) if (p.pose_id[i] <= p.pose_id[i - 1]) return bad("pose_id not strictly ascending");
    if (p.point_id)
        for (int i = 1; i < p.n_points; ++i) if (p.point_id[i] <= p.point_id[i - 1]) return bad("point_id not strictly ascending");
    // three branch-light sweeps (the first two vectorise): bounds, order, longest run of one landmark
    const int *ep = p.edge_pose, *el = p.edge_point;
    const unsigned nP = (unsigned)p.n_poses, nL = (unsigned)p.n_points;
    unsigned oob = 0;
    for (int e = 0; e < p.n_edges; ++e) oob |= (unsigned)((unsigned)ep[e] >= nP) | (unsigned)((unsigned)el[e] >= nL);
    if (oob) return bad("edge index out of range");
    unsigned unsorted = 0, dup = 0;
    for (int e = 1; e < p.n_edges; ++e) {
        unsorted |= (unsigned)(el[e] < el[e - 1]) | ((unsigned)(el[e] == el[e - 1]) & (unsigned)(ep[e] < ep[e - 1]));
        dup |= (unsigned)(el[e] == el[e - 1]) & (unsigned)(ep[e] == ep[e - 1]);
    }
    // one observation per (point, pose): the reference's std::map<pid, FeatureBA> per feature cannot hold two
    // (Optimizer.h:52), and the kernels give every pose of a landmark its own accumulator
    if (dup) return bad("duplicate (point, pose) edge");
    const bool srt = unsorted == 0;
    int maxrun = p.n_edges > 0 ? 1 : 0;
    if (srt) {
        int start = 0;
        for (int e = 1; e < p.n_edges; ++e)
            if (el[e] != el[e - 1]) { maxrun = std::max(maxrun, e - start); start = e; }
        maxrun = std::max(maxrun, p.n_edges - start);
    }
    *sorted = srt;
    *max_degree = srt ? maxrun : -1;
    return VISFS_BA_OK;
}

Batch make_batch(visfs_ba_handle *h) {
    Batch b{};
    b.n_win = h->n_win; b.n_chunks = h->n_chunks;
    b.tot_pose = h->tot_pose; b.tot_point = h->tot_point; b.tot_edge = h->tot_edge;
    b.win = h->in.win; b.st = h->d_st.as<LMState>(); b.chunks = h->in.chunks;
    b.pose = h->d_pose.as<double>(); b.point = h->d_point.as<double>();
    b.pose_flags = h->d_pose_flags.as<uint8_t>(); b.lm_flags = h->d_lm_flags.as<uint8_t>();
    b.pose_hidx = h->d_pose_hidx.as<int>(); b.pose_active = h->d_pose_active.as<int>();
    b.point_hidx = h->d_point_hidx.as<int>(); b.lm_edge_off = h->d_lm_edge_off.as<int>();
    b.obs_u = h->d_obs_u.as<double>(); b.obs_v = h->d_obs_v.as<double>(); b.obs_r = h->d_obs_r.as<double>();
    b.edge_pose = h->d_edge_pose.as<int>(); b.edge_point = h->d_edge_point.as<int>();
    b.edge_orig = h->sorted ? nullptr : h->d_edge_orig.as<int>();
    b.covis = h->d_covis.as<unsigned>(); b.part = h->d_part.as<double>(); b.part2 = h->d_part2.as<double>();
    b.lm_sum = (h->use_ws && !h->use_ds && !getenv("VISFS_BA_NO_LMSUM")) ? h->d_lm_sum.as<double>() : nullptr;
    b.xp = h->d_xp.as<double>(); b.n_running = h->d_n_running.as<int>(); b.ctl_count = h->d_ctl_count.as<int>();
    b.dbg = nullptr; b.dbg_lambda = -1.0;
    b.tot_link = h->tot_link;
    b.link_win = h->d_link_win.as<int>(); b.link_from = h->d_link_from.as<int>(); b.link_to = h->d_link_to.as<int>();
    b.link_m = h->d_link_m.as<double>(); b.link_lin = h->d_link_lin.as<double>();
    b.tiles = h->d_tiles.as<Tile>(); b.chunk_tile_off = h->d_tile_off.as<int>();
    b.chunk_regular = (h->use_ws && !h->use_ds && !getenv("VISFS_BA_NO_REGULAR")) ? h->d_chunk_regular.as<int>() : nullptr;
    b.wtiles = h->d_wtiles.as<Tile>(); b.chunk_wtile_off = h->d_wtile_off.as<int>();
    b.sky_first = h->d_sky_first.as<int>(); b.sky_off = h->d_sky_off.as<long long>();
    b.col_ptr = h->d_col_ptr.as<int>(); b.col_rows = h->d_col_rows.as<int>();
    b.red = h->d_red.as<double>(); b.red_g_off = 0; b.red_bp_off = 0; b.hdiag = h->d_hdiag.as<double>();
    return b;
}

dim3 grid2(int items, int n_win) {
    int gx = (items + 255) / 256;
    gx = std::max(1, std::min(gx, 1024));
    return dim3((unsigned)gx, (unsigned)n_win);
}

// inputs that are already on the device (resident local map, visfs_ba_window_solve): sizes come with the problem structs,
// the arrays are written into the staging layout by `emit` (kernels on h->stream) instead of being packed and copied
struct DevInput {
    int max_degree = 0, n_fixed = 0;
    std::function<int(visfs_ba_handle *)> emit;
};

// ---- upload: validate, pack into pinned staging, H2D, device-side preparation ---------------------
int upload(visfs_ba_handle *h, int n, const visfs_ba_problem *probs, const DevInput *dev = nullptr) {
    h->resident = false; h->has_run = false;
    if (n <= 0 || !probs) return h->fail(VISFS_BA_ERR_INVALID, "empty batch");
    if (n > 65535) return h->fail(VISFS_BA_ERR_INVALID, "at most 65535 windows per batch");
    CK(cudaSetDevice(h->device));
    h->win.assign(n, WinDesc{});
    long long tp = 0, tl = 0, te = 0, tk = 0;
    int max_pose = 0, max_point = 0, max_edge = 0, max_iter = 0, max_free = 0;
    bool all_sorted = true, any_large = false, any_part = false;
    int max_deg_seen = 0;
    for (int w = 0; w < n; ++w) {
        const visfs_ba_problem &p = probs[w];
        bool srt = true; int deg = dev ? dev->max_degree : 0;
        if (!dev) {
            const int st = validate(h, p, w, &srt, &deg);
            if (st != VISFS_BA_OK) return st;
        }
        const bool part = (p.flags & VISFS_BA_FLAG_PARTITIONED) != 0;
        const bool big = part || p.n_poses > kMaxSmallPoses || getenv("VISFS_BA_FORCE_LARGE");
        if (big && n != 1)
            return h->fail(VISFS_BA_ERR_UNSUPPORTED, "windows with more than 32 poses and partitioned problems are solved one per call");
        any_large = any_large || big;
        any_part = any_part || part;
        // (the degree limit of the large path is checked on the device so that all ranks of a partitioned run agree)
        if (!big && deg > kTileEdges) return h->fail(VISFS_BA_ERR_UNSUPPORTED, "landmark observed by more poses than one tile holds (kTileEdges)");
        all_sorted = all_sorted && srt;
        max_deg_seen = (deg < 0 || max_deg_seen < 0) ? -1 : std::max(max_deg_seen, deg);
        WinDesc &d = h->win[w];
        d.pose_off = (int)tp; d.n_pose = p.n_poses; d.point_off = (int)tl; d.n_point = p.n_points;
        d.edge_off = (int)te; d.n_edge = p.n_edges;
        const bool single = (p.flags & VISFS_BA_FLAG_SINGLE_PASS) != 0;
        d.max_iter = single ? p.iterations : p.iterations / 2;
        if (d.max_iter < 0) d.max_iter = 0;
        d.solver = p.solver; d.trust = p.trust_region; d.flags = p.flags; d.large = big ? 1 : 0;
        d.fx = p.fx; d.fy = p.fy; d.cx = p.cx; d.cy = p.cy; d.bf = p.bf;
        d.inv_pv = 1.0 / p.pixel_variance; d.delta = p.huber_delta;
        d.link_off = (int)tk; d.n_link = p.n_links; d.inv_ov = p.n_links > 0 ? 1.0 / p.odometry_variance : 0.0;
        tk += p.n_links;
        tp += p.n_poses; tl += p.n_points; te += p.n_edges;
        {
            int nfix = dev ? dev->n_fixed : 0;
            if (p.pose_fixed) for (int i = 0; i < p.n_poses; ++i) nfix += p.pose_fixed[i] != 0;
            max_free = std::max(max_free, p.n_poses - nfix);   // upper bound of the free poses F of any pass
        }
        max_pose = std::max(max_pose, p.n_poses); max_point = std::max(max_point, p.n_points);
        max_edge = std::max(max_edge, p.n_edges); max_iter = std::max(max_iter, d.max_iter);
        if (tp > 0x3fffffff || tl > 0x3fffffff || te > 0x3fffffff) return h->fail(VISFS_BA_ERR_INVALID, "batch too large");
    }
    h->n_win = n; h->tot_pose = (int)tp; h->tot_point = (int)tl; h->tot_edge = (int)te; h->tot_link = (int)tk;
    h->max_pose = max_pose; h->max_iter = max_iter; h->sorted = all_sorted;
    h->large = any_large; h->partitioned = any_part;
    h->max_deg = max_deg_seen;
    if (any_part && h->comm_ranks > 1 && !h->comm) return h->fail(VISFS_BA_ERR_INVALID, "partitioned problem without visfs_ba_comm_init");
    if (any_large && probs[0].solver == VISFS_BA_SOLVER_PCG && h->pcg_grid < 1)
        return h->fail(VISFS_BA_ERR_UNSUPPORTED, "Optimizer/Solver=2 (PCG) on a large window needs cooperative kernel launches");
    h->grid_lm_x = std::max(1, std::min((max_point + 255) / 256, 1024));
    h->grid_edge_x = std::max(1, std::min((max_edge + 255) / 256, 1024));

    // ---- work decomposition.  A chunk is a landmark range of one window (one CTA).  The warp-specialised build
    // kernel runs one CTA per SM, so the number of chunks is chosen to fill whole waves of `sm_count` CTAs:
    //   one window : up to sm_count chunks of >= 16 landmarks, clusters of up to 8 CTAs sum their partials
    //   a batch    : c chunks per window with c in 1..16 maximising n*c / (ceil(n*c / sm_count) * sm_count)
    h->use_ws = (max_pose <= ws::kMaxPosesWs) && (max_free <= ws::kMaxFreeWs) && !getenv("VISFS_BA_NO_WS") && !any_large;
    h->use_ds = h->use_ws && max_pose <= ds::kMaxPosesDs && getenv("VISFS_BA_USE_DS") != nullptr;
    const int sms = std::max(h->sm_count, 8);
    int per_window = 1;
    bool wide_clusters = false;
    if (n == 1) {
        int cap_pw = (sms * 7 / 8) / 4 * 4;
        // one wave of clusters of 8 when the device holds >= 12 of them at once (B200: 15 -> 120 chunks; C1 0.754 -> 0.724 ms)
        if (h->use_ws && h->max_clusters8 >= 12 && !getenv("VISFS_BA_CLUSTER") && !getenv("VISFS_BA_NO_CLUSTER")) { cap_pw = h->max_clusters8 * 8; wide_clusters = true; }
        if (const char *e = getenv("VISFS_BA_PERWINDOW")) cap_pw = std::max(1, atoi(e));   // (experiments)
        per_window = std::max(1, std::min((max_point + kTileLm / 2 - 1) / (kTileLm / 2), cap_pw));
        if (const char *e = getenv("VISFS_BA_CHUNKS")) per_window = std::max(1, atoi(e));   // experiments
    } else {
        double best = -1.0;
        for (int c = 1; c <= 16; ++c) {
            if (c > 1 && (max_point + c - 1) / c < kTileLm / 2) break;
            const long long tot = (long long)n * c;
            const double eff = (double)tot / (double)(((tot + sms - 1) / sms) * sms);
            if (eff > best + 0.03) { best = eff; per_window = c; }
        }
    }
    int cl = 1;
    // clusters only for the single-window latency case: 4 CTAs per cluster still fit one wave (33 clusters of 4 are
    // co-resident at 195 KB of shared memory per CTA; clusters of 8 drop that to 15) and cut k_solve's input 4x
    int cl_max = wide_clusters ? 8 : 4;
    if (const char *e = getenv("VISFS_BA_CLUSTER")) cl_max = std::max(1, std::min(atoi(e), 8));
    if (h->use_ws && n == 1 && !getenv("VISFS_BA_NO_CLUSTER")) while (cl < cl_max && cl * 2 <= per_window) cl *= 2;
    h->cluster = cl;
    h->chunks.clear();
    long long part_total = 0;
    for (int w = 0; w < n; ++w) {
        WinDesc &d = h->win[w];
        if (d.large) {   // the large path needs no chunk table: its kernels stride over the landmarks, k_control reads one folded partial
            d.chunk_off = 0; d.n_chunks = 1; d.layout = 2; d.n_parts = 0; d.part_stride = 8; d.part_off = 0;
            continue;
        }
        d.chunk_off = (int)h->chunks.size();
        const int lm_per_chunk = std::max(1, (d.n_point + per_window - 1) / per_window);
        for (int l0 = 0; l0 < d.n_point; l0 += lm_per_chunk)
            h->chunks.push_back(Chunk{w, d.point_off + l0, d.point_off + std::min(d.n_point, l0 + lm_per_chunk)});
        while (((int)h->chunks.size() - d.chunk_off) % cl != 0 || (int)h->chunks.size() == d.chunk_off)
            h->chunks.push_back(Chunk{w, d.point_off + d.n_point, d.point_off + d.n_point});   // empty padding chunk
        d.n_chunks = (int)h->chunks.size() - d.chunk_off;
        d.layout = h->use_ws ? 1 : 0;
        d.n_parts = h->use_ws ? d.n_chunks / cl : d.n_chunks;
        const int Fm = std::min(d.n_pose, kMaxSmallPoses);
        d.part_stride = std::max(std::max(Fm * (Fm - 1) / 2 * 36 + Fm * kHStride, ws::part_len(Fm)), 8);
        d.part_off = part_total;
        part_total += (long long)d.part_stride * std::max(d.n_chunks, 1);
    }
    h->n_chunks = (int)h->chunks.size();
    {
        const int nmax = 6 * std::min(max_pose, kMaxSmallPoses);
        h->solve_smem = sizeof(double) * ((size_t)nmax * (nmax + 1) / 2 + 8 * (size_t)nmax + 36 * (size_t)(nmax / 6) + 32);
    }

    // device buffers
    const size_t P = (size_t)std::max<long long>(tp, 1), L = (size_t)std::max<long long>(tl, 1), E = (size_t)std::max<long long>(te, 1);
    CK(h->d_st.reserve(sizeof(LMState) * n));
    CK(h->d_pose.reserve(sizeof(double) * 2 * P * kPoseStride)); CK(h->d_point.reserve(sizeof(double) * 2 * L * 3));
    CK(h->d_pose_flags.reserve(P)); CK(h->d_lm_flags.reserve(L));
    CK(h->d_pose_hidx.reserve(sizeof(int) * P)); CK(h->d_pose_active.reserve(sizeof(int) * P));
    CK(h->d_point_hidx.reserve(sizeof(int) * L)); CK(h->d_lm_edge_off.reserve(sizeof(int) * (L + 1)));
    CK(h->d_obs_u.reserve(sizeof(double) * E)); CK(h->d_obs_v.reserve(sizeof(double) * E)); CK(h->d_obs_r.reserve(sizeof(double) * E));
    CK(h->d_edge_pose.reserve(sizeof(int) * E)); CK(h->d_edge_point.reserve(sizeof(int) * E));
    CK(h->d_covis.reserve(sizeof(unsigned) * P));
    CK(h->d_part.reserve(sizeof(double) * (size_t)std::max<long long>(part_total, 8)));
    if (any_large) {
        // persistent grids: 2 (build, 77 KB of shared memory per CTA) / 4 (update) CTAs per SM, warps stride over the landmarks
        const int want = (max_point + lg::kWarpsL - 1) / lg::kWarpsL;
        h->grid_build_l = std::max(1, std::min(want, sms * 2));
        h->grid_update_l = std::max(1, std::min(want, sms * 4));
        CK(h->d_sky_first.reserve(sizeof(int) * (P + 1))); CK(h->d_sky_off.reserve(sizeof(long long) * (P + 2)));
        CK(h->d_col_ptr.reserve(sizeof(int) * (P + 2))); CK(h->d_col_cnt.reserve(sizeof(int) * (P + 2)));
        CK(h->d_hdiag.reserve(sizeof(double) * 6 * P)); CK(h->d_scal.reserve(sizeof(double) * 8));
        CK(h->d_info.reserve(sizeof(long long) * 4)); CK(h->d_cnt.reserve(sizeof(int) * 4));
    }
    CK(h->d_part2.reserve(sizeof(double) * 2 * std::max(std::max(h->n_chunks, 1), std::max(h->grid_build_l, h->grid_update_l))));
    if (h->use_ws) CK(h->d_lm_sum.reserve(sizeof(double) * 9 * L));
    if (h->use_ws) CK(h->d_chunk_regular.reserve(sizeof(int) * (size_t)std::max(h->n_chunks, 1)));
    CK(h->d_xp.reserve(sizeof(double) * 6 * P)); CK(h->d_n_running.reserve(sizeof(int) * 4)); CK(h->d_ctl_count.reserve(sizeof(int) * (size_t)n));
    const size_t max_tiles = 2 * E / (ds::kEdges + 1) + L / ds::kLm + 2 * (size_t)h->n_chunks + 8;   // (bound for either tile shape)
    CK(h->d_tiles.reserve(sizeof(Tile) * max_tiles));
    CK(h->d_tile_off.reserve(sizeof(int) * (h->n_chunks + 2))); CK(h->d_tile_cnt.reserve(sizeof(int) * (h->n_chunks + 2)));
    CK(h->d_wtiles.reserve(sizeof(Tile) * (L + (size_t)h->n_chunks + 8))); CK(h->d_wtile_off.reserve(sizeof(int) * (h->n_chunks + 2)));
    CK(h->d_out_level.reserve(E));   // (visfs_ba_structure_build stages the caller's edge levels here)

    if (tk > 0) {   // odometry links: a few dozen per window, staged through pageable vectors
        std::vector<int> lw((size_t)tk), lf((size_t)tk), lt((size_t)tk);
        std::vector<double> lm(7 * (size_t)tk);
        for (int w = 0; w < n; ++w) {
            const visfs_ba_problem &p = probs[w];
            const WinDesc &d = h->win[w];
            for (int k = 0; k < p.n_links; ++k) {
                lw[(size_t)d.link_off + k] = w; lf[(size_t)d.link_off + k] = p.link_from[k]; lt[(size_t)d.link_off + k] = p.link_to[k];
                double *m = &lm[7 * ((size_t)d.link_off + k)];
                memcpy(m, p.link_tq + 7 * (size_t)k, 7 * sizeof(double));
                // g2o::SE3Quat::normalizeRotation, run by the constructor the measurement goes through at Optimizer.cpp:140
                if (m[6] < 0.0) for (int i = 3; i < 7; ++i) m[i] = -m[i];
                const double qn = std::sqrt(m[3] * m[3] + m[4] * m[4] + m[5] * m[5] + m[6] * m[6]);
                for (int i = 3; i < 7; ++i) m[i] /= qn;
            }
        }
        CK(h->d_link_win.reserve(sizeof(int) * (size_t)tk)); CK(h->d_link_from.reserve(sizeof(int) * (size_t)tk));
        CK(h->d_link_to.reserve(sizeof(int) * (size_t)tk)); CK(h->d_link_m.reserve(sizeof(double) * 7 * (size_t)tk));
        CK(h->d_link_lin.reserve(sizeof(double) * kLinkStride * (size_t)tk));
        CK(cudaMemcpyAsync(h->d_link_win.p, lw.data(), sizeof(int) * (size_t)tk, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(h->d_link_from.p, lf.data(), sizeof(int) * (size_t)tk, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(h->d_link_to.p, lt.data(), sizeof(int) * (size_t)tk, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(h->d_link_m.p, lm.data(), sizeof(double) * 7 * (size_t)tk, cudaMemcpyHostToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));   // the host vectors go out of scope
    }

    // pack into one pinned staging buffer (layout: pose | point | obs | epose | epoint | pfix | lfix | ekind)
    auto al16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
    // observations travel as floats when every window offers them that way (edge_obs_f32), else as doubles
    bool obs_f32 = te > 0;
    for (int w = 0; w < n; ++w) if (probs[w].n_edges > 0 && !probs[w].edge_obs_f32 && !dev) obs_f32 = false;
    const size_t obs_elem = obs_f32 ? sizeof(float) : sizeof(double);
    const size_t o_pose = 0, o_point = o_pose + sizeof(double) * 7 * P, o_obs = o_point + sizeof(double) * 3 * L,
                 o_epose = o_obs + obs_elem * 3 * E, o_epoint = o_epose + sizeof(int) * E,
                 o_pfix = o_epoint + sizeof(int) * E, o_lfix = o_pfix + P, o_ekind = o_lfix + L, o_win = al16(o_ekind + E),
                 o_chunks = al16(o_win + sizeof(WinDesc) * n), o_end = al16(o_chunks + sizeof(Chunk) * std::max(h->n_chunks, 1));
    CK(h->h_stage.reserve(o_end));
    CK(h->d_in.reserve(o_end));
    {
        char *db = h->d_in.as<char>();
        h->in.pose = reinterpret_cast<double *>(db + o_pose); h->in.point = reinterpret_cast<double *>(db + o_point);
        h->in.obs = reinterpret_cast<double *>(db + o_obs); h->in.epose = reinterpret_cast<int *>(db + o_epose);
        h->in.epoint = reinterpret_cast<int *>(db + o_epoint); h->in.pfix = reinterpret_cast<uint8_t *>(db + o_pfix);
        h->in.lfix = reinterpret_cast<uint8_t *>(db + o_lfix); h->in.ekind = reinterpret_cast<uint8_t *>(db + o_ekind);
        h->in.win = reinterpret_cast<WinDesc *>(db + o_win); h->in.chunks = reinterpret_cast<Chunk *>(db + o_chunks);
    }
    char *sg = h->h_stage.as<char>();
    cudaStream_t s = h->stream;
    {
        // Every piece (one array of one window) either is DMA'd from the caller's memory (page-locked source, worth its
        // own copy) or is packed into the staging buffer; adjacent staged pieces leave as one copy.
        char *db = h->d_in.as<char>();
        // (the groups of a pipelined batch keep the staging route when the rank has 16 host cores for the packing: one large
        //  DMA transfer per group moves at 45 GB/s, the many array-sized ones of the direct route at 20 GB/s; see pick_groups)
        const bool allow_direct = !getenv("VISFS_BA_NO_DIRECT") && (!h->is_sub || h->direct_groups);
        constexpr size_t kDirectMin = 32 * 1024;
        size_t run_lo = 0, run_hi = 0;
        int direct = 0;
        cudaError_t cerr = cudaSuccess;
        auto flush = [&]() {
            if (run_hi > run_lo && cerr == cudaSuccess) cerr = cudaMemcpyAsync(db + run_lo, sg + run_lo, run_hi - run_lo, cudaMemcpyHostToDevice, s);
            run_lo = run_hi = 0;
        };
        auto stage_at = [&](size_t off, size_t bytes) -> char * {   // reserve [off, off + bytes) of the staging buffer in the current run
            if (off != run_hi || run_hi == run_lo) { flush(); run_lo = off; }
            run_hi = off + bytes;
            return sg + off;
        };
        auto put = [&](size_t off, const void *src, size_t bytes) {
            if (!bytes) return;
            if (allow_direct && bytes >= kDirectMin && host_is_pinned(src)) {
                if (cerr == cudaSuccess) cerr = cudaMemcpyAsync(db + off, src, bytes, cudaMemcpyHostToDevice, s);
                ++direct;
            } else {
                memcpy(stage_at(off, bytes), src, bytes);
            }
        };
        auto put_flags = [&](size_t off, const uint8_t *src, size_t bytes) {   // optional byte arrays: absent = zeros
            if (!bytes) return;
            if (src) memcpy(stage_at(off, bytes), src, bytes); else memset(stage_at(off, bytes), 0, bytes);
        };
        if (!dev) {
        for (int w = 0; w < n; ++w) put(o_pose + sizeof(double) * 7 * h->win[w].pose_off, probs[w].pose_tq, sizeof(double) * 7 * probs[w].n_poses);
        for (int w = 0; w < n; ++w) put(o_point + sizeof(double) * 3 * h->win[w].point_off, probs[w].point_xyz, sizeof(double) * 3 * probs[w].n_points);
        for (int w = 0; w < n; ++w) {
            const visfs_ba_problem &p = probs[w];
            const size_t at = o_obs + obs_elem * 3 * h->win[w].edge_off;
            if (obs_f32) put(at, p.edge_obs_f32, sizeof(float) * 3 * p.n_edges);
            else if (!p.edge_obs_f32) put(at, p.edge_obs, sizeof(double) * 3 * p.n_edges);
            else if (p.n_edges) {   // a float window in a batch that travels as doubles: widened while packing
                double *dst = reinterpret_cast<double *>(stage_at(at, sizeof(double) * 3 * p.n_edges));
                for (size_t i = 0; i < 3 * (size_t)p.n_edges; ++i) dst[i] = (double)p.edge_obs_f32[i];
            }
        }
        for (int w = 0; w < n; ++w) put(o_epose + sizeof(int) * h->win[w].edge_off, probs[w].edge_pose, sizeof(int) * probs[w].n_edges);
        for (int w = 0; w < n; ++w) put(o_epoint + sizeof(int) * h->win[w].edge_off, probs[w].edge_point, sizeof(int) * probs[w].n_edges);
        for (int w = 0; w < n; ++w) put_flags(o_pfix + h->win[w].pose_off, probs[w].pose_fixed, probs[w].n_poses);
        for (int w = 0; w < n; ++w) put_flags(o_lfix + h->win[w].point_off, probs[w].point_fixed, probs[w].n_points);
        for (int w = 0; w < n; ++w) put_flags(o_ekind + h->win[w].edge_off, probs[w].edge_kind, probs[w].n_edges);
        }
        memcpy(stage_at(o_win, sizeof(WinDesc) * n), h->win.data(), sizeof(WinDesc) * n);
        if (h->n_chunks) memcpy(stage_at(o_chunks, sizeof(Chunk) * h->n_chunks), h->chunks.data(), sizeof(Chunk) * h->n_chunks);
        flush();
        CK(cerr);
        h->direct_h2d = direct;
        if (dev) { const int st = dev->emit(h); if (st) return st; }
    }

    h->h2d_bytes = (dev ? (int64_t)(o_end - o_win) : (int64_t)o_end) + (int64_t)tk * (3 * (int64_t)sizeof(int) + 7 * (int64_t)sizeof(double));
    // device-side preparation: (optional) stable sort by (window, point, pose), SoA split, CSR offsets
    const int *perm = nullptr;
    if (!all_sorted && te > 0) {
        if (max_point >= (1 << 24)) return h->fail(VISFS_BA_ERR_UNSUPPORTED, "unsorted edge lists need fewer than 2^24 points per window");
        CK(h->d_keys.reserve(sizeof(unsigned long long) * E)); CK(h->d_keys2.reserve(sizeof(unsigned long long) * E));
        CK(h->d_perm.reserve(sizeof(int) * E)); CK(h->d_edge_orig.reserve(sizeof(int) * E));
        // keys / identity are produced on the host side of the staging copy (cheap, and only on this path)
        std::vector<unsigned long long> keys(E);
        std::vector<int> ident(E);
        for (int w = 0; w < n; ++w) {
            const visfs_ba_problem &p = probs[w];
            const WinDesc &d = h->win[w];
            for (int e = 0; e < p.n_edges; ++e) {
                keys[d.edge_off + e] = ((unsigned long long)w << 48) | ((unsigned long long)p.edge_point[e] << 24) | (unsigned long long)p.edge_pose[e];
                ident[d.edge_off + e] = e;
            }
        }
        CK(cudaMemcpyAsync(h->d_keys.p, keys.data(), sizeof(unsigned long long) * E, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(h->d_perm.p, ident.data(), sizeof(int) * E, cudaMemcpyHostToDevice, s));
        size_t tmp_bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, h->d_keys.as<unsigned long long>(), h->d_keys2.as<unsigned long long>(),
                                        h->d_perm.as<int>(), h->d_edge_orig.as<int>(), (int)te, 0, 64, s);
        CK(h->d_tmp.reserve(tmp_bytes));
        CK(cub::DeviceRadixSort::SortPairs(h->d_tmp.p, tmp_bytes, h->d_keys.as<unsigned long long>(), h->d_keys2.as<unsigned long long>(),
                                           h->d_perm.as<int>(), h->d_edge_orig.as<int>(), (int)te, 0, 64, s));
        // duplicates of an unsorted list are adjacent after the sort
        CK(cudaMemsetAsync(h->d_n_running.as<int>() + 1, 0, sizeof(int), s));
        k_dup_keys<<<std::max(1, std::min(((int)te + 255) / 256, 1024)), 256, 0, s>>>(h->d_keys2.as<unsigned long long>(), (int)te, h->d_n_running.as<int>() + 1);
        int n_dup = 0;
        CK(cudaMemcpyAsync(&n_dup, h->d_n_running.as<int>() + 1, sizeof(int), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));  // host vectors above go out of scope
        if (n_dup) return h->fail(VISFS_BA_ERR_INVALID, "duplicate (point, pose) edge in an unsorted edge list");
        perm = h->d_edge_orig.as<int>();
    }
    h->batch = make_batch(h);
    {   // few windows with many partial systems each: fold them with a wide kernel before the one-CTA k_solve
        int max_parts = 0, max_stride = 0;
        long long volume = 0;   // doubles one k_solve CTA would have to pull through its SM
        for (const WinDesc &d : h->win) {
            if (d.large) continue;
            max_parts = std::max(max_parts, d.n_parts); max_stride = std::max(max_stride, d.part_stride);
            volume = std::max(volume, (long long)d.n_parts * d.part_stride);
        }
        // worth an extra launch (~3 us) from about 0.8 MB per window (C2: 1.9 MB, C1: 0.5 MB)
        h->batch.parts_reduced = (n <= 8 && max_parts >= 4 && volume >= (getenv("VISFS_BA_REDUCE_MIN") ? atoll(getenv("VISFS_BA_REDUCE_MIN")) : 100000) && !getenv("VISFS_BA_NO_REDUCE")) ? 1 : 0;
        h->reduce_grid = std::max(1, std::min((max_stride + 255) / 256, 64));
    }
    h->batch_ctl = h->batch;
    h->batch_ctl.part2 = h->d_scal.as<double>() + 2;
    Batch &B = h->batch;
    if (te > 0) {
        k_prepare_edges<<<grid2(max_edge, n), 256, 0, s>>>(B, obs_f32 ? nullptr : h->in.obs, obs_f32 ? reinterpret_cast<const float *>(h->in.obs) : nullptr, h->in.epose, h->in.epoint,
                                                          h->in.ekind, perm, h->d_obs_u.as<double>(),
                                                          h->d_obs_v.as<double>(), h->d_obs_r.as<double>(), h->d_edge_point.as<int>());
    }
    LAUNCH_CHECK("k_prepare_edges");
    k_lm_offsets<<<grid2(max_point + 1, n), 256, 0, s>>>(B, h->d_lm_edge_off.as<int>());
    LAUNCH_CHECK("k_lm_offsets");
    {   // tile table: count per chunk, exclusive scan, fill (all on the stream, no host round trip)
        const int nc1 = h->n_chunks + 1;
        const int tile_lm = h->use_ds ? ds::kLm : kTileLm, tile_edges = h->use_ds ? ds::kEdges : kTileEdges;
        k_count_tiles<<<(nc1 + 127) / 128, 128, 0, s>>>(B, h->d_tile_cnt.as<int>(), tile_lm, tile_edges);
        size_t tmp_bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, h->d_tile_cnt.as<int>(), h->d_tile_off.as<int>(), nc1, s);
        CK(h->d_tmp2.reserve(tmp_bytes));
        CK(cub::DeviceScan::ExclusiveSum(h->d_tmp2.p, tmp_bytes, h->d_tile_cnt.as<int>(), h->d_tile_off.as<int>(), nc1, s));
        if (h->n_chunks) k_fill_tiles<<<(h->n_chunks + 127) / 128, 128, 0, s>>>(B, h->d_tile_off.as<int>(), h->d_tiles.as<Tile>(), tile_lm, tile_edges);
        if (h->n_chunks && B.chunk_regular) k_chunk_regular<<<h->n_chunks, 128, 0, s>>>(B, h->d_chunk_regular.as<int>());
        // warp tiles of k_update (d_tile_cnt is reused: the scan above has consumed it)
        k_count_wtiles<<<(nc1 + 127) / 128, 128, 0, s>>>(B, h->d_tile_cnt.as<int>());
        CK(cub::DeviceScan::ExclusiveSum(h->d_tmp2.p, tmp_bytes, h->d_tile_cnt.as<int>(), h->d_wtile_off.as<int>(), nc1, s));
        if (h->n_chunks) k_fill_wtiles<<<(h->n_chunks + 127) / 128, 128, 0, s>>>(B, h->d_wtile_off.as<int>(), h->d_wtiles.as<Tile>());
        LAUNCH_CHECK("tile tables");
    }
    CK(cudaGetLastError());
    h->resident = true;
    return VISFS_BA_OK;
}

int reset_state(visfs_ba_handle *h) {
    Batch &B = h->batch;
    const int items = std::max(std::max(h->tot_pose, h->tot_point * 3), std::max(h->tot_edge, h->n_win));
    const int gx = std::max(1, std::min((items + 255) / 256, 4096));
    k_reset<<<gx, 256, 0, h->stream>>>(B, h->in.pose, h->in.point, h->in.pfix, h->in.lfix);
    LAUNCH_CHECK("k_reset");
    CK(cudaMemsetAsync(h->d_n_running.p, 0, sizeof(int) * 4, h->stream));
    CK(cudaMemsetAsync(h->d_ctl_count.p, 0, sizeof(int) * (size_t)std::max(h->n_win, 1), h->stream));
    CK(cudaGetLastError());
    h->launches += 1;
    return VISFS_BA_OK;
}

int run_structure(visfs_ba_handle *h, bool want_covis = false) {
    Batch &B = h->batch;
    cudaStream_t s = h->stream;
    CK(cudaMemsetAsync(h->d_pose_active.p, 0, sizeof(int) * std::max(h->tot_pose, 1), s));
    const dim3 glm((unsigned)h->grid_lm_x, (unsigned)h->n_win);
    const int gw = (h->n_win + 127) / 128;
    k_struct_lm<<<glm, 256, 0, s>>>(B);
    k_struct_pose<<<gw, 128, 0, s>>>(B);
    k_struct_count<<<glm, 256, 0, s>>>(B, want_covis ? 1 : 0);
    k_struct_finish<<<gw, 128, 0, s>>>(B);
    LAUNCH_CHECK("k_struct_*");
    CK(cudaGetLastError());
    h->launches += 4;
    return VISFS_BA_OK;
}

template <int MODE>
int launch_build(visfs_ba_handle *h) {
    if (h->n_chunks == 0) return VISFS_BA_OK;
    h->launches += 1;
    if (MODE == MODE_BUILD && h->use_ws) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)h->n_chunks);
        cfg.blockDim = dim3(h->use_ds ? ds::kThreadsDs : ws::kThreadsWs);
        cfg.dynamicSmemBytes = h->use_ds ? sizeof(ds::Smem) : sizeof(ws::Smem);
        cfg.stream = h->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)h->cluster;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if (h->use_ds) CK(cudaLaunchKernelEx(&cfg, ds::k_build_ds, h->batch, h->cluster));
        else CK(cudaLaunchKernelEx(&cfg, ws::k_build_ws, h->batch, h->cluster));
        return VISFS_BA_OK;
    }
    const size_t smem = sizeof(BuildSmemT<MODE>);
    if (h->max_pose <= 23) k_build<MODE, 1><<<h->n_chunks, kThreads, smem, h->stream>>>(h->batch);
    else k_build<MODE, 2><<<h->n_chunks, kThreads, smem, h->stream>>>(h->batch);
    return VISFS_BA_OK;
}

int enqueue_body(visfs_ba_handle *h) {
    int ev = ev_begin(h, EV_BUILD);
    { const int st = launch_build<MODE_BUILD>(h); if (st) return st; }
    LAUNCH_CHECK(h->use_ds ? "k_build_ds" : (h->use_ws ? "k_build_ws" : "k_build"));
    ev_end(h, ev);
    ev = ev_begin(h, EV_SOLVE);
    if (h->tot_link > 0) { k_link_lin<<<(h->tot_link + 63) / 64, 64, 0, h->stream>>>(h->batch); h->launches += 1; }
    if (h->batch.parts_reduced) {
        k_reduce_parts<<<dim3((unsigned)h->reduce_grid, (unsigned)h->n_win), 256, 0, h->stream>>>(h->batch);
        h->launches += 1;
    }
    if (h->n_win >= 2 * h->sm_count) k_solve2<<<h->n_win, kSolveThreads, h->solve_smem, h->stream>>>(h->batch);
    else k_solve<<<h->n_win, kSolveThreads, h->solve_smem, h->stream>>>(h->batch);
    LAUNCH_CHECK("k_solve");
    ev_end(h, ev);
    ev = ev_begin(h, EV_UPDATE);
    if (h->n_chunks) k_update<<<h->n_chunks, kUpdThreads, sizeof(UpdateSmem), h->stream>>>(h->batch);   // (its last CTA per window runs the LM controller)
    else k_control<<<h->n_win, 32, 0, h->stream>>>(h->batch);
    LAUNCH_CHECK("k_update");
    ev_end(h, ev);
    h->launches += 3;
    return VISFS_BA_OK;
}


// ---- NCCL, bound at run time (a single-GPU user never needs the library) ---------------------------------------
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
    bool load() {
        if (lib) return true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);   // resolves to the copy the process already holds (e.g. torch's), else the system one
            if (lib) break;
        }
        if (!lib) { error = std::string("dlopen libnccl.so.2: ") + dlerror(); return false; }
        GetUniqueId = reinterpret_cast<decltype(GetUniqueId)>(dlsym(lib, "ncclGetUniqueId"));
        CommInitRank = reinterpret_cast<decltype(CommInitRank)>(dlsym(lib, "ncclCommInitRank"));
        AllReduce = reinterpret_cast<decltype(AllReduce)>(dlsym(lib, "ncclAllReduce"));
        CommDestroy = reinterpret_cast<decltype(CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
        GetErrorString = reinterpret_cast<decltype(GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
        if (!GetUniqueId || !CommInitRank || !AllReduce || !CommDestroy || !GetErrorString) {
            error = "libnccl.so.2 lacks a required symbol";
            lib = nullptr;
            return false;
        }
        return true;
    }
};
NcclApi g_nccl;

int allreduce(visfs_ba_handle *h, void *buf, size_t count, ncclDataType_t dt, ncclRedOp_t op) {
    if (!h->comm || h->comm_ranks <= 1 || !h->partitioned || count == 0) return VISFS_BA_OK;
    const ncclResult_t r = g_nccl.AllReduce(buf, buf, count, dt, op, h->comm, h->stream);
    if (r != ncclSuccess) return h->fail(VISFS_BA_ERR_CUDA, std::string("ncclAllReduce: ") + g_nccl.GetErrorString(r));
    h->allreduce_calls += 1;
    h->allreduce_bytes += (int64_t)count * (dt == ncclFloat64 ? 8 : 4);
    return VISFS_BA_OK;
}

// ---- large-window path ---------------------------------------------------------------------------------------------

// Schedule of k_solve_front (ba_large.cuh), built on the host from the envelope once per pass: which rows enter the
// front at which column, the shared-memory slot each row holds, and the skyline blocks to bring in for the entering rows.
// Leaves use_front = false when the front does not fit the slot matrix.
int plan_front(visfs_ba_handle *h, int F) {
    cudaStream_t s = h->stream;
    std::vector<int> first((size_t)F);
    CK(cudaMemcpyAsync(first.data(), h->d_sky_first.p, sizeof(int) * (size_t)F, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const int SL = lg::kFrontSlots;
    std::vector<int> eptr((size_t)F + 1, 0), erow((size_t)F), ebase((size_t)F), row_len((size_t)F), row_off((size_t)F);
    long long off = 0;
    for (int r = 0; r < F; ++r) {
        row_len[r] = r - first[r];
        row_off[r] = (int)off;
        off += row_len[r] + 1;
        eptr[(size_t)first[r] + 1] += 1;
    }
    if (off != h->n_sky) return h->fail(VISFS_BA_ERR_CUDA, "front plan: envelope size differs from the device's");
    for (int k = 0; k < F; ++k) eptr[k + 1] += eptr[k];
    {
        std::vector<int> fill(eptr.begin(), eptr.end() - 1);
        for (int r = 0; r < F; ++r) {   // ascending r inside a column
            const int i = fill[first[r]]++;
            erow[i] = r;
            ebase[i] = row_off[r] - first[r];
        }
    }
    // column structure (rows r > k with first[r] <= k, ascending) — the same lists k_col_fill builds on the device
    std::vector<int> cptr((size_t)F + 1, 0);
    for (int r = 0; r < F; ++r) for (int k = first[r]; k < r; ++k) cptr[(size_t)k + 1] += 1;
    for (int k = 0; k < F; ++k) cptr[k + 1] += cptr[k];
    std::vector<int> crow((size_t)cptr[F]);
    {
        std::vector<int> fill(cptr.begin(), cptr.end() - 1);
        for (int r = 0; r < F; ++r) for (int k = first[r]; k < r; ++k) crow[fill[k]++] = r;
    }
    // slots: a row takes one when it enters and gives it back one column after it was the pivot
    std::vector<unsigned char> slot((size_t)F, 0);
    std::vector<int> free_slots;
    for (int q = SL - 1; q >= 0; --q) free_slots.push_back(q);
    for (int k = 0; k < F; ++k) {
        for (int i = eptr[k]; i < eptr[k + 1]; ++i) {
            if (free_slots.empty()) return VISFS_BA_OK;   // front too wide: the other solvers take over
            slot[(size_t)erow[i]] = (unsigned char)free_slots.back();
            free_slots.pop_back();
        }
        if (k >= 1) free_slots.push_back(slot[(size_t)k - 1]);
    }
    std::vector<unsigned char> fr_slot(crow.size());
    for (size_t i = 0; i < crow.size(); ++i) fr_slot[i] = slot[(size_t)crow[i]];
    // skyline blocks to bring in for the rows entering at column k: against the pivot and every row of the column structure
    std::vector<int> lptr((size_t)F + 1, 0), lsrc;
    std::vector<unsigned short> ldst;
    lsrc.reserve((size_t)h->n_sky); ldst.reserve((size_t)h->n_sky);
    for (int k = 0; k < F; ++k) {
        for (int i = eptr[k]; i < eptr[k + 1]; ++i) {
            const int r = erow[i];
            auto emit = [&](int q) {
                if (first[q] == k && q > r) return;   // both enter here: emitted once, by the larger row
                const int hi = std::max(r, q), lo = std::min(r, q);
                lsrc.push_back(row_off[hi] + (lo - first[hi]));
                ldst.push_back((unsigned short)(slot[(size_t)hi] * lg::kFrontPitch + slot[(size_t)lo]));
            };
            emit(k);
            for (int c = cptr[k]; c < cptr[k + 1]; ++c) emit(crow[c]);
        }
        lptr[(size_t)k + 1] = (int)lsrc.size();
    }
    // one device buffer: [pslot | fr_slot | ldst | eptr | erow | ebase | lptr | lsrc | row_len | row_off]
    auto al = [](size_t v) { return (v + 15) & ~(size_t)15; };
    const size_t o_pslot = 0, o_fr = al(o_pslot + (size_t)F), o_ldst = al(o_fr + fr_slot.size()), o_eptr = al(o_ldst + 2 * ldst.size()),
                 o_erow = al(o_eptr + 4 * ((size_t)F + 1)), o_ebase = al(o_erow + 4 * (size_t)F), o_lptr = al(o_ebase + 4 * (size_t)F),
                 o_lsrc = al(o_lptr + 4 * ((size_t)F + 1)), o_rlen = al(o_lsrc + 4 * lsrc.size()), o_roff = al(o_rlen + 4 * (size_t)F),
                 o_end = al(o_roff + 4 * (size_t)F);
    std::vector<unsigned char> pack(o_end, 0);
    memcpy(pack.data() + o_pslot, slot.data(), (size_t)F);
    if (!fr_slot.empty()) memcpy(pack.data() + o_fr, fr_slot.data(), fr_slot.size());
    if (!ldst.empty()) memcpy(pack.data() + o_ldst, ldst.data(), 2 * ldst.size());
    memcpy(pack.data() + o_eptr, eptr.data(), 4 * ((size_t)F + 1));
    memcpy(pack.data() + o_erow, erow.data(), 4 * (size_t)F);
    memcpy(pack.data() + o_ebase, ebase.data(), 4 * (size_t)F);
    memcpy(pack.data() + o_lptr, lptr.data(), 4 * ((size_t)F + 1));
    if (!lsrc.empty()) memcpy(pack.data() + o_lsrc, lsrc.data(), 4 * lsrc.size());
    memcpy(pack.data() + o_rlen, row_len.data(), 4 * (size_t)F);
    memcpy(pack.data() + o_roff, row_off.data(), 4 * (size_t)F);
    CK(h->d_plan.reserve(o_end));
    CK(cudaMemcpyAsync(h->d_plan.p, pack.data(), o_end, cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));
    const unsigned char *base = h->d_plan.as<unsigned char>();
    lg::FrontPlan &P = h->front_plan;
    P.pslot = base + o_pslot; P.fr_slot = base + o_fr;
    P.ld_dst = reinterpret_cast<const unsigned short *>(base + o_ldst);
    P.ent_ptr = reinterpret_cast<const int *>(base + o_eptr); P.ent_row = reinterpret_cast<const int *>(base + o_erow);
    P.ent_base = reinterpret_cast<const int *>(base + o_ebase); P.ld_ptr = reinterpret_cast<const int *>(base + o_lptr);
    P.ld_src = reinterpret_cast<const int *>(base + o_lsrc); P.row_len = reinterpret_cast<const int *>(base + o_rlen);
    P.row_off = reinterpret_cast<const int *>(base + o_roff);
    h->use_front = true;
    return VISFS_BA_OK;
}


// Plan of the multifrontal solver (ba_mf.cuh), built on the host from the envelope once per pass.  Rows whose envelope is
// longer than mf::kMaxBand blocks are "arrows" (loop closures): they are eliminated last and sit in every front's boundary.
// The remaining rows form a band of half-width w <= mf::kMaxBand; the chain is cut by nested dissection into leaves of
// <= 9 poses and separators of w poses.  Leaves use_mf = false when the structure does not fit (the other solvers take over).
int plan_mf(visfs_ba_handle *h, int F) {
    cudaStream_t s = h->stream;
    h->use_mf = false;
    std::vector<int> first((size_t)F);
    CK(cudaMemcpyAsync(first.data(), h->d_sky_first.p, sizeof(int) * (size_t)F, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    std::vector<int> comp, arrows, before((size_t)F + 1, 0);   // before[r] = non-arrow rows with index < r
    for (int r = 0; r < F; ++r) {
        const bool arrow = r - first[r] > mf::kMaxBand;
        before[(size_t)r + 1] = before[r] + (arrow ? 0 : 1);
        (arrow ? arrows : comp).push_back(r);
    }
    if ((int)arrows.size() > mf::kMaxArrow) return VISFS_BA_OK;
    const int N = (int)comp.size();
    int w = 1;
    for (int q = 0; q < N; ++q) w = std::max(w, q - before[first[comp[q]]]);
    if (w > mf::kMaxBand || N < 4 * w) return VISFS_BA_OK;
    // which columns every arrow really couples with (device: the landmarks' pose lists; every rank of a partitioned run
    // sees its own landmarks only, so the maps are merged with an integer MAX all-reduce)
    const int n_arrow = (int)arrows.size();
    std::vector<int> touch((size_t)n_arrow * F, 0);
    if (n_arrow > 0) {
        std::vector<int> arrow_of((size_t)F, -1);
        for (int a = 0; a < n_arrow; ++a) arrow_of[(size_t)arrows[a]] = a;
        CK(h->d_mf_touch.reserve(sizeof(int) * ((size_t)F + (size_t)n_arrow * F)));
        int *d_arrow_of = h->d_mf_touch.as<int>(), *d_touch = d_arrow_of + F;
        CK(cudaMemcpyAsync(d_arrow_of, arrow_of.data(), sizeof(int) * (size_t)F, cudaMemcpyHostToDevice, s));
        CK(cudaMemsetAsync(d_touch, 0, sizeof(int) * (size_t)n_arrow * F, s));
        mf::k_arrow_touch<<<std::max(1, std::min((h->tot_point + 255) / 256, 1024)), 256, 0, s>>>(h->batch, d_arrow_of, d_touch, F);
        CK(cudaGetLastError());
        const int st_ar = allreduce(h, d_touch, (size_t)n_arrow * F, ncclInt32, ncclMax);
        if (st_ar) return st_ar;
        CK(cudaMemcpyAsync(touch.data(), d_touch, sizeof(int) * (size_t)n_arrow * F, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
    }

    struct Node { int lo, hi, e0, e1, level, parent; std::vector<int> kids; unsigned reach = 0; };   // reach: arrows coupled with the subtree
    std::vector<Node> nodes;
    // recursive bisection of the compressed range [lo, hi): returns the node that eliminates its separator (or the leaf)
    std::vector<std::array<int, 3>> stack;   // (lo, hi, parent) — children are created after their parent, levels fixed below
    struct Rec { static int build(std::vector<Node> &nodes, int lo, int hi, int w) {
        const int L = hi - lo;
        Node nd{lo, hi, lo, hi, 0, -1, {}, 0};
        if (L <= 9 || L <= w + 1) { nodes.push_back(nd); return (int)nodes.size() - 1; }
        const int m = lo + (L - w) / 2;
        nd.e0 = m; nd.e1 = m + w;
        int k0 = -1, k1 = -1;
        if (m > lo) k0 = build(nodes, lo, m, w);
        if (hi > m + w) k1 = build(nodes, m + w, hi, w);
        nd.level = 1 + std::max(k0 >= 0 ? nodes[k0].level : -1, k1 >= 0 ? nodes[k1].level : -1);
        if (k0 >= 0) nd.kids.push_back(k0);
        if (k1 >= 0) nd.kids.push_back(k1);
        nodes.push_back(nd);
        const int id = (int)nodes.size() - 1;
        for (int k : nodes[id].kids) nodes[k].parent = id;
        return id;
    } };
    int top = Rec::build(nodes, 0, N, w);
    if (!arrows.empty()) {
        Node nd{0, N, -1, -1, nodes[top].level + 1, -1, {top}, 0};   // e0 = -1: eliminates the arrows
        nodes.push_back(nd);
        nodes[top].parent = (int)nodes.size() - 1;
        top = (int)nodes.size() - 1;
    }
    // arrows a front can reach: those coupled with a column it eliminates, and whatever its children reach (children have
    // smaller node indices than their parents: one ascending sweep)
    for (size_t i = 0; i < nodes.size(); ++i) {
        Node &nd = nodes[i];
        if (nd.e0 >= 0)
            for (int a = 0; a < n_arrow; ++a)
                for (int q = nd.e0; q < nd.e1 && !(nd.reach >> a & 1u); ++q)
                    if (touch[(size_t)a * F + comp[q]]) nd.reach |= 1u << a;
        for (int k : nd.kids) nd.reach |= nodes[(size_t)k].reach;
    }
    // problems in level order (children before parents; ascending node index inside a level keeps the order deterministic)
    const int n_nodes = (int)nodes.size();
    std::vector<int> order(n_nodes), newid(n_nodes);
    for (int i = 0; i < n_nodes; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return nodes[a].level < nodes[b].level; });
    for (int i = 0; i < n_nodes; ++i) newid[order[i]] = i;
    const int n_level = nodes[top].level + 1;
    h->mf_level_off.assign((size_t)n_level + 1, 0);
    for (int i = 0; i < n_nodes; ++i) h->mf_level_off[(size_t)nodes[order[i]].level + 1] += 1;
    for (int l = 0; l < n_level; ++l) h->mf_level_off[l + 1] += h->mf_level_off[l];

    std::vector<mf::Prob> prob(n_nodes);
    std::vector<int> idx, child, pmap;
    std::vector<std::vector<int>> local(n_nodes);   // per node: hessian indices of its local blocks
    long long d_len = 0, linv_len = 0;
    for (int i = 0; i < n_nodes; ++i) {
        const Node &nd = nodes[order[i]];
        std::vector<int> &loc = local[order[i]];
        int ne = 0;
        if (nd.e0 < 0) { loc = arrows; ne = (int)arrows.size(); }
        else {
            for (int q = nd.e0; q < nd.e1; ++q) loc.push_back(comp[q]);
            ne = nd.e1 - nd.e0;
            for (int q = std::max(0, nd.lo - w); q < nd.lo; ++q) loc.push_back(comp[q]);
            for (int q = nd.hi; q < std::min(N, nd.hi + w); ++q) loc.push_back(comp[q]);
            for (int a = 0; a < n_arrow; ++a) if (nd.reach >> a & 1u) loc.push_back(arrows[a]);
        }
        mf::Prob &p = prob[i];
        p.ne = ne; p.nb = (int)loc.size() - ne;
        if (p.ne > 2 * mf::kMaxBand || p.nb > 2 * mf::kMaxBand + mf::kMaxArrow || p.ne < 1) return VISFS_BA_OK;
        const int n_t = 6 * (p.ne + p.nb);
        p.LD = (n_t + 32 + 7) / 8 * 8;
        p.d_off = d_len; d_len += ((long long)p.LD * p.LD + 15) / 16 * 16;
        p.linv_off = linv_len; linv_len += (long long)((6 * p.ne + dn::kNB - 1) / dn::kNB) * dn::kNB * dn::kNB;
        p.idx_off = (int)idx.size();
        idx.insert(idx.end(), loc.begin(), loc.end());
        p.child_off = (int)child.size(); p.n_child = (int)nd.kids.size();
        std::vector<int> kids;
        for (int k : nd.kids) kids.push_back(newid[k]);
        std::sort(kids.begin(), kids.end());
        child.insert(child.end(), kids.begin(), kids.end());
        p.map_off = 0;
    }
    // boundary of a child -> position in its parent's front
    std::vector<int> where((size_t)F, -1);
    for (int i = 0; i < n_nodes; ++i) {
        const Node &nd = nodes[order[i]];
        if (nd.parent < 0) continue;
        const std::vector<int> &ploc = local[nd.parent];
        for (size_t k = 0; k < ploc.size(); ++k) where[(size_t)ploc[k]] = (int)k;
        prob[i].map_off = (int)pmap.size();
        const std::vector<int> &loc = local[order[i]];
        for (int b = prob[i].ne; b < (int)loc.size(); ++b) {
            if (where[(size_t)loc[b]] < 0) return h->fail(VISFS_BA_ERR_CUDA, "multifrontal plan: a boundary block is missing from the parent's front");
            pmap.push_back(where[(size_t)loc[b]]);
        }
        for (size_t k = 0; k < ploc.size(); ++k) where[(size_t)ploc[k]] = -1;
    }
    if (d_len * (long long)sizeof(double) > ((long long)3 << 30)) return VISFS_BA_OK;
    // one metadata buffer: [prob | idx | child | pmap]
    auto al = [](size_t v) { return (v + 15) & ~(size_t)15; };
    const size_t o_prob = 0, o_idx = al(sizeof(mf::Prob) * prob.size()), o_child = al(o_idx + 4 * idx.size()),
                 o_pmap = al(o_child + 4 * std::max<size_t>(child.size(), 1)), o_end = al(o_pmap + 4 * std::max<size_t>(pmap.size(), 1));
    std::vector<unsigned char> pack(o_end, 0);
    memcpy(pack.data() + o_prob, prob.data(), sizeof(mf::Prob) * prob.size());
    memcpy(pack.data() + o_idx, idx.data(), 4 * idx.size());
    if (!child.empty()) memcpy(pack.data() + o_child, child.data(), 4 * child.size());
    if (!pmap.empty()) memcpy(pack.data() + o_pmap, pmap.data(), 4 * pmap.size());
    CK(h->d_mf_meta.reserve(o_end));
    CK(cudaMemcpyAsync(h->d_mf_meta.p, pack.data(), o_end, cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));
    CK(h->d_mf_fronts.reserve(sizeof(double) * (size_t)(d_len + linv_len + 6 * (long long)F + 64)));
    const unsigned char *mb = h->d_mf_meta.as<unsigned char>();
    mf::Plan &P = h->mf_plan;
    P.prob = reinterpret_cast<const mf::Prob *>(mb + o_prob); P.idx = reinterpret_cast<const int *>(mb + o_idx);
    P.child = reinterpret_cast<const int *>(mb + o_child); P.pmap = reinterpret_cast<const int *>(mb + o_pmap);
    P.fronts = h->d_mf_fronts.as<double>(); P.linvt = P.fronts + d_len; P.x = P.linvt + linv_len;
    P.flag = h->d_cnt.as<int>() + 2;
    P.prof = nullptr;
    if (getenv("VISFS_BA_DENSE_PROF")) {
        CK(h->d_dense_prof.reserve(sizeof(long long) * 2600));
        CK(cudaMemsetAsync(h->d_dense_prof.p, 0, sizeof(long long) * 2600, s));
        P.prof = h->d_dense_prof.as<long long>();
    }
    h->use_mf = true;
    return VISFS_BA_OK;
}

int dev_scan(visfs_ba_handle *h, const int *in, int *out, int n, bool inclusive) {
    size_t tb = 0;
    if (inclusive) cub::DeviceScan::InclusiveSum(nullptr, tb, in, out, n, h->stream);
    else cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, n, h->stream);
    CK(h->d_sort_tmp.reserve(tb));
    if (inclusive) CK(cub::DeviceScan::InclusiveSum(h->d_sort_tmp.p, tb, in, out, n, h->stream));
    else CK(cub::DeviceScan::ExclusiveSum(h->d_sort_tmp.p, tb, in, out, n, h->stream));
    return VISFS_BA_OK;
}

// Band chunks of a large window (ba_band.cuh), once per pass after the envelope is laid out.  Needs the landmarks sorted by
// their first pose (d_lm_key2 = sorted keys, d_lm_rec = sorted records).  Leaves use_band = false when the map has no band
// structure (landmarks seen from all over the trajectory: C5).
int prep_band(visfs_ba_handle *h) {
    cudaStream_t s = h->stream;
    Batch &B = h->batch;
    h->use_band = false; h->band_rest = false;
    const int L = h->tot_point, E = h->tot_edge;
    const bool force = getenv("VISFS_BA_BAND_FORCE") != nullptr;   // tests: band chunks on maps of any size and shape
    if ((L < 4096 && !force) || L <= 0 || E <= 0 || h->n_sky >= 0x7fffffffLL / 36) return VISFS_BA_OK;
    const int g = std::max(1, std::min((L + 256) / 256, 4 * h->sm_count));
    const size_t n1 = ((size_t)L + 1 + 3) & ~(size_t)3;
    CK(h->d_bd_scan.reserve(sizeof(int) * 6 * n1));
    int *deg = h->d_bd_scan.as<int>(), *sorted_off = deg + n1, *newkey = sorted_off + n1, *rank = newkey + n1, *start = rank + n1, *cid = start + n1;
    int4 *rec = h->d_lm_rec.as<int4>();
    bd::k_band_deg<<<g, 256, 0, s>>>(rec, h->d_lm_key2.as<int>(), L, deg, newkey);
    int st;
    if ((st = dev_scan(h, deg, sorted_off, L + 1, false))) return st;
    if ((st = dev_scan(h, newkey, rank, L + 1, true))) return st;
    int *hs = h->h_small.as<int>() + 16;   // bytes 64 .. 80 of the 128-byte pinned scratch
    // first-pose values per chunk: one when a value has >= ~256 landmarks (measured best on C4, ba_band.cuh); a rank of a
    // partitioned run holds 1 / N of every value's landmarks, so it takes N values per chunk to keep the chunks at a size
    // where the per-chunk prologue / epilogue does not dominate (bounded by the 19 poses a chunk may touch)
    CK(cudaMemcpyAsync(hs, rank + (L - 1), sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const int n_keys = std::max(hs[0], 1);
    int keys = (int)std::lround(256.0 * n_keys / L), max_lm = bd::kBandMaxLm;
    keys = std::max(bd::kBandKeys, std::min(keys, std::max(1, ws::kBandPoses - std::max(h->max_deg, 1))));
    if (const char *e = getenv("VISFS_BA_BAND_KEYS")) keys = std::max(1, atoi(e));
    if (const char *e = getenv("VISFS_BA_BAND_MAXLM")) max_lm = std::max(32, atoi(e));
    bd::k_band_gstart<<<g, 256, 0, s>>>(rank, L, keys, start);
    {   // cid (free until the chunk ids are formed) = start of every landmark's group: inclusive max scan
        size_t tb = 0;
        cub::DeviceScan::InclusiveScan(nullptr, tb, start, cid, cub::Max(), L + 1, s);
        CK(h->d_sort_tmp.reserve(tb));
        CK(cub::DeviceScan::InclusiveScan(h->d_sort_tmp.p, tb, start, cid, cub::Max(), L + 1, s));
    }
    bd::k_band_flags<<<g, 256, 0, s>>>(cid, L, start, max_lm);
    if ((st = dev_scan(h, start, cid, L + 1, true))) return st;
    CK(cudaMemcpyAsync(hs, cid + L, sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const int nc = hs[0];
    h->launches += 6;
    if (nc <= 0 || nc >= (1 << 22) || (size_t)nc * ws::kBandPartStride * sizeof(double) > ((size_t)8 << 30)) return VISFS_BA_OK;   // (partials <= 8 GB)
    // per-chunk tables: [chunk (4 ints) | chunk_pose (19) | ntiles | tile_off | npair | pair_off | npose | pose_off] (nc + 1 each) | counts
    const size_t c1 = ((size_t)nc + 1 + 3) & ~(size_t)3;
    CK(h->d_bd_chunk.reserve(sizeof(int) * (c1 * (6 + ws::kBandPoses + 10) + 4)));
    int *cb = h->d_bd_chunk.as<int>();
    ws::BandChunk *chunk = reinterpret_cast<ws::BandChunk *>(cb);
    int *chunk_pose = cb + 6 * c1, *ntiles = chunk_pose + ws::kBandPoses * c1, *tile_off = ntiles + c1, *npair = tile_off + c1,
        *pair_off = npair + c1, *npose = pair_off + c1, *pose_off = npose + c1, *cost = pose_off + c1, *cidx = cost + c1,
        *cost2 = cidx + c1, *order = cost2 + c1, *counts = order + c1;
    CK(cudaMemsetAsync(counts, 0, sizeof(int) * 4, s));
    CK(h->d_bd_edge.reserve(sizeof(int) * 3 * (size_t)E));
    CK(h->d_bd_obs.reserve(sizeof(double) * 3 * (size_t)E));
    int *s_pw = h->d_bd_edge.as<int>(), *s_gl = s_pw + E, *s_sl = s_gl + E;
    double *s_ou = h->d_bd_obs.as<double>(), *s_ov = s_ou + E, *s_our = s_ov + E;
    const int gc = std::max(1, (nc + 1 + 127) / 128);
    bd::k_band_ranges<<<g, 256, 0, s>>>(start, cid, L, chunk);
    bd::k_band_chunk<<<nc, 256, 0, s>>>(B, rec, sorted_off, chunk, chunk_pose, s_pw, s_gl, s_sl, s_ou, s_ov, s_our, counts);
    bd::k_band_count_tiles<<<gc, 128, 0, s>>>(chunk, nc, sorted_off, ntiles, npair, npose, cost, cidx);
    {
        size_t tb = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tb, cost, cost2, cidx, order, nc, 0, 32, s);
        CK(h->d_sort_tmp.reserve(tb));
        CK(cub::DeviceRadixSort::SortPairs(h->d_sort_tmp.p, tb, cost, cost2, cidx, order, nc, 0, 32, s));
    }
    if ((st = dev_scan(h, ntiles, tile_off, nc + 1, false))) return st;
    if ((st = dev_scan(h, npair, pair_off, nc + 1, false))) return st;
    if ((st = dev_scan(h, npose, pose_off, nc + 1, false))) return st;
    CK(cudaMemcpyAsync(hs + 0, tile_off + nc, sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(hs + 1, pair_off + nc, sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(hs + 2, pose_off + nc, sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(hs + 3, counts, sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const int n_tiles = hs[0], n_pair = hs[1], n_poseent = hs[2], n_in = hs[3];
    h->launches += 6;
    if (getenv("VISFS_BA_VERBOSE"))
        fprintf(stderr, "[visfs_ba] band chunks: %d first-pose values per chunk, %d chunks, %d tiles, %d of %d landmarks inside, %d + %d gather entries\n", keys, nc, n_tiles, n_in, L, n_pair, n_poseent);
    // worth it when most of the map is inside band chunks of a useful size
    if (n_tiles <= 0 || n_pair <= 0 || n_in <= 0) return VISFS_BA_OK;
    if (!force && (n_in * 2 < L || (long long)n_in < 32LL * nc)) return VISFS_BA_OK;
    const int n_ent = n_pair + n_poseent;
    CK(h->d_bd_tiles.reserve(sizeof(Tile) * (size_t)n_tiles));
    CK(h->d_bd_ent.reserve(sizeof(int) * (5 * (size_t)n_ent + 12)));
    CK(h->d_bd_val.reserve(sizeof(unsigned long long) * 2 * (size_t)n_ent));
    CK(h->d_bd_part.reserve(sizeof(double) * (size_t)nc * ws::kBandPartStride));
    CK(h->d_bd_part2.reserve(sizeof(double) * 2 * (size_t)nc));
    int *key = h->d_bd_ent.as<int>(), *key2 = key + n_ent, *flag = key2 + n_ent, *sid = flag + n_ent + 1, *seg_start = sid + n_ent + 1;
    // (flag / sid / seg_start have n_ent + 1 entries: every key may be a segment of its own)
    unsigned long long *val = h->d_bd_val.as<unsigned long long>(), *val2 = val + n_ent;
    bd::k_band_fill_tiles<<<gc, 128, 0, s>>>(chunk, nc, sorted_off, tile_off, h->d_bd_tiles.as<Tile>());
    bd::k_band_entries<<<nc, 128, 0, s>>>(B, chunk, nc, chunk_pose, pair_off, n_pair, pose_off, key, val);
    {
        size_t tb = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tb, key, key2, val, val2, n_ent, 0, 32, s);
        CK(h->d_sort_tmp.reserve(tb));
        CK(cub::DeviceRadixSort::SortPairs(h->d_sort_tmp.p, tb, key, key2, val, val2, n_ent, 0, 32, s));
    }
    const int ge = std::max(1, std::min((n_ent + 256) / 256, 4 * h->sm_count));
    bd::k_band_seg_flags<<<ge, 256, 0, s>>>(key2, n_ent, flag);
    if ((st = dev_scan(h, flag, sid, n_ent + 1, false))) return st;
    CK(cudaMemcpyAsync(hs, sid + n_ent, sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const int n_seg = hs[0];
    if (n_seg <= 0 || n_seg > n_ent) return VISFS_BA_OK;
    CK(h->d_bd_scan.reserve(sizeof(int) * 6 * n1));   // (no-op: keeps the scan arrays alive for the pass)
    bd::k_band_seg_starts<<<ge, 256, 0, s>>>(flag, sid, n_ent, seg_start);
    CK(cudaGetLastError());
    h->launches += 6;
    ws::Band &bd_ = h->band;
    bd_.chunk = chunk; bd_.chunk_pose = chunk_pose; bd_.sorted_off = sorted_off; bd_.s_pw = s_pw; bd_.s_gl = s_gl; bd_.s_sl = s_sl;
    bd_.s_ou = s_ou; bd_.s_ov = s_ov; bd_.s_our = s_our; bd_.lm_rec = rec; bd_.tiles = h->d_bd_tiles.as<Tile>(); bd_.chunk_tile_off = tile_off; bd_.order = order;
    bd_.part = h->d_bd_part.as<double>(); bd_.n_chunk = nc;
    h->band_key = key2; h->band_val = val2; h->band_seg = seg_start;
    h->band_n_chunk = nc; h->band_n_seg = n_seg;
    h->band_rest = n_in < L;
    h->band_n_rest = L - n_in;
    if (h->band_rest) {   // what the chunks left over, compacted in sorted order (deg / newkey are free again)
        CK(h->d_bd_rest.reserve(sizeof(int4) * (size_t)(L - n_in)));
        bd::k_band_rest_flag<<<g, 256, 0, s>>>(rec, L, newkey);
        if ((st = dev_scan(h, newkey, deg, L + 1, false))) return st;
        bd::k_band_rest_fill<<<g, 256, 0, s>>>(rec, newkey, deg, L, h->d_bd_rest.as<int4>());
        CK(cudaGetLastError());
        h->launches += 3;
    }
    h->use_band = true;
    return VISFS_BA_OK;
}

int run_structure_large(visfs_ba_handle *h) {
    Batch &B = h->batch;
    cudaStream_t s = h->stream;
    int st;
    CK(cudaMemsetAsync(h->d_pose_active.p, 0, sizeof(int) * std::max(h->tot_pose, 1), s));
    const dim3 glm((unsigned)h->grid_lm_x, 1u);
    k_struct_lm<<<glm, 256, 0, s>>>(B);
    if ((st = allreduce(h, h->d_pose_active.p, (size_t)h->tot_pose, ncclInt32, ncclMax))) return st;
    k_struct_pose<<<1, 128, 0, s>>>(B);
    k_struct_count<<<glm, 256, 0, s>>>(B, 0);
    k_struct_finish<<<1, 128, 0, s>>>(B);
    if (h->partitioned && h->comm_ranks > 1) {   // landmarks in the Hessian: sum over ranks; device-side error: any rank
        lg::k_get_counts<<<1, 32, 0, s>>>(B, h->d_cnt.as<int>());
        if ((st = allreduce(h, h->d_cnt.p, 2, ncclInt32, ncclSum))) return st;
        lg::k_set_counts<<<1, 32, 0, s>>>(B, h->d_cnt.as<int>());
        h->launches += 2;
    }
    // landmarks in the order of their first pose for the run-aggregating build kernel (degree <= 10 only)
    h->use_run = h->max_deg >= 0 && h->max_deg <= lg::kRunDeg && h->tot_point > 0 && !getenv("VISFS_BA_NO_RUN");
    if (h->use_run) {
        const size_t L = (size_t)h->tot_point;
        CK(h->d_lm_key.reserve(sizeof(int) * L)); CK(h->d_lm_key2.reserve(sizeof(int) * L));
        CK(h->d_lm_idx.reserve(sizeof(int) * L)); CK(h->d_lm_order.reserve(sizeof(int) * L));
        lg::k_lm_first<<<glm, 256, 0, s>>>(B, h->d_lm_key.as<int>(), h->d_lm_idx.as<int>());
        size_t tb = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tb, h->d_lm_key.as<int>(), h->d_lm_key2.as<int>(), h->d_lm_idx.as<int>(),
                                        h->d_lm_order.as<int>(), (int)L, 0, 32, s);
        CK(h->d_sort_tmp.reserve(tb));
        CK(cub::DeviceRadixSort::SortPairs(h->d_sort_tmp.p, tb, h->d_lm_key.as<int>(), h->d_lm_key2.as<int>(), h->d_lm_idx.as<int>(),
                                           h->d_lm_order.as<int>(), (int)L, 0, 32, s));
        CK(h->d_lm_rec.reserve(sizeof(int4) * L));
        CK(cudaMemsetAsync(h->d_info.as<long long>() + 3, 0, sizeof(long long), s));
        lg::k_run_prep<<<glm, 256, 0, s>>>(B, h->d_lm_order.as<int>(), h->d_lm_rec.as<int4>(), reinterpret_cast<int *>(h->d_info.as<long long>() + 3));
        h->launches += 3;
    }
    const int gp = std::max(1, std::min((h->tot_pose + 255) / 256, 64));
    lg::k_sky_init<<<gp, 256, 0, s>>>(B);
    lg::k_sky_first<<<glm, 256, 0, s>>>(B);
    if (h->tot_link > 0) lg::k_sky_links<<<(h->tot_link + 127) / 128, 128, 0, s>>>(B);
    if ((st = allreduce(h, h->d_sky_first.p, (size_t)h->tot_pose, ncclInt32, ncclMin))) return st;
    lg::k_sky_layout<<<1, 1024, 0, s>>>(B, h->d_col_cnt.as<int>(), h->d_info.as<long long>());
    lg::k_col_count<<<std::max(1, std::min(h->tot_pose, 1024)), 128, 0, s>>>(B, h->d_col_cnt.as<int>());
    lg::k_col_scan<<<1, 1024, 0, s>>>(B, h->d_col_cnt.as<int>(), h->d_info.as<long long>());
    long long *info = h->h_small.as<long long>() + 2;
    CK(cudaMemcpyAsync(info, h->d_info.p, sizeof(long long) * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    h->n_sky = info[0];
    h->max_front = (int)info[2];
    const bool sorted_lm = h->use_run;
    if (h->use_run) {   // worth it only when landmarks that touch the same poses follow each other: runs of >= 4 on average
        const long long n_runs = (long long)(int)(info[3] & 0xffffffffLL);
        if (n_runs * 4 > h->tot_point) h->use_run = false;
    }
    h->use_band = false;
    if (sorted_lm && !getenv("VISFS_BA_NO_BAND")) {
        const int st2 = prep_band(h);
        if (st2) return st2;
    }
    const long long F = info[1];
    h->st_F_hint = (int)F;
    const size_t red_len = (size_t)h->n_sky * 36 + 12 * (size_t)F;
    CK(h->d_red.reserve(sizeof(double) * std::max<size_t>(red_len, 8)));
    CK(h->d_col_rows.reserve(sizeof(int) * (size_t)std::max<long long>(h->n_sky, 1)));
    Batch *both[2] = {&h->batch, &h->batch_ctl};
    for (Batch *b : both) {
        b->red = h->d_red.as<double>(); b->col_rows = h->d_col_rows.as<int>();
        b->red_g_off = h->n_sky * 36; b->red_bp_off = h->n_sky * 36 + 6 * F;
    }
    lg::k_col_fill<<<std::max(1, std::min((h->tot_pose + 7) / 8, 512)), 256, 0, s>>>(h->batch);
    h->use_dense = false;
    if (F > 0 && h->max_front > 32 && h->dense_grid > 0 && !getenv("VISFS_BA_NO_DENSE")) {
        // dense contraction: the envelope holds at least a quarter of the full lower triangle (C5: all of it) and D fits easily
        const int n = 6 * (int)F, NP = dn::dense_pad(n);
        const int own = (n + 1 + h->dense_grid - 1) / h->dense_grid;
        const size_t smem_chol = sizeof(double) * ((size_t)(dn::kNB + own + 1) * dn::kTP + (size_t)(dn::kNB + own + 1) * 6 + 2 * dn::kNB);
        const int npan = (n + dn::kNB - 1) / dn::kNB;
        if (h->n_sky * 4 >= F * (F + 1) / 2 && (size_t)NP * NP * sizeof(double) <= ((size_t)1 << 30) && own + 1 <= dn::kMaxOwn &&
            smem_chol <= 200 * 1024) {
            const size_t extra = (size_t)npan * dn::kNB * dn::kNB + (size_t)n + 64;
            CK(h->d_dense.reserve(sizeof(double) * ((size_t)NP * NP + extra)));
            h->dense.D = h->d_dense.as<double>(); h->dense.n = n; h->dense.LD = NP; h->dense.NP = NP;
            h->dense.Linvt = h->dense.D + (size_t)NP * NP; h->dense.x = h->dense.Linvt + (size_t)npan * dn::kNB * dn::kNB;
            h->dense.flag = h->d_cnt.as<int>() + 2;
            h->dense.prof = nullptr;
            if (getenv("VISFS_BA_DENSE_PROF")) {   // phase timing of the last k_dense_chol launch, printed by run_resident
                CK(h->d_dense_prof.reserve(sizeof(long long) * 2600));
                CK(cudaMemsetAsync(h->d_dense_prof.p, 0, sizeof(long long) * 2600, s));
                h->dense.prof = h->d_dense_prof.as<long long>();
            }
            h->dense_smem = smem_chol;
            h->use_dense = true;
        }
    }
    h->use_mf = false;
    {
        int mf_min = 256;   // below that the one-CTA frontal solver is as fast
        if (const char *e = getenv("VISFS_BA_MF_MIN")) mf_min = atoi(e);
        if (!h->use_dense && F >= mf_min && !getenv("VISFS_BA_NO_MF")) {
            const int st2 = plan_mf(h, (int)F);
            if (st2) return st2;
        }
    }
    h->use_front = false;
    if (!h->use_dense && !h->use_mf && F > 0 && F <= lg::kFrontMaxF && h->max_front + 2 <= lg::kFrontSlots && h->n_sky < 0x7fffffffLL / 36 && !getenv("VISFS_BA_NO_FRONT")) {
        const int st2 = plan_front(h, (int)F);
        if (st2) return st2;
    }
    CK(cudaGetLastError());
    h->launches += 11;
    return VISFS_BA_OK;
}

size_t red_doubles(const visfs_ba_handle *h) { return (size_t)h->batch.red_bp_off + (size_t)(h->batch.red_bp_off - h->batch.red_g_off); }

int enqueue_build_large(visfs_ba_handle *h) {
    int ev = ev_begin(h, EV_BUILD);
    CK(cudaMemsetAsync(h->d_red.p, 0, sizeof(double) * red_doubles(h), h->stream));
    if (h->use_band) {
        ws::k_build_band<<<h->band_n_chunk, ws::kThreadsWs, sizeof(ws::Smem), h->stream>>>(h->batch, h->band);
        bd::k_band_gather<<<(h->band_n_seg + 6) / 7, 252, 0, h->stream>>>(h->batch, h->band, h->band_key, h->band_val, h->band_seg, h->band_n_seg);
        if (h->band_rest)
            lg::k_build_large_run<<<std::max(1, std::min(h->sm_count, (h->band_n_rest + 15) / 16)), lg::kThreadsL, sizeof(lg::RunSmem), h->stream>>>(
                h->batch, h->d_bd_rest.as<int4>(), h->band_n_rest);
        h->launches += h->band_rest ? 2 : 1;
    } else if (h->use_run)
        lg::k_build_large_run<<<std::max(1, std::min(h->sm_count, (h->tot_point + 63) / 64)), lg::kThreadsL, sizeof(lg::RunSmem), h->stream>>>(
            h->batch, h->d_lm_rec.as<int4>(), h->tot_point);
    else
        lg::k_build_large<false><<<h->grid_build_l, lg::kThreadsL, sizeof(lg::BuildSmemL), h->stream>>>(h->batch);
    if (h->tot_link > 0) {   // odometry links: every rank evaluates them (identical poses), ONE rank adds them to the sums
        k_link_lin<<<(h->tot_link + 63) / 64, 64, 0, h->stream>>>(h->batch);
        if (!h->partitioned || h->comm_rank == 0) lg::k_link_add_large<<<(h->tot_link + 63) / 64, 64, 0, h->stream>>>(h->batch);
        h->launches += 2;
    }
    ev_end(h, ev);
    h->launches += 1;
    if (h->partitioned) {
        ev = ev_begin(h, EV_OTHER);
        const int st = allreduce(h, h->d_red.p, red_doubles(h), ncclFloat64, ncclSum);
        ev_end(h, ev);
        if (st) return st;
    }
    return VISFS_BA_OK;
}

#define DBG_SYNC(what)                                                                              \
    do {                                                                                            \
        if (getenv("VISFS_BA_SYNC_DEBUG")) {                                                        \
            cudaError_t e__ = cudaStreamSynchronize(h->stream);                                     \
            if (e__ != cudaSuccess) return h->cuda_fail(e__, what);                                 \
        }                                                                                           \
    } while (0)

// dn::k_dense_back as one cluster of dn::kBackCluster CTAs
int launch_dense_back(visfs_ba_handle *h, const dn::DenseMat &M) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)dn::kBackCluster);
    cfg.blockDim = dim3(dn::kBackThreads);
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)dn::kBackCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, dn::k_dense_back, h->batch, M));
    return VISFS_BA_OK;
}

int enqueue_rest_large(visfs_ba_handle *h) {
    int st;
    DBG_SYNC("before solve (build / allreduce)");
    int ev = ev_begin(h, EV_SOLVE);
    CK(cudaMemsetAsync(h->d_cnt.as<int>() + 2, 0, sizeof(int), h->stream));   // Cholesky failure flag
    if (h->win[0].solver == VISFS_BA_SOLVER_PCG) {
        const size_t F = (size_t)std::max(h->tot_pose, 1);
        CK(h->d_pcg.reserve(sizeof(double) * (5 * 6 * F + 36 * F + 2 * (size_t)h->pcg_grid + 8)));
        double *work = h->d_pcg.as<double>();
        void *args[] = {(void *)&h->batch, (void *)&work};
        CK(cudaLaunchCooperativeKernel((const void *)lg::k_solve_pcg, dim3((unsigned)h->pcg_grid), dim3(lg::kPcgThreads), args, 0, h->stream));
    } else if (h->use_dense) {
        // dense window: blocked Cholesky with the trailing updates on the FP64 tensor pipe (ba_dense.cuh)
        const dn::DenseMat &M = h->dense;
        CK(cudaMemsetAsync(M.D, 0, sizeof(double) * (size_t)M.NP * M.LD, h->stream));
        dn::k_dense_fill<<<std::max(1, std::min(M.n / 6, 4 * h->sm_count)), 256, 0, h->stream>>>(h->batch, M);
        void *args[] = {(void *)&h->batch, (void *)&h->dense};
        CK(cudaLaunchCooperativeKernel((const void *)dn::k_dense_chol, dim3((unsigned)h->dense_grid), dim3(dn::kThreadsD), args, h->dense_smem, h->stream));
        { const int stb = launch_dense_back(h, M); if (stb) return stb; }
        h->launches += 2;
    } else if (h->use_mf) {
        // long banded system: nested-dissection multifrontal Cholesky, one launch per tree level (ba_mf.cuh)
        const size_t smem_f = sizeof(double) * ((size_t)mf::kMaxRows * dn::kTP + (size_t)mf::kMaxRows * 6);
        const size_t smem_b = sizeof(double) * ((size_t)dn::kNB * dn::kTP + 2 * dn::kNB + 6 * (2 * mf::kMaxBand + mf::kMaxArrow));
        const int nl = (int)h->mf_level_off.size() - 1;
        for (int l = 0; l < nl; ++l)
            mf::k_mf_factor<<<h->mf_level_off[l + 1] - h->mf_level_off[l], dn::kThreadsD, smem_f, h->stream>>>(h->batch, h->mf_plan, h->mf_level_off[l], l);
        for (int l = nl - 1; l >= 0; --l)
            mf::k_mf_back<<<h->mf_level_off[l + 1] - h->mf_level_off[l], dn::kThreadsD, smem_b, h->stream>>>(h->batch, h->mf_plan, h->mf_level_off[l]);
        dn::DenseMat M{};
        M.n = 6 * h->st_F_hint; M.x = h->mf_plan.x; M.flag = h->mf_plan.flag;
        { const int stb = launch_dense_back(h, M); if (stb) return stb; }
        h->launches += 2 * nl;
    } else if (h->use_front) {
        lg::k_solve_front<<<1, lg::kSolveThreadsL, sizeof(lg::FrontSmem), h->stream>>>(h->batch, h->front_plan);
    } else if (h->max_front > 32 && h->coop_grid > 1 && !getenv("VISFS_BA_NO_COOP")) {
        // wide fronts (dense windows): the trailing update of a column is spread over the whole GPU
        int *flag = h->d_cnt.as<int>() + 2;
        void *args[] = {(void *)&h->batch, (void *)&flag};
        CK(cudaLaunchCooperativeKernel((const void *)lg::k_solve_large<true>, dim3((unsigned)h->coop_grid), dim3(lg::kSolveThreadsL),
                                       args, 0, h->stream));
    } else {
        lg::k_solve_large<false><<<1, lg::kSolveThreadsL, 0, h->stream>>>(h->batch, h->d_cnt.as<int>() + 2);
    }
    if (h->tot_link > 0) { lg::k_link_chi_large<<<1, 128, 0, h->stream>>>(h->batch); h->launches += 1; }
    ev_end(h, ev);
    DBG_SYNC("solve kernel");
    ev = ev_begin(h, EV_UPDATE);
    if (h->use_band && !getenv("VISFS_BA_NO_BAND_UPDATE")) {
        if (h->band_rest) lg::k_update_large<<<h->grid_update_l, lg::kThreadsL, 0, h->stream>>>(h->batch, 1);
        bd::k_update_band<<<h->band_n_chunk, kUpdThreads, 0, h->stream>>>(h->batch, h->band, h->d_bd_part2.as<double>());
        ev_end(h, ev);
        ev = ev_begin(h, EV_OTHER);
        bd::k_band_fold<<<1, 256, 0, h->stream>>>(h->batch, h->batch.part2, h->band_rest ? h->grid_update_l : 0, h->d_bd_part2.as<double>(),
                                                 h->band_n_chunk, h->d_scal.as<double>() + 2);
        h->launches += h->band_rest ? 1 : 0;
    } else {
        lg::k_update_large<<<h->grid_update_l, lg::kThreadsL, 0, h->stream>>>(h->batch, 0);
        ev_end(h, ev);
        ev = ev_begin(h, EV_OTHER);
        lg::k_fold_part2<<<1, 256, 0, h->stream>>>(h->batch, h->grid_update_l, 0, h->d_scal.as<double>() + 2);
    }
    if ((st = allreduce(h, h->d_scal.as<double>() + 2, 2, ncclFloat64, ncclSum))) return st;
    k_control<<<1, 32, 0, h->stream>>>(h->batch_ctl);
    ev_end(h, ev);
    DBG_SYNC("update / fold / control");
    h->launches += 5;
    return VISFS_BA_OK;
}

int enqueue_body_large(visfs_ba_handle *h) {
    const int st = enqueue_build_large(h);
    return st ? st : enqueue_rest_large(h);
}

int init_pass_large(visfs_ba_handle *h) {
    cudaStream_t s = h->stream;
    int st = run_structure_large(h);
    if (st) return st;
    k_sync_buffers<<<dim3((unsigned)h->grid_lm_x, 1u), 256, 0, s>>>(h->batch);
    CK(cudaMemsetAsync(h->d_hdiag.p, 0, sizeof(double) * 6 * std::max(h->tot_pose, 1), s));
    lg::k_build_large<true><<<h->grid_build_l, lg::kThreadsL, sizeof(lg::BuildSmemL), s>>>(h->batch);
    lg::k_fold_part2<<<1, 256, 0, s>>>(h->batch, h->grid_build_l, 1, h->d_scal.as<double>());
    if (h->tot_link > 0) {
        k_link_lin<<<(h->tot_link + 63) / 64, 64, 0, s>>>(h->batch);
        if (!h->partitioned || h->comm_rank == 0) lg::k_link_init_large<<<1, 128, 0, s>>>(h->batch, h->d_scal.as<double>());
        h->launches += 2;
    }
    if ((st = allreduce(h, h->d_hdiag.p, 6 * (size_t)h->tot_pose, ncclFloat64, ncclSum))) return st;
    if ((st = allreduce(h, h->d_scal.as<double>(), 1, ncclFloat64, ncclSum))) return st;
    if ((st = allreduce(h, h->d_scal.as<double>() + 1, 1, ncclFloat64, ncclMax))) return st;
    lg::k_control_init_large<<<1, 256, 0, s>>>(h->batch, h->d_scal.as<double>());
    h->launches += 4;
    CK(cudaGetLastError());
    return VISFS_BA_OK;
}

int run_pass(visfs_ba_handle *h, int pass) {
    cudaStream_t s = h->stream;
    Batch &B = h->batch;
    const int gw = (h->n_win + 127) / 128;
    int ev = ev_begin(h, EV_OTHER);
    k_begin_pass<<<gw, 128, 0, s>>>(B, pass);
    int st;
    if (h->large) {
        if ((st = init_pass_large(h))) return st;
    } else {
        if ((st = run_structure(h))) return st;
        k_sync_buffers<<<dim3((unsigned)h->grid_lm_x, (unsigned)h->n_win), 256, 0, s>>>(B);
        LAUNCH_CHECK("structure kernels");
        if (h->n_chunks) k_init<<<h->n_chunks, kUpdThreads, sizeof(InitSmem), s>>>(B);
        LAUNCH_CHECK("k_init");
        if (h->tot_link > 0) { k_link_lin<<<(h->tot_link + 63) / 64, 64, 0, s>>>(B); h->launches += 1; }
        k_control_init<<<h->n_win, kCtlInitThreads, 0, s>>>(B);
        LAUNCH_CHECK("k_control_init");
        h->launches += 3;
    }
    ev_end(h, ev);
    CK(cudaGetLastError());
    int *running = h->h_small.as<int>();
    int bodies = 0;
    const int cap = 10 * std::max(h->max_iter, 1) + 2;
    int burst = h->max_iter;
    while (burst > 0 && bodies < cap) {
        for (int k = 0; k < burst; ++k)
            if ((st = h->large ? enqueue_body_large(h) : enqueue_body(h))) return st;
        bodies += burst;
        CK(cudaMemcpyAsync(running, h->d_n_running.p, sizeof(int), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if (*running <= 0) break;
        burst = 2;  // some window rejected a step: keep going two trials at a time
    }
    ev = ev_begin(h, EV_OTHER);
    k_end_pass<<<gw, 128, 0, s>>>(B, pass);
    if (pass == 0 && h->tot_edge > 0) k_cull<<<dim3((unsigned)((h->grid_edge_x + 3) / 4), (unsigned)h->n_win), 256, 0, s>>>(B);   // (four edges per thread)
    ev_end(h, ev);
    h->launches += (pass == 0 && h->tot_edge > 0) ? 2 : 1;
    CK(cudaGetLastError());
    return VISFS_BA_OK;
}

int run_resident(visfs_ba_handle *h) {
    if (!h->resident) return h->fail(VISFS_BA_ERR_INVALID, "no batch uploaded");
    CK(cudaSetDevice(h->device));
    CK(h->h_small.reserve(128));
    h->ev_next = 0;
    for (auto &v : h->ev_used) v.clear();
    h->launches = 0;
    h->d2h_bytes = 0;
    CK(cudaEventRecord(h->ev_t0, h->stream));
    int st = reset_state(h);
    if (st) return st;
    for (int pass = 0; pass < 2; ++pass) {
        st = run_pass(h, pass);
        if (st) return st;
    }
    CK(cudaEventRecord(h->ev_t1, h->stream));
    // state back to the host (small), also synchronises the run
    h->st_host.resize(h->n_win);
    CK(cudaMemcpyAsync(h->st_host.data(), h->d_st.p, sizeof(LMState) * h->n_win, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->has_run = true;
    if (h->use_mf && h->mf_plan.prof) {
        std::vector<long long> pr(2600);
        cudaMemcpy(pr.data(), h->mf_plan.prof, sizeof(long long) * pr.size(), cudaMemcpyDeviceToHost);
        const int nl = (int)h->mf_level_off.size() - 1;
        for (int l = 0; l < nl; ++l) {
            const long long *q = pr.data() + 8 * l;
            fprintf(stderr, "[visfs_ba] mf level %d (%d fronts, CTA 0): tables %.1f  gather assembly %.1f  fence %.1f  panel %.1f  corner update %.1f  rest %.1f  total %.1f us\n", l,
                    h->mf_level_off[l + 1] - h->mf_level_off[l], (q[1] - q[0]) * 1e-3, (q[2] - q[1]) * 1e-3, (q[3] - q[2]) * 1e-3, (q[4] - q[3]) * 1e-3,
                    (q[5] - q[4]) * 1e-3, (q[6] - q[5]) * 1e-3, (q[6] - q[0]) * 1e-3);
        }
    }
    if (h->use_dense && h->dense.prof) {
        const int np = std::min(512, (h->dense.n + dn::kNB - 1) / dn::kNB);
        std::vector<long long> pr(2600);
        cudaMemcpy(pr.data(), h->dense.prof, sizeof(long long) * pr.size(), cudaMemcpyDeviceToHost);
        double a = 0, s1 = 0, b = 0, s2 = 0;
        for (int k = 0; k + 1 < np; ++k) { a += pr[5 * k + 1] - pr[5 * k]; s1 += pr[5 * k + 2] - pr[5 * k + 1]; b += pr[5 * k + 3] - pr[5 * k + 2]; s2 += pr[5 * k + 4] - pr[5 * k + 3]; }
        fprintf(stderr, "[visfs_ba] k_dense_chol (last launch, CTA 0, %d panels): phase A %.1f us, sync %.1f us, phase B %.1f us, sync %.1f us, total %.1f us\n",
                np, a * 1e-3, s1 * 1e-3, b * 1e-3, s2 * 1e-3, (pr[5 * (np - 1) + 1] - pr[0]) * 1e-3);
        fprintf(stderr, "[visfs_ba]   back-substitution (inside k_dense_chol): %.1f us\n", pr[2530] * 1e-3);
    }

    visfs_ba_timing &t = h->timing;
    t = visfs_ba_timing{};
    t.kernel_launches = h->launches;
    t.h2d_bytes = h->h2d_bytes;
    h->d2h_bytes += (int64_t)sizeof(LMState) * h->n_win;
    t.d2h_bytes = h->d2h_bytes;
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev_t0, h->ev_t1);
    t.total_ms = ms;
    double *cls_ms[EV_CLASSES] = {&t.build_ms, &t.solve_ms, &t.update_ms, &t.other_ms};
    int64_t *cls_n[EV_CLASSES] = {&t.build_launches, &t.solve_launches, &t.update_launches, &t.other_launches};
    for (int c = 0; c < EV_CLASSES; ++c)
        for (auto &ab : h->ev_used[c]) {
            cudaEventElapsedTime(&ms, h->ev_pool[ab.first], h->ev_pool[ab.second]);
            *cls_ms[c] += ms;
            *cls_n[c] += 1;
        }
    for (int k = 0; k < 6; ++k) t.solve_clocks[k] = h->st_host[0].t_solve[k];
    for (int w = 0; w < h->n_win; ++w) {
        const LMState &s = h->st_host[w];
        const WinDesc &d = h->win[w];
        const int64_t trials = (int64_t)s.trials_run[0] + s.trials_run[1];
        t.lm_iterations += (int64_t)s.iterations_run[0] + s.iterations_run[1];
        t.lm_trials += trials;
        // edges a trial linearises: all of the window in pass 1, those that survived the cull (Optimizer.cpp:283-297) in pass 2
        const int64_t edge_trials = (int64_t)s.trials_run[0] * d.n_edge + (int64_t)s.trials_run[1] * std::max(0, d.n_edge - s.n_outliers);
        t.edge_trials += edge_trials;
        // build reads every active edge record (32 B), every landmark (24 B) and every pose (56 B) once;
        // update re-reads them and writes every landmark (24 B)
        t.alg_bytes_build += 32LL * edge_trials + trials * (24LL * d.n_point + 56LL * d.n_pose);
        t.alg_bytes_update += 32LL * edge_trials + trials * (48LL * d.n_point + 56LL * d.n_pose);
    }
    return VISFS_BA_OK;
}

int download(visfs_ba_handle *h, int n, visfs_ba_result *res) {
    if (!h->resident || !h->has_run) return h->fail(VISFS_BA_ERR_INVALID, "nothing to download");
    if (n != h->n_win || !res) return h->fail(VISFS_BA_ERR_INVALID, "result count does not match the uploaded batch");
    CK(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    const int items = std::max(h->max_pose * 7, std::max(h->grid_lm_x * 256 * 3, h->grid_edge_x * 256));
    const size_t P = (size_t)h->tot_pose, L = (size_t)h->tot_point, E = (size_t)h->tot_edge;
    const size_t o_pose = 0, o_point = sizeof(double) * 7 * P, o_level = o_point + sizeof(double) * 3 * L, o_end = o_level + E;
    CK(h->d_out_pose.reserve(o_end + 8));      // one output buffer [pose | point | level], one D2H copy
    char *dout = h->d_out_pose.as<char>();
    k_export<<<grid2(items, h->n_win), 256, 0, s>>>(h->batch, reinterpret_cast<double *>(dout + o_pose), reinterpret_cast<double *>(dout + o_point),
                                                   reinterpret_cast<uint8_t *>(dout + o_level));
    CK(cudaGetLastError());
    CK(h->h_out.reserve(o_end + 8));
    char *ho = h->h_out.as<char>();
    // results go straight into page-locked caller arrays when every window offers them for its landmarks and edge levels
    bool direct = !getenv("VISFS_BA_NO_DIRECT") && n > 0 && (!h->is_sub || h->direct_groups);
    for (int w = 0; w < n && direct; ++w) {
        const WinDesc &d = h->win[w];
        const visfs_ba_result &r = res[w];
        if (d.n_point && r.point_xyz && !host_is_pinned(r.point_xyz)) direct = false;
        if (d.n_edge && r.edge_level && !host_is_pinned(r.edge_level)) direct = false;
    }
    if (direct) {
        if (o_point) CK(cudaMemcpyAsync(ho, dout, o_point, cudaMemcpyDeviceToHost, s));   // poses: small, staged
        for (int w = 0; w < n; ++w) {
            const WinDesc &d = h->win[w];
            const visfs_ba_result &r = res[w];
            if (r.point_xyz && d.n_point)
                CK(cudaMemcpyAsync(r.point_xyz, dout + o_point + sizeof(double) * 3 * d.point_off, sizeof(double) * 3 * d.n_point, cudaMemcpyDeviceToHost, s));
            if (r.edge_level && d.n_edge) CK(cudaMemcpyAsync(r.edge_level, dout + o_level + d.edge_off, d.n_edge, cudaMemcpyDeviceToHost, s));
        }
    } else if (o_end) {
        CK(cudaMemcpyAsync(ho, dout, o_end, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaStreamSynchronize(s));
    h->d2h_bytes += (int64_t)o_end;
    h->timing.d2h_bytes = h->d2h_bytes;
    h->timing.kernel_launches = h->launches + 1;
    for (int w = 0; w < n; ++w) {
        const WinDesc &d = h->win[w];
        const LMState &st = h->st_host[w];
        visfs_ba_result &r = res[w];
        if (r.pose_tq && d.n_pose) memcpy(r.pose_tq, ho + o_pose + sizeof(double) * 7 * d.pose_off, sizeof(double) * 7 * d.n_pose);
        if (!direct) {
            if (r.point_xyz && d.n_point) memcpy(r.point_xyz, ho + o_point + sizeof(double) * 3 * d.point_off, sizeof(double) * 3 * d.n_point);
            if (r.edge_level && d.n_edge) memcpy(r.edge_level, ho + o_level + d.edge_off, d.n_edge);
        }
        r.status = st.status; r.n_outliers = st.n_outliers;
        for (int k = 0; k < 2; ++k) {
            r.iterations_run[k] = st.iterations_run[k]; r.trials_run[k] = st.trials_run[k]; r.stop_reason[k] = st.stop[k];
            r.n_free_poses[k] = st.nF[k]; r.n_free_points[k] = st.nNL[k]; r.lambda_final[k] = st.lambda_final[k];
        }
        if (st.stop[1] == VISFS_BA_STOP_NOT_RUN) { r.n_free_poses[1] = r.n_free_points[1] = 0; r.lambda_final[1] = 0; }
        r.chi2_initial = st.chi_initial; r.chi2_pass1 = st.chi_pass[0]; r.chi2_final = st.chi_pass[1];
        r.chi2_last_trial = st.chi_last_trial;
        if (st.status == VISFS_BA_ERR_UNSUPPORTED || st.status == VISFS_BA_ERR_INVALID || st.status == VISFS_BA_ERR_CUDA) {
            char buf[200];   // failures found on the device (k_struct_lm's degree check, ...): leave text for visfs_ba_last_error
            snprintf(buf, sizeof buf, "problem %d: rejected on the device with status %d (%s)", h->idx_base + w, st.status,
                     st.status == VISFS_BA_ERR_UNSUPPORTED ? "a landmark is observed by more poses than the kernels hold: 192 per small window, 32 per large window"
                                                           : "device-side check");
            h->error = buf;
        }
    }
    return VISFS_BA_OK;
}

}  // namespace


// ================================================================================================
// resident local map (visfs_ba_window_*, ba_window.cuh)
// ================================================================================================
struct visfs_ba_window {
    visfs_ba_handle *h = nullptr;
    visfs_ba_window_config cfg{};
    // host mirror of the structure (ids, slots, which observation lives where); the numbers live on the device only
    std::map<int64_t, int> frame_slot;                 // signature id -> slot, ascending ids = pose order
    std::vector<int> free_frames;
    std::unordered_map<int64_t, int> point_slot;
    std::map<int64_t, int> point_order;                // ascending feature id -> slot (std::map order of Optimizer.cpp:156)
    std::vector<int> free_points;
    bool order_dirty = true;
    // the observation pool lives on the device only (append-only, tombstones); the host knows how many entries it holds
    int n_pool = 0;
    std::vector<int> stamp;                            // per feature slot: the insert_frame call that last saw it (duplicate check)
    int stamp_now = 0;
    std::vector<int64_t> point_id_of_slot, frame_id_of_slot;
    struct Link { int64_t from, to; double tq[7]; };
    std::vector<Link> links;                           // odometry links by frame id (visfs_ba_window_set_links)
    double odometry_variance = 0.0;
    int64_t h2d_total = 0;
    // device
    DevBuf d_frame_tq, d_frame_pose, d_pose_slot, d_point_xyz, d_point_fixed, d_point_id, d_order, d_ob_point, d_ob_frame, d_ob_obs,
        d_ob_kind, d_ob_dead, d_cnt, d_act, d_rank, d_rank_of_slot, d_slot_of_rank, d_key, d_key2, d_val, d_val2, d_counters, d_scan_tmp,
        d_sort_tmp, d_pose_out, d_outliers, d_list, d_mask, d_compact;
    PinBuf h_small, h_stage;
    DevBuf d_stage;
    cudaEvent_t stage_done = nullptr;              // the last asynchronous copy out of h_stage
    // page-locked staging for a delta: waits until the previous delta has left, returns the buffer
    char *stage(size_t bytes) {
        if (stage_done) cudaEventSynchronize(stage_done);
        if (h_stage.reserve(bytes) != cudaSuccess || d_stage.reserve(bytes) != cudaSuccess) return nullptr;
        return h_stage.as<char>();
    }
};

namespace {

int win_fail(visfs_ba_window *w, int st, const std::string &msg) { return w->h->fail(st, "window: " + msg); }

// the device pool is append-only; when it is full the live observations are packed to its front on the device
int win_compact(visfs_ba_window *w) {
    visfs_ba_handle *h = w->h;
    cudaStream_t s = h->stream;
    const int n = w->n_pool;
    if (n == 0) return VISFS_BA_OK;
    const size_t N = (size_t)n;
    // scratch: live (n + 1) | pos (n + 1) | point | frame | obs (3 n floats) | kind
    const size_t o_live = 0, o_pos = o_live + 4 * (N + 1), o_p = o_pos + 4 * (N + 1), o_f = o_p + 4 * N, o_o = o_f + 4 * N, o_k = o_o + 12 * N;
    CK(w->d_compact.reserve(o_k + N + 16));
    char *base = w->d_compact.as<char>();
    int *live = reinterpret_cast<int *>(base + o_live), *pos = reinterpret_cast<int *>(base + o_pos);
    const int g = std::max(1, std::min((n + 256) / 256, 1024));
    wn::k_win_live<<<g, 256, 0, s>>>(w->d_ob_dead.as<uint8_t>(), n, live);
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, live, pos, n + 1, s);
    CK(w->d_scan_tmp.reserve(tb));
    CK(cub::DeviceScan::ExclusiveSum(w->d_scan_tmp.p, tb, live, pos, n + 1, s));
    wn::k_win_compact<<<g, 256, 0, s>>>(live, pos, n, w->d_ob_point.as<int>(), w->d_ob_frame.as<int>(), w->d_ob_obs.as<float>(), w->d_ob_kind.as<uint8_t>(),
                                        reinterpret_cast<int *>(base + o_p), reinterpret_cast<int *>(base + o_f), reinterpret_cast<float *>(base + o_o),
                                        reinterpret_cast<uint8_t *>(base + o_k));
    int *cnt_h = w->h_small.as<int>() + 2 * w->cfg.max_frames;
    CK(cudaMemcpyAsync(cnt_h, pos + n, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const size_t m = (size_t)cnt_h[0];
    if (m) {
        CK(cudaMemcpyAsync(w->d_ob_point.p, base + o_p, 4 * m, cudaMemcpyDeviceToDevice, s));
        CK(cudaMemcpyAsync(w->d_ob_frame.p, base + o_f, 4 * m, cudaMemcpyDeviceToDevice, s));
        CK(cudaMemcpyAsync(w->d_ob_obs.p, base + o_o, 12 * m, cudaMemcpyDeviceToDevice, s));
        CK(cudaMemcpyAsync(w->d_ob_kind.p, base + o_k, m, cudaMemcpyDeviceToDevice, s));
    }
    CK(cudaMemsetAsync(w->d_ob_dead.p, 0, (size_t)w->cfg.max_observations, s));
    w->n_pool = (int)m;
    return VISFS_BA_OK;
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

int visfs_ba_abi_version(void) { return VISFS_BA_ABI_VERSION; }

const char *visfs_ba_last_error(const visfs_ba_handle *h) { return h ? h->error.c_str() : g_create_error.c_str(); }

int visfs_ba_create(const visfs_ba_config *cfg, visfs_ba_handle **out) {
    if (!out) return VISFS_BA_ERR_INVALID;
    *out = nullptr;
    if (cfg && cfg->abi_version != VISFS_BA_ABI_VERSION) { g_create_error = "ABI version mismatch"; return VISFS_BA_ERR_INVALID; }
    const int dev = cfg ? cfg->device : 0;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_create_error = std::string("no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e);
        return VISFS_BA_ERR_CUDA;
    }
    if (dev < 0 || dev >= count) { g_create_error = "device ordinal out of range"; return VISFS_BA_ERR_INVALID; }
    if ((e = cudaSetDevice(dev)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); return VISFS_BA_ERR_CUDA; }
    visfs_ba_handle *h = new visfs_ba_handle();
    h->device = dev;
    h->profile = cfg && cfg->profile_kernels != 0;
    cudaDeviceProp prop{};
    cudaGetDeviceProperties(&prop, dev);
    h->sm_count = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : 148;
    if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreate(&h->ev_t0)) != cudaSuccess || (e = cudaEventCreate(&h->ev_t1)) != cudaSuccess) {
        g_create_error = cudaGetErrorString(e);
        delete h;
        return VISFS_BA_ERR_CUDA;
    }
    const int smem_build = (int)sizeof(BuildSmemT<MODE_BUILD>), smem_update = (int)sizeof(UpdateSmem);
    cudaFuncSetAttribute(k_build<MODE_BUILD, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_build);
    cudaFuncSetAttribute(k_build<MODE_BUILD, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_build);
    cudaFuncSetAttribute(k_update, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_update);
    e = cudaFuncSetAttribute(ws::k_build_ws, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ws::Smem));
    cudaFuncSetAttribute(ds::k_build_ds, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ds::Smem));
    cudaFuncSetAttribute(ws::k_build_band, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ws::Smem));
    {   // how many clusters of 8 k_build_ws CTAs the device keeps resident at once (B200: 15): the single-window decomposition
        // uses clusters of 8 when one wave of them covers the window (k_solve then reads half as many partial systems)
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(8 * 16); cfg.blockDim = dim3(ws::kThreadsWs); cfg.dynamicSmemBytes = sizeof(ws::Smem);
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 8; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int nc = 0;
        if (cudaOccupancyMaxActiveClusters(&nc, ws::k_build_ws, &cfg) != cudaSuccess) { nc = 0; cudaGetLastError(); }
        h->max_clusters8 = nc;
    }
    if (getenv("VISFS_BA_VERBOSE")) {
        cudaFuncAttributes fa{};
        cudaFuncGetAttributes(&fa, ws::k_build_ws);
        fprintf(stderr, "[visfs_ba] k_build_ws: set smem %zu -> %s; regs %d maxThreads %d static smem %zu maxDyn %d\n", sizeof(ws::Smem),
                cudaGetErrorString(e), fa.numRegs, fa.maxThreadsPerBlock, fa.sharedSizeBytes, fa.maxDynamicSharedSizeBytes);
        for (int cl = 1; cl <= 8; cl *= 2) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(8 * 16); cfg.blockDim = dim3(ws::kThreadsWs); cfg.dynamicSmemBytes = sizeof(ws::Smem);
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            int nc = -1;
            cudaError_t e2 = cudaOccupancyMaxActiveClusters(&nc, ws::k_build_ws, &cfg);
            fprintf(stderr, "[visfs_ba]   cluster %d: max active clusters %d (%s)\n", cl, nc, cudaGetErrorString(e2));
        }
        {
            cudaFuncAttributes fd{};
            cudaError_t e3 = cudaFuncSetAttribute(ds::k_build_ds, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ds::Smem));
            cudaFuncGetAttributes(&fd, ds::k_build_ds);
            int nb = -1;
            cudaError_t e4 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, ds::k_build_ds, ds::kThreadsDs, sizeof(ds::Smem));
            fprintf(stderr, "[visfs_ba] k_build_ds: set smem %zu -> %s; regs %d maxThreads %d static smem %zu maxDyn %d; blocks/SM %d (%s)\n", sizeof(ds::Smem),
                    cudaGetErrorString(e3), fd.numRegs, fd.maxThreadsPerBlock, fd.sharedSizeBytes, fd.maxDynamicSharedSizeBytes, nb, cudaGetErrorString(e4));
        }
        cudaFuncGetAttributes(&fa, k_solve);
        fprintf(stderr, "[visfs_ba] k_solve: regs %d maxThreads %d static smem %zu\n", fa.numRegs, fa.maxThreadsPerBlock, fa.sharedSizeBytes);
        cudaGetLastError();
    }
    {
        int per_sm = 0, coop = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lg::k_solve_large<true>, lg::kSolveThreadsL, 0);
        h->coop_grid = coop ? h->sm_count * std::max(per_sm, 0) : 0;
        if (per_sm > 1) h->coop_grid = h->sm_count;   // one CTA per SM is enough
        int per_sm_pcg = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_pcg, lg::k_solve_pcg, lg::kPcgThreads, 0);
        h->pcg_grid = (coop && per_sm_pcg > 0) ? h->sm_count : 0;
    }
    {
        int coop = 0, per_sm = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        cudaFuncSetAttribute(dn::k_dense_chol, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(dn::k_dense_back, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(mf::k_mf_factor, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dn::k_dense_chol, dn::kThreadsD, 200 * 1024);
        h->dense_grid = (coop && per_sm > 0) ? h->sm_count : 0;
    }
    cudaFuncSetAttribute(lg::k_solve_front, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(lg::FrontSmem));
    cudaFuncSetAttribute(lg::k_build_large<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(lg::BuildSmemL));
    cudaFuncSetAttribute(lg::k_build_large<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(lg::BuildSmemL));
    e = cudaFuncSetAttribute(k_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_solve2, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) {
        g_create_error = std::string("kernel image not usable on this device (built for sm_100a): ") + cudaGetErrorString(e);
        cudaStreamDestroy(h->stream);
        delete h;
        return VISFS_BA_ERR_CUDA;
    }
    *out = h;
    return VISFS_BA_OK;
}

void visfs_ba_destroy(visfs_ba_handle *h) {
    if (!h) return;
    for (visfs_ba_handle *s : h->subs) visfs_ba_destroy(s);
    h->subs.clear();
    visfs_ba_comm_destroy(h);
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    {
        DevBuf *lb[] = {&h->d_sky_first, &h->d_sky_off, &h->d_col_ptr, &h->d_col_cnt, &h->d_col_rows, &h->d_red, &h->d_hdiag,
                        &h->d_scal, &h->d_info, &h->d_cnt, &h->d_plan, &h->d_pcg, &h->d_lm_key, &h->d_lm_key2, &h->d_lm_idx,
                        &h->d_lm_order, &h->d_sort_tmp, &h->d_lm_rec, &h->d_dense, &h->d_dense_prof, &h->d_mf_meta, &h->d_mf_fronts, &h->d_mf_touch};
        for (DevBuf *b : lb) b->release();
    }
    DevBuf *bufs[] = {&h->d_in, &h->d_st, &h->d_pose, &h->d_point, &h->d_pose_flags, &h->d_lm_flags, &h->d_pose_hidx,
                      &h->d_pose_active, &h->d_point_hidx, &h->d_lm_edge_off, &h->d_obs_u, &h->d_obs_v, &h->d_obs_r, &h->d_edge_pose,
                      &h->d_edge_point, &h->d_edge_orig, &h->d_covis, &h->d_part, &h->d_part2, &h->d_xp, &h->d_n_running, &h->d_ctl_count,
                      &h->d_out_pose, &h->d_out_point, &h->d_out_level, &h->d_tmp, &h->d_tmp2, &h->d_keys, &h->d_keys2, &h->d_perm, &h->d_tiles, &h->d_tile_off, &h->d_tile_cnt, &h->d_wtiles, &h->d_wtile_off,
                      &h->d_link_win, &h->d_link_from, &h->d_link_to, &h->d_link_m, &h->d_link_lin};
    for (DevBuf *b : bufs) b->release();
    h->h_stage.release(); h->h_out.release(); h->h_small.release();
    for (cudaEvent_t ev : h->ev_pool) cudaEventDestroy(ev);
    if (h->ev_t0) cudaEventDestroy(h->ev_t0);
    if (h->ev_t1) cudaEventDestroy(h->ev_t1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int visfs_ba_upload(visfs_ba_handle *h, int32_t n, const visfs_ba_problem *problems) {
    if (!h) return VISFS_BA_ERR_INVALID;
    const int st = upload(h, n, problems);
    // copies that read the caller's page-locked arrays directly must have finished when this returns
    if (st == VISFS_BA_OK && h->direct_h2d > 0 && cudaStreamSynchronize(h->stream) != cudaSuccess)
        return h->fail(VISFS_BA_ERR_CUDA, "upload: copy from page-locked caller memory failed");
    return st;
}

int visfs_ba_run_resident(visfs_ba_handle *h) {
    if (!h) return VISFS_BA_ERR_INVALID;
    return run_resident(h);
}

int visfs_ba_download(visfs_ba_handle *h, int32_t n, visfs_ba_result *results) {
    if (!h) return VISFS_BA_ERR_INVALID;
    return download(h, n, results);
}

int visfs_ba_get_timing(const visfs_ba_handle *h, visfs_ba_timing *out) {
    if (!h || !out) return VISFS_BA_ERR_INVALID;
    *out = h->timing;
    return VISFS_BA_OK;
}

// How many pipeline groups a batch of n windows is cut into (1 = the plain single-stream path).
// How a batch is pipelined: the number of groups, and whether the groups DMA page-locked caller arrays directly.
// One host thread per group; several ranks of one job (torchrun exports LOCAL_WORLD_SIZE) share the host cores.
// Measured on C3 x 512 with page-locked inputs, ms per batch (staging route / direct route, best group count of each):
//   16 cores 27.9 (16 groups) / 28.7 (8)   8 cores 30.3 (8) / 28.8 (8)   4 cores 41.3 (8) / 30.1 (8)   2 cores 55.3 (8) / 37.3 (8)
// so with fewer than 16 cores for this rank the packing threads are the bottleneck and the direct route wins.
static int pick_groups(const visfs_ba_handle *h, int n, const visfs_ba_problem *problems, bool *direct) {
    *direct = false;
    if (h->is_sub || n < 32) return 1;
    for (int w = 0; w < n; ++w)
        if (problems[w].flags & VISFS_BA_FLAG_PARTITIONED) return 1;
    int share = 16;
    const int hw = (int)std::thread::hardware_concurrency();
    if (hw > 0) {
        share = hw;
        if (const char *lws = getenv("LOCAL_WORLD_SIZE")) share = std::max(1, hw / std::max(atoi(lws), 1));
    }
    auto obs_of = [](const visfs_ba_problem &p) -> const void * { return p.edge_obs_f32 ? (const void *)p.edge_obs_f32 : (const void *)p.edge_obs; };
    const bool pinned = problems[0].n_edges > 0 && problems[n - 1].n_edges > 0 && host_is_pinned(obs_of(problems[0])) &&
                        host_is_pinned(obs_of(problems[n - 1]));
    int g;
    if (getenv("VISFS_BA_DIRECT_GROUPS") || (pinned && share < 16 && !getenv("VISFS_BA_NO_DIRECT"))) {
        *direct = true;
        g = 8;
    } else {
        g = std::min(20, std::max(4, share + share / 4));   // (a few more threads than cores: 16 cores, C3 x 512: 16 groups 26.8 ms, 20 groups 26.1 ms)
    }
    if (const char *e = getenv("VISFS_BA_GROUPS")) g = atoi(e);
    g = std::min(g, n / 16);
    return std::max(g, 1);
}

int visfs_ba_solve_batch(visfs_ba_handle *h, int32_t n, const visfs_ba_problem *problems, visfs_ba_result *results) {
    if (!h) return VISFS_BA_ERR_INVALID;
    if (n <= 0 || !problems || !results) return h->fail(VISFS_BA_ERR_INVALID, "empty batch");
    bool direct_groups = false;
    const int groups = pick_groups(h, n, problems, &direct_groups);
    if (groups <= 1) {
        int st = upload(h, n, problems);
        if (st) return st;
        st = run_resident(h);
        if (st) return st;
        return download(h, n, results);
    }
    while ((int)h->subs.size() < groups) {
        visfs_ba_config cfg{};
        cfg.abi_version = VISFS_BA_ABI_VERSION; cfg.device = h->device; cfg.profile_kernels = 0;
        visfs_ba_handle *s = nullptr;
        const int st = visfs_ba_create(&cfg, &s);
        if (st != VISFS_BA_OK) return h->fail(st, std::string("pipeline sub-handle: ") + g_create_error);
        s->is_sub = true;
        h->subs.push_back(s);
    }
    for (visfs_ba_handle *s : h->subs) s->direct_groups = direct_groups;
    h->resident = false; h->has_run = false;
    std::vector<int> status(groups, VISFS_BA_OK);
    std::vector<std::thread> workers;
    const auto t_call = std::chrono::steady_clock::now();
    const bool trace = getenv("VISFS_BA_TRACE") != nullptr;
    // Group sizes grow geometrically (x 1.22 per group, over the first half of the groups).  Every group's host thread starts packing at once, so a small first
    // group puts the GPU to work early; more important, groups of unequal size do not run in lock-step: the latency-bound
    // phases of one (k_solve, control) overlap the throughput-bound ones of another (k_build_ws).  Measured on C3 x 512 with
    // 16 groups: 33.0 ms uniform -> 27.8 ms; with the packing removed entirely uniform groups still take 30.7 ms.
    std::vector<int> sizes(groups, 0);
    {
        double ramp = 1.22;
        if (const char *e = getenv("VISFS_BA_RAMP")) ramp = std::max(1.0, atof(e));
        std::vector<double> wgt(groups);
        double tot = 0.0;
        int flat = std::max(4, groups / 2);   // the ramp stops half way: the large half is equal, so no single group is left running alone at the end
        if (const char *e = getenv("VISFS_BA_RAMP_FLAT")) flat = std::max(1, atoi(e));
        for (int g = 0; g < groups; ++g) { wgt[g] = std::pow(ramp, std::min(g, flat)); tot += wgt[g]; }
        int given = 0;
        for (int g = 0; g < groups; ++g) { sizes[g] = std::max(1, (int)std::floor(n * wgt[g] / tot)); given += sizes[g]; }
        for (int g = groups - 1; given != n; g = (g + groups - 1) % groups) {   // hand the rounding remainder to the large end
            if (given < n) { ++sizes[g]; ++given; }
            else if (sizes[g] > 1) { --sizes[g]; --given; }
        }
    }
    int off = 0;
    for (int g = 0; g < groups; ++g) {
        const int cnt = sizes[g];
        visfs_ba_handle *s = h->subs[g];
        s->idx_base = off;
        workers.emplace_back([s, cnt, off, problems, results, &status, g, t_call, trace]() {
            auto ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call).count(); };
            const double t0 = ms();
            int st = upload(s, cnt, problems + off);
            const double t1 = ms();
            if (trace) { cudaStreamSynchronize(s->stream); }
            const double t1b = ms();
            if (!st) st = run_resident(s);
            const double t2 = ms();
            if (!st) st = download(s, cnt, results + off);
            const double t3 = ms();
            if (trace) fprintf(stderr, "[visfs_ba] group %2d: start %.2f upload-issued %.2f h2d-done %.2f run-done %.2f download-done %.2f ms\n", g, t0, t1, t1b, t2, t3);
            status[g] = st;
        });
        off += cnt;
    }
    for (auto &t : workers) t.join();
    visfs_ba_timing &t = h->timing;
    t = visfs_ba_timing{};
    for (int g = 0; g < groups; ++g) {
        const visfs_ba_timing &u = h->subs[g]->timing;
        t.total_ms = std::max(t.total_ms, u.total_ms);
        t.lm_iterations += u.lm_iterations; t.lm_trials += u.lm_trials; t.edge_trials += u.edge_trials;
        t.alg_bytes_build += u.alg_bytes_build; t.alg_bytes_update += u.alg_bytes_update;
        t.kernel_launches += u.kernel_launches; t.h2d_bytes += u.h2d_bytes; t.d2h_bytes += u.d2h_bytes;
    }
    for (int g = 0; g < groups; ++g)
        if (status[g]) return h->fail(status[g], h->subs[g]->error);
    return VISFS_BA_OK;
}

int visfs_ba_solve(visfs_ba_handle *h, const visfs_ba_problem *problem, visfs_ba_result *result) {
    if (!h || !problem || !result) return VISFS_BA_ERR_INVALID;
    const int st = visfs_ba_solve_batch(h, 1, problem, result);
    if (st) return st;
    return result->status;
}

int visfs_ba_linearize(visfs_ba_handle *h, const visfs_ba_problem *problem, visfs_ba_linearization *out) {
    if (!h || !problem || !out) return VISFS_BA_ERR_INVALID;
    int st = upload(h, 1, problem);
    if (st) return st;
    st = reset_state(h);
    if (st) return st;
    const size_t E = (size_t)std::max(h->tot_edge, 1);
    CK(h->d_tmp.reserve(sizeof(double) * 33 * E));
    double *base = h->d_tmp.as<double>();
    double *d_err = base, *d_chi = base + 3 * E, *d_rho = base + 4 * E, *d_w = base + 5 * E, *d_jl = base + 6 * E, *d_jp = base + 15 * E;
    k_linearize_debug<<<grid2(h->tot_edge, 1), 256, 0, h->stream>>>(h->batch, d_err, d_chi, d_rho, d_w, d_jl, d_jp);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    const size_t n = (size_t)h->tot_edge;
    if (out->error) CK(cudaMemcpy(out->error, d_err, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost));
    if (out->chi2) CK(cudaMemcpy(out->chi2, d_chi, sizeof(double) * n, cudaMemcpyDeviceToHost));
    if (out->rho) CK(cudaMemcpy(out->rho, d_rho, sizeof(double) * n, cudaMemcpyDeviceToHost));
    if (out->weight) CK(cudaMemcpy(out->weight, d_w, sizeof(double) * n, cudaMemcpyDeviceToHost));
    if (out->J_point) CK(cudaMemcpy(out->J_point, d_jl, sizeof(double) * 9 * n, cudaMemcpyDeviceToHost));
    if (out->J_pose) CK(cudaMemcpy(out->J_pose, d_jp, sizeof(double) * 18 * n, cudaMemcpyDeviceToHost));
    return VISFS_BA_OK;
}

int visfs_ba_structure_build(visfs_ba_handle *h, const visfs_ba_problem *problem, visfs_ba_structure *out) {
    if (!h || !problem || !out) return VISFS_BA_ERR_INVALID;
    int st = upload(h, 1, problem);
    if (st) return st;
    st = reset_state(h);
    if (st) return st;
    cudaStream_t s = h->stream;
    const size_t P = (size_t)std::max(h->tot_pose, 1), L = (size_t)std::max(h->tot_point, 1), E = (size_t)std::max(h->tot_edge, 1);
    if (out->edge_level && h->tot_edge) {
        CK(cudaMemcpyAsync(h->d_out_level.p, out->edge_level, h->tot_edge, cudaMemcpyHostToDevice, s));
        k_mark_levels<<<grid2(h->tot_edge, 1), 256, 0, s>>>(h->batch, h->d_out_level.as<uint8_t>());
    }
    k_begin_pass<<<1, 128, 0, s>>>(h->batch, 0);
    CK(h->h_small.reserve(128));
    st = h->large ? run_structure_large(h) : run_structure(h, true);
    if (st) return st;
    // landmark hessian indices: exclusive scan of the in-Hessian flags
    const int cap = std::max(out->schur_capacity, 0);
    const size_t ints = L /*flags*/ + L /*scan*/ + L /*hidx*/ + 2 * E /*hpl*/ + 2 * (size_t)std::max(cap, 1) + 8;
    DevBuf scratch, cubtmp;
    CK(scratch.reserve(sizeof(int) * ints + E + 64));
    int *d_flag = scratch.as<int>(), *d_scan = d_flag + L, *d_hidx = d_scan + L, *d_row = d_hidx + L, *d_col = d_row + E,
        *d_srow = d_col + E, *d_scol = d_srow + std::max(cap, 1), *d_cnt = d_scol + std::max(cap, 1);
    uint8_t *d_act = reinterpret_cast<uint8_t *>(d_cnt + 8);
    std::vector<uint8_t> lmf(L);
    CK(cudaMemcpyAsync(lmf.data(), h->d_lm_flags.p, h->tot_point, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    std::vector<int> flags(L, 0);
    for (int l = 0; l < h->tot_point; ++l) flags[l] = (lmf[l] & kInHessian) ? 1 : 0;
    CK(cudaMemcpyAsync(d_flag, flags.data(), sizeof(int) * L, cudaMemcpyHostToDevice, s));
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_flag, d_scan, (int)L, s);
    CK(cubtmp.reserve(tmp_bytes));
    CK(cub::DeviceScan::ExclusiveSum(cubtmp.p, tmp_bytes, d_flag, d_scan, (int)L, s));
    k_structure_export<<<grid2(std::max(h->tot_point, h->tot_edge), 1), 256, 0, s>>>(h->batch, d_scan, d_hidx, d_act, d_row, d_col);
    std::vector<unsigned long long> lkeys;   // large path: unique (col,row) keys, sorted
    if (!h->large) {
        k_schur_pattern<<<1, 32, 0, s>>>(h->batch, 0, d_srow, d_scol, cap, d_cnt);
    } else {
        // pattern = diagonal + every pose pair of every landmark in the Hessian: emit keys, sort, unique
        DevBuf pc, po, kb, kb2, t2;
        CK(pc.reserve(sizeof(int) * (L + 1))); CK(po.reserve(sizeof(int) * (L + 1)));
        lg::k_pattern_count<<<grid2(h->tot_point, 1), 256, 0, s>>>(h->batch, pc.as<int>());
        CK(cudaMemsetAsync(pc.as<int>() + h->tot_point, 0, sizeof(int), s));
        size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, pc.as<int>(), po.as<int>(), h->tot_point + 1, s);
        CK(t2.reserve(tb));
        CK(cub::DeviceScan::ExclusiveSum(t2.p, tb, pc.as<int>(), po.as<int>(), h->tot_point + 1, s));
        int total = 0;
        CK(cudaMemcpyAsync(&total, po.as<int>() + h->tot_point, sizeof(int), cudaMemcpyDeviceToHost, s));
        LMState st0;
        CK(cudaMemcpyAsync(&st0, h->d_st.p, sizeof(LMState), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        const size_t nk = (size_t)total + (size_t)st0.F;
        CK(kb.reserve(sizeof(unsigned long long) * std::max<size_t>(nk, 1))); CK(kb2.reserve(sizeof(unsigned long long) * std::max<size_t>(nk, 1)));
        lg::k_pattern_emit<<<grid2(std::max(h->tot_point, st0.F), 1), 256, 0, s>>>(h->batch, po.as<int>(), kb.as<unsigned long long>(), st0.F);
        if (nk) {
            tb = 0;
            cub::DeviceRadixSort::SortKeys(nullptr, tb, kb.as<unsigned long long>(), kb2.as<unsigned long long>(), (int)nk, 0, 64, s);
            CK(t2.reserve(tb));
            CK(cub::DeviceRadixSort::SortKeys(t2.p, tb, kb.as<unsigned long long>(), kb2.as<unsigned long long>(), (int)nk, 0, 64, s));
        }
        lkeys.resize(nk);
        if (nk) CK(cudaMemcpyAsync(lkeys.data(), kb2.p, sizeof(unsigned long long) * nk, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if (problem->n_links > 0 && st0.F > 0) {   // pose-pose blocks of the odometry links
            std::vector<int> hidx((size_t)h->tot_pose);
            CK(cudaMemcpy(hidx.data(), h->d_pose_hidx.p, sizeof(int) * (size_t)h->tot_pose, cudaMemcpyDeviceToHost));
            for (int k = 0; k < problem->n_links; ++k) {
                const int hi = hidx[(size_t)problem->link_from[k]], hj = hidx[(size_t)problem->link_to[k]];
                if (hi >= 0 && hj >= 0)
                    lkeys.push_back((unsigned long long)std::max(hi, hj) * (unsigned long long)st0.F + (unsigned long long)std::min(hi, hj));
            }
            std::sort(lkeys.begin(), lkeys.end());
        }
        lkeys.erase(std::unique(lkeys.begin(), lkeys.end()), lkeys.end());
        const int nuniq = (int)lkeys.size();
        std::vector<int> hr(std::max(std::min(nuniq, cap), 1)), hc(std::max(std::min(nuniq, cap), 1));
        for (int i = 0; i < std::min(nuniq, cap); ++i) {
            hc[i] = (int)(lkeys[i] / (unsigned long long)std::max(st0.F, 1));
            hr[i] = (int)(lkeys[i] % (unsigned long long)std::max(st0.F, 1));
        }
        if (std::min(nuniq, cap) > 0) {
            CK(cudaMemcpyAsync(d_srow, hr.data(), sizeof(int) * std::min(nuniq, cap), cudaMemcpyHostToDevice, s));
            CK(cudaMemcpyAsync(d_scol, hc.data(), sizeof(int) * std::min(nuniq, cap), cudaMemcpyHostToDevice, s));
        }
        CK(cudaMemcpyAsync(d_cnt, &nuniq, sizeof(int), cudaMemcpyHostToDevice, s));
        CK(cudaStreamSynchronize(s));
        }
    CK(cudaGetLastError());
    std::vector<LMState> sth(1);
    CK(cudaMemcpyAsync(sth.data(), h->d_st.p, sizeof(LMState), cudaMemcpyDeviceToHost, s));
    int cnt = 0;
    CK(cudaMemcpyAsync(&cnt, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, s));
    if (out->pose_hidx && h->tot_pose) CK(cudaMemcpyAsync(out->pose_hidx, h->d_pose_hidx.p, sizeof(int) * h->tot_pose, cudaMemcpyDeviceToHost, s));
    if (out->point_hidx && h->tot_point) CK(cudaMemcpyAsync(out->point_hidx, d_hidx, sizeof(int) * h->tot_point, cudaMemcpyDeviceToHost, s));
    if (out->edge_active && h->tot_edge) CK(cudaMemcpyAsync(out->edge_active, d_act, h->tot_edge, cudaMemcpyDeviceToHost, s));
    if (out->hpl_row && h->tot_edge) CK(cudaMemcpyAsync(out->hpl_row, d_row, sizeof(int) * h->tot_edge, cudaMemcpyDeviceToHost, s));
    if (out->hpl_col && h->tot_edge) CK(cudaMemcpyAsync(out->hpl_col, d_col, sizeof(int) * h->tot_edge, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const int ncopy = std::min(cnt, cap);
    if (out->schur_rows && ncopy) CK(cudaMemcpy(out->schur_rows, d_srow, sizeof(int) * ncopy, cudaMemcpyDeviceToHost));
    if (out->schur_cols && ncopy) CK(cudaMemcpy(out->schur_cols, d_scol, sizeof(int) * ncopy, cudaMemcpyDeviceToHost));
    out->n_schur_blocks = cnt;
    out->n_free_poses = sth[0].F;
    out->n_free_points = sth[0].NL;
    int nact = 0, nhpl = 0;
    if (h->tot_edge) {
        std::vector<uint8_t> act(E);
        std::vector<int> row(E);
        CK(cudaMemcpy(act.data(), d_act, h->tot_edge, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(row.data(), d_row, sizeof(int) * h->tot_edge, cudaMemcpyDeviceToHost));
        for (int e = 0; e < h->tot_edge; ++e) { nact += act[e]; nhpl += row[e] >= 0; }
    }
    out->n_active_edges = nact;
    out->n_hpl_blocks = nhpl;
    scratch.release(); cubtmp.release();
    (void)P;
    return VISFS_BA_OK;
}

int visfs_ba_debug_trial(visfs_ba_handle *h, const visfs_ba_problem *problem, double lambda, double *S_dense, double *b_s,
                         double *x_pose, double *trial_points, int32_t *n_out, double *chi2_out, double *lambda_out,
                         double *trial_chi2_out) {
    if (!h || !problem) return VISFS_BA_ERR_INVALID;
    int st = upload(h, 1, problem);
    if (st) return st;
    st = reset_state(h);
    if (st) return st;
    cudaStream_t s = h->stream;
    k_begin_pass<<<1, 128, 0, s>>>(h->batch, 0);
    CK(h->h_small.reserve(128));
    LMState before;
    DevBuf dbg;
    struct DbgGuard {   // the damping / capture overrides never outlive this call, whichever way it returns
        visfs_ba_handle *h;
        ~DbgGuard() { h->batch.dbg = nullptr; h->batch.dbg_lambda = -1.0; h->batch_ctl.dbg = nullptr; h->batch_ctl.dbg_lambda = -1.0; }
    } dbg_guard{h};
    const int nmax = 6 * h->tot_pose;
    const size_t ntri_max = (size_t)nmax * (nmax + 1) / 2;
    std::vector<double> denseS, denseB;
    if (h->large) {
        if ((st = init_pass_large(h))) return st;
        CK(cudaMemcpyAsync(&before, h->d_st.p, sizeof(LMState), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        const int nn = 6 * before.F;
        h->batch.dbg_lambda = lambda; h->batch_ctl.dbg_lambda = lambda;
        if ((st = enqueue_build_large(h))) return st;
        CK(dbg.reserve(sizeof(double) * ((size_t)nn * nn + nn + 8)));
        const double lam_used = (lambda >= 0.0) ? lambda : before.lambda;
        lg::k_sky_to_dense<<<256, 256, 0, s>>>(h->batch, problem->trust_region == 0 ? lam_used : 0.0, dbg.as<double>(),
                                               dbg.as<double>() + (size_t)nn * nn);
        denseS.resize((size_t)nn * nn); denseB.resize(nn);
        if (nn) {
            CK(cudaMemcpyAsync(denseS.data(), dbg.p, sizeof(double) * (size_t)nn * nn, cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(denseB.data(), dbg.as<double>() + (size_t)nn * nn, sizeof(double) * nn, cudaMemcpyDeviceToHost, s));
        }
        if ((st = enqueue_rest_large(h))) return st;
        h->batch.dbg_lambda = -1.0; h->batch_ctl.dbg_lambda = -1.0;
    } else {
        st = run_structure(h);
        if (st) return st;
        k_sync_buffers<<<dim3((unsigned)h->grid_lm_x, 1u), 256, 0, s>>>(h->batch);
        LAUNCH_CHECK("k_sync_buffers");
        if (h->n_chunks) k_init<<<h->n_chunks, kUpdThreads, sizeof(InitSmem), s>>>(h->batch);
        LAUNCH_CHECK("k_init");
        if (h->tot_link > 0) k_link_lin<<<(h->tot_link + 63) / 64, 64, 0, s>>>(h->batch);
        k_control_init<<<1, kCtlInitThreads, 0, s>>>(h->batch);
        LAUNCH_CHECK("k_control_init");
        CK(cudaMemcpyAsync(&before, h->d_st.p, sizeof(LMState), cudaMemcpyDeviceToHost, s));
        CK(dbg.reserve(sizeof(double) * (ntri_max + nmax + 8)));
        CK(cudaMemsetAsync(dbg.p, 0, sizeof(double) * (ntri_max + nmax + 8), s));
        Batch saved = h->batch;
        h->batch.dbg = dbg.as<double>();
        h->batch.dbg_lambda = lambda;
        enqueue_body(h);
        h->batch = saved;
    }
    CK(cudaGetLastError());
    LMState after;
    CK(cudaMemcpyAsync(&after, h->d_st.p, sizeof(LMState), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const int n = 6 * after.F;
    const size_t ntri = (size_t)n * (n + 1) / 2;
    if (n_out) *n_out = n;
    if (h->large) {
        if (S_dense) memcpy(S_dense, denseS.data(), sizeof(double) * (size_t)n * n);
        if (b_s) memcpy(b_s, denseB.data(), sizeof(double) * n);
    } else {
        std::vector<double> packed(ntri + n + 1);
        if (ntri + n) CK(cudaMemcpy(packed.data(), dbg.p, sizeof(double) * (ntri + n), cudaMemcpyDeviceToHost));
        if (S_dense)
            for (int r = 0; r < n; ++r)
                for (int c = 0; c <= r; ++c) {
                    const double v = packed[(size_t)r * (r + 1) / 2 + c];
                    S_dense[(size_t)r * n + c] = v;
                    S_dense[(size_t)c * n + r] = v;
                }
        if (b_s) for (int i = 0; i < n; ++i) b_s[i] = packed[ntri + i];
    }
    if (x_pose && n) CK(cudaMemcpy(x_pose, h->d_xp.p, sizeof(double) * n, cudaMemcpyDeviceToHost));
    const int trial_buf = 1 - before.cur;  // the trial was written into the buffer that was not accepted before the body
    if (trial_points && h->tot_point)
        CK(cudaMemcpy(trial_points, h->d_point.as<double>() + (size_t)trial_buf * h->tot_point * 3, sizeof(double) * 3 * h->tot_point,
                      cudaMemcpyDeviceToHost));
    if (chi2_out) *chi2_out = before.chi_initial;
    if (lambda_out) *lambda_out = (lambda >= 0.0) ? lambda : before.lambda;
    if (trial_chi2_out) *trial_chi2_out = after.chi_last_trial;
    dbg.release();
    return VISFS_BA_OK;
}

int visfs_ba_debug_pose_oplus(visfs_ba_handle *h, int32_t n, const double *tq_in, const double *delta, double *tq_out) {
    if (!h) return VISFS_BA_ERR_INVALID;
    if (n <= 0 || !tq_in || !delta || !tq_out) return h->fail(VISFS_BA_ERR_INVALID, "debug_pose_oplus: bad arguments");
    CK(cudaSetDevice(h->device));
    DevBuf buf;
    CK(buf.reserve(sizeof(double) * 20 * (size_t)n));
    double *d_in = buf.as<double>(), *d_delta = d_in + 7 * (size_t)n, *d_out = d_delta + 6 * (size_t)n;
    CK(cudaMemcpyAsync(d_in, tq_in, sizeof(double) * 7 * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_delta, delta, sizeof(double) * 6 * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    k_debug_pose_oplus<<<(n + 63) / 64, 64, 0, h->stream>>>(n, d_in, d_delta, d_out);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(tq_out, d_out, sizeof(double) * 7 * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return VISFS_BA_OK;
}

int visfs_ba_debug_link_linearize(visfs_ba_handle *h, int32_t n, const double *from_tq, const double *to_tq, const double *meas_tq,
                                  double *err, double *J_from, double *J_to) {
    if (!h) return VISFS_BA_ERR_INVALID;
    if (n <= 0 || !from_tq || !to_tq || !meas_tq || !err || !J_from || !J_to) return h->fail(VISFS_BA_ERR_INVALID, "debug_link_linearize: bad arguments");
    CK(cudaSetDevice(h->device));
    DevBuf buf;
    const size_t N = (size_t)n;
    CK(buf.reserve(sizeof(double) * (21 + 6 + 72) * N));
    double *d_a = buf.as<double>(), *d_b = d_a + 7 * N, *d_m = d_b + 7 * N, *d_e = d_m + 7 * N, *d_ji = d_e + 6 * N, *d_jj = d_ji + 36 * N;
    CK(cudaMemcpyAsync(d_a, from_tq, sizeof(double) * 7 * N, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_b, to_tq, sizeof(double) * 7 * N, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_m, meas_tq, sizeof(double) * 7 * N, cudaMemcpyHostToDevice, h->stream));
    k_debug_link<<<(n + 63) / 64, 64, 0, h->stream>>>(n, d_a, d_b, d_m, d_e, d_ji, d_jj);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(err, d_e, sizeof(double) * 6 * N, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(J_from, d_ji, sizeof(double) * 36 * N, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(J_to, d_jj, sizeof(double) * 36 * N, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return VISFS_BA_OK;
}

int visfs_ba_comm_unique_id(void *id_out) {
    if (!id_out) return VISFS_BA_ERR_INVALID;
    if (!g_nccl.load()) { g_create_error = g_nccl.error; return VISFS_BA_ERR_CUDA; }
    static_assert(sizeof(ncclUniqueId) <= VISFS_BA_COMM_ID_BYTES, "ncclUniqueId does not fit VISFS_BA_COMM_ID_BYTES");
    ncclUniqueId id;
    const ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) { g_create_error = g_nccl.GetErrorString(r); return VISFS_BA_ERR_CUDA; }
    memset(id_out, 0, VISFS_BA_COMM_ID_BYTES);
    memcpy(id_out, &id, sizeof id);
    return VISFS_BA_OK;
}

int visfs_ba_comm_init(visfs_ba_handle *h, int32_t n_ranks, int32_t rank, const void *id_in) {
    if (!h) return VISFS_BA_ERR_INVALID;
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return h->fail(VISFS_BA_ERR_INVALID, "bad rank / world size");
    if (h->comm) visfs_ba_comm_destroy(h);
    h->comm_ranks = n_ranks; h->comm_rank = rank;
    if (n_ranks == 1) return VISFS_BA_OK;
    if (!id_in) return h->fail(VISFS_BA_ERR_INVALID, "null unique id");
    if (!g_nccl.load()) return h->fail(VISFS_BA_ERR_CUDA, g_nccl.error);
    CK(cudaSetDevice(h->device));
    ncclUniqueId id;
    memcpy(&id, id_in, sizeof id);
    const ncclResult_t r = g_nccl.CommInitRank(&h->comm, n_ranks, id, rank);
    if (r != ncclSuccess) { h->comm = nullptr; return h->fail(VISFS_BA_ERR_CUDA, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r)); }
    return VISFS_BA_OK;
}

int visfs_ba_comm_destroy(visfs_ba_handle *h) {
    if (!h) return VISFS_BA_ERR_INVALID;
    if (h->comm) {
        cudaSetDevice(h->device);
        cudaStreamSynchronize(h->stream);
        g_nccl.CommDestroy(h->comm);
        h->comm = nullptr;
    }
    h->comm_ranks = 1; h->comm_rank = 0;
    return VISFS_BA_OK;
}


// ---- resident local map -----------------------------------------------------------------------------------------
int visfs_ba_window_create(visfs_ba_handle *h, const visfs_ba_window_config *cfg, visfs_ba_window **out) {
    if (!h || !out) return VISFS_BA_ERR_INVALID;
    *out = nullptr;
    if (!cfg || cfg->max_frames < 2 || cfg->max_frames > kMaxSmallPoses || cfg->max_points < 1 || cfg->max_observations < 1 ||
        !(cfg->pixel_variance > 0.0))
        return h->fail(VISFS_BA_ERR_INVALID, "window: bad configuration (2..32 frames, positive capacities, pixel_variance > 0)");
    CK(cudaSetDevice(h->device));
    visfs_ba_window *w = new visfs_ba_window();
    w->h = h; w->cfg = *cfg;
    const size_t F = (size_t)cfg->max_frames, L = (size_t)cfg->max_points, E = (size_t)cfg->max_observations;
    for (int i = (int)F - 1; i >= 0; --i) w->free_frames.push_back(i);
    for (int i = (int)L - 1; i >= 0; --i) w->free_points.push_back(i);
    w->stamp.assign(L, 0);
    w->point_id_of_slot.assign(L, -1); w->frame_id_of_slot.assign(F, -1);
    cudaError_t e = cudaSuccess;
    auto R = [&](DevBuf &b, size_t bytes) { if (e == cudaSuccess) e = b.reserve(bytes); };
    R(w->d_frame_tq, 56 * F); R(w->d_frame_pose, 4 * F); R(w->d_pose_slot, 4 * F); R(w->d_point_xyz, 24 * L); R(w->d_point_fixed, L);
    R(w->d_point_id, 8 * L); R(w->d_order, 4 * L); R(w->d_ob_point, 4 * E); R(w->d_ob_frame, 4 * E); R(w->d_ob_obs, 12 * E);
    R(w->d_ob_kind, E); R(w->d_ob_dead, E); R(w->d_cnt, 4 * L); R(w->d_act, 4 * L); R(w->d_rank, 4 * L); R(w->d_rank_of_slot, 4 * L);
    R(w->d_slot_of_rank, 4 * L); R(w->d_key, 4 * E); R(w->d_key2, 4 * E); R(w->d_val, 4 * E); R(w->d_val2, 4 * E); R(w->d_counters, 16);
    R(w->d_pose_out, 56 * F); R(w->d_outliers, 8 * E); R(w->d_list, 8 * E); R(w->d_mask, 4 * L);
    if (e == cudaSuccess) e = w->h_small.reserve(4096 + 56 * F + 8 * E);
    // everything a solve or a delta may need later is allocated now: cudaMalloc / cudaFree / cudaMallocHost in the middle of a
    // session cost milliseconds once the process holds many streams and large page-locked buffers
    if (e == cudaSuccess) {
        size_t tb_scan = 0, tb_sort = 0, tb_scan2 = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb_scan, (const int *)nullptr, (int *)nullptr, (int)L, h->stream);
        cub::DeviceScan::ExclusiveSum(nullptr, tb_scan2, (const int *)nullptr, (int *)nullptr, (int)E + 1, h->stream);
        cub::DeviceRadixSort::SortPairs(nullptr, tb_sort, (const unsigned *)nullptr, (unsigned *)nullptr, (const int *)nullptr, (int *)nullptr, (int)E, 0, 32, h->stream);
        R(w->d_scan_tmp, std::max(tb_scan, tb_scan2)); R(w->d_sort_tmp, tb_sort);
        R(w->d_compact, 25 * E + 64);
        const size_t stage_bytes = 64 + 21 * std::max<size_t>(E / 4, 1024) + 37 * std::min<size_t>(L, 16384);
        R(w->d_stage, stage_bytes);
        if (e == cudaSuccess) e = w->h_stage.reserve(stage_bytes);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&w->stage_done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMemsetAsync(w->d_ob_dead.p, 0, E, h->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(w->d_mask.p, 0, 4 * L, h->stream);
    if (e != cudaSuccess) { const int st = h->cuda_fail(e, "window allocation"); delete w; return st; }
    *out = w;
    return VISFS_BA_OK;
}

void visfs_ba_window_destroy(visfs_ba_window *w) {
    if (!w) return;
    cudaSetDevice(w->h->device);
    cudaStreamSynchronize(w->h->stream);
    if (w->stage_done) cudaEventDestroy(w->stage_done);
    delete w;
}

int64_t visfs_ba_window_h2d_bytes_total(const visfs_ba_window *w) { return w ? w->h2d_total : 0; }

int visfs_ba_window_set_points(visfs_ba_window *w, int32_t n, const int64_t *point_id, const double *xyz, const uint8_t *fixed) {
    if (!w) return VISFS_BA_ERR_INVALID;
    visfs_ba_handle *h = w->h;
    if (n < 0 || (n > 0 && (!point_id || !xyz))) return win_fail(w, VISFS_BA_ERR_INVALID, "set_points: null arrays");
    if (n == 0) return VISFS_BA_OK;
    CK(cudaSetDevice(h->device));
    int fresh = 0;
    for (int i = 0; i < n; ++i) fresh += w->point_slot.find(point_id[i]) == w->point_slot.end();
    if (fresh > (int)w->free_points.size()) return win_fail(w, VISFS_BA_ERR_INVALID, "set_points: more features than max_points");
    // one packed delta [slot | xyz | id | fixed] -> one copy -> scatter on the device
    const size_t N = (size_t)n, o_slot = 0, o_xyz = (4 * N + 15) & ~(size_t)15, o_id = o_xyz + 24 * N, o_fix = o_id + 8 * N, bytes = o_fix + N;
    char *sg = w->stage(bytes);
    if (!sg) return win_fail(w, VISFS_BA_ERR_CUDA, "set_points: staging allocation failed");
    int *slots = reinterpret_cast<int *>(sg + o_slot);
    for (int i = 0; i < n; ++i) {
        auto it = w->point_slot.find(point_id[i]);
        if (it == w->point_slot.end()) {
            const int slot = w->free_points.back(); w->free_points.pop_back();
            w->point_slot.emplace(point_id[i], slot); w->point_order.emplace(point_id[i], slot);
            w->point_id_of_slot[(size_t)slot] = point_id[i];
            w->order_dirty = true;
            slots[i] = slot;
        } else slots[i] = it->second;
        sg[o_fix + (size_t)i] = (char)(fixed ? fixed[i] : 0);
    }
    memcpy(sg + o_xyz, xyz, 24 * N);
    memcpy(sg + o_id, point_id, 8 * N);
    cudaStream_t s = h->stream;
    char *dg = w->d_stage.as<char>();
    CK(cudaMemcpyAsync(dg, sg, bytes, cudaMemcpyHostToDevice, s));
    wn::k_win_scatter_points<<<std::max(1, std::min((n + 127) / 128, 64)), 128, 0, s>>>(
        n, reinterpret_cast<const int *>(dg + o_slot), reinterpret_cast<const double *>(dg + o_xyz), reinterpret_cast<const uint8_t *>(dg + o_fix),
        reinterpret_cast<const long long *>(dg + o_id), w->d_point_xyz.as<double>(), w->d_point_fixed.as<uint8_t>(), w->d_point_id.as<long long>());
    CK(cudaGetLastError());
    CK(cudaEventRecord(w->stage_done, s));
    w->h2d_total += (int64_t)bytes;
    return VISFS_BA_OK;
}

int visfs_ba_window_insert_frame(visfs_ba_window *w, int64_t frame_id, const double *pose_tq, int32_t n_obs, const int64_t *point_id,
                                 const float *obs_uvr, const uint8_t *kind) {
    if (!w) return VISFS_BA_ERR_INVALID;
    visfs_ba_handle *h = w->h;
    if (!pose_tq || n_obs < 0 || (n_obs > 0 && (!point_id || !obs_uvr))) return win_fail(w, VISFS_BA_ERR_INVALID, "insert_frame: null arrays");
    if (w->frame_slot.count(frame_id)) return win_fail(w, VISFS_BA_ERR_INVALID, "insert_frame: the frame is already in the window");
    if (w->free_frames.empty()) return win_fail(w, VISFS_BA_ERR_INVALID, "insert_frame: window full (remove a frame first, LocalMap::removeSignature)");
    CK(cudaSetDevice(h->device));
    const int fslot = w->free_frames.back();
    const size_t n = (size_t)n_obs;
    // one packed delta [pose | feature slots | frame slots | observations | kinds] -> one staging buffer, contiguous appends.
    // The feature slots are looked up straight into it; nothing of the window changes before the whole list has been validated.
    const size_t o_pose = 0, o_ps = 64, o_fs = o_ps + 4 * n, o_obs = o_fs + 4 * n, o_kind = o_obs + 12 * n, bytes = o_kind + n;
    char *sg = w->stage(bytes);
    if (!sg) return win_fail(w, VISFS_BA_ERR_CUDA, "insert_frame: staging allocation failed");
    int *ps = reinterpret_cast<int *>(sg + o_ps), *fs = reinterpret_cast<int *>(sg + o_fs);
    const int now = ++w->stamp_now;
    for (size_t i = 0; i < n; ++i) {
        auto it = w->point_slot.find(point_id[i]);
        if (it == w->point_slot.end()) return win_fail(w, VISFS_BA_ERR_INVALID, "insert_frame: observation of a feature that was never set (visfs_ba_window_set_points)");
        const int slot = it->second;
        if (w->stamp[(size_t)slot] == now) return win_fail(w, VISFS_BA_ERR_INVALID, "insert_frame: two observations of one feature in one frame");
        w->stamp[(size_t)slot] = now;
        ps[i] = slot; fs[i] = fslot;
    }
    if (w->n_pool + n_obs > w->cfg.max_observations) {
        const int st = win_compact(w);
        if (st) return st;
        if (w->n_pool + n_obs > w->cfg.max_observations) return win_fail(w, VISFS_BA_ERR_INVALID, "insert_frame: observation pool full (max_observations)");
    }
    w->free_frames.pop_back();
    w->frame_slot.emplace(frame_id, fslot);
    w->frame_id_of_slot[(size_t)fslot] = frame_id;
    cudaStream_t s = h->stream;
    const size_t at = (size_t)w->n_pool;
    memcpy(sg + o_pose, pose_tq, 56);
    if (n) {
        memcpy(sg + o_obs, obs_uvr, 12 * n);
        if (kind) memcpy(sg + o_kind, kind, n); else memset(sg + o_kind, 0, n);
    }
    CK(cudaMemcpyAsync(w->d_frame_tq.as<double>() + 7 * (size_t)fslot, sg + o_pose, 56, cudaMemcpyHostToDevice, s));
    if (n) {
        CK(cudaMemcpyAsync(w->d_ob_point.as<int>() + at, sg + o_ps, 4 * n, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(w->d_ob_frame.as<int>() + at, sg + o_fs, 4 * n, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(w->d_ob_obs.as<float>() + 3 * at, sg + o_obs, 12 * n, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(w->d_ob_kind.as<uint8_t>() + at, sg + o_kind, n, cudaMemcpyHostToDevice, s));
    }
    CK(cudaEventRecord(w->stage_done, s));
    w->n_pool += n_obs;
    w->h2d_total += 56 + (int64_t)(21 * n);
    return VISFS_BA_OK;
}

int visfs_ba_window_remove_frame(visfs_ba_window *w, int64_t frame_id) {
    if (!w) return VISFS_BA_ERR_INVALID;
    visfs_ba_handle *h = w->h;
    auto it = w->frame_slot.find(frame_id);
    if (it == w->frame_slot.end()) return win_fail(w, VISFS_BA_ERR_INVALID, "remove_frame: no such frame");
    CK(cudaSetDevice(h->device));
    const int fslot = it->second;
    if (w->n_pool > 0)
        wn::k_win_kill_frame<<<std::max(1, std::min((w->n_pool + 255) / 256, 256)), 256, 0, h->stream>>>(w->d_ob_frame.as<int>(), w->d_ob_dead.as<uint8_t>(), w->n_pool, fslot);
    CK(cudaGetLastError());
    w->frame_slot.erase(it);
    w->frame_id_of_slot[(size_t)fslot] = -1;
    w->free_frames.push_back(fslot);
    return VISFS_BA_OK;
}

// (feature slot [, frame slot]) pairs -> tombstones in the pool: masks set, one pass over the pool, masks cleared (ba_window.cuh)
static int win_kill(visfs_ba_window *w, const std::vector<int> &pslots, const std::vector<int> *fslots) {
    visfs_ba_handle *h = w->h;
    const size_t n = pslots.size();
    if (n == 0 || w->n_pool == 0) return VISFS_BA_OK;
    const size_t bytes = 4 * n * (fslots ? 2 : 1);
    char *sg = w->stage(bytes);
    if (!sg) return win_fail(w, VISFS_BA_ERR_CUDA, "staging allocation failed");
    memcpy(sg, pslots.data(), 4 * n);
    if (fslots) memcpy(sg + 4 * n, fslots->data(), 4 * n);
    cudaStream_t s = h->stream;
    CK(cudaMemcpyAsync(w->d_list.p, sg, bytes, cudaMemcpyHostToDevice, s));
    const int *dp = w->d_list.as<int>(), *df = fslots ? dp + n : nullptr;
    const int gl = std::max(1, std::min(((int)n + 255) / 256, 64)), gp = std::max(1, std::min((w->n_pool + 255) / 256, 256));
    wn::k_win_mask_set<<<gl, 256, 0, s>>>(w->d_mask.as<unsigned>(), dp, df, (int)n);
    wn::k_win_mask_kill<<<gp, 256, 0, s>>>(w->d_mask.as<unsigned>(), w->d_ob_point.as<int>(), w->d_ob_frame.as<int>(), w->d_ob_dead.as<uint8_t>(), w->n_pool);
    wn::k_win_mask_clear<<<gl, 256, 0, s>>>(w->d_mask.as<unsigned>(), dp, (int)n);
    CK(cudaGetLastError());
    CK(cudaEventRecord(w->stage_done, s));
    w->h2d_total += (int64_t)bytes;
    return VISFS_BA_OK;
}

int visfs_ba_window_remove_points(visfs_ba_window *w, int32_t n, const int64_t *point_id) {
    if (!w) return VISFS_BA_ERR_INVALID;
    if (n < 0 || (n > 0 && !point_id)) return win_fail(w, VISFS_BA_ERR_INVALID, "remove_points: null array");
    if (cudaSetDevice(w->h->device) != cudaSuccess) return VISFS_BA_ERR_CUDA;
    for (int k = 0; k < n; ++k)
        if (w->point_slot.find(point_id[k]) == w->point_slot.end()) return win_fail(w, VISFS_BA_ERR_INVALID, "remove_points: no such feature");
    std::vector<int> list;
    list.reserve((size_t)n);
    for (int k = 0; k < n; ++k) {
        auto it = w->point_slot.find(point_id[k]);
        if (it == w->point_slot.end()) continue;   // listed twice
        const int slot = it->second;
        list.push_back(slot);
        w->point_order.erase(point_id[k]);
        w->point_slot.erase(it);
        w->point_id_of_slot[(size_t)slot] = -1;
        w->order_dirty = true;
    }
    // the tombstones are set before the slots can be handed out again (same stream)
    const int st = win_kill(w, list, nullptr);
    for (int slot : list) w->free_points.push_back(slot);
    return st;
}

int visfs_ba_window_remove_observations(visfs_ba_window *w, int32_t n, const int64_t *point_id, const int64_t *frame_id) {
    if (!w) return VISFS_BA_ERR_INVALID;
    if (n < 0 || (n > 0 && (!point_id || !frame_id))) return win_fail(w, VISFS_BA_ERR_INVALID, "remove_observations: null arrays");
    if (cudaSetDevice(w->h->device) != cudaSuccess) return VISFS_BA_ERR_CUDA;
    std::vector<int> ps, fs;
    ps.reserve((size_t)n); fs.reserve((size_t)n);
    for (int k = 0; k < n; ++k) {
        auto ip = w->point_slot.find(point_id[k]);
        auto jf = w->frame_slot.find(frame_id[k]);
        if (ip == w->point_slot.end() || jf == w->frame_slot.end()) continue;      // LocalMap.cpp:219-224 logs and goes on
        ps.push_back(ip->second); fs.push_back(jf->second);
    }
    return win_kill(w, ps, &fs);
}

int visfs_ba_window_set_poses(visfs_ba_window *w, int32_t n, const int64_t *frame_id, const double *pose_tq) {
    if (!w) return VISFS_BA_ERR_INVALID;
    visfs_ba_handle *h = w->h;
    if (n < 0 || (n > 0 && (!frame_id || !pose_tq))) return win_fail(w, VISFS_BA_ERR_INVALID, "set_poses: null arrays");
    CK(cudaSetDevice(h->device));
    for (int k = 0; k < n; ++k) {
        auto it = w->frame_slot.find(frame_id[k]);
        if (it == w->frame_slot.end()) return win_fail(w, VISFS_BA_ERR_INVALID, "set_poses: no such frame");
        CK(cudaMemcpyAsync(w->d_frame_tq.as<double>() + 7 * (size_t)it->second, pose_tq + 7 * (size_t)k, 56, cudaMemcpyHostToDevice, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    w->h2d_total += 56 * (int64_t)n;
    return VISFS_BA_OK;
}

int visfs_ba_window_set_links(visfs_ba_window *w, int32_t n, const int64_t *from_frame_id, const int64_t *to_frame_id, const double *link_tq,
                              double odometry_variance) {
    if (!w) return VISFS_BA_ERR_INVALID;
    if (n < 0 || (n > 0 && (!from_frame_id || !to_frame_id || !link_tq))) return win_fail(w, VISFS_BA_ERR_INVALID, "set_links: null arrays");
    if (n > 0 && !(odometry_variance > 0.0)) return win_fail(w, VISFS_BA_ERR_INVALID, "set_links: odometry_variance must be positive");
    for (int k = 0; k < n; ++k) {
        if (from_frame_id[k] == to_frame_id[k]) return win_fail(w, VISFS_BA_ERR_INVALID, "set_links: a link from a frame to itself");
        double nq = 0.0;
        for (int a = 0; a < 7; ++a) {
            if (!std::isfinite(link_tq[7 * (size_t)k + a])) return win_fail(w, VISFS_BA_ERR_INVALID, "set_links: non-finite measurement");
            if (a >= 3) nq += link_tq[7 * (size_t)k + a] * link_tq[7 * (size_t)k + a];
        }
        if (!(nq > 0.0)) return win_fail(w, VISFS_BA_ERR_INVALID, "set_links: zero quaternion");
    }
    w->links.resize((size_t)n);
    for (int k = 0; k < n; ++k) {
        w->links[(size_t)k].from = from_frame_id[k]; w->links[(size_t)k].to = to_frame_id[k];
        memcpy(w->links[(size_t)k].tq, link_tq + 7 * (size_t)k, 56);
    }
    w->odometry_variance = odometry_variance;
    return VISFS_BA_OK;
}

int visfs_ba_window_get_points(visfs_ba_window *w, int32_t n, const int64_t *point_id, double *xyz_out) {
    if (!w) return VISFS_BA_ERR_INVALID;
    visfs_ba_handle *h = w->h;
    if (n < 0 || (n > 0 && (!point_id || !xyz_out))) return win_fail(w, VISFS_BA_ERR_INVALID, "get_points: null arrays");
    CK(cudaSetDevice(h->device));
    for (int k = 0; k < n; ++k) {
        auto it = w->point_slot.find(point_id[k]);
        if (it == w->point_slot.end()) return win_fail(w, VISFS_BA_ERR_INVALID, "get_points: no such feature");
        CK(cudaMemcpyAsync(xyz_out + 3 * (size_t)k, w->d_point_xyz.as<double>() + 3 * (size_t)it->second, 24, cudaMemcpyDeviceToHost, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    return VISFS_BA_OK;
}

int visfs_ba_window_solve(visfs_ba_window *w, int64_t root_frame_id, visfs_ba_window_result *res) {
    if (!w || !res) return VISFS_BA_ERR_INVALID;
    visfs_ba_handle *h = w->h;
    CK(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    const bool wtrace = getenv("VISFS_BA_WIN_TRACE") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a2, std::chrono::steady_clock::time_point b2) { return std::chrono::duration<double, std::milli>(b2 - a2).count(); };
    const auto tw0 = now();
    const int P = (int)w->frame_slot.size(), n_pool = w->n_pool, n_order = (int)w->point_order.size();
    const int maxF = w->cfg.max_frames;
    int64_t h2d = 0;
    // ---- the one table a solve sends: frame slot -> pose index (ascending signature id), pose index -> slot
    int *tab = w->h_small.as<int>();
    int root_pose = -1;
    for (int i = 0; i < maxF; ++i) tab[i] = -1;
    std::vector<int32_t> link_from, link_to;
    std::vector<double> link_tq;
    {
        int k = 0;
        std::unordered_map<int64_t, int> pose_of;
        for (const auto &kv : w->frame_slot) {
            tab[kv.second] = k; tab[maxF + k] = kv.second;
            if (kv.first == root_frame_id) root_pose = k;
            if (res->frame_id) res->frame_id[k] = kv.first;
            pose_of[kv.first] = k;
            ++k;
        }
        for (const auto &lk : w->links) {   // the links whose two frames are in the window (Optimizer.cpp:130: uContains(_poses, ...))
            const auto a = pose_of.find(lk.from), b = pose_of.find(lk.to);
            if (a == pose_of.end() || b == pose_of.end()) continue;
            link_from.push_back(a->second); link_to.push_back(b->second);
            link_tq.insert(link_tq.end(), lk.tq, lk.tq + 7);
        }
    }
    CK(cudaMemcpyAsync(w->d_frame_pose.p, tab, 4 * (size_t)maxF, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(w->d_pose_slot.p, tab + maxF, 4 * (size_t)std::max(P, 1), cudaMemcpyHostToDevice, s));
    h2d += 4 * (maxF + P);
    std::vector<int> order;
    if (w->order_dirty) {   // only when a feature entered or left since the last solve
        order.reserve((size_t)n_order);
        for (const auto &kv : w->point_order) order.push_back(kv.second);
        if (n_order) CK(cudaMemcpyAsync(w->d_order.p, order.data(), 4 * (size_t)n_order, cudaMemcpyHostToDevice, s));
        h2d += 4 * n_order;
    }
    wn::Win W{};
    W.frame_tq = w->d_frame_tq.as<double>(); W.frame_pose = w->d_frame_pose.as<int>(); W.pose_slot = w->d_pose_slot.as<int>();
    W.point_xyz = w->d_point_xyz.as<double>(); W.point_fixed = w->d_point_fixed.as<uint8_t>(); W.point_id = w->d_point_id.as<long long>();
    W.order = w->d_order.as<int>(); W.n_order = n_order;
    W.ob_point = w->d_ob_point.as<int>(); W.ob_frame = w->d_ob_frame.as<int>(); W.ob_obs = w->d_ob_obs.as<float>();
    W.ob_kind = w->d_ob_kind.as<uint8_t>(); W.ob_dead = w->d_ob_dead.as<uint8_t>(); W.n_pool = n_pool;
    W.cnt = w->d_cnt.as<int>(); W.act = w->d_act.as<int>(); W.rank = w->d_rank.as<int>();
    W.rank_of_slot = w->d_rank_of_slot.as<int>(); W.slot_of_rank = w->d_slot_of_rank.as<int>();
    W.key = w->d_key.as<unsigned>(); W.key_sorted = w->d_key2.as<unsigned>();
    W.val = w->d_val.as<int>(); W.val_sorted = w->d_val2.as<int>(); W.counters = w->d_counters.as<int>();
    int L = 0, E = 0;
    if (n_pool > 0 && n_order > 0 && P > 0) {
        const int gp = std::max(1, std::min((n_pool + 255) / 256, 1024)), go = std::max(1, std::min((n_order + 255) / 256, 1024));
        CK(cudaMemsetAsync(w->d_cnt.p, 0, 4 * (size_t)w->cfg.max_points, s));
        CK(cudaMemsetAsync(w->d_counters.p, 0, 16, s));
        wn::k_win_count<<<gp, 256, 0, s>>>(W);
        wn::k_win_act<<<go, 256, 0, s>>>(W);
        size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, W.act, W.rank, n_order, s);
        CK(w->d_scan_tmp.reserve(tb));
        CK(cub::DeviceScan::ExclusiveSum(w->d_scan_tmp.p, tb, W.act, W.rank, n_order, s));
        wn::k_win_rank<<<go, 256, 0, s>>>(W);
        wn::k_win_keys<<<gp, 256, 0, s>>>(W);
        tb = 0;   // (all 32 bits: the keys of observations that take no part are 0xffffffff and must sort last)
        cub::DeviceRadixSort::SortPairs(nullptr, tb, W.key, W.key_sorted, W.val, W.val_sorted, n_pool, 0, 32, s);
        CK(w->d_sort_tmp.reserve(tb));
        CK(cub::DeviceRadixSort::SortPairs(w->d_sort_tmp.p, tb, W.key, W.key_sorted, W.val, W.val_sorted, n_pool, 0, 32, s));
        int *cnt_h = w->h_small.as<int>() + 2 * maxF;
        CK(cudaMemcpyAsync(cnt_h, w->d_counters.p, 8, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        L = cnt_h[0]; E = cnt_h[1];
        h->launches = 0;
    } else {
        CK(cudaStreamSynchronize(s));
    }
    w->order_dirty = false;
    const auto tw1 = now();
    // ---- the window as one problem whose arrays are already on the device
    visfs_ba_problem pb{};
    pb.n_poses = P; pb.n_points = L; pb.n_edges = E;
    pb.fx = w->cfg.fx; pb.fy = w->cfg.fy; pb.cx = w->cfg.cx; pb.cy = w->cfg.cy; pb.bf = w->cfg.bf;
    pb.pixel_variance = w->cfg.pixel_variance; pb.huber_delta = w->cfg.huber_delta;
    pb.iterations = w->cfg.iterations; pb.solver = w->cfg.solver; pb.trust_region = w->cfg.trust_region;
    pb.n_links = (int32_t)link_from.size();   // (the links travel from the host like in a full call: a few dozen bytes each)
    pb.link_from = link_from.data(); pb.link_to = link_to.data(); pb.link_tq = link_tq.data();
    pb.odometry_variance = w->odometry_variance;
    DevInput dev;
    dev.max_degree = std::max(P, 1); dev.n_fixed = root_pose >= 0 ? 1 : 0;
    dev.emit = [&](visfs_ba_handle *hh) -> int {
        const int items = std::max(std::max(P * 7, L), E);
        wn::k_win_emit<<<std::max(1, std::min((items + 255) / 256, 1024)), 256, 0, hh->stream>>>(
            W, P, L, E, root_pose, hh->in.pose, hh->in.point, reinterpret_cast<float *>(hh->in.obs), hh->in.epose, hh->in.epoint, hh->in.pfix,
            hh->in.lfix, hh->in.ekind);
        return cudaGetLastError() == cudaSuccess ? VISFS_BA_OK : VISFS_BA_ERR_CUDA;
    };
    *res = visfs_ba_window_result{res->frame_id, res->pose_tq, res->outlier_point_id, res->outlier_frame_id, res->outlier_capacity};
    res->n_frames = P; res->n_points = L; res->n_edges = E;
    if (P == 0) return win_fail(w, VISFS_BA_ERR_INVALID, "solve: the window holds no frame");
    int st = upload(h, 1, &pb, &dev);
    if (st) return st;
    h2d += h->h2d_bytes;
    const auto tw2 = now();
    st = run_resident(h);
    if (st) return st;
    const auto tw3 = now();
    // ---- results: poses back, write-back into the resident state, outlier list
    const int cap = std::max(0, std::min(res->outlier_capacity, w->cfg.max_observations));
    const int items = std::max(std::max(P * 7, L), E);
    wn::k_win_finish<<<std::max(1, std::min((items + 255) / 256, 1024)), 256, 0, s>>>(W, h->batch, P, L, E, 1, w->d_pose_out.as<double>(),
                                                                                     w->d_outliers.as<int>(), cap);
    CK(cudaGetLastError());
    double *pose_h = reinterpret_cast<double *>(w->h_small.as<char>() + 4096);
    int *out_h = reinterpret_cast<int *>(w->h_small.as<char>() + 4096 + 56 * (size_t)maxF);
    int *cnt_h = w->h_small.as<int>() + 2 * maxF;
    CK(cudaMemcpyAsync(pose_h, w->d_pose_out.p, 56 * (size_t)P, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(cnt_h, w->d_counters.p, 12, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const int n_out = (E > 0) ? cnt_h[2] : 0;
    const int n_copy = std::min(n_out, cap);
    if (n_copy > 0) {
        CK(cudaMemcpyAsync(out_h, w->d_outliers.p, 8 * (size_t)n_copy, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
    }
    if (res->pose_tq) memcpy(res->pose_tq, pose_h, 56 * (size_t)P);
    {   // (feature id, signature id) pairs, sorted: the order of Optimizer.cpp:283-297 (features ascending, frames ascending)
        std::vector<std::pair<int64_t, int64_t>> outl((size_t)n_copy);
        for (int k = 0; k < n_copy; ++k) outl[(size_t)k] = {w->point_id_of_slot[(size_t)out_h[2 * k]], w->frame_id_of_slot[(size_t)out_h[2 * k + 1]]};
        std::sort(outl.begin(), outl.end());
        for (int k = 0; k < n_copy; ++k) {
            if (res->outlier_point_id) res->outlier_point_id[k] = outl[(size_t)k].first;
            if (res->outlier_frame_id) res->outlier_frame_id[k] = outl[(size_t)k].second;
        }
    }
    const LMState &ls = h->st_host[0];
    res->n_outliers = n_out; res->status = ls.status;
    for (int k = 0; k < 2; ++k) { res->iterations_run[k] = ls.iterations_run[k]; res->trials_run[k] = ls.trials_run[k]; res->stop_reason[k] = ls.stop[k]; }
    res->chi2_initial = ls.chi_initial; res->chi2_pass1 = ls.chi_pass[0]; res->chi2_final = ls.chi_pass[1];
    res->h2d_bytes = h2d;
    res->d2h_bytes = 56 * (int64_t)P + 8 * (int64_t)n_copy + 20 + (int64_t)sizeof(LMState);
    w->h2d_total += h2d;
    if (wtrace) fprintf(stderr, "[visfs_ba] window solve: prep %.3f upload %.3f run %.3f finish %.3f ms\n", ms(tw0, tw1), ms(tw1, tw2), ms(tw2, tw3), ms(tw3, now()));
    return ls.status;
}

void *visfs_ba_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}

void visfs_ba_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

int visfs_ba_probe_fp64(visfs_ba_handle *h, double *tflops_out) {
    if (!h || !tflops_out) return VISFS_BA_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    const int blocks = h->sm_count * 8, threads = 256, iters = 1 << 15;
    CK(h->d_tmp.reserve(sizeof(double) * (size_t)blocks * threads));
    k_probe_fp64<<<blocks, threads, 0, h->stream>>>(h->d_tmp.as<double>(), 1024);
    double best = 0.0;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(h->ev_t0, h->stream));
        k_probe_fp64<<<blocks, threads, 0, h->stream>>>(h->d_tmp.as<double>(), iters);
        CK(cudaEventRecord(h->ev_t1, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        float ms = 0;
        cudaEventElapsedTime(&ms, h->ev_t0, h->ev_t1);
        const double flops = 2.0 * 8.0 * (double)iters * (double)blocks * threads;
        best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    *tflops_out = best;
    return VISFS_BA_OK;
}

}  // extern "C"

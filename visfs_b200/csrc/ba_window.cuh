// ba_window.cuh — device side of the resident local map (include/visfs_ba.h, visfs_ba_window_*).
//
// State kept in HBM between calls: a frame table (T_cw per slot), a feature table (xyz, fixed flag, id per slot) and an
// append-only observation pool (feature slot, frame slot, (u, v, u_right) as floats, kind, dead flag).  A solve turns that
// state into the solver's input layout ON THE DEVICE — the job LocalMap::getSignaturePoses / getFeaturePosesAndObservations
// (corelib/src/LocalMap.cpp:228-236, 274-294) and the graph construction of Optimizer.cpp:100-223 do with std::maps:
//   k_win_count    live observations per feature (a feature takes part when observed more than once, LocalMap.cpp:277)
//   k_win_rank     features in ascending id order -> dense index (the std::map order of Optimizer.cpp:156)
//   k_win_keys     sort key (feature index, pose index) of every live observation: g2o's insertion order
//   (cub radix sort)
//   k_win_emit     poses, points, sorted edges into the staging layout of the batch upload
//   k_win_finish   write-back of Optimizer.cpp:320-358 into the resident state, outlier list of Optimizer.cpp:283-297
#pragma once
#include "ba_kernels.cuh"

namespace visfs {
namespace wn {

struct Win {
    // frames
    double *frame_tq;            // [max_frames][7]
    const int *frame_pose;       // [max_frames] slot -> pose index of this solve, -1: not in the window
    const int *pose_slot;        // [P] pose index -> slot
    // features
    double *point_xyz;           // [max_points][3]
    const uint8_t *point_fixed;  // [max_points]
    const long long *point_id;   // [max_points]
    const int *order;            // [n_order] feature slots in ascending id order
    int n_order;
    // observation pool
    const int *ob_point, *ob_frame;      // [n_pool] slots
    const float *ob_obs;                 // [n_pool][3]
    const uint8_t *ob_kind, *ob_dead;    // [n_pool]
    int n_pool;
    // per-solve scratch
    int *cnt;                    // [max_points] live observations per feature slot
    int *act, *rank;             // [n_order] takes part / dense index (exclusive scan of act)
    int *rank_of_slot;           // [max_points] dense index or -1
    int *slot_of_rank;           // [max_points]
    unsigned *key, *key_sorted;  // (feature index * 32 + pose index), 0xffffffff = not part of this solve
    int *val, *val_sorted;
    int *counters;               // [0] features in the window, [1] edges, [2] outliers
};

__global__ void k_win_count(Win W) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < W.n_pool; i += gridDim.x * blockDim.x)
        if (!W.ob_dead[i] && W.frame_pose[W.ob_frame[i]] >= 0) atomicAdd(&W.cnt[W.ob_point[i]], 1);
}

__global__ void k_win_act(Win W) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < W.n_order; j += gridDim.x * blockDim.x)
        W.act[j] = W.cnt[W.order[j]] > 1 ? 1 : 0;      // feature.second.getObservedTimes() > 1 (LocalMap.cpp:277)
}

__global__ void k_win_rank(Win W) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < W.n_order; j += gridDim.x * blockDim.x) {
        const int slot = W.order[j];
        if (W.act[j]) { W.rank_of_slot[slot] = W.rank[j]; W.slot_of_rank[W.rank[j]] = slot; }
        else W.rank_of_slot[slot] = -1;
        if (j == W.n_order - 1) W.counters[0] = W.rank[j] + W.act[j];
    }
}

__global__ void k_win_keys(Win W) {
    int n = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < W.n_pool; i += gridDim.x * blockDim.x) {
        unsigned k = 0xffffffffu;
        if (!W.ob_dead[i]) {
            const int p = W.frame_pose[W.ob_frame[i]], r = W.rank_of_slot[W.ob_point[i]];
            if (p >= 0 && r >= 0) { k = ((unsigned)r << 5) | (unsigned)p; ++n; }      // at most 32 frames (kMaxSmallPoses)
        }
        W.key[i] = k;
        W.val[i] = i;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n += __shfl_down_sync(0xffffffffu, n, o);
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(&W.counters[1], n);
}

// staging layout of the batch upload (visfs_ba.cu: pose | point | obs | edge pose | edge point | flags)
__global__ void k_win_emit(Win W, int P, int L, int E, int root_pose, double *pose, double *point, float *obs, int *epose, int *epoint,
                           uint8_t *pfix, uint8_t *lfix, uint8_t *ekind) {
    const int t0 = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    for (int i = t0; i < P * 7; i += stride) pose[i] = W.frame_tq[W.pose_slot[i / 7] * 7 + i % 7];
    for (int i = t0; i < P; i += stride) pfix[i] = (i == root_pose) ? 1 : 0;
    for (int r = t0; r < L; r += stride) {
        const int slot = W.slot_of_rank[r];
        point[3 * r] = W.point_xyz[3 * slot]; point[3 * r + 1] = W.point_xyz[3 * slot + 1]; point[3 * r + 2] = W.point_xyz[3 * slot + 2];
        lfix[r] = W.point_fixed[slot];
    }
    for (int e = t0; e < E; e += stride) {
        const int i = W.val_sorted[e];
        obs[3 * e] = W.ob_obs[3 * i]; obs[3 * e + 1] = W.ob_obs[3 * i + 1]; obs[3 * e + 2] = W.ob_obs[3 * i + 2];
        epose[e] = W.frame_pose[W.ob_frame[i]];
        epoint[e] = W.rank_of_slot[W.ob_point[i]];
        ekind[e] = W.ob_kind[i];
    }
}

// after the solve: optimised poses into the frame table and the result array; features that took part move when displaced by
// less than 5 m (Optimizer.cpp:343-351; fixed ones never moved); culled observations as (feature slot, frame slot) pairs
__global__ void k_win_finish(Win W, Batch B, int P, int L, int E, int write_back, double *pose_out, int *outliers, int outlier_cap) {
    const int t0 = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    const int cur = B.st[0].cur;
    const bool ok = B.st[0].status == 0;
    const double *ps = B.pose + (size_t)cur * B.tot_pose * kPoseStride;
    for (int i = t0; i < P * 7; i += stride) {
        const double v = ps[(i / 7) * kPoseStride + i % 7];
        pose_out[i] = v;
        if (write_back && ok) W.frame_tq[W.pose_slot[i / 7] * 7 + i % 7] = v;
    }
    const double *qs = B.point + (size_t)cur * B.tot_point * 3;
    if (write_back && ok)
        for (int r = t0; r < L; r += stride) {
            const int slot = W.slot_of_rank[r];
            const double dx = W.point_xyz[3 * slot] - qs[3 * r], dy = W.point_xyz[3 * slot + 1] - qs[3 * r + 1], dz = W.point_xyz[3 * slot + 2] - qs[3 * r + 2];
            if (sqrt(dx * dx + dy * dy + dz * dz) < 5.0) {   // uNorm (utilite/include/Math.h:248-251)
                W.point_xyz[3 * slot] = qs[3 * r]; W.point_xyz[3 * slot + 1] = qs[3 * r + 1]; W.point_xyz[3 * slot + 2] = qs[3 * r + 2];
            }
        }
    for (int e = t0; e < E; e += stride)
        if (B.edge_pose[e] & kCulledBit) {
            const int at = atomicAdd(&W.counters[2], 1);
            if (at < outlier_cap) { const int i = W.val_sorted[e]; outliers[2 * at] = W.ob_point[i]; outliers[2 * at + 1] = W.ob_frame[i]; }
        }
}

__global__ void k_win_kill_frame(int *ob_frame, uint8_t *ob_dead, int n_pool, int fslot) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pool; i += gridDim.x * blockDim.x)
        if (ob_frame[i] == fslot) ob_dead[i] = 1;
}

// deltas arrive packed in one staging copy: scatter them into the tables
__global__ void k_win_scatter_points(int n, const int *slot, const double *xyz, const uint8_t *fixed, const long long *id, double *point_xyz,
                                     uint8_t *point_fixed, long long *point_id) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int s = slot[i];
        point_xyz[3 * s] = xyz[3 * i]; point_xyz[3 * s + 1] = xyz[3 * i + 1]; point_xyz[3 * s + 2] = xyz[3 * i + 2];
        point_fixed[s] = fixed[i];
        point_id[s] = id[i];
    }
}

// Removing features / single observations without a host copy of the pool: the delta sets bits in mask[feature slot] (bit f =
// frame slot f, all bits = the whole feature), one pass over the pool marks what the masks cover, the masks are cleared again.
__global__ void k_win_mask_set(unsigned *mask, const int *pslot, const int *fslot /* or null: all frames */, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        atomicOr(&mask[pslot[i]], fslot ? (1u << fslot[i]) : 0xffffffffu);
}
__global__ void k_win_mask_kill(const unsigned *mask, const int *ob_point, const int *ob_frame, uint8_t *ob_dead, int n_pool) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pool; i += gridDim.x * blockDim.x)
        if ((mask[ob_point[i]] >> ob_frame[i]) & 1u) ob_dead[i] = 1;
}
__global__ void k_win_mask_clear(unsigned *mask, const int *pslot, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) mask[pslot[i]] = 0u;
}

// Compaction of the append-only pool on the device: pos = exclusive scan of the live flags
__global__ void k_win_live(const uint8_t *ob_dead, int n_pool, int *live) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= n_pool; i += gridDim.x * blockDim.x) live[i] = (i < n_pool && !ob_dead[i]) ? 1 : 0;
}
__global__ void k_win_compact(const int *live, const int *pos, int n_pool, const int *ob_point, const int *ob_frame, const float *ob_obs,
                              const uint8_t *ob_kind, int *t_point, int *t_frame, float *t_obs, uint8_t *t_kind) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pool; i += gridDim.x * blockDim.x) {
        if (!live[i]) continue;
        const int d = pos[i];
        t_point[d] = ob_point[i]; t_frame[d] = ob_frame[i]; t_kind[d] = ob_kind[i];
        t_obs[3 * d] = ob_obs[3 * i]; t_obs[3 * d + 1] = ob_obs[3 * i + 1]; t_obs[3 * d + 2] = ob_obs[3 * i + 2];
    }
}

}  // namespace wn
}  // namespace visfs

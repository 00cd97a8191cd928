// ba_band.cuh — band chunks: the build of a LARGE window (block-skyline path, ba_large.cuh) through the warp-specialised
// shared-memory kernel of the small windows (ws::k_build_band, ba_build_ws.cuh), without atomics.
//
// In a map whose landmarks are seen by a few consecutive key frames (a trajectory, BASELINE config C4) the landmarks that
// start at the same few frames touch only a handful of poses.  Once per pass:
//   1. the landmarks are already sorted by their first pose (k_lm_first + radix sort, ba_large.cuh);
//   2. k_band_flags / scans cut that order into chunks: kBandKeys consecutive first-pose values per chunk, at most
//      kBandMaxLm landmarks;
//   3. k_band_chunk lists the poses of every chunk in ascending order (<= kBandPoses, else the chunk is left to
//      lg::k_build_large_run), numbers the chunk's edges locally and copies the edge records into the sorted order;
//   4. the tile table of the chunks (walk_tiles over the sorted offsets);
//   5. k_band_entries + one radix sort: for every skyline block, the chunk slots that contribute to it, in chunk order.
// Per LM trial: ws::k_build_band (one CTA per chunk -> one partial system per chunk), then k_band_gather adds the partials
// into the reduce buffer — every destination has ONE owner thread that adds its sources in chunk order, so the sums are
// reproducible and there is no red.global.add.f64 on this path.
#pragma once
#include "ba_build_ws.cuh"

namespace visfs {
namespace bd {

using ws::Band;
using ws::BandChunk;
using ws::kBandPartStride;
using ws::kBandPoses;

constexpr int kBandKeys = 1;          // first-pose values per chunk.  Measured on C4 (ms per build launch): 1 -> 0.90, 2 -> 1.00, 3 -> 1.08,
                                      // 4 -> 1.15: inside a tile all landmarks start at the same frame, so the pose pairs that
                                      // no landmark of the tile covers leave their owner threads idle
constexpr int kBandMaxLm = 2048;      // landmarks per chunk
constexpr int kChunkCache = 12032;    // edges of a chunk whose pose indices k_band_chunk keeps in shared memory
constexpr int kBandFlag = 0x80;       // lm_rec.w: this landmark is NOT in a band chunk (lg::k_build_large_run takes it)

// degree of the landmarks in the sorted order (-> exclusive scan = sorted_off) and "a new first pose starts here"
__global__ void k_band_deg(const int4 *__restrict__ rec, const int *__restrict__ key, int L, int *deg, int *newkey) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= L; i += gridDim.x * blockDim.x) {
        deg[i] = (i < L) ? rec[i].z : 0;
        newkey[i] = (i < L && (i == 0 || key[i] != key[i - 1])) ? 1 : 0;
    }
}

// chunk starts: rank = inclusive scan of newkey (1-based number of the landmark's first-pose value)
__global__ void k_band_flags(const int *__restrict__ rank, int L, int *start, int keys, int max_lm) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= L; i += gridDim.x * blockDim.x)
        start[i] = (i < L && (i == 0 || (rank[i] - 1) / keys != (rank[i - 1] - 1) / keys || i % max_lm == 0)) ? 1 : 0;
}

// cid = INCLUSIVE scan of start: landmark i belongs to chunk cid[i] - 1; the chunk's range
__global__ void k_band_ranges(const int *__restrict__ start, const int *__restrict__ cid, int L, BandChunk *chunk) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L; i += gridDim.x * blockDim.x) {
        if (start[i]) chunk[cid[i] - 1].lm0 = i;
        if (i == L - 1 || start[i + 1]) chunk[cid[i] - 1].lm1 = i + 1;
    }
}

// one CTA per chunk: its poses in ascending order (rounds of "smallest pose above the last one"), the local pose number of
// every edge, the sorted edge tables.  A chunk with more than kBandPoses poses gets n_pose = -1 and its landmarks the flag.
__global__ void __launch_bounds__(256) k_band_chunk(Batch B, int4 *rec, const int *__restrict__ sorted_off, BandChunk *chunk, int *chunk_pose,
                                                    int *s_pw, int *s_gl, int *s_sl, double *s_ou, double *s_ov, double *s_our,
                                                    int *counts /* [0] landmarks in band chunks */) {
    __shared__ int s_min[8];
    __shared__ int s_list[kBandPoses + 1];
    __shared__ int s_p[kChunkCache];
    const int c = blockIdx.x, tid = threadIdx.x;
    const int lm0 = chunk[c].lm0, lm1 = chunk[c].lm1;
    const WinDesc &wd = B.win[0];
    const int eb = sorted_off[lm0], ne = sorted_off[lm1] - eb;
    const bool cached = ne <= kChunkCache;
    if (cached) {
        for (int i = lm0 + tid; i < lm1; i += 256) {
            const int4 r = rec[i];
            const int base = sorted_off[i] - eb;
            for (int k = 0; k < r.z; ++k) s_p[base + k] = B.edge_pose[r.y + k] & kPoseMask;
        }
        __syncthreads();
    }
    int last = -1, n = 0;
    for (;;) {
        int mn = 0x7fffffff;
        if (cached) {
            for (int k = tid; k < ne; k += 256) { const int p = s_p[k]; if (p > last) mn = min(mn, p); }
        } else {
            for (int i = lm0 + tid; i < lm1; i += 256) {
                const int4 r = rec[i];
                for (int k = 0; k < r.z; ++k) {
                    const int p = B.edge_pose[r.y + k] & kPoseMask;
                    if (p > last) mn = min(mn, p);
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        if ((tid & 31) == 0) s_min[tid >> 5] = mn;
        __syncthreads();
        mn = s_min[0];
#pragma unroll
        for (int w = 1; w < 8; ++w) mn = min(mn, s_min[w]);
        __syncthreads();
        if (mn == 0x7fffffff) break;
        if (n == kBandPoses) { n = -1; break; }
        if (tid == 0) s_list[n] = mn;
        ++n;
        last = mn;
    }
    __syncthreads();
    if (tid == 0) {
        int F = 0;
        for (int i = 0; i < n; ++i) {
            chunk_pose[c * kBandPoses + i] = s_list[i];
            if (B.pose_hidx[wd.pose_off + s_list[i]] >= 0) ++F;
        }
        chunk[c].n_pose = n;
        chunk[c].F = (n >= 0) ? F : 0;
        if (n >= 0) atomicAdd(&counts[0], lm1 - lm0);
    }
    for (int i = lm0 + tid; i < lm1; i += 256) {
        int4 r = rec[i];
        const int base = sorted_off[i];
        for (int k = 0; k < r.z && n >= 0; ++k) {
            const int e = r.y + k, pw = B.edge_pose[e], p = pw & kPoseMask;
            int loc = 0;
            while (loc < n - 1 && s_list[loc] != p) ++loc;
            s_pw[base + k] = (pw & ~kPoseMask) | loc;
            s_gl[base + k] = wd.point_off + B.edge_point[e];
            s_sl[base + k] = i;
            s_ou[base + k] = B.obs_u[e]; s_ov[base + k] = B.obs_v[e]; s_our[base + k] = B.obs_r[e];
        }
        if (n < 0) { r.w |= kBandFlag; rec[i] = r; }
    }
}

// the landmarks outside the band chunks, compacted in sorted order for lg::k_build_large_run (pos = exclusive scan of flag)
__global__ void k_band_rest_flag(const int4 *__restrict__ rec, int L, int *flag) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= L; i += gridDim.x * blockDim.x) flag[i] = (i < L && (rec[i].w & kBandFlag)) ? 1 : 0;
}
__global__ void k_band_rest_fill(const int4 *__restrict__ rec, const int *__restrict__ flag, const int *__restrict__ pos, int L, int4 *rest) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L; i += gridDim.x * blockDim.x)
        if (flag[i]) rest[pos[i]] = rec[i];
}

__global__ void k_band_count_tiles(const BandChunk *__restrict__ chunk, int n_chunk, const int *__restrict__ sorted_off, int *ntiles, int *npair, int *npose) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > n_chunk) return;
    const bool ok = c < n_chunk && chunk[c].n_pose >= 0;
    ntiles[c] = ok ? walk_tiles<false>(sorted_off, chunk[c].lm0, chunk[c].lm1, nullptr, kTileLm, kTileEdges) : 0;
    npair[c] = ok ? chunk[c].F * (chunk[c].F + 1) / 2 : 0;   // gather entries: pair blocks ...
    npose[c] = ok ? chunk[c].F : 0;                           // ... and per-pose sums
}

__global__ void k_band_fill_tiles(const BandChunk *__restrict__ chunk, int n_chunk, const int *__restrict__ sorted_off, const int *__restrict__ tile_off,
                                  Tile *tiles) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunk || chunk[c].n_pose < 0) return;
    walk_tiles<true>(sorted_off, chunk[c].lm0, chunk[c].lm1, tiles + tile_off[c], kTileLm, kTileEdges);
}

// gather entries of chunk c: key = destination skyline block, value = (hessian index of the row pose << 32) | chunk << 10 |
// kind << 8 | index, kind 0: block pair (i < j), 1: diagonal pair (i, i), 2: the per-pose sums of free pose i.  The pair
// entries of all chunks come first in the input, then the per-pose entries: the stable sort keeps that order per block.
__global__ void k_band_entries(Batch B, const BandChunk *__restrict__ chunk, int n_chunk, const int *__restrict__ chunk_pose,
                               const int *__restrict__ ent_off, int n_ent_pairs_total, const int *__restrict__ pose_ent_off, int *key,
                               unsigned long long *val) {
    const int c = blockIdx.x;
    if (chunk[c].n_pose < 0) return;
    __shared__ int s_h[kBandPoses];
    const WinDesc &wd = B.win[0];
    if (threadIdx.x == 0) {
        int f = 0;
        for (int i = 0; i < chunk[c].n_pose; ++i) {
            const int h = B.pose_hidx[wd.pose_off + chunk_pose[c * kBandPoses + i]];
            if (h >= 0) s_h[f++] = h;
        }
    }
    __syncthreads();
    const int F = chunk[c].F, np = F * (F + 1) / 2;
    for (int pt = threadIdx.x; pt < np + F; pt += blockDim.x) {
        if (pt < np) {
            int i = 0, base = 0;
            while (base + (F - i) <= pt) { base += F - i; ++i; }
            const int j = i + (pt - base);
            const int ha = min(s_h[i], s_h[j]), hb = max(s_h[i], s_h[j]);
            const long long blk = B.sky_off[hb] - B.sky_first[hb] + ha;
            const int o = ent_off[c] + pt;
            key[o] = (int)blk;
            val[o] = ((unsigned long long)(unsigned)(s_h[i] > s_h[j] ? 1 : 0) << 63) | ((unsigned long long)(unsigned)hb << 32) |
                     ((unsigned long long)c << 10) | ((unsigned long long)(i == j ? 1 : 0) << 8) | (unsigned long long)pt;
        } else {
            const int i = pt - np, h = s_h[i];
            const long long blk = B.sky_off[h] - B.sky_first[h] + h;
            const int o = n_ent_pairs_total + pose_ent_off[c] + i;
            key[o] = (int)blk;
            val[o] = ((unsigned long long)(unsigned)h << 32) | ((unsigned long long)c << 10) | (2ull << 8) | (unsigned long long)i;
        }
    }
}

// segment heads of the sorted keys
__global__ void k_band_seg_flags(const int *__restrict__ key, int n, int *flag) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += gridDim.x * blockDim.x)
        flag[i] = (i < n && (i == 0 || key[i] != key[i - 1])) ? 1 : 0;
}
__global__ void k_band_seg_starts(const int *__restrict__ flag, const int *__restrict__ sid, int n, int *seg_start) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += gridDim.x * blockDim.x) {
        if (i < n && flag[i]) seg_start[sid[i]] = i;
        if (i == n) seg_start[sid[n]] = n;
    }
}

// per LM trial: every destination element has one owner thread (36 threads per skyline block that receives anything), the
// sources are added in sorted (= chunk) order.  Destination element q = row * 6 + col of the lower block (h_b, h_a); a pair
// block of ws::k_build_band holds Yn_i W_j^T as [row of pose i][col of pose j], so element (row, col) of the lower block
// (h_j, h_i) is its entry [col][row].  Diagonal blocks keep their lower triangle (row >= col), like lg::k_build_large_run.
__global__ void __launch_bounds__(252) k_band_gather(Batch B, Band bd, const int *__restrict__ key, const unsigned long long *__restrict__ val,
                                                     const int *__restrict__ seg_start, int n_seg) {
    const LMState &st = B.st[0];
    if (st.done) return;
    const int s = blockIdx.x * 7 + threadIdx.x / 36, q = threadIdx.x % 36;
    if (s >= n_seg) return;
    const int row = q / 6, col = q - row * 6;
    const int e0 = seg_start[s], e1 = seg_start[s + 1];
    double acc = 0.0, accv = 0.0;
    int h = -1;
    for (int e = e0; e < e1; ++e) {
        const unsigned long long v = val[e];
        const int c = (int)((v >> 10) & 0x3fffff), kind = (int)((v >> 8) & 3), idx = (int)(v & 0xff);
        const double *part = bd.part + (size_t)c * kBandPartStride;
        if (kind == 0) {
            const bool swapped = (v >> 63) != 0;   // the chunk numbers its poses in ascending order, so this never happens; kept general
            acc += part[idx * 36 + (swapped ? row * 6 + col : col * 6 + row)];
        } else if (kind == 1) {
            if (row >= col) acc += part[idx * 36 + col * 6 + row];
        } else {
            const int F = bd.chunk[c].F;
            const double *pa = part + F * (F + 1) / 2 * 36 + idx * kHStride;
            if (row >= col) acc += pa[hd_index(col, row)];
            if (q < 12) accv += pa[21 + q];
            h = (int)((v >> 32) & 0x7fffffff);
        }
    }
    B.red[(size_t)key[e0] * 36 + q] += acc;
    if (h >= 0 && q < 12) {
        double *dst = B.red + (q < 6 ? B.red_g_off : B.red_bp_off);
        dst[6 * (size_t)h + (q < 6 ? q : q - 6)] += accv;
    }
}

}  // namespace bd
}  // namespace visfs

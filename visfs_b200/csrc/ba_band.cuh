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
constexpr int kBandMaxLm = 608;       // landmarks per chunk (32 tiles of 19 landmarks x 10 edges): bounds the longest chunk of a launch
constexpr int kChunkCache = 12032;    // edges of a chunk whose pose indices k_band_chunk keeps in shared memory
constexpr int kBandFlag = 0x80;       // lm_rec.w: this landmark is NOT in a band chunk (lg::k_build_large_run takes it)
constexpr uint8_t kInBand = 0x40;     // lm_flags: this landmark IS in a band chunk (lg::k_update_large skips it); rewritten every pass by k_struct_lm

// degree of the landmarks in the sorted order (-> exclusive scan = sorted_off) and "a new first pose starts here"
__global__ void k_band_deg(const int4 *__restrict__ rec, const int *__restrict__ key, int L, int *deg, int *newkey) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= L; i += gridDim.x * blockDim.x) {
        deg[i] = (i < L) ? rec[i].z : 0;
        newkey[i] = (i < L && (i == 0 || key[i] != key[i - 1])) ? 1 : 0;
    }
}

// chunk starts: rank = inclusive scan of newkey (1-based number of the landmark's first-pose value).  A group = `keys`
// consecutive values; gstart[i] = i where a group starts, else 0 -> an inclusive MAX scan gives every landmark the start of its
// group, and groups longer than max_lm are cut every max_lm landmarks counted from there.
__global__ void k_band_gstart(const int *__restrict__ rank, int L, int keys, int *gstart) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= L; i += gridDim.x * blockDim.x)
        gstart[i] = (i < L && i > 0 && (rank[i] - 1) / keys != (rank[i - 1] - 1) / keys) ? i : 0;
}
__global__ void k_band_flags(const int *__restrict__ gpos, int L, int *start, int max_lm) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= L; i += gridDim.x * blockDim.x)
        start[i] = (i < L && (i - gpos[i]) % max_lm == 0) ? 1 : 0;
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
    __shared__ int s_ref[kMaxSmallPoses + 1];   // degree and local pose numbers of the chunk's first landmark: the others are compared with it
    if (tid == 0) {
        int F = 0;
        for (int i = 0; i < n; ++i) {
            chunk_pose[c * kBandPoses + i] = s_list[i];
            if (B.pose_hidx[wd.pose_off + s_list[i]] >= 0) ++F;
        }
        chunk[c].n_pose = n;
        chunk[c].F = (n >= 0) ? F : 0;
        if (n >= 0) atomicAdd(&counts[0], lm1 - lm0);
        const int4 r0 = rec[lm0];
        s_ref[0] = (n >= 0 && r0.z <= kMaxSmallPoses) ? r0.z : -1;
        for (int k = 0; k < r0.z && s_ref[0] >= 0; ++k) {
            const int p = B.edge_pose[r0.y + k] & kPoseMask;
            int loc = 0;
            while (loc < n - 1 && s_list[loc] != p) ++loc;
            s_ref[1 + k] = loc;
        }
    }
    __syncthreads();
    bool same = s_ref[0] > 0;
    for (int i = lm0 + tid; i < lm1; i += 256) {
        int4 r = rec[i];
        const int base = sorted_off[i];
        if (r.z != s_ref[0]) same = false;
        for (int k = 0; k < r.z && n >= 0; ++k) {
            const int e = r.y + k, pw = B.edge_pose[e], p = pw & kPoseMask;
            int loc = 0;
            while (loc < n - 1 && s_list[loc] != p) ++loc;
            if (same && loc != s_ref[1 + k]) same = false;
            s_pw[base + k] = (pw & ~kPoseMask) | loc;
            s_gl[base + k] = wd.point_off + B.edge_point[e];
            s_sl[base + k] = i;
            s_ou[base + k] = B.obs_u[e]; s_ov[base + k] = B.obs_v[e]; s_our[base + k] = B.obs_r[e];
        }
        if (n < 0) { r.w |= kBandFlag; rec[i] = r; }
        else B.lm_flags[wd.point_off + r.x] |= kInBand;
    }
    same = __syncthreads_and(same);
    if (tid == 0) { chunk[c].regular = (same && n >= 0) ? 1 : 0; chunk[c].pad = 0; }
}

// the landmarks outside the band chunks, compacted in sorted order for lg::k_build_large_run (pos = exclusive scan of flag)
__global__ void k_band_rest_flag(const int4 *__restrict__ rec, int L, int *flag) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= L; i += gridDim.x * blockDim.x) flag[i] = (i < L && (rec[i].w & kBandFlag)) ? 1 : 0;
}
__global__ void k_band_rest_fill(const int4 *__restrict__ rec, const int *__restrict__ flag, const int *__restrict__ pos, int L, int4 *rest) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L; i += gridDim.x * blockDim.x)
        if (flag[i]) rest[pos[i]] = rec[i];
}

__global__ void k_band_count_tiles(const BandChunk *__restrict__ chunk, int n_chunk, const int *__restrict__ sorted_off, int *ntiles, int *npair, int *npose,
                                   int *cost, int *cidx) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > n_chunk) return;
    const bool ok = c < n_chunk && chunk[c].n_pose >= 0;
    ntiles[c] = ok ? walk_tiles<false>(sorted_off, chunk[c].lm0, chunk[c].lm1, nullptr, kTileLm, kTileEdges) : 0;
    // launch order: longest first.  A tile costs more the more block pairs its chunk has (fewer landmark groups in stage C).
    if (c < n_chunk) { cost[c] = 0x7fffffff - (ok ? ntiles[c] * (64 + chunk[c].F * (chunk[c].F + 1) / 2) : 0); cidx[c] = c; }
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

// ------------------------------------------------------------------------------------------------
// k_update_band: lg::k_update_large on the band chunks — landmark back-substitution, point oplus, chi2 of the trial state —
// with the chunk's poses (accepted and trial), pose steps and flags in shared memory under their chunk-local numbers and
// the edge records read from the sorted copies.  One CTA per chunk, every warp walks a contiguous piece of the chunk's
// landmarks in warp tiles (whole landmarks, <= 32 edges); offsets and records of the NEXT tile are requested while the
// current one is computed.  part2[2 c], part2[2 c + 1] = chi2 / scale partials of chunk c.
// ------------------------------------------------------------------------------------------------
struct UpdBandSmem {
    double pose[ws::kMaxPosesWs * kPoseSm];
    double poseT[ws::kMaxPosesWs * kPoseSm];
    double xp[ws::kMaxPosesWs * 6];
    double H[kUpdWarps][32 * kHs];
    double lm[kUpdWarps][kWtLm * 12];
    double red[32];
    int lmoff[kUpdWarps][kWtLm + 1];
    int hidx[ws::kMaxPosesWs];
    unsigned char pflag[ws::kMaxPosesWs];
};

__global__ void __launch_bounds__(kUpdThreads, 2) k_update_band(Batch B, Band bd, double *part2) {
    __shared__ UpdBandSmem sm;
    const WinDesc &wd = B.win[0];
    const LMState &st = B.st[0];
    if (st.done) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = bd.order[blockIdx.x];   // longest chunks first
    const BandChunk bc = bd.chunk[c];
    if (bc.n_pose < 0) {
        if (tid == 0) { part2[2 * (size_t)c] = 0.0; part2[2 * (size_t)c + 1] = 0.0; }
        return;
    }
    const int *__restrict__ cpose = bd.chunk_pose + (size_t)c * kBandPoses;
    const int cur = st.cur;
    const double lambda = (wd.trust == 0) ? (B.dbg_lambda >= 0.0 ? B.dbg_lambda : st.lambda) : 0.0;
    const Intr K = load_intr(wd);
    const double *__restrict__ gpose = B.pose + ((size_t)cur * B.tot_pose + wd.pose_off) * kPoseStride;
    const double *__restrict__ gposeT = B.pose + ((size_t)(1 - cur) * B.tot_pose + wd.pose_off) * kPoseStride;
    const double *__restrict__ gpoint = B.point + (size_t)cur * B.tot_point * 3;
    double *__restrict__ gpointT = B.point + (size_t)(1 - cur) * B.tot_point * 3;
    for (int i = tid; i < bc.n_pose * kPoseStride; i += kUpdThreads) {
        const int j = (i >> 4) * kPoseSm + (i & 15);
        const size_t g = (size_t)cpose[i >> 4] * kPoseStride + (i & 15);
        sm.pose[j] = gpose[g]; sm.poseT[j] = gposeT[g];
    }
    if (tid < 32) {
        const int h = (tid < bc.n_pose) ? B.pose_hidx[wd.pose_off + cpose[tid]] : -1;
        if (tid < bc.n_pose) {
            sm.hidx[tid] = h;
            sm.pflag[tid] = B.pose_flags[wd.pose_off + cpose[tid]];
#pragma unroll
            for (int a = 0; a < 6; ++a) sm.xp[tid * 6 + a] = (h >= 0) ? B.xp[(size_t)wd.pose_off * 6 + 6 * (size_t)h + a] : 0.0;
        }
    }
    __syncthreads();
    double *Hs = sm.H[warp], *Ls = sm.lm[warp];
    int *lmoff = sm.lmoff[warp];
    double chi_acc = 0.0, scale_acc = 0.0;

    const int n_lm = bc.lm1 - bc.lm0, per = (n_lm + kUpdWarps - 1) / kUpdWarps;
    const int l_beg = min(bc.lm0 + warp * per, bc.lm1), l_end = min(l_beg + per, bc.lm1);
    // lane l: offset of landmark lt + l (lanes 0 .. 8) and its record (landmark, first edge, degree, flags)
    auto req_off = [&](int lt) { return (lt + lane <= l_end && lane <= kWtLm) ? bd.sorted_off[lt + lane] : 0x7fffffff; };
    auto req_rec = [&](int lt) { return (lt + lane < l_end && lane < kWtLm) ? bd.lm_rec[lt + lane] : make_int4(0, 0, 0, 0); };
    int off_abs = req_off(l_beg);
    int4 rec = req_rec(l_beg);
    for (int lt = l_beg; lt < l_end;) {
        const int navail = min(kWtLm, l_end - lt);
        const int e0 = __shfl_sync(0xffffffffu, off_abs, 0);
        const int off_l = (lane <= navail) ? off_abs - e0 : 0x7fff;
        int ntl = 1;
#pragma unroll
        for (int l = 2; l <= kWtLm; ++l) {
            const int v = __shfl_sync(0xffffffffu, off_l, l);
            ntl += (l <= navail && v <= 32) ? 1 : 0;
        }
        const int ne = min(__shfl_sync(0xffffffffu, off_l, ntl), 32);
        const int lf_l = (lane < ntl) ? rec.w : 0;
        const int gl_l = rec.x;
        const int g_of = __shfl_sync(0xffffffffu, gl_l, min(lane / 3, kWtLm - 1));
        const double pt = (lane < 3 * ntl) ? gpoint[3 * (size_t)g_of + (lane - 3 * (lane / 3))] : 0.0;
        int pw = 0;
        double ou = 0, ov = 0, our = 0;
        if (lane < ne) {
            const int k = e0 + lane;
            pw = bd.s_pw[k];
            ou = bd.s_ou[k]; ov = bd.s_ov[k]; our = bd.s_our[k];
        }
        // the next tile's offsets and records: requested now, used in the next iteration
        const int off_next = req_off(lt + ntl);
        const int4 rec_next = req_rec(lt + ntl);
        if (lane <= ntl) lmoff[lane] = min(off_l, 32);
        int tl = 0;
#pragma unroll
        for (int l = 1; l < kWtLm; ++l) {
            const int v = __shfl_sync(0xffffffffu, off_l, l);
            tl += (l < ntl && v <= lane) ? 1 : 0;
        }
        const int lf = __shfl_sync(0xffffffffu, lf_l, tl);
        const int gl = __shfl_sync(0xffffffffu, gl_l, tl);
        const double px = __shfl_sync(0xffffffffu, pt, 3 * tl), py = __shfl_sync(0xffffffffu, pt, 3 * tl + 1),
                     pz = __shfl_sync(0xffffffffu, pt, 3 * tl + 2);
        bool act = false, mono = false, lmfree = false;
        int p = 0;
        double hl[12];
#pragma unroll
        for (int q = 0; q < 12; ++q) hl[q] = 0.0;
        if (lane < ne) {
            p = pw & kPoseMask;
            mono = (pw & kMonoBit) != 0;
            lmfree = (lf & kInHessian) != 0;
            act = !(pw & kCulledBit) && !((lf & kFixed) && (sm.pflag[p] & kFixed));
            if (act && lmfree)
                upd_edge_terms(sm.pose + p * kPoseSm, px, py, pz, ou, ov, our, mono, K, sm.hidx[p] >= 0 ? sm.xp + p * 6 : nullptr, hl);
#pragma unroll
            for (int q = 0; q < 12; ++q) Hs[lane * kHs + q] = hl[q];
        }
        __syncwarp();
        for (int task = lane; task < ntl * 12; task += 32) {
            const int l = task / 12, q = task - l * 12;
            double sacc = 0.0;
            for (int e = lmoff[l]; e < lmoff[l + 1]; ++e) sacc += Hs[e * kHs + q];
            Ls[task] = sacc;
        }
        __syncwarp();
        if (lane < ne) {
            double np0 = px, np1 = py, np2 = pz;
            if (lmfree) {
                double xl[3];
                const double sc = upd_point_step(Ls + tl * 12, lambda, xl);
                np0 = px + xl[0]; np1 = py + xl[1]; np2 = pz + xl[2];
                if (lane == lmoff[tl]) {      // first edge of the landmark: owner of the point
                    gpointT[3 * (size_t)gl] = np0; gpointT[3 * (size_t)gl + 1] = np1; gpointT[3 * (size_t)gl + 2] = np2;
                    scale_acc += sc;
                }
            }
            if (act) {
                double r0, r1, r2, rho, wgt;
                edge_residual(sm.poseT + p * kPoseSm, np0, np1, np2, ou, ov, our, mono, K, r0, r1, r2);
                huber((r0 * r0 + r1 * r1 + r2 * r2) * K.inv_pv, K.delta, rho, wgt);
                chi_acc += rho;
            }
        }
        __syncwarp();
        lt += ntl;
        off_abs = off_next; rec = rec_next;
    }
    const double chi = block_sum(chi_acc, sm.red);
    const double sc = block_sum(scale_acc, sm.red);
    if (tid == 0) { part2[2 * (size_t)c] = chi; part2[2 * (size_t)c + 1] = sc; }
}

// chi2 / scale partials of lg::k_update_large (na CTAs, may be 0) and of k_update_band (nb chunks), in that order -> out[0], out[1]
__global__ void k_band_fold(Batch B, const double *__restrict__ a, int na, const double *__restrict__ b, int nb, double *out) {
    __shared__ double red[32];
    if (B.st[0].done) return;
    double x = 0.0, y = 0.0;
    for (int i = threadIdx.x; i < na + nb; i += blockDim.x) {
        const double *src = (i < na) ? a + 2 * (size_t)i : b + 2 * (size_t)(i - na);
        x += src[0]; y += src[1];
    }
    const double sx = block_sum(x, red);
    const double sy = block_sum(y, red);
    if (threadIdx.x == 0) { out[0] = sx; out[1] = sy; }
}

}  // namespace bd
}  // namespace visfs

// pe_blob.cu -- +/- cutoff thresholding fused with 26-connected component labelling of the difference map.
//
// Replaces  createFullCrsList (pdb_eda/cutils.pyx:185-203) + createCrsLists (pdb_eda/cutils.pyx:41-70) +
// DensityBlob.fromCrsList (pdb_eda/ccp4.py:522-545), i.e. DensityMatrix.createFullBlobList (pdb_eda/ccp4.py:463-485),
// for the green (+cutoff) and red (-cutoff) lists of pdb_eda/densityAnalysis.py:392-412 in one pass over the map.
//
// Design (B200).  The map is read exactly once, by an HBM-bound streaming kernel that thresholds both signs and
// writes two bit planes (1 bit per voxel and class, 1/16 of the bytes read).  The bit planes are stored in the
// REFERENCE'S scan order -- column slowest, section fastest, 32 sections per word -- so that
//   * a popcount prefix sum over the words gives every foreground voxel its position in createFullCrsList's list
//     (deterministic, order-preserving compaction without sorting), and
//   * union-find over those positions with "smaller id wins" makes every blob's root its first voxel in that
//     order, so ranking the roots yields the reference's blob order directly (canonical min-index relabelling).
// Everything after the streaming kernel touches only the bit planes (N/8 bytes) and the sparse foreground.
// Runs of set bits along the section axis are linked at initialisation (parent = predecessor), the remaining
// 12 predecessor neighbours are merged with lock-free atomicMin hooking.
//
// The transposition (memory order is column-fastest, bit order is section-fastest) costs nothing: a thread owns
// 4 adjacent columns of one row and walks 32 sections, so the 32 loads it issues are independent 16-byte loads
// that are contiguous across the warp (512 B per warp per section), and the word it builds is complete in
// registers.
#include <cooperative_groups.h>
#include "pe_common.cuh"

namespace cg = cooperative_groups;

namespace pe {

constexpr int kBmpTx = 32;         // threads along columns (x VEC columns each)
constexpr int kBmpTy = 8;          // threads along rows
constexpr int kSparseThreads = 512;
constexpr int kSparseMaxBlocks = 1024;  // size of the per-block partial-sum arrays

struct BlobPlan {
    int U0, U1, U2, W;    // unique columns, rows, sections; words per (column,row)
    int64_t nwords;       // U0*U1*W words per class
    int64_t nwords_pad;   // the same rounded up to 64: distance between the two bit planes
    int64_t cap;          // foreground capacity per class
    // workspace carve-up
    int64_t off_bmp, off_base, off_parent, off_rank, off_sums, total;
};

static BlobPlan make_plan(const pe_geom *g, int64_t cap) {
    BlobPlan p;
    p.U0 = g->unique_ncrs[0];
    p.U1 = g->unique_ncrs[1];
    p.U2 = g->unique_ncrs[2];
    p.W = ((p.U2 + 31) / 32 + 7) / 8 * 8;  // padded: see threshold_bitmap_kernel
    p.nwords = (int64_t)p.U0 * p.U1 * p.W;
    p.nwords_pad = align_up(p.nwords, 64);
    p.cap = cap;
    int64_t o = 0;
    p.off_bmp = o;
    o += align_up(2 * p.nwords_pad * 4, 256);
    p.off_base = o;
    o += align_up(2 * p.nwords_pad * 4, 256);
    p.off_parent = o;
    o += align_up(2 * cap * 4, 256);
    p.off_rank = o;
    o += align_up(2 * cap * 4, 256);
    p.off_sums = o;
    o += align_up(2 * kSparseMaxBlocks * 4, 256);
    p.total = o;
    return p;
}

// ------------------------------------------------------------------------------------------------ K1: stream
// W (words per (column, row)) is padded to a multiple of 8, so the kWordsPerThread = 4 words a thread produces for
// one (column, row) are one aligned 16-byte store, and the two thread blocks that share a 32-byte sector are
// neighbours in launch order (the word-group index is blockIdx.x, the fastest-varying block index): every bit-plane
// sector reaches DRAM as one full write.  Scattered 4-byte word stores cost a read-modify-write per sector once the
// bit planes no longer fit L2 (768^3 ran at 28 % of the HBM peak, tune log in profiles/r01_threshold_tuning.md).
constexpr int kWordsPerThread = 4;

template <int VEC>
__global__ void __launch_bounds__(384)
    threshold_bitmap_kernel(const float *__restrict__ rho, int NC, int NR, int U0, int U1, int U2, int W, float cpos,
                            float cneg, bool use_pos, bool use_neg, uint32_t *__restrict__ bmp_pos,
                            uint32_t *__restrict__ bmp_neg) {
    const int c = (blockIdx.y * blockDim.x + threadIdx.x) * VEC;
    const int r = blockIdx.z * blockDim.y + threadIdx.y;
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.y == 0) {  // padding between / after the planes
        const int64_t nwords = (int64_t)U0 * U1 * W;
        for (int64_t i = nwords + threadIdx.x; i < nwords + 64; i += blockDim.x) {
            if (i < (nwords + 63) / 64 * 64) {
                bmp_pos[i] = 0u;
                bmp_neg[i] = 0u;
            }
        }
    }
    if (c >= U0 || r >= U1) return;
    const int w_begin = blockIdx.x * kWordsPerThread;
    const int64_t plane = (int64_t)NR * NC;
    const float *col = rho + (int64_t)r * NC + c;
    uint32_t accp[VEC][kWordsPerThread], accn[VEC][kWordsPerThread];
#pragma unroll
    for (int q = 0; q < kWordsPerThread; ++q) {
        const int w = w_begin + q;
        uint32_t pos[VEC], neg[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) pos[i] = neg[i] = 0u;
        const int s0 = w * 32;
        const int nbits = min(32, U2 - s0);  // <= 0 in the padding words
        const float *p = col + (int64_t)s0 * plane;
        if (nbits == 32) {
            if (VEC == 4) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {  // two batches of 16 independent 16-byte loads
                    float4 v[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        v[j] = __ldg(reinterpret_cast<const float4 *>(p + (int64_t)(h * 16 + j) * plane));
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float e[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            pos[i] |= (e[i] >= cpos) ? (1u << (h * 16 + j)) : 0u;
                            neg[i] |= (e[i] <= cneg) ? (1u << (h * 16 + j)) : 0u;
                        }
                    }
                }
            } else {
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __ldg(p + (int64_t)j * plane);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    pos[0] |= (v[j] >= cpos) ? (1u << j) : 0u;
                    neg[0] |= (v[j] <= cneg) ? (1u << j) : 0u;
                }
            }
        } else {
            for (int j = 0; j < nbits; ++j) {
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    const float e = (i == 0 || c + i < NC) ? __ldg(p + (int64_t)j * plane + i) : 0.f;
                    pos[i] |= (e >= cpos) ? (1u << j) : 0u;
                    neg[i] |= (e <= cneg) ? (1u << j) : 0u;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            accp[i][q] = use_pos ? pos[i] : 0u;
            accn[i][q] = use_neg ? neg[i] : 0u;
        }
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        if (c + i < U0) {
            const int64_t widx = ((int64_t)(c + i) * U1 + r) * W + w_begin;  // multiple of 4 words: 16-byte aligned
            *reinterpret_cast<uint4 *>(bmp_pos + widx) = make_uint4(accp[i][0], accp[i][1], accp[i][2], accp[i][3]);
            *reinterpret_cast<uint4 *>(bmp_neg + widx) = make_uint4(accn[i][0], accn[i][1], accn[i][2], accn[i][3]);
        }
    }
}

// ------------------------------------------------------------------------------------------------ sparse stage
// Everything after the streaming kernel touches only the bit planes and the sparse foreground, for BOTH signs at
// once: the planes are scanned as one concatenated word array, so green voxels occupy positions [0, n0) and red
// voxels [n0, n0 + n1) of one index space, and one union-find / one root ranking serves both.  The stages are
// separated by grid-wide barriers inside ONE cooperative kernel (grid = resident blocks, cg::grid_group::sync)
// instead of ten tiny launches per sign.
// Stage boundaries of the last blob_sparse_kernel launch (globaltimer ns, written by block 0): a cheap built-in
// diagnostic, read with pe_blob_stage_times().
__device__ unsigned long long g_stage_ns[12];
__device__ __forceinline__ void stamp(int i) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_stage_ns[i] = t;
    }
}

struct SparseArgs {
    const float *rho;
    int NC, NR, U1, U2, W;
    int64_t nwords, nwords_pad, cap, cap_blobs;
    const uint32_t *bmp;   // two planes, nwords_pad apart
    uint32_t *base;        // per word: position of its first set bit in the concatenated voxel list
    uint32_t *parent;      // 2 * cap
    uint32_t *rank;        // 2 * cap (valid at roots)
    uint32_t *sums;        // 2 * kSparseMaxBlocks block partials
    int64_t *counts;       // n_fg0, n_blobs0, n_fg1, n_blobs1, overflow
    uint32_t *key;         // outputs, class k at k * cap
    float *value;
    int32_t *label;
    double *stats;         // class k at k * cap_blobs * 8
    bool use_pos, use_neg;
};

__device__ __forceinline__ uint32_t block_sum_u32(uint32_t v, uint32_t *smem /* >= 32 */) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = (uint32_t)warp_sum((int)v);
    __syncthreads();
    if (lane == 0) smem[w] = v;
    __syncthreads();
    uint32_t t = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += smem[i];
    return t;
}

// Exclusive prefix of sums[0 .. nb) at index b (every block scans the few hundred partials itself).
__device__ __forceinline__ void grid_prefix(const uint32_t *sums, int nb, int b, uint32_t *smem, uint32_t &before, uint32_t &total) {
    uint32_t mine = 0, all = 0;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) {
        const uint32_t v = sums[i];
        all += v;
        if (i < b) mine += v;
    }
    before = block_sum_u32(mine, smem);
    total = block_sum_u32(all, smem);
}

// Positions of the voxels of neighbour column `nwidx` (same class) that are 26-adjacent to section bit b of word w:
// the voxel at the same section if set (its s-1 / s+1 neighbours are chained to it already), else those at s-1 and s+1.
__device__ __forceinline__ void neighbours_in_column(const uint32_t *__restrict__ bmp, const uint32_t *__restrict__ base,
                                                     int64_t nwidx, int w, int b, int W, uint32_t *nb, int &cnt) {
    const uint32_t B = bmp[nwidx];
    if ((B >> b) & 1u) {
        nb[cnt++] = base[nwidx] + (uint32_t)__popc(B & ((1u << b) - 1u));
        return;
    }
    if (b > 0) {  // section s-1
        if ((B >> (b - 1)) & 1u) nb[cnt++] = base[nwidx] + (uint32_t)__popc(B & ((1u << (b - 1)) - 1u));
    } else if (w > 0) {
        const uint32_t Bm = bmp[nwidx - 1];
        if (Bm >> 31) nb[cnt++] = base[nwidx - 1] + (uint32_t)__popc(Bm & 0x7fffffffu);
    }
    if (b < 31) {  // section s+1
        if ((B >> (b + 1)) & 1u) nb[cnt++] = base[nwidx] + (uint32_t)__popc(B & ((1u << (b + 1)) - 1u));
    } else if (w + 1 < W) {
        const uint32_t Bp = bmp[nwidx + 1];
        if (Bp & 1u) nb[cnt++] = base[nwidx + 1];
    }
}

__global__ void __launch_bounds__(kSparseThreads, 3) blob_sparse_kernel(const __grid_constant__ pe_geom g, const SparseArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ uint32_t smem[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nb = gridDim.x, b = blockIdx.x;
    const int warps_per_block = blockDim.x >> 5;
    const int64_t total_words = 2 * a.nwords_pad;
    // contiguous word segment of this warp (multiple of 32 words)
    const int64_t nwarps = (int64_t)nb * warps_per_block;
    const int64_t seg = (((total_words + nwarps - 1) / nwarps) + 31) / 32 * 32;
    const int64_t gw = (int64_t)b * warps_per_block + warp;
    const int64_t w_begin = min(gw * seg, total_words), w_end = min(w_begin + seg, total_words);
    stamp(0);

    // ---- S1a: popcount of every warp segment -> block partials
    uint32_t cnt = 0;
    for (int64_t i = w_begin + lane; i < w_end; i += 32) cnt += (uint32_t)__popc(a.bmp[i]);
    const uint32_t warp_cnt = (uint32_t)warp_sum((int)cnt);
    __shared__ uint32_t warp_cnts[32];
    if (lane == 0) warp_cnts[warp] = warp_cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int i = 0; i < warps_per_block; ++i) t += warp_cnts[i];
        a.sums[b] = t;
    }
    grid.sync();
    stamp(1);

    // ---- S1b + S2 (one pass): position of every word's first voxel, then key / initial parent of every foreground
    // voxel.  Keys go to a temporary array indexed by the combined position (the rank buffer, free until S5), because
    // the class-1 output offset needs n0, which is only known grid-wide after the next barrier.  Words are fetched
    // four iterations ahead so that the L2 latency of the loads overlaps.
    uint32_t before, n_all;
    grid_prefix(a.sums, nb, b, smem, before, n_all);
    uint32_t run = before;
    for (int i = 0; i < warp; ++i) run += warp_cnts[i];
    uint32_t *keyc = a.rank;
    const uint32_t pcap = (uint32_t)(2 * a.cap);
    uint32_t carry = (w_begin > 0 && w_begin < w_end) ? a.bmp[w_begin - 1] : 0u;  // the word before the current one (lane 0's predecessor)
    for (int64_t i0 = w_begin; i0 < w_end; i0 += 128) {
        uint32_t wq[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int64_t widx = i0 + 32 * q + lane;
            wq[q] = widx < w_end ? a.bmp[widx] : 0u;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int64_t widx = i0 + 32 * q + lane;
            const uint32_t word = wq[q];
            const int c0 = __popc(word);
            const uint32_t p_first = run + (uint32_t)warp_excl_scan(c0, lane);
            run += (uint32_t)__shfl_sync(kFull, (int)(p_first - run) + c0, 31);
            uint32_t prev = __shfl_up_sync(kFull, word, 1);
            if (lane == 0) prev = carry;
            carry = __shfl_sync(kFull, word, 31);
            if (widx >= w_end) continue;
            a.base[widx] = p_first;
            if (widx == a.nwords_pad) a.counts[0] = (int64_t)p_first;  // everything before plane 1 is class 0
            if (!word) continue;
            const int k = widx >= a.nwords_pad ? 1 : 0;
            const int64_t local = widx - (k ? a.nwords_pad : 0);
            const int w = (int)(local % a.W);
            const int64_t colrow = local / a.W;
            const bool prev_last = (w > 0) && (prev >> 31);
            // bits that start a run of consecutive sections inside this word (a run entering from the previous word
            // has no start bit here: its voxels point at the previous word's last voxel, one hop from that run's start)
            const uint32_t starts = word & ~((word << 1) | (prev_last ? 1u : 0u));
            const uint32_t keybase = (uint32_t)(colrow * a.U2 + (int64_t)w * 32);
            uint32_t rest = word, p = p_first;
            while (rest) {
                const int bit = __ffs(rest) - 1;
                rest &= rest - 1;
                // parent = first voxel of the run (path-compressed chaining: finds stay O(1) instead of O(run length))
                const uint32_t below = starts & ((2u << bit) - 1u);
                uint32_t par;
                if (below) {
                    const int sb = 31 - __clz(below);
                    par = p_first + (uint32_t)__popc(word & ((1u << sb) - 1u));
                } else {
                    par = p_first - 1;  // the run continues from the previous word (prev_last is set)
                }
                if (p < pcap) {  // beyond the capacity the overflow flag is raised after the barrier
                    keyc[p] = keybase + (uint32_t)bit;
                    a.parent[p] = par;
                }
                ++p;
            }
        }
    }
    grid.sync();
    stamp(2);
    const int64_t n0 = a.counts[0];
    const int64_t n = (int64_t)n_all;
    const int64_t n1 = n - n0;
    const bool overflow = n0 > a.cap || n1 > a.cap;
    if (b == 0 && threadIdx.x == 0) {
        a.counts[2] = n1;
        if (overflow) a.counts[4] = 1;
    }
    if (overflow) return;  // grid-uniform
    stamp(3);

    // ---- S3: hook the 12 predecessor neighbours outside the voxel's own column
    const int64_t gstride = (int64_t)nb * blockDim.x;
    const int64_t gtid = (int64_t)b * blockDim.x + threadIdx.x;
    for (int64_t i = gtid; i < n; i += gstride) {
        const int k = i >= n0 ? 1 : 0;
        const uint32_t p = (uint32_t)i;
        const uint32_t kk = a.rank[i];  // the key parked by S2
        a.key[(int64_t)k * a.cap + (i - (k ? n0 : 0))] = kk;
        const int s = (int)(kk % (uint32_t)a.U2);
        const uint32_t colrow = kk / (uint32_t)a.U2;
        const int r = (int)(colrow % (uint32_t)a.U1), c = (int)(colrow / (uint32_t)a.U1);
        const int w = s >> 5, bit = s & 31;
        const uint32_t *bmp = a.bmp + (k ? a.nwords_pad : 0);
        const uint32_t *base = a.base + (k ? a.nwords_pad : 0);
        // the density of the voxel (one independent gather per thread, in flight while the neighbours are looked up)
        a.value[(int64_t)k * a.cap + (i - (k ? n0 : 0))] = __ldg(a.rho + ((int64_t)s * a.NR + r) * a.NC + c);
        // the 12 predecessor neighbours outside the voxel's own column: columns (c-1, r-1..r+1) and (c, r-1)
        uint32_t nbr[8];
        int cnt = 0;
        if (c > 0) {
            const int64_t rowbase = (int64_t)(c - 1) * a.U1;
            if (r > 0) neighbours_in_column(bmp, base, (rowbase + r - 1) * a.W + w, w, bit, a.W, nbr, cnt);
            neighbours_in_column(bmp, base, (rowbase + r) * a.W + w, w, bit, a.W, nbr, cnt);
            if (r + 1 < a.U1) neighbours_in_column(bmp, base, (rowbase + r + 1) * a.W + w, w, bit, a.W, nbr, cnt);
        }
        if (r > 0) neighbours_in_column(bmp, base, ((int64_t)c * a.U1 + r - 1) * a.W + w, w, bit, a.W, nbr, cnt);
        // one union site for all lanes: the finds and hooks of a warp proceed in lock step instead of once per case
#pragma unroll 1
        for (int j = 0; j < cnt; ++j) uf_union(a.parent, p, nbr[j]);
    }
    grid.sync();
    stamp(4);

    // ---- S4: flatten; S5: rank the roots (blob number = rank of the blob's first voxel)
    const int64_t vseg = (((n + nwarps - 1) / nwarps) + 31) / 32 * 32;
    const int64_t v_begin = min(gw * vseg, n), v_end = min(v_begin + vseg, n);
    uint32_t roots = 0;
    for (int64_t i = v_begin + lane; i < v_end; i += 32) {
        const uint32_t root = uf_find(a.parent, (uint32_t)i);
        a.parent[i] = root;  // still a valid ancestor for concurrent finds
        roots += root == (uint32_t)i ? 1u : 0u;
    }
    const uint32_t warp_roots = (uint32_t)warp_sum((int)roots);
    __syncthreads();
    if (lane == 0) warp_cnts[warp] = warp_roots;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int i = 0; i < warps_per_block; ++i) t += warp_cnts[i];
        a.sums[kSparseMaxBlocks + b] = t;
    }
    grid.sync();
    stamp(5);
    uint32_t roots_before, n_roots;
    grid_prefix(a.sums + kSparseMaxBlocks, nb, b, smem, roots_before, n_roots);
    uint32_t rrun = roots_before;
    for (int i = 0; i < warp; ++i) rrun += warp_cnts[i];
    for (int64_t i0 = v_begin; i0 < v_end; i0 += 32) {
        const int64_t i = i0 + lane;
        const int isroot = (i < v_end && a.parent[i] == (uint32_t)i) ? 1 : 0;
        const uint32_t ex = rrun + (uint32_t)warp_excl_scan(isroot, lane);
        if (isroot) a.rank[i] = ex;
        if (i < v_end && i == n0) a.counts[1] = (int64_t)ex;  // roots among the class-0 voxels
        rrun += (uint32_t)__shfl_sync(kFull, (int)(ex - rrun) + isroot, 31);
    }
    if (b == 0 && threadIdx.x == 0 && n0 == n) a.counts[1] = (int64_t)n_roots;
    grid.sync();
    stamp(6);
    const int64_t nb0 = a.counts[1];
    const int64_t nb1 = (int64_t)n_roots - nb0;
    if (b == 0 && threadIdx.x == 0) a.counts[3] = nb1;
    const bool blob_overflow = nb0 > a.cap_blobs || nb1 > a.cap_blobs;
    if (blob_overflow) {
        if (b == 0 && threadIdx.x == 0) a.counts[4] = 1;
        return;  // grid-uniform
    }

    // ---- S6: zero the per-blob sums, then labels + sums (DensityBlob.fromCrsList, pdb_eda/ccp4.py:534-545)
    for (int64_t i = gtid; i < nb0 * 8; i += gstride) a.stats[i] = 0.0;
    for (int64_t i = gtid; i < nb1 * 8; i += gstride) a.stats[a.cap_blobs * 8 + i] = 0.0;
    grid.sync();
    stamp(7);
    for (int64_t i0 = gtid - lane; i0 < n; i0 += gstride) {  // warp-uniform trip count
        const int64_t i = i0 + lane;
        const bool live = i < n;
        int32_t blob = -1;
        int k = 0;
        double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (live) {
            k = i >= n0 ? 1 : 0;
            const int64_t oi = (int64_t)k * a.cap + (i - (k ? n0 : 0));
            const uint32_t gr = a.rank[a.parent[i]];
            blob = (int32_t)(gr - (k ? (uint32_t)nb0 : 0u));
            a.label[oi] = blob;
            blob += k ? (int32_t)a.cap_blobs : 0;  // row of the combined stats table
            const uint32_t kk = a.key[oi];
            const int s = (int)(kk % (uint32_t)a.U2);
            const uint32_t colrow = kk / (uint32_t)a.U2;
            const int r = (int)(colrow % (uint32_t)a.U1), c = (int)(colrow / (uint32_t)a.U1);
            double x, y, z;
            crs2xyz(g, c, r, s, x, y, z);
            const double d = (double)a.value[oi];
            v[0] = 1.0;
            v[1] = d;
            v[2] = __dmul_rn(d, x);
            v[3] = __dmul_rn(d, y);
            v[4] = __dmul_rn(d, z);
            v[5] = x;
            v[6] = y;
            v[7] = z;
        }
        // segmented reduction over runs of equal blob id (runs along the section axis are the common case)
        const int32_t prev = __shfl_up_sync(kFull, blob, 1);
        const bool head = (lane == 0) || (prev != blob);
        const unsigned heads = __ballot_sync(kFull, head);
        const int segno = __popc(heads & (0xffffffffu >> (31 - lane)));
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int seg_o = __shfl_down_sync(kFull, segno, o);
            const bool take = (lane + o < 32) && (seg_o == segno);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const double other = __shfl_down_sync(kFull, v[q], o);
                if (take) v[q] += other;
            }
        }
        if (live && head) {
            double *st = a.stats + (int64_t)blob * 8;
#pragma unroll
            for (int q = 0; q < 8; ++q) atomicAdd(st + q, v[q]);
        }
    }
    stamp(8);
}

}  // namespace pe

using namespace pe;

extern "C" {

int pe_blob_stage_times(unsigned long long *out12) {
    PE_CHECK_ARG(out12 != nullptr, "pe_blob_stage_times: null pointer");
    PE_CUDA(cudaMemcpyFromSymbol(out12, g_stage_ns, sizeof(unsigned long long) * 12));
    return PE_OK;
}

int64_t pe_blob_workspace_bytes(const pe_geom *g, int64_t cap_voxels) {
    if (!g || cap_voxels < 0) return -1;
    return make_plan(g, cap_voxels).total;
}

int pe_blob_label(const pe_geom *g, const float *d_rho, float cut_pos, float cut_neg, int64_t cap_voxels,
                  int64_t cap_blobs, int64_t *d_counts, uint32_t *d_key, float *d_value, int32_t *d_label,
                  double *d_stats, void *d_ws, void *stream) {
    if (int rc = check_geom(g)) return rc;
    PE_CHECK_ARG(d_rho && d_counts && d_key && d_value && d_label && d_stats && d_ws, "pe_blob_label: null pointer");
    PE_CHECK_ARG(cap_voxels > 0 && cap_blobs > 0, "pe_blob_label: capacities must be positive");
    PE_CHECK_ARG(cap_voxels < (1ll << 30), "pe_blob_label: cap_voxels must be below 2^30");
    PE_CHECK_ARG(!(cut_pos < 0.f) && !(cut_neg > 0.f), "pe_blob_label: cut_pos must be >= 0 and cut_neg <= 0");
    PE_CHECK_ARG(cut_pos == cut_pos && cut_neg == cut_neg, "pe_blob_label: NaN cutoff");
    const BlobPlan p = make_plan(g, cap_voxels);
    PE_CHECK_ARG((int64_t)p.U0 * p.U1 * p.U2 < (1ll << 32), "pe_blob_label: unique volume too large for 32-bit keys");
    cudaStream_t st = (cudaStream_t)stream;
    char *ws = (char *)d_ws;
    const int NC = g->ncrs[0], NR = g->ncrs[1];
    uint32_t *bmp = (uint32_t *)(ws + p.off_bmp);
    const bool use_pos = cut_pos > 0.f, use_neg = cut_neg < 0.f;

    PE_CUDA(cudaMemsetAsync(d_counts, 0, 5 * sizeof(int64_t), st));
    // K1: the only pass over the map
    {
        const bool vec4 = (NC % 4 == 0) && (((uintptr_t)d_rho & 15u) == 0);
        const int vec = vec4 ? 4 : 1;
        const int tx = kBmpTx, ty = kBmpTy;  // block shape makes no measurable difference (profiles/r01_threshold_tuning.md)
        dim3 block(tx, ty, 1);
        dim3 grid(p.W / kWordsPerThread, (p.U0 + tx * vec - 1) / (tx * vec), (p.U1 + ty - 1) / ty);
        PE_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "pe_blob_label: map too large for the launch grid");
        if (vec4)
            PE_LAUNCH("threshold_bitmap_kernel", st, threshold_bitmap_kernel<4><<<grid, block, 0, st>>>(
                d_rho, NC, NR, p.U0, p.U1, p.U2, p.W, cut_pos, cut_neg, use_pos, use_neg, bmp, bmp + p.nwords_pad));
        else
            PE_LAUNCH("threshold_bitmap_kernel", st, threshold_bitmap_kernel<1><<<grid, block, 0, st>>>(
                d_rho, NC, NR, p.U0, p.U1, p.U2, p.W, cut_pos, cut_neg, use_pos, use_neg, bmp, bmp + p.nwords_pad));
        PE_LAUNCH_CHECK();
    }
    // sparse stage: one cooperative kernel for both signs
    SparseArgs a;
    a.rho = d_rho;
    a.NC = NC;
    a.NR = NR;
    a.U1 = p.U1;
    a.U2 = p.U2;
    a.W = p.W;
    a.nwords = p.nwords;
    a.nwords_pad = p.nwords_pad;
    a.cap = p.cap;
    a.cap_blobs = cap_blobs;
    a.bmp = bmp;
    a.base = (uint32_t *)(ws + p.off_base);
    a.parent = (uint32_t *)(ws + p.off_parent);
    a.rank = (uint32_t *)(ws + p.off_rank);
    a.sums = (uint32_t *)(ws + p.off_sums);
    a.counts = d_counts;
    a.key = d_key;
    a.value = d_value;
    a.label = d_label;
    a.stats = d_stats;
    a.use_pos = use_pos;
    a.use_neg = use_neg;
    static int blocks_per_sm = 0;
    if (blocks_per_sm == 0) {
        PE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, blob_sparse_kernel, kSparseThreads, 0));
        PE_CHECK_ARG(blocks_per_sm > 0, "pe_blob_label: the sparse kernel does not fit an SM");
    }
    int nblocks = sm_count() * (blocks_per_sm < 3 ? blocks_per_sm : 3);
    if (nblocks > kSparseMaxBlocks) nblocks = kSparseMaxBlocks;
    pe_geom geom = *g;
    void *args[] = {(void *)&geom, (void *)&a};
    PE_LAUNCH("blob_sparse_kernel", st,
              PE_CUDA(cudaLaunchCooperativeKernel((const void *)blob_sparse_kernel, dim3(nblocks), dim3(kSparseThreads), args, 0, st)));
    PE_LAUNCH_CHECK();
    return PE_OK;
}

}  // extern "C"

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
#include "pe_common.cuh"

namespace pe {

constexpr int kBmpTx = 32;         // threads along columns (x VEC columns each)
constexpr int kBmpTy = 8;          // threads along rows
constexpr int kChunkWords = 4;     // 32-section words per thread (grid.z splits the section axis)
constexpr int kSparseThreads = 256;

struct BlobPlan {
    int U0, U1, U2, W;   // unique columns, rows, sections; words per (column,row)
    int64_t nwords;      // U0*U1*W per class
    int64_t cap;         // foreground capacity per class
    // workspace carve-up (per class k: base + k*stride)
    int64_t off_bmp, off_base, off_parent, off_flag, off_rank, off_scan, total;
};

static BlobPlan make_plan(const pe_geom *g, int64_t cap) {
    BlobPlan p;
    p.U0 = g->unique_ncrs[0];
    p.U1 = g->unique_ncrs[1];
    p.U2 = g->unique_ncrs[2];
    p.W = (p.U2 + 31) / 32;
    p.nwords = (int64_t)p.U0 * p.U1 * p.W;
    p.cap = cap;
    int64_t o = 0;
    p.off_bmp = o;
    o += 2 * align_up(p.nwords * 4, 256);
    p.off_base = o;
    o += 2 * align_up(p.nwords * 4, 256);
    p.off_parent = o;
    o += 2 * align_up(cap * 4, 256);
    p.off_flag = o;
    o += 2 * align_up(cap * 4, 256);
    p.off_rank = o;
    o += 2 * align_up(cap * 4, 256);
    p.off_scan = o;
    const int64_t larger = p.nwords > cap ? p.nwords : cap;
    o += scan_ws_bytes(larger);
    p.total = o;
    return p;
}

// ------------------------------------------------------------------------------------------------ K1: stream
template <int VEC>
__global__ void __launch_bounds__(kBmpTx *kBmpTy)
    threshold_bitmap_kernel(const float *__restrict__ rho, int NC, int NR, int U0, int U1, int U2, int W, float cpos,
                            float cneg, bool use_pos, bool use_neg, uint32_t *__restrict__ bmp_pos,
                            uint32_t *__restrict__ bmp_neg) {
    const int c = (blockIdx.x * kBmpTx + threadIdx.x) * VEC;
    const int r = blockIdx.y * kBmpTy + threadIdx.y;
    if (c >= U0 || r >= U1) return;
    const int w_begin = blockIdx.z * kChunkWords;
    const int w_end = min(W, w_begin + kChunkWords);
    const int64_t plane = (int64_t)NR * NC;
    const float *col = rho + (int64_t)r * NC + c;
    for (int w = w_begin; w < w_end; ++w) {
        uint32_t pos[VEC], neg[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) pos[i] = neg[i] = 0u;
        const int s0 = w * 32;
        const int nbits = min(32, U2 - s0);
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
            if (c + i < U0) {
                const int64_t widx = ((int64_t)(c + i) * U1 + r) * W + w;
                bmp_pos[widx] = use_pos ? pos[i] : 0u;
                bmp_neg[widx] = use_neg ? neg[i] : 0u;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ K3: sparse init
// One thread per bitmap word: writes key / density / initial parent for each of its set bits.
__global__ void __launch_bounds__(kSparseThreads)
    blob_init_kernel(const float *__restrict__ rho, int NC, int NR, int U1, int U2, int W, int64_t nwords, int64_t cap,
                     const uint32_t *__restrict__ bmp, const uint32_t *__restrict__ base, const int64_t *__restrict__ d_nfg,
                     int64_t *__restrict__ d_overflow, uint32_t *__restrict__ key, float *__restrict__ value,
                     uint32_t *__restrict__ parent) {
    if (*d_nfg > cap) {
        if (blockIdx.x == 0 && threadIdx.x == 0) *d_overflow = 1;
        return;
    }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t widx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; widx < nwords; widx += stride) {
        uint32_t word = bmp[widx];
        if (!word) continue;
        const int w = (int)(widx % W);
        const int64_t colrow = widx / W;
        const int r = (int)(colrow % U1), c = (int)(colrow / U1);
        const bool prev_last = (w > 0) && (bmp[widx - 1] >> 31);
        const uint32_t linked = (word << 1) | (prev_last ? 1u : 0u);  // bit b set: voxel b-1 of this column is foreground
        uint32_t p = base[widx];
        const uint32_t keybase = (uint32_t)(colrow * U2 + (int64_t)w * 32);
        uint32_t rest = word;
        while (rest) {
            const int b = __ffs(rest) - 1;
            rest &= rest - 1;
            const int s = w * 32 + b;
            key[p] = keybase + (uint32_t)b;
            value[p] = __ldg(rho + ((int64_t)s * NR + r) * NC + c);
            parent[p] = ((linked >> b) & 1u) ? p - 1 : p;
            ++p;
        }
    }
}

// ------------------------------------------------------------------------------------------------ K4: merge
__device__ __forceinline__ void merge_with_column(uint32_t *parent, const uint32_t *__restrict__ bmp,
                                                  const uint32_t *__restrict__ base, int64_t nwidx, int w, int b, int W,
                                                  uint32_t p) {
    const uint32_t B = bmp[nwidx];
    if ((B >> b) & 1u) {  // same section: its s-1 / s+1 neighbours are chained to it already
        uf_union(parent, p, base[nwidx] + (uint32_t)__popc(B & ((1u << b) - 1u)));
        return;
    }
    // section s-1
    if (b > 0) {
        if ((B >> (b - 1)) & 1u) uf_union(parent, p, base[nwidx] + (uint32_t)__popc(B & ((1u << (b - 1)) - 1u)));
    } else if (w > 0) {
        const uint32_t Bm = bmp[nwidx - 1];
        if (Bm >> 31) uf_union(parent, p, base[nwidx - 1] + (uint32_t)__popc(Bm & 0x7fffffffu));
    }
    // section s+1
    if (b < 31) {
        if ((B >> (b + 1)) & 1u) uf_union(parent, p, base[nwidx] + (uint32_t)__popc(B & ((1u << (b + 1)) - 1u)));
    } else if (w + 1 < W) {
        const uint32_t Bp = bmp[nwidx + 1];
        if (Bp & 1u) uf_union(parent, p, base[nwidx + 1]);
    }
}

__global__ void __launch_bounds__(kSparseThreads)
    blob_merge_kernel(int U1, int U2, int W, int64_t cap, const uint32_t *__restrict__ bmp, const uint32_t *__restrict__ base,
                      const int64_t *__restrict__ d_nfg, const uint32_t *__restrict__ key, uint32_t *parent) {
    const int64_t n = *d_nfg;
    if (n > cap) return;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t p = (uint32_t)i;
        const uint32_t k = key[p];
        const int s = (int)(k % (uint32_t)U2);
        const uint32_t colrow = k / (uint32_t)U2;
        const int r = (int)(colrow % (uint32_t)U1), c = (int)(colrow / (uint32_t)U1);
        const int w = s >> 5, b = s & 31;
        // the 12 predecessor neighbours outside the voxel's own column: columns (c-1, r-1..r+1) and (c, r-1)
        if (c > 0) {
            const int64_t rowbase = (int64_t)(c - 1) * U1;
            if (r > 0) merge_with_column(parent, bmp, base, (rowbase + r - 1) * W + w, w, b, W, p);
            merge_with_column(parent, bmp, base, (rowbase + r) * W + w, w, b, W, p);
            if (r + 1 < U1) merge_with_column(parent, bmp, base, (rowbase + r + 1) * W + w, w, b, W, p);
        }
        if (r > 0) merge_with_column(parent, bmp, base, ((int64_t)c * U1 + r - 1) * W + w, w, b, W, p);
    }
}

// ------------------------------------------------------------------------------------------------ K5: flatten
__global__ void __launch_bounds__(kSparseThreads)
    blob_flatten_kernel(int64_t cap, const int64_t *__restrict__ d_nfg, uint32_t *parent, uint32_t *__restrict__ flag) {
    const int64_t n = *d_nfg;
    if (n > cap) return;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t root = uf_find(parent, (uint32_t)i);
        parent[i] = root;
        flag[i] = root == (uint32_t)i ? 1u : 0u;
    }
}

__global__ void __launch_bounds__(kSparseThreads)
    blob_zero_stats_kernel(int64_t cap_blobs, const int64_t *__restrict__ d_nblobs, int64_t *__restrict__ d_overflow,
                           double *__restrict__ stats) {
    int64_t n = *d_nblobs;
    if (n > cap_blobs) {
        if (blockIdx.x == 0 && threadIdx.x == 0) *d_overflow = 1;
        n = cap_blobs;
    }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * 8; i += stride) stats[i] = 0.0;
}

// ------------------------------------------------------------------------------------------------ K7: labels + stats
// label[p] = rank of p's root; per-blob sums of DensityBlob.fromCrsList (pdb_eda/ccp4.py:534-545).  Lanes that hold
// consecutive voxels of one blob (the common case: runs along the section axis) are combined by a segmented warp
// reduction before the float64 atomics.
__global__ void __launch_bounds__(kSparseThreads)
    blob_stats_kernel(const __grid_constant__ pe_geom g, int64_t cap, int64_t cap_blobs, const int64_t *__restrict__ d_nfg,
                      const uint32_t *__restrict__ key, const float *__restrict__ value, const uint32_t *__restrict__ parent,
                      const uint32_t *__restrict__ rank, int32_t *__restrict__ label, double *__restrict__ stats) {
    const int64_t n = *d_nfg;
    if (n > cap) return;
    const int U1 = g.unique_ncrs[1], U2 = g.unique_ncrs[2];
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t start = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t i0 = start - lane; i0 < n; i0 += stride) {  // warp-uniform trip count
        const int64_t i = i0 + lane;
        const bool live = i < n;
        int32_t blob = -1;
        double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (live) {
            blob = (int32_t)rank[parent[i]];
            label[i] = blob;
            const uint32_t k = key[i];
            const int s = (int)(k % (uint32_t)U2);
            const uint32_t colrow = k / (uint32_t)U2;
            const int r = (int)(colrow % (uint32_t)U1), c = (int)(colrow / (uint32_t)U1);
            double x, y, z;
            crs2xyz(g, c, r, s, x, y, z);
            const double d = (double)value[i];
            v[0] = 1.0;
            v[1] = d;
            v[2] = __dmul_rn(d, x);
            v[3] = __dmul_rn(d, y);
            v[4] = __dmul_rn(d, z);
            v[5] = x;
            v[6] = y;
            v[7] = z;
        }
        // segmented reduction over runs of equal blob id
        const int32_t prev = __shfl_up_sync(kFull, blob, 1);
        const bool head = (lane == 0) || (prev != blob);
        const unsigned heads = __ballot_sync(kFull, head);
        const int seg = __popc(heads & (0xffffffffu >> (31 - lane)));  // run number of this lane (1-based)
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int seg_o = __shfl_down_sync(kFull, seg, o);
            const bool take = (lane + o < 32) && (seg_o == seg);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const double other = __shfl_down_sync(kFull, v[q], o);
                if (take) v[q] += other;
            }
        }
        if (live && head && blob < cap_blobs) {
            double *st = stats + (int64_t)blob * 8;
#pragma unroll
            for (int q = 0; q < 8; ++q) atomicAdd(st + q, v[q]);
        }
    }
}

static int sparse_grid() { return sm_count() * 8; }

}  // namespace pe

using namespace pe;

extern "C" {

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
    PE_CHECK_ARG(cap_voxels < (1ll << 31), "pe_blob_label: cap_voxels must be below 2^31");
    PE_CHECK_ARG(!(cut_pos < 0.f) && !(cut_neg > 0.f), "pe_blob_label: cut_pos must be >= 0 and cut_neg <= 0");
    PE_CHECK_ARG(cut_pos == cut_pos && cut_neg == cut_neg, "pe_blob_label: NaN cutoff");
    const BlobPlan p = make_plan(g, cap_voxels);
    PE_CHECK_ARG((int64_t)p.U0 * p.U1 * p.U2 < (1ll << 32), "pe_blob_label: unique volume too large for 32-bit keys");
    cudaStream_t st = (cudaStream_t)stream;
    char *ws = (char *)d_ws;
    const int NC = g->ncrs[0], NR = g->ncrs[1];
    const int64_t bmp_stride = align_up(p.nwords * 4, 256), cap_stride = align_up(p.cap * 4, 256);
    uint32_t *bmp[2], *base[2], *parent[2], *flag[2], *rank[2];
    for (int k = 0; k < 2; ++k) {
        bmp[k] = (uint32_t *)(ws + p.off_bmp + k * bmp_stride);
        base[k] = (uint32_t *)(ws + p.off_base + k * bmp_stride);
        parent[k] = (uint32_t *)(ws + p.off_parent + k * cap_stride);
        flag[k] = (uint32_t *)(ws + p.off_flag + k * cap_stride);
        rank[k] = (uint32_t *)(ws + p.off_rank + k * cap_stride);
    }
    void *scan_ws = ws + p.off_scan;
    const bool use_pos = cut_pos > 0.f, use_neg = cut_neg < 0.f;

    PE_CUDA(cudaMemsetAsync(d_counts, 0, 5 * sizeof(int64_t), st));
    // K1: the only pass over the map
    {
        const bool vec4 = (NC % 4 == 0) && (((uintptr_t)d_rho & 15u) == 0);
        const int vec = vec4 ? 4 : 1;
        dim3 block(kBmpTx, kBmpTy, 1);
        dim3 grid((p.U0 + kBmpTx * vec - 1) / (kBmpTx * vec), (p.U1 + kBmpTy - 1) / kBmpTy, (p.W + kChunkWords - 1) / kChunkWords);
        PE_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "pe_blob_label: map too large for the launch grid");
        if (vec4)
            PE_LAUNCH("threshold_bitmap_kernel", st, threshold_bitmap_kernel<4><<<grid, block, 0, st>>>(d_rho, NC, NR, p.U0, p.U1, p.U2, p.W, cut_pos, cut_neg, use_pos,
                                                               use_neg, bmp[0], bmp[1]));
        else
            PE_LAUNCH("threshold_bitmap_kernel", st, threshold_bitmap_kernel<1><<<grid, block, 0, st>>>(d_rho, NC, NR, p.U0, p.U1, p.U2, p.W, cut_pos, cut_neg, use_pos,
                                                               use_neg, bmp[0], bmp[1]));
        PE_LAUNCH_CHECK();
    }
    const int sg = sparse_grid();
    for (int k = 0; k < 2; ++k) {
        if (!(k == 0 ? use_pos : use_neg)) continue;
        int64_t *d_nfg = d_counts + 2 * k, *d_nblobs = d_counts + 2 * k + 1, *d_overflow = d_counts + 4;
        uint32_t *key = d_key + (int64_t)k * cap_voxels;
        float *value = d_value + (int64_t)k * cap_voxels;
        int32_t *label = d_label + (int64_t)k * cap_voxels;
        double *stats = d_stats + (int64_t)k * cap_blobs * 8;
        // K2: position of every foreground voxel in the reference's list order
        if (int rc = exclusive_scan_u32(bmp[k], base[k], p.nwords, nullptr, d_nfg, scan_ws, st, true)) return rc;
        PE_LAUNCH("blob_init_kernel", st, blob_init_kernel<<<sg, kSparseThreads, 0, st>>>(d_rho, NC, NR, p.U1, p.U2, p.W, p.nwords, p.cap, bmp[k], base[k], d_nfg,
                                                        d_overflow, key, value, parent[k]));
        PE_LAUNCH("blob_merge_kernel", st, blob_merge_kernel<<<sg, kSparseThreads, 0, st>>>(p.U1, p.U2, p.W, p.cap, bmp[k], base[k], d_nfg, key, parent[k]));
        PE_LAUNCH("blob_flatten_kernel", st, blob_flatten_kernel<<<sg, kSparseThreads, 0, st>>>(p.cap, d_nfg, parent[k], flag[k]));
        PE_LAUNCH_CHECK();
        // K6: blob number = rank of its root among roots
        if (int rc = exclusive_scan_u32(flag[k], rank[k], p.cap, d_nfg, d_nblobs, scan_ws, st, false)) return rc;
        PE_LAUNCH("blob_zero_stats_kernel", st, blob_zero_stats_kernel<<<sg, kSparseThreads, 0, st>>>(cap_blobs, d_nblobs, d_overflow, stats));
        PE_LAUNCH("blob_stats_kernel", st, blob_stats_kernel<<<sg, kSparseThreads, 0, st>>>(*g, p.cap, cap_blobs, d_nfg, key, value, parent[k], rank[k], label, stats));
        PE_LAUNCH_CHECK();
    }
    return PE_OK;
}

}  // extern "C"

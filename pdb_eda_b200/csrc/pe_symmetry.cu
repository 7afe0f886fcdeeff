// pe_symmetry.cu -- symmetry-operator x lattice-translation expansion with box culling, and the nearest
// (symmetry) atom of every blob centroid.
//
//   pe_symmetry_expand   createSymmetryAtoms                 pdb_eda/cutils.pyx:73-103
//                        (driven by _calculateSymmetryAtoms, pdb_eda/densityAnalysis.py:885-912)
//   pe_nearest_atom      cdist + argmin/min per blob         pdb_eda/densityAnalysis.py:932-937
//
// Expansion: candidate q = (image * n_ops + op) * n_atoms + atom is exactly the reference's loop order
// (itertools.product([-1,0,1]^3, ops) outer, atoms inner), so an order-preserving compaction of the keep flags
// (flag kernel -> exclusive scan -> scatter kernel) reproduces the reference's list without sorting.
// The arithmetic per image is  (np.dot(R, x) + t) + np.dot(orthoMat, (i,j,k))  in float64: the 3x3 product is
// evaluated in the host BLAS's order (pe_geom.mv_perm / mv_fma), the lattice shifts are formed on the host by
// numpy itself.
//
// Nearest atom: brute force in float64, 8 blobs per CTA so each atom coordinate fetched from L2 is used 8 times;
// squared distances are compared first and the (correctly rounded) square roots only when a candidate improves,
// which keeps np.argmin's first-minimum-of-the-rooted-values semantics.
#include "pe_common.cuh"

namespace pe {

constexpr int kSymThreads = 256;

__device__ __forceinline__ bool sym_image(const pe_geom &g, const double *__restrict__ xyz, const double *__restrict__ rot,
                                          const double *__restrict__ shift, int n_ops, int n_atoms, int64_t q,
                                          double lo0, double lo1, double lo2, double hi0, double hi1, double hi2,
                                          double &px, double &py, double &pz) {
    const int a = (int)(q % n_atoms);
    const int io = (int)(q / n_atoms);
    const int op = io % n_ops, img = io / n_ops;
    const double x = xyz[3 * a], y = xyz[3 * a + 1], z = xyz[3 * a + 2];
    if (img == 13 && op == 0) {  // symmetry == (0,0,0,0): every atom, coordinate untouched (pdb_eda/cutils.pyx:94-95)
        px = x;
        py = y;
        pz = z;
        return true;
    }
    const double *m = rot + 12 * op;
    const double *sh = shift + 3 * img;
    px = __dadd_rn(__dadd_rn(mv_row(m + 0, x, y, z, g), m[3]), sh[0]);
    py = __dadd_rn(__dadd_rn(mv_row(m + 4, x, y, z, g), m[7]), sh[1]);
    pz = __dadd_rn(__dadd_rn(mv_row(m + 8, x, y, z, g), m[11]), sh[2]);
    return lo0 <= px && px <= hi0 && lo1 <= py && py <= hi1 && lo2 <= pz && pz <= hi2;
}

__global__ void __launch_bounds__(kSymThreads)
    sym_flag_kernel(const __grid_constant__ pe_geom g, const double *__restrict__ xyz, const double *__restrict__ rot,
                    const double *__restrict__ shift, int n_ops, int n_atoms, int64_t total, double lo0, double lo1,
                    double lo2, double hi0, double hi1, double hi2, uint32_t *__restrict__ flag) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += stride) {
        double px, py, pz;
        flag[q] = sym_image(g, xyz, rot, shift, n_ops, n_atoms, q, lo0, lo1, lo2, hi0, hi1, hi2, px, py, pz) ? 1u : 0u;
    }
}

__global__ void __launch_bounds__(kSymThreads)
    sym_scatter_kernel(const __grid_constant__ pe_geom g, const double *__restrict__ xyz, const double *__restrict__ rot,
                       const double *__restrict__ shift, int n_ops, int n_atoms, int64_t total, double lo0, double lo1,
                       double lo2, double hi0, double hi1, double hi2, const uint32_t *__restrict__ flag,
                       const uint32_t *__restrict__ pos, int64_t cap, int32_t *__restrict__ out_atom,
                       int32_t *__restrict__ out_image, double *__restrict__ out_xyz) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += stride) {
        if (!flag[q]) continue;
        const int64_t m = pos[q];
        if (m >= cap) continue;
        double px, py, pz;
        sym_image(g, xyz, rot, shift, n_ops, n_atoms, q, lo0, lo1, lo2, hi0, hi1, hi2, px, py, pz);
        out_atom[m] = (int32_t)(q % n_atoms);
        out_image[m] = (int32_t)(q / n_atoms);
        out_xyz[3 * m] = px;
        out_xyz[3 * m + 1] = py;
        out_xyz[3 * m + 2] = pz;
    }
}

// ------------------------------------------------------------------------------------------------ nearest atom
constexpr int kNearThreads = 256;
constexpr int kNearBlobs = 8;

struct Best {
    double d2;    // squared distance of the current best
    double dist;  // its correctly rounded square root
    int idx;
};

__device__ __forceinline__ void best_offer(Best &b, double d2, int idx) {
    if (d2 < b.d2) {  // sqrt is monotone: only a smaller square can give a smaller root
        const double dist = __dsqrt_rn(d2);
        if (dist < b.dist) {  // strict: equal roots keep the earlier index (np.argmin)
            b.dist = dist;
            b.idx = idx;
        }
        b.d2 = d2;  // safe either way: roots of anything between are equal to b.dist
    }
}

__device__ __forceinline__ void best_merge(Best &b, double dist, int idx) {
    if (idx >= 0 && (dist < b.dist || (dist == b.dist && idx < b.idx) || b.idx < 0)) {
        b.dist = dist;
        b.idx = idx;
    }
}

__global__ void __launch_bounds__(kNearThreads)
    nearest_kernel(int64_t n_blobs, const double *__restrict__ centroid, int64_t n_atoms, const double *__restrict__ coords,
                   int32_t *__restrict__ out_idx, double *__restrict__ out_dist) {
    __shared__ double s_dist[kNearThreads / 32][kNearBlobs];
    __shared__ int s_idx[kNearThreads / 32][kNearBlobs];
    const int64_t b0 = (int64_t)blockIdx.x * kNearBlobs;
    double ux[kNearBlobs], uy[kNearBlobs], uz[kNearBlobs];
    Best best[kNearBlobs];
#pragma unroll
    for (int j = 0; j < kNearBlobs; ++j) {
        const int64_t b = b0 + j < n_blobs ? b0 + j : n_blobs - 1;
        ux[j] = centroid[3 * b];
        uy[j] = centroid[3 * b + 1];
        uz[j] = centroid[3 * b + 2];
        best[j].d2 = INFINITY;
        best[j].dist = INFINITY;
        best[j].idx = -1;
    }
    for (int64_t a = threadIdx.x; a < n_atoms; a += kNearThreads) {
        const double vx = coords[3 * a], vy = coords[3 * a + 1], vz = coords[3 * a + 2];
#pragma unroll
        for (int j = 0; j < kNearBlobs; ++j) {
            // scipy's euclidean kernel: s = 0; s += (u_k - v_k)^2 for k = 0, 1, 2; sqrt(s)
            const double d0 = __dsub_rn(ux[j], vx), d1 = __dsub_rn(uy[j], vy), d2 = __dsub_rn(uz[j], vz);
            const double s = __dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2));
            best_offer(best[j], s, (int)a);
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < kNearBlobs; ++j) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double od = __shfl_xor_sync(kFull, best[j].dist, o);
            const int oi = __shfl_xor_sync(kFull, best[j].idx, o);
            best_merge(best[j], od, oi);
        }
        if (lane == 0) {
            s_dist[warp][j] = best[j].dist;
            s_idx[warp][j] = best[j].idx;
        }
    }
    __syncthreads();
    if (threadIdx.x < kNearBlobs) {
        const int j = threadIdx.x;
        Best b;
        b.d2 = INFINITY;
        b.dist = INFINITY;
        b.idx = -1;
        for (int w = 0; w < kNearThreads / 32; ++w) best_merge(b, s_dist[w][j], s_idx[w][j]);
        if (b0 + j < n_blobs) {
            out_idx[b0 + j] = b.idx;
            out_dist[b0 + j] = b.dist;
        }
    }
}

}  // namespace pe

using namespace pe;

extern "C" {

int64_t pe_symmetry_workspace_bytes(int32_t n_atoms, int32_t n_ops) {
    if (n_atoms < 0 || n_ops < 0) return -1;
    const int64_t total = 27ll * n_ops * n_atoms;
    return 2 * align_up(total * 4, 256) + scan_ws_bytes(total);
}

int pe_symmetry_expand(const pe_geom *g, int32_t n_atoms, const double *d_xyz, int32_t n_ops, const double *d_rot,
                       const double *d_shift, const double *lo, const double *hi, int64_t cap, int64_t *d_count,
                       int32_t *d_atom, int32_t *d_image, double *d_out_xyz, void *d_ws, void *stream) {
    if (int rc = check_geom(g)) return rc;
    PE_CHECK_ARG(n_atoms >= 0 && n_ops >= 0 && cap >= 0, "pe_symmetry_expand: negative size");
    PE_CHECK_ARG(d_count != nullptr, "pe_symmetry_expand: d_count is null");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t total = 27ll * n_ops * n_atoms;
    if (total == 0) {
        PE_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int64_t), st));
        return PE_OK;
    }
    PE_CHECK_ARG(d_xyz && d_rot && d_shift && lo && hi && d_ws, "pe_symmetry_expand: null pointer");
    PE_CHECK_ARG(cap == 0 || (d_atom && d_image && d_out_xyz), "pe_symmetry_expand: null output pointer");
    PE_CHECK_ARG(total < (1ll << 32), "pe_symmetry_expand: more than 2^32 candidate images");
    char *ws = (char *)d_ws;
    uint32_t *flag = (uint32_t *)ws;
    uint32_t *pos = (uint32_t *)(ws + align_up(total * 4, 256));
    void *scan_ws = ws + 2 * align_up(total * 4, 256);
    int64_t blocks64 = (total + kSymThreads - 1) / kSymThreads;
    const int64_t max_blocks = (int64_t)sm_count() * 16;
    const int blocks = (int)(blocks64 < max_blocks ? blocks64 : max_blocks);
    PE_LAUNCH("sym_flag_kernel", st, sym_flag_kernel<<<blocks, kSymThreads, 0, st>>>(*g, d_xyz, d_rot, d_shift, n_ops, n_atoms, total, lo[0], lo[1], lo[2], hi[0],
                                                    hi[1], hi[2], flag));
    PE_LAUNCH_CHECK();
    if (int rc = exclusive_scan_u32(flag, pos, total, nullptr, d_count, scan_ws, st, false)) return rc;
    if (cap > 0) {
        PE_LAUNCH("sym_scatter_kernel", st, sym_scatter_kernel<<<blocks, kSymThreads, 0, st>>>(*g, d_xyz, d_rot, d_shift, n_ops, n_atoms, total, lo[0], lo[1], lo[2],
                                                           hi[0], hi[1], hi[2], flag, pos, cap, d_atom, d_image, d_out_xyz));
        PE_LAUNCH_CHECK();
    }
    return PE_OK;
}

int pe_nearest_atom(int64_t n_blobs, const double *d_centroid, int64_t n_atoms, const double *d_coords, int32_t *d_idx,
                    double *d_dist, void *stream) {
    PE_CHECK_ARG(n_blobs >= 0 && n_atoms >= 0, "pe_nearest_atom: negative size");
    if (n_blobs == 0) return PE_OK;
    PE_CHECK_ARG(n_atoms > 0, "pe_nearest_atom: no atoms to search (np.argmin of an empty row raises)");
    PE_CHECK_ARG(n_atoms < (1ll << 31), "pe_nearest_atom: too many atoms");
    PE_CHECK_ARG(d_centroid && d_coords && d_idx && d_dist, "pe_nearest_atom: null pointer");
    const int64_t blocks = (n_blobs + kNearBlobs - 1) / kNearBlobs;
    PE_CHECK_ARG(blocks < (1ll << 31), "pe_nearest_atom: too many blobs");
    PE_LAUNCH("nearest_kernel", (cudaStream_t)stream, nearest_kernel<<<(unsigned)blocks, kNearThreads, 0, (cudaStream_t)stream>>>(n_blobs, d_centroid, n_atoms, d_coords, d_idx,
                                                                                d_dist));
    PE_LAUNCH_CHECK();
    return PE_OK;
}

}  // extern "C"

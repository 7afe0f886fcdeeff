// pe_map.cu -- whole-map reductions and batched point conversions / lookups.
//
//   pe_map_mean_std   DensityMatrix.meanDensity / stdDensity    pdb_eda/ccp4.py:343-363 (np.mean / np.std, ddof 0)
//   pe_map_sum_abs    sumOfAbs / getTotalAbsDensity             pdb_eda/cutils.pyx:28-39, pdb_eda/ccp4.py:365-376
//   pe_point_density  getPointDensityFromCrs / testValidCrs     pdb_eda/cutils.pyx:125-167
//   pe_xyz2crs / pe_crs2xyz                                      pdb_eda/ccp4.py:288-316
//
// The reductions are HBM-bound streams (4 B per voxel, read once per pass): float4 loads, float64 accumulation,
// fixed grid and a fixed-order second stage so results are run-to-run deterministic.
#include "pe_common.cuh"

namespace pe {

constexpr int kRedThreads = 256;
constexpr int kRedBlocksPerSm = 8;

enum RedOp { RED_SUM = 0, RED_SQDEV = 1, RED_ABS_ABOVE = 2 };

template <int OP>
__device__ __forceinline__ double red_term(float v, double mean, float cut) {
    if (OP == RED_SUM) return (double)v;
    if (OP == RED_SQDEV) {
        const double d = (double)v - mean;
        return d * d;
    }
    const float a = fabsf(v);
    return a > cut ? (double)a : 0.0;
}

__device__ __forceinline__ double block_sum(double v) {
    __shared__ double warp_part[kRedThreads / 32];
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x < 32) {
        t = threadIdx.x < kRedThreads / 32 ? warp_part[threadIdx.x] : 0.0;
        t = warp_sum(t);
    }
    return t;  // valid in thread 0
}

// d_mean_in: optional device scalar (the mean of pass 1) for RED_SQDEV.
template <int OP>
__global__ void __launch_bounds__(kRedThreads) reduce_partial(const float *__restrict__ rho, int64_t n,
                                                               const double *__restrict__ d_mean_in, float cut,
                                                               double *__restrict__ partial) {
    const double mean = (OP == RED_SQDEV) ? d_mean_in[0] : 0.0;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    double acc0 = 0.0, acc1 = 0.0;
    const bool aligned = ((uintptr_t)rho & 15u) == 0;
    int64_t nvec = aligned ? n / 4 : 0;
    const float4 *rho4 = reinterpret_cast<const float4 *>(rho);
    int64_t i = tid;
    for (; i + nthreads < nvec; i += 2 * nthreads) {  // two independent 16-byte loads in flight per thread
        const float4 a = __ldg(rho4 + i);
        const float4 b = __ldg(rho4 + i + nthreads);
        acc0 += red_term<OP>(a.x, mean, cut) + red_term<OP>(a.y, mean, cut);
        acc1 += red_term<OP>(a.z, mean, cut) + red_term<OP>(a.w, mean, cut);
        acc0 += red_term<OP>(b.x, mean, cut) + red_term<OP>(b.y, mean, cut);
        acc1 += red_term<OP>(b.z, mean, cut) + red_term<OP>(b.w, mean, cut);
    }
    for (; i < nvec; i += nthreads) {
        const float4 a = __ldg(rho4 + i);
        acc0 += red_term<OP>(a.x, mean, cut) + red_term<OP>(a.y, mean, cut);
        acc1 += red_term<OP>(a.z, mean, cut) + red_term<OP>(a.w, mean, cut);
    }
    for (int64_t j = nvec * 4 + tid; j < n; j += nthreads) acc0 += red_term<OP>(__ldg(rho + j), mean, cut);
    const double s = block_sum(acc0 + acc1);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// Second stage: one block adds the partials in a fixed order.  mode 0: out[0] = sum / n (mean);
// mode 1: out[1] = sqrt(sum / n) (std); mode 2: out[0] = sum.
__global__ void __launch_bounds__(kRedThreads) reduce_final(const double *__restrict__ partial, int nparts, int64_t n,
                                                             int mode, double *__restrict__ out) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) acc += partial[i];
    const double s = block_sum(acc);
    if (threadIdx.x == 0) {
        if (mode == 0)
            out[0] = s / (double)n;
        else if (mode == 1)
            out[1] = sqrt(s / (double)n);
        else
            out[0] = s;
    }
}

static int red_blocks() { return sm_count() * kRedBlocksPerSm; }

// float64 input variant of RED_ABS_ABOVE (sumOfAbs over an arbitrary numeric array).
__global__ void __launch_bounds__(kRedThreads) reduce_abs_f64(const double *__restrict__ v, int64_t n, float cut,
                                                               double *__restrict__ partial) {
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    const double c = (double)cut;
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nthreads) {
        const double a = fabs(v[i]);
        acc += a > c ? a : 0.0;
    }
    const double s = block_sum(acc);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void point_density_kernel(pe_geom g, const float *__restrict__ rho, int64_t n,
                                     const int32_t *__restrict__ crs, float *__restrict__ out,
                                     uint8_t *__restrict__ valid) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = wrap_index(crs[3 * i + 0], g.ncrs[0], g.crs_interval[0]);
    const int r = wrap_index(crs[3 * i + 1], g.ncrs[1], g.crs_interval[1]);
    const int s = wrap_index(crs[3 * i + 2], g.ncrs[2], g.crs_interval[2]);
    const bool ok = (c | r | s) >= 0;
    if (out) out[i] = ok ? __ldg(rho + ((int64_t)s * g.ncrs[1] + r) * g.ncrs[0] + c) : 0.0f;
    if (valid) valid[i] = ok ? 1 : 0;
}

__global__ void xyz2crs_kernel(pe_geom g, int64_t n, const double *__restrict__ xyz, int32_t *__restrict__ crs) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c, r, s;
    xyz2crs(g, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], c, r, s);
    crs[3 * i] = c;
    crs[3 * i + 1] = r;
    crs[3 * i + 2] = s;
}

__global__ void crs2xyz_kernel(pe_geom g, int64_t n, const int32_t *__restrict__ crs, double *__restrict__ xyz) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x, y, z;
    crs2xyz(g, crs[3 * i], crs[3 * i + 1], crs[3 * i + 2], x, y, z);
    xyz[3 * i] = x;
    xyz[3 * i + 1] = y;
    xyz[3 * i + 2] = z;
}

// check_geom_shape: the header fields alone -- for entry points that use a geometry only for index <-> coordinate arithmetic
// (pe_slab_relabel takes the WHOLE map's geometry while its voxels live in slabs on several GPUs); check_geom adds the limit of
// the entry points that address the stored voxels with 32-bit element offsets.
int check_geom_shape(const pe_geom *g, int64_t *nvox_out) {
    PE_CHECK_ARG(g != nullptr, "geometry pointer is null");
    int64_t nvox = 1;
    for (int a = 0; a < 3; ++a) {
        PE_CHECK_ARG(g->ncrs[a] > 0, "ncrs[%d] = %d must be positive", a, g->ncrs[a]);
        PE_CHECK_ARG(g->crs_interval[a] > 0 && g->xyz_interval[a] > 0, "sampling interval of axis %d must be positive", a);
        PE_CHECK_ARG(g->unique_ncrs[a] > 0 && g->unique_ncrs[a] <= g->ncrs[a], "unique_ncrs[%d] out of range", a);
        PE_CHECK_ARG(g->map2xyz[a] >= 0 && g->map2xyz[a] < 3 && g->map2crs[a] >= 0 && g->map2crs[a] < 3,
                     "axis permutation entry %d out of range", a);
        PE_CHECK_ARG(g->mv_perm[a] >= 0 && g->mv_perm[a] < 3, "mv_perm[%d] out of range", a);
        nvox *= g->ncrs[a];
    }
    PE_CHECK_ARG(g->map2crs[g->map2xyz[0]] == 0 && g->map2crs[g->map2xyz[1]] == 1 && g->map2crs[g->map2xyz[2]] == 2,
                 "map2xyz / map2crs are not inverse permutations");
    if (nvox_out) *nvox_out = nvox;
    return PE_OK;
}

int check_geom(const pe_geom *g) {
    int64_t nvox = 0;
    if (int rc = check_geom_shape(g, &nvox)) return rc;
    PE_CHECK_ARG(nvox < (1ll << 31),
                 "maps with 2^31 or more stored voxels are not supported in one call (%lld): label them in slabs (pe_slab_*)", (long long)nvox);
    return PE_OK;
}

}  // namespace pe

using namespace pe;

extern "C" {

int64_t pe_stats_workspace_bytes(void) { return align_up((int64_t)red_blocks() * sizeof(double) + 64, 256); }

int pe_map_mean_std(const float *d_rho, int64_t n, double *d_out, void *d_ws, void *stream) {
    PE_CHECK_ARG(d_rho && d_out && d_ws && n > 0, "pe_map_mean_std: null pointer or empty map");
    cudaStream_t st = (cudaStream_t)stream;
    double *partial = (double *)d_ws;
    const int nb = red_blocks();
    PE_LAUNCH("reduce_partial", st, reduce_partial<RED_SUM><<<nb, kRedThreads, 0, st>>>(d_rho, n, nullptr, 0.f, partial));
    PE_LAUNCH("reduce_final", st, reduce_final<<<1, kRedThreads, 0, st>>>(partial, nb, n, 0, d_out));
    PE_LAUNCH("reduce_partial", st, reduce_partial<RED_SQDEV><<<nb, kRedThreads, 0, st>>>(d_rho, n, d_out, 0.f, partial));
    PE_LAUNCH("reduce_final", st, reduce_final<<<1, kRedThreads, 0, st>>>(partial, nb, n, 1, d_out));
    PE_LAUNCH_CHECK();
    return PE_OK;
}

int pe_map_sum_abs(const float *d_rho, int64_t n, float cutoff, double *d_out, void *d_ws, void *stream) {
    PE_CHECK_ARG(d_rho && d_out && d_ws && n > 0, "pe_map_sum_abs: null pointer or empty map");
    cudaStream_t st = (cudaStream_t)stream;
    double *partial = (double *)d_ws;
    const int nb = red_blocks();
    PE_LAUNCH("reduce_partial", st, reduce_partial<RED_ABS_ABOVE><<<nb, kRedThreads, 0, st>>>(d_rho, n, nullptr, cutoff, partial));
    PE_LAUNCH("reduce_final", st, reduce_final<<<1, kRedThreads, 0, st>>>(partial, nb, n, 2, d_out));
    PE_LAUNCH_CHECK();
    return PE_OK;
}

int pe_sum_abs_f64(const double *d_values, int64_t n, float cutoff, double *d_out, void *d_ws, void *stream) {
    PE_CHECK_ARG(d_values && d_out && d_ws && n > 0, "pe_sum_abs_f64: null pointer or empty array");
    cudaStream_t st = (cudaStream_t)stream;
    double *partial = (double *)d_ws;
    const int nb = red_blocks();
    PE_LAUNCH("reduce_abs_f64", st, reduce_abs_f64<<<nb, kRedThreads, 0, st>>>(d_values, n, cutoff, partial));
    PE_LAUNCH("reduce_final", st, reduce_final<<<1, kRedThreads, 0, st>>>(partial, nb, n, 2, d_out));
    PE_LAUNCH_CHECK();
    return PE_OK;
}

int pe_point_density(const pe_geom *g, const float *d_rho, int64_t n, const int32_t *d_crs, float *d_rho_out,
                     uint8_t *d_valid, void *stream) {
    if (int rc = check_geom(g)) return rc;
    PE_CHECK_ARG(n >= 0 && (n == 0 || (d_rho && d_crs)), "pe_point_density: null pointer");
    if (n == 0) return PE_OK;
    const int threads = 256;
    PE_LAUNCH("point_density_kernel", (cudaStream_t)stream, point_density_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
        *g, d_rho, n, d_crs, d_rho_out, d_valid));
    PE_LAUNCH_CHECK();
    return PE_OK;
}

int pe_xyz2crs(const pe_geom *g, int64_t n, const double *d_xyz, int32_t *d_crs, void *stream) {
    if (int rc = check_geom(g)) return rc;
    PE_CHECK_ARG(n >= 0 && (n == 0 || (d_xyz && d_crs)), "pe_xyz2crs: null pointer");
    if (n == 0) return PE_OK;
    const int threads = 256;
    PE_LAUNCH("xyz2crs_kernel", (cudaStream_t)stream, xyz2crs_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(*g, n, d_xyz, d_crs));
    PE_LAUNCH_CHECK();
    return PE_OK;
}

int pe_crs2xyz(const pe_geom *g, int64_t n, const int32_t *d_crs, double *d_xyz, void *stream) {
    if (int rc = check_geom(g)) return rc;
    PE_CHECK_ARG(n >= 0 && (n == 0 || (d_xyz && d_crs)), "pe_crs2xyz: null pointer");
    if (n == 0) return PE_OK;
    const int threads = 256;
    PE_LAUNCH("crs2xyz_kernel", (cudaStream_t)stream, crs2xyz_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(*g, n, d_crs, d_xyz));
    PE_LAUNCH_CHECK();
    return PE_OK;
}

}  // extern "C"

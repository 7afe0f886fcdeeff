// pe_cloudstats.cu -- the per-atom-type statistics block of aggregateCloud (pdb_eda/densityAnalysis.py:734-766) for every
// structure of a batch, on the device.
//
// The reference builds a numpy structured array of the contributing atoms and then, per atom type, takes np.nanmedian of a
// dozen columns and fits scipy.stats.linregress(log(bfactor), domain fraction) with its p-value test.  On a batch of thousands
// of structures that block would leave the GPU idle behind host sorts, so it runs here: one CTA per structure, one warp per
// atom type present in it.  Atoms are visited through a permutation that lists a structure's atoms type by type (static, built
// by the host once per batch), so a (structure, type) group is a contiguous segment.  Medians are exact order statistics by
// radix select on the IEEE bit patterns (np.nanmedian: NaNs ignored, mean of the two middle values for an even count); sums
// run in a fixed order, so results are run-to-run deterministic.
#include <limits.h>
#include <math.h>
#include "pe_select.cuh"

namespace pe {

// Regularised incomplete beta function I_x(a, b) (continued fraction, modified Lentz): the two-sided p-value of the
// slope's t statistic is I_{df / (df + t^2)}(df / 2, 1 / 2) = 2 * stdtr(df, -|t|), which is what scipy.stats.linregress
// reports and pdb_eda/densityAnalysis.py:733 compares with 0.05.
__device__ double beta_cf(double a, double b, double x) {
    const double tiny = 1.0e-300, eps = 1.0e-16;
    const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0, d = 1.0 - qab * x / qap;
    if (fabs(d) < tiny) d = tiny;
    d = 1.0 / d;
    double h = d;
    for (int m = 1; m <= 5000; ++m) {
        const double m2 = 2.0 * m;
        double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1.0 + aa * d;
        if (fabs(d) < tiny) d = tiny;
        c = 1.0 + aa / c;
        if (fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        h *= d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1.0 + aa * d;
        if (fabs(d) < tiny) d = tiny;
        c = 1.0 + aa / c;
        if (fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < eps) break;
    }
    return h;
}
__device__ double beta_inc(double a, double b, double x) {
    if (!(x > 0.0)) return 0.0;
    if (!(x < 1.0)) return 1.0;
    const double bt = exp(lgamma(a + b) - lgamma(a) - lgamma(b) + a * log(x) + b * log1p(-x));
    if (x < (a + 1.0) / (a + b + 2.0)) return bt * beta_cf(a, b, x) / a;
    return 1.0 - bt * beta_cf(b, a, 1.0 - x) / b;
}

constexpr int kStatThreads = 1024;
constexpr int kStatCols = 14;  // kept rows, ten statistics, contributing atoms, completely overlapped atoms, spare
// scratch columns (structure-of-arrays over the permuted atom order, n_atoms doubles each)
enum { C_DER = 0, C_NV, C_CD, C_BF, C_ADJ, C_DOM, C_COR, C_KEEP, C_FLAGS, kScratchCols };

// A: gathers the per-atom inputs into permuted (type by type) order, so that every later pass streams contiguous memory.
__global__ void __launch_bounds__(kStatThreads)
    stats_prepare_kernel(int n_atoms, const int32_t *__restrict__ perm, const double *__restrict__ atom_out,
                         const double *__restrict__ atom_static, double *__restrict__ sc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_atoms) return;
    const int a = perm[i];
    const double *rec = atom_out + (int64_t)a * 8;
    const int flags = (int)rec[7];
    const bool acc = (flags & 1) != 0;
    const int64_t n = n_atoms;
    sc[C_DER * n + i] = acc ? rec[3] / atom_static[3 * a] / atom_static[3 * a + 1] : nan("");  // totalDensity / electrons / occupancy (:641)
    sc[C_NV * n + i] = rec[1];
    sc[C_CD * n + i] = acc ? rec[2] : nan("");
    sc[C_BF * n + i] = atom_static[3 * a + 2];
    sc[C_KEEP * n + i] = acc ? 1.0 : 0.0;
    sc[C_FLAGS * n + i] = (double)flags;
}

// B: the second centroid filter (:746-748): rows with centroid_distance < nanmedian + 2 nanstd survive, unless every distance
// is NaN.  One CTA per structure (its atoms are the positions [atom_begin, atom_end) of the permuted order as well).
__global__ void __launch_bounds__(kStatThreads)
    stats_filter_kernel(const pe_batch_map *__restrict__ maps, int n_atoms, const double *__restrict__ map_out, double min_total_electrons,
                        double *__restrict__ sc, double *__restrict__ map_stats) {
    __shared__ SelectShared sel;
    __shared__ double red[kStatThreads / 32];
    const int map_id = blockIdx.x;
    const int a0 = maps[map_id].atom_begin, a1 = maps[map_id].atom_end;
    const int64_t n = n_atoms;
    const double *cd = sc + C_CD * n;
    double *keep = sc + C_KEEP * n;
    const double *mo = map_out + (int64_t)map_id * 8;
    double *ms = map_stats + (int64_t)map_id * 4;
    const bool ok = mo[2] >= min_total_electrons;  // otherwise aggregateCloud returns before the statistics (:726)
    double cnt = 0.0, sum = 0.0;
    for (int i = a0 + threadIdx.x; i < a1; i += blockDim.x)
        if (!isnan(cd[i])) {
            cnt += 1.0;
            sum += cd[i];
        }
    cnt = block_sum_fixed(cnt, red);
    sum = block_sum_fixed(sum, red);
    const double mean = sum / cnt;
    double sq = 0.0;
    for (int i = a0 + threadIdx.x; i < a1; i += blockDim.x)
        if (!isnan(cd[i])) {
            const double d = cd[i] - mean;
            sq += d * d;
        }
    sq = block_sum_fixed(sq, red);
    const double med = block_nanmedian(a1 - a0, [&](int i) { return isnan(cd[a0 + i]) ? kNoKey : order_key(cd[a0 + i]); }, sel);
    const double cutoff = med + sqrt(sq / cnt) * 2.0;
    const bool all_nan = cnt == 0.0;
    double kept = 0.0;
    for (int i = a0 + threadIdx.x; i < a1; i += blockDim.x) {
        if (keep[i] != 0.0 && !all_nan && !(cd[i] < cutoff)) keep[i] = 0.0;
        kept += keep[i];
    }
    kept = block_sum_fixed(kept, red);
    if (threadIdx.x == 0) {
        ms[0] = ok ? kept : 0.0;
        ms[1] = ok ? mo[1] / mo[2] : nan("");
        ms[2] = cutoff;
        ms[3] = ok ? 1.0 : 0.0;
    }
}

// C: one CTA per (structure, atom type) segment.  Each median stages the column's keys in shared memory when the segment
// fits (stage_cap keys), so the nine select passes do not go back to L2.
__global__ void __launch_bounds__(kStatThreads)
    stats_segment_kernel(int n_atoms, const double *__restrict__ map_out, const double *__restrict__ map_stats,
                         const int32_t *__restrict__ seg_map, const int32_t *__restrict__ seg_type, const int32_t *__restrict__ seg_begin,
                         const int32_t *__restrict__ seg_end, const double *__restrict__ unit_volume,
                         const double *__restrict__ current_slopes, int stage_cap, double *__restrict__ sc, double *__restrict__ seg_out) {
    extern __shared__ __align__(16) unsigned long long stage[];
    __shared__ SelectShared sel;
    __shared__ double red[kStatThreads / 32];
    __shared__ double bcast[2];
    const int seg = blockIdx.x;
    const int b = seg_begin[seg], len = seg_end[seg] - b;
    const int map_id = seg_map[seg];
    const int64_t n_all = n_atoms;
    const double *der = sc + C_DER * n_all + b, *nv = sc + C_NV * n_all + b, *cd = sc + C_CD * n_all + b;
    const double *keep = sc + C_KEEP * n_all + b, *flags = sc + C_FLAGS * n_all + b;
    double *bf = sc + C_BF * n_all + b, *adj = sc + C_ADJ * n_all + b, *dom = sc + C_DOM * n_all + b, *cor = sc + C_COR * n_all + b;
    double *so = seg_out + (int64_t)seg * kStatCols;
    const bool ok = map_stats[(int64_t)map_id * 4 + 3] != 0.0;
    const double ratio = map_stats[(int64_t)map_id * 4 + 1];
    const double uvol = unit_volume[map_id];
    const int tid = threadIdx.x;
    const bool staged = len <= stage_cap;

    double n_keep = 0.0, n_acc = 0.0, n_full = 0.0;
    for (int i = tid; i < len; i += blockDim.x) {
        n_keep += keep[i];
        const int fl = (int)flags[i];  // 1: contributes, 3: and touches all its contributing bonded atoms (:653-659)
        n_acc += (double)(fl & 1);
        n_full += (double)((fl >> 1) & 1);
    }
    n_keep = block_sum_fixed(n_keep, red);
    n_acc = block_sum_fixed(n_acc, red);
    n_full = block_sum_fixed(n_full, red);
    if (tid == 0) {
        so[11] = n_acc;
        so[12] = n_full;
        so[13] = 0.0;
    }
    if (n_keep == 0.0 || !ok) {
        if (tid < 11) so[tid] = tid == 0 ? 0.0 : nan("");
        return;
    }
    const int n = (int)n_keep;
    // np.nanmedian over the kept rows of value(i), optionally restricted by use(i)
    auto median = [&](auto value, auto use) -> double {
        if (staged) {
            for (int i = tid; i < len; i += blockDim.x) {
                const double v = value(i);
                stage[i] = (keep[i] != 0.0 && use(i) && !isnan(v)) ? order_key(v) : kNoKey;
            }
            __syncthreads();
            const double m = block_nanmedian(len, [&](int i) { return stage[i]; }, sel);
            __syncthreads();
            return m;
        }
        return block_nanmedian(len, [&](int i) {
            const double v = value(i);
            return (keep[i] != 0.0 && use(i) && !isnan(v)) ? order_key(v) : kNoKey; }, sel);
    };
    auto all = [](int) { return true; };
    const double med_nv = median([&](int i) { return nv[i]; }, all);
    for (int i = tid; i < len; i += blockDim.x)
        if (keep[i] != 0.0) adj[i] = der[i] / nv[i] * med_nv;  // adj_density_electron_ratio (:752)
    __syncthreads();
    const double med_der = median([&](int i) { return der[i]; }, all);
    const double med_cd = median([&](int i) { return cd[i]; }, all);
    const double med_adj = median([&](int i) { return adj[i]; }, all);
    const double med_vol = median([&](int i) { return nv[i] * uvol; }, all);
    const double med_bf = median([&](int i) { return bf[i]; }, [&](int i) { return bf[i] > 0.0; });
    // b-factors <= 0 take the type's median (:757); x = log(bfactor), y = domain fraction of the adjusted ratio
    double sx = 0.0, sy = 0.0;
    int first_i = INT_MAX;
    for (int i = tid; i < len; i += blockDim.x)
        if (keep[i] != 0.0) {
            if (bf[i] <= 0.0) bf[i] = med_bf;
            const double y = (adj[i] - ratio) / ratio;
            dom[i] = y;
            sx += log(bf[i]);
            sy += y;
            first_i = min(first_i, i);
        }
    __syncthreads();
    sx = block_sum_fixed(sx, red);
    sy = block_sum_fixed(sy, red);
    // the first kept b-factor, then "are they all one value" (len(np.unique(bfactor)) == 1; NaN == NaN there)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) first_i = min(first_i, __shfl_xor_sync(kFull, first_i, o));
    __shared__ int first_w[kStatThreads / 32];
    if ((tid & 31) == 0) first_w[tid >> 5] = first_i;
    __syncthreads();
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) first_i = min(first_i, first_w[w]);
    const double first_bf = bf[first_i];
    const double xm = sx / n, ym = sy / n;
    double sxx = 0.0, syy = 0.0, sxy = 0.0;
    int differs = 0;
    for (int i = tid; i < len; i += blockDim.x)
        if (keep[i] != 0.0) {
            const double dx = log(bf[i]) - xm, dy = dom[i] - ym;
            sxx += dx * dx;
            syy += dy * dy;
            sxy += dx * dy;
            differs |= (bf[i] == first_bf || (isnan(bf[i]) && isnan(first_bf))) ? 0 : 1;
        }
    sxx = block_sum_fixed(sxx, red) / n;
    syy = block_sum_fixed(syy, red) / n;
    sxy = block_sum_fixed(sxy, red) / n;
    const bool one_value = __syncthreads_or(differs) == 0;
    const double current = current_slopes[seg_type[seg]];
    double slope = current;
    if (n > 2 && !one_value) {  // calcSlope (:729-733)
        if (tid == 0) {
            double r;
            if (sxx == 0.0 || syy == 0.0)
                r = sxy == 0.0 ? nan("") : 0.0;
            else
                r = fmin(fmax(sxy / sqrt(sxx * syy), -1.0), 1.0);
            const double df = (double)(n - 2);
            const double t = r * sqrt(df / ((1.0 - r + 1.0e-20) * (1.0 + r + 1.0e-20)));
            double p = nan("");
            if (!isnan(t)) p = isinf(t) ? 0.0 : beta_inc(0.5 * df, 0.5, df / (df + t * t));
            bcast[0] = (p > 0.05) ? current : sxy / sxx;
        }
        __syncthreads();
        slope = bcast[0];
    }
    // b-factor correction (:761-764)
    const double lmed = log(med_bf);
    for (int i = tid; i < len; i += blockDim.x)
        if (keep[i] != 0.0) cor[i] = dom[i] - (log(bf[i]) - lmed) * slope;
    __syncthreads();
    const double med_dom = median([&](int i) { return dom[i]; }, all);
    const double med_cor = median([&](int i) { return cor[i]; }, all);
    const double med_crr = median([&](int i) { return cor[i] * ratio + ratio; }, all);
    if (tid == 0) {
        so[0] = (double)n;
        so[1] = med_nv;
        so[2] = med_der;
        so[3] = med_cd;
        so[4] = med_adj;
        so[5] = med_vol;
        so[6] = med_bf;
        so[7] = slope;
        so[8] = med_dom;
        so[9] = med_cor;
        so[10] = med_crr;
    }
}

}  // namespace pe

using namespace pe;

extern "C" {

int pe_cloud_statistics(int32_t n_maps, const pe_batch_map *d_maps, int32_t n_atoms, const double *d_atom_out, const double *d_map_out,
                        const double *d_atom_static, const int32_t *d_perm, int32_t n_segments, const int32_t *d_seg_map,
                        const int32_t *d_seg_type, const int32_t *d_seg_begin, const int32_t *d_seg_end, int32_t max_segment,
                        const double *d_unit_volume, const double *d_current_slopes, double min_total_electrons, double *d_scratch,
                        double *d_seg_out, double *d_map_stats, void *stream) {
    PE_CHECK_ARG(n_maps >= 0 && n_atoms >= 0 && n_segments >= 0 && max_segment >= 0, "pe_cloud_statistics: negative size");
    if (n_maps == 0) return PE_OK;
    PE_CHECK_ARG(d_maps && d_map_out && d_unit_volume && d_current_slopes && d_map_stats, "pe_cloud_statistics: null pointer");
    PE_CHECK_ARG(n_atoms == 0 || (d_atom_out && d_atom_static && d_perm && d_scratch), "pe_cloud_statistics: null pointer");
    PE_CHECK_ARG(n_segments == 0 || (d_seg_map && d_seg_type && d_seg_begin && d_seg_end && d_seg_out), "pe_cloud_statistics: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_atoms > 0)
        PE_LAUNCH("stats_prepare_kernel", st, stats_prepare_kernel<<<(n_atoms + kStatThreads - 1) / kStatThreads, kStatThreads, 0, st>>>(
            n_atoms, d_perm, d_atom_out, d_atom_static, d_scratch));
    PE_LAUNCH("stats_filter_kernel", st, stats_filter_kernel<<<n_maps, kStatThreads, 0, st>>>(d_maps, n_atoms, d_map_out, min_total_electrons,
                                                                                       d_scratch, d_map_stats));
    if (n_segments > 0) {
        const int stage_cap = max_segment < 12288 ? max_segment : 12288;  // <= 96 KB of keys per CTA
        const size_t smem = (size_t)(stage_cap > 0 ? stage_cap : 1) * sizeof(unsigned long long);
        PE_CUDA(cudaFuncSetAttribute(stats_segment_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PE_LAUNCH("stats_segment_kernel", st, stats_segment_kernel<<<n_segments, kStatThreads, smem, st>>>(
            n_atoms, d_map_out, d_map_stats, d_seg_map, d_seg_type, d_seg_begin, d_seg_end, d_unit_volume, d_current_slopes, stage_cap,
            d_scratch, d_seg_out));
    }
    PE_LAUNCH_CHECK();
    return PE_OK;
}

}  // extern "C"

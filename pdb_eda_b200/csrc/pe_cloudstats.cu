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
#include "pe_common.cuh"

namespace pe {

// monotone map of a double's bits to an unsigned key (negative values included)
__device__ __forceinline__ unsigned long long order_key(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v + 0.0);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double order_value(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

// Warp-level nanmedian of value(i) over i in [begin, end) with use(i) (NaNs skipped).  hist: 256 words of shared memory
// owned by the warp.  Returns NaN when nothing qualifies.
template <class V, class U>
__device__ double warp_nanmedian(int begin, int end, V value, U use, unsigned int *hist, int lane) {
    int n = 0;
    for (int i = begin + lane; i < end; i += 32)
        if (use(i) && !isnan(value(i))) ++n;
    n = warp_sum(n);
    if (n == 0) return nan("");
    double mid[2] = {0.0, 0.0};
    for (int which = 0; which < 2; ++which) {
        if (which == 1 && (n & 1)) {
            mid[1] = mid[0];
            break;
        }
        unsigned int want = which == 0 ? (unsigned)(n - 1) / 2u : (unsigned)n / 2u;  // 0-based rank
        unsigned long long prefix = 0ull;
        for (int shift = 56; shift >= 0; shift -= 8) {
            for (int k = lane; k < 256; k += 32) hist[k] = 0u;
            __syncwarp();
            const unsigned long long himask = shift == 56 ? 0ull : (~0ull << (shift + 8));
            for (int i = begin + lane; i < end; i += 32) {
                if (!use(i)) continue;
                const double v = value(i);
                if (isnan(v)) continue;
                const unsigned long long key = order_key(v);
                if ((key & himask) == prefix) atomicAdd(&hist[(key >> shift) & 0xffu], 1u);
            }
            __syncwarp();
            unsigned int mine = 0;  // lane l owns bins 8l .. 8l+7
#pragma unroll
            for (int k = 0; k < 8; ++k) mine += hist[8 * lane + k];
            const unsigned int before = (unsigned int)warp_excl_scan((int)mine, lane);
            const bool holds = want >= before && want < before + mine;
            const int src = __ffs(__ballot_sync(kFull, holds)) - 1;
            unsigned int bin = 0, rank = 0;
            if (lane == src) {
                unsigned int acc = before;
                int k = 0;
                for (; k < 7; ++k) {
                    if (acc + hist[8 * lane + k] > want) break;
                    acc += hist[8 * lane + k];
                }
                bin = (unsigned int)(8 * lane + k);
                rank = want - acc;
            }
            bin = __shfl_sync(kFull, bin, src);
            want = __shfl_sync(kFull, rank, src);
            prefix |= (unsigned long long)bin << shift;
            __syncwarp();
        }
        mid[which] = order_value(prefix);
    }
    return (n & 1) ? mid[0] : (mid[0] + mid[1]) / 2.0;
}

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

constexpr int kStatWarps = 8;
constexpr int kStatCols = 14;  // kept rows, ten statistics, contributing atoms, completely overlapped atoms, spare

__device__ __forceinline__ double stat_block_sum(double v, double *scratch) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    double tot = 0.0;
    for (int k = 0; k < kStatWarps; ++k) tot += scratch[k];
    return tot;
}

__global__ void __launch_bounds__(kStatWarps * 32)
    atom_stats_kernel(const pe_batch_map *__restrict__ maps, const double *__restrict__ atom_out, const double *__restrict__ map_out,
                      const double *__restrict__ atom_static /* n x 3: electrons, occupancy, bfactor */,
                      const int32_t *__restrict__ perm, const int32_t *__restrict__ map_seg_ptr,
                      const int32_t *__restrict__ seg_type, const int32_t *__restrict__ seg_begin, const int32_t *__restrict__ seg_end,
                      const double *__restrict__ unit_volume, const double *__restrict__ current_slopes, double min_total_electrons,
                      double *__restrict__ scratch /* n x 6 */, double *__restrict__ seg_out, double *__restrict__ map_stats /* n_maps x 4 */) {
    __shared__ unsigned int hist[kStatWarps][256];
    __shared__ double red[kStatWarps];
    __shared__ double cutoff_s;
    __shared__ int all_nan_s;
    const int map_id = blockIdx.x;
    const pe_batch_map *m = maps + map_id;
    const int a0 = m->atom_begin, a1 = m->atom_end;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double *mo = map_out + (int64_t)map_id * 8;
    double *ms = map_stats + (int64_t)map_id * 4;
    const double total_e = mo[2], total_d = mo[1];
    const bool ok = total_e >= min_total_electrons;  // otherwise aggregateCloud returns before the statistics (:726)
    const double ratio = total_d / total_e;
    const double uvol = unit_volume[map_id];
    // scratch columns per atom: 0 density_electron_ratio, 1 adj_density_electron_ratio, 2 bfactor (<= 0 replaced),
    // 3 domain_fraction, 4 corrected_fraction, 5 keep flag
    auto rec = [&](int a, int k) { return atom_out[(int64_t)a * 8 + k]; };
    for (int a = a0 + threadIdx.x; a < a1; a += blockDim.x) {
        double *sc = scratch + (int64_t)a * 6;
        const bool acc = ((int)rec(a, 7) & 1) != 0;
        sc[0] = acc ? rec(a, 3) / atom_static[3 * a] / atom_static[3 * a + 1] : 0.0;  // totalDensity / electrons / occupancy (:641)
        sc[5] = acc ? 1.0 : 0.0;
    }
    __syncthreads();
    // centroid filter (:746-748): keep rows with centroid_distance < nanmedian + 2 nanstd unless every distance is NaN
    {
        double cnt = 0.0, sum = 0.0;
        for (int a = a0 + threadIdx.x; a < a1; a += blockDim.x)
            if (scratch[(int64_t)a * 6 + 5] != 0.0 && !isnan(rec(a, 2))) {
                cnt += 1.0;
                sum += rec(a, 2);
            }
        cnt = stat_block_sum(cnt, red);
        sum = stat_block_sum(sum, red);
        const double mean = sum / cnt;
        double sq = 0.0;
        for (int a = a0 + threadIdx.x; a < a1; a += blockDim.x)
            if (scratch[(int64_t)a * 6 + 5] != 0.0 && !isnan(rec(a, 2))) {
                const double d = rec(a, 2) - mean;
                sq += d * d;
            }
        sq = stat_block_sum(sq, red);
        if (warp == 0) {
            const double med = warp_nanmedian(a0, a1, [&](int a) { return rec(a, 2); },
                                              [&](int a) { return scratch[(int64_t)a * 6 + 5] != 0.0; }, hist[0], lane);
            if (lane == 0) {
                cutoff_s = med + sqrt(sq / cnt) * 2.0;
                all_nan_s = cnt == 0.0 ? 1 : 0;
            }
        }
        __syncthreads();
        double kept = 0.0;
        for (int a = a0 + threadIdx.x; a < a1; a += blockDim.x) {
            double *sc = scratch + (int64_t)a * 6;
            if (sc[5] != 0.0 && !all_nan_s && !(rec(a, 2) < cutoff_s)) sc[5] = 0.0;
            kept += sc[5];
        }
        kept = stat_block_sum(kept, red);
        if (threadIdx.x == 0) {
            ms[0] = ok ? kept : 0.0;
            ms[1] = ok ? ratio : nan("");
            ms[2] = cutoff_s;
            ms[3] = ok ? 1.0 : 0.0;
        }
    }
    __syncthreads();
    // one warp per atom type of the structure
    for (int seg = map_seg_ptr[map_id] + warp; seg < map_seg_ptr[map_id + 1]; seg += kStatWarps) {
        const int b = seg_begin[seg], e = seg_end[seg];
        double *so = seg_out + (int64_t)seg * kStatCols;
        unsigned int *h = hist[warp];
        auto keep = [&](int i) { return scratch[(int64_t)perm[i] * 6 + 5] != 0.0; };
        int n = 0, n_acc = 0, n_full = 0;
        for (int i = b + lane; i < e; i += 32) {
            n += keep(i) ? 1 : 0;
            const int fl = (int)rec(perm[i], 7);  // 1: contributes, 3: and touches all its contributing bonded atoms (:653-659)
            n_acc += fl & 1;
            n_full += (fl >> 1) & 1;
        }
        n = warp_sum(n);
        n_acc = warp_sum(n_acc);
        n_full = warp_sum(n_full);
        if (lane == 0) {
            so[11] = (double)n_acc;
            so[12] = (double)n_full;
            so[13] = 0.0;
        }
        if (n == 0 || !ok) {
            if (lane < 11) so[lane] = lane == 0 ? 0.0 : nan("");
            continue;
        }
        const double med_nv = warp_nanmedian(b, e, [&](int i) { return rec(perm[i], 1); }, keep, h, lane);
        for (int i = b + lane; i < e; i += 32)
            if (keep(i)) {
                const int a = perm[i];
                scratch[(int64_t)a * 6 + 1] = scratch[(int64_t)a * 6] / rec(a, 1) * med_nv;  // adj_density_electron_ratio (:752)
            }
        __syncwarp();
        const double med_der = warp_nanmedian(b, e, [&](int i) { return scratch[(int64_t)perm[i] * 6]; }, keep, h, lane);
        const double med_cd = warp_nanmedian(b, e, [&](int i) { return rec(perm[i], 2); }, keep, h, lane);
        const double med_adj = warp_nanmedian(b, e, [&](int i) { return scratch[(int64_t)perm[i] * 6 + 1]; }, keep, h, lane);
        const double med_vol = warp_nanmedian(b, e, [&](int i) { return rec(perm[i], 1) * uvol; }, keep, h, lane);
        const double med_bf = warp_nanmedian(b, e, [&](int i) { return atom_static[3 * perm[i] + 2]; },
                                             [&](int i) { return keep(i) && atom_static[3 * perm[i] + 2] > 0.0; }, h, lane);
        // b-factors <= 0 take the type's median (:757); x = log(bfactor), y = domain fraction of the adjusted ratio
        double sx = 0.0, sy = 0.0;
        double first_bf = 0.0;
        int first_i = INT_MAX;
        for (int i = b + lane; i < e; i += 32)
            if (keep(i)) {
                const int a = perm[i];
                double bf = atom_static[3 * a + 2];
                if (bf <= 0.0) bf = med_bf;
                scratch[(int64_t)a * 6 + 2] = bf;
                const double y = (scratch[(int64_t)a * 6 + 1] - ratio) / ratio;
                scratch[(int64_t)a * 6 + 3] = y;
                sx += log(bf);
                sy += y;
                if (i < first_i) {
                    first_i = i;
                    first_bf = bf;
                }
            }
        __syncwarp();
        sx = warp_sum(sx);
        sy = warp_sum(sy);
        // the first kept b-factor of the segment, then "are they all one value" (len(np.unique(bfactor)) == 1; NaN == NaN there)
        int fi = first_i;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) fi = min(fi, __shfl_xor_sync(kFull, fi, o));
        const int owner = __ffs(__ballot_sync(kFull, first_i == fi)) - 1;
        first_bf = __shfl_sync(kFull, first_bf, owner);
        const double xm = sx / n, ym = sy / n;
        double sxx = 0.0, syy = 0.0, sxy = 0.0;
        int differs = 0;
        for (int i = b + lane; i < e; i += 32)
            if (keep(i)) {
                const int a = perm[i];
                const double bf = scratch[(int64_t)a * 6 + 2];
                const double dx = log(bf) - xm, dy = scratch[(int64_t)a * 6 + 3] - ym;
                sxx += dx * dx;
                syy += dy * dy;
                sxy += dx * dy;
                differs |= (bf == first_bf || (isnan(bf) && isnan(first_bf))) ? 0 : 1;
            }
        sxx = warp_sum(sxx) / n;
        syy = warp_sum(syy) / n;
        sxy = warp_sum(sxy) / n;
        const bool one_value = !__any_sync(kFull, differs != 0);
        const double current = current_slopes[seg_type[seg]];
        double slope = current;
        if (n > 2 && !one_value) {  // calcSlope (:729-733)
            double r;
            if (sxx == 0.0 || syy == 0.0)
                r = sxy == 0.0 ? nan("") : 0.0;
            else
                r = fmin(fmax(sxy / sqrt(sxx * syy), -1.0), 1.0);
            const double df = (double)(n - 2);
            const double t = r * sqrt(df / ((1.0 - r + 1.0e-20) * (1.0 + r + 1.0e-20)));
            double p = nan("");
            if (lane == 0 && !isnan(t)) p = isinf(t) ? 0.0 : beta_inc(0.5 * df, 0.5, df / (df + t * t));
            p = __shfl_sync(kFull, p, 0);
            slope = (p > 0.05) ? current : sxy / sxx;
        }
        // b-factor correction (:761-764)
        const double lmed = log(med_bf);
        for (int i = b + lane; i < e; i += 32)
            if (keep(i)) {
                const int a = perm[i];
                scratch[(int64_t)a * 6 + 4] = scratch[(int64_t)a * 6 + 3] - (log(scratch[(int64_t)a * 6 + 2]) - lmed) * slope;
            }
        __syncwarp();
        const double med_dom = warp_nanmedian(b, e, [&](int i) { return scratch[(int64_t)perm[i] * 6 + 3]; }, keep, h, lane);
        const double med_cor = warp_nanmedian(b, e, [&](int i) { return scratch[(int64_t)perm[i] * 6 + 4]; }, keep, h, lane);
        const double med_crr = warp_nanmedian(b, e, [&](int i) { return scratch[(int64_t)perm[i] * 6 + 4] * ratio + ratio; }, keep, h, lane);
        if (lane == 0) {
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
}

}  // namespace pe

using namespace pe;

extern "C" {

int pe_cloud_statistics(int32_t n_maps, const pe_batch_map *d_maps, int32_t n_atoms, const double *d_atom_out, const double *d_map_out,
                        const double *d_atom_static, const int32_t *d_perm, const int32_t *d_map_seg_ptr, int32_t n_segments,
                        const int32_t *d_seg_type, const int32_t *d_seg_begin, const int32_t *d_seg_end, const double *d_unit_volume,
                        const double *d_current_slopes, double min_total_electrons, double *d_scratch, double *d_seg_out,
                        double *d_map_stats, void *stream) {
    PE_CHECK_ARG(n_maps >= 0 && n_atoms >= 0 && n_segments >= 0, "pe_cloud_statistics: negative size");
    if (n_maps == 0) return PE_OK;
    PE_CHECK_ARG(d_maps && d_map_out && d_map_seg_ptr && d_unit_volume && d_current_slopes && d_map_stats, "pe_cloud_statistics: null pointer");
    PE_CHECK_ARG(n_atoms == 0 || (d_atom_out && d_atom_static && d_perm && d_scratch), "pe_cloud_statistics: null pointer");
    PE_CHECK_ARG(n_segments == 0 || (d_seg_type && d_seg_begin && d_seg_end && d_seg_out), "pe_cloud_statistics: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    PE_LAUNCH("atom_stats_kernel", st, atom_stats_kernel<<<n_maps, kStatWarps * 32, 0, st>>>(
        d_maps, d_atom_out, d_map_out, d_atom_static, d_perm, d_map_seg_ptr, d_seg_type, d_seg_begin, d_seg_end, d_unit_volume,
        d_current_slopes, min_total_electrons, d_scratch, d_seg_out, d_map_stats));
    PE_LAUNCH_CHECK();
    return PE_OK;
}

}  // extern "C"

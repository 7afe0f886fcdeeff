// pe_select.cuh -- exact block-wide order statistics (np.nanmedian) by radix select on IEEE bit patterns.
//
// Used by the centroid-distance cutoffs of aggregateCloud (pdb_eda/densityAnalysis.py:609, :746-748) and by the per-atom-type
// medians of its statistics block (:749-766).  A median is the element of rank (n-1)/2 -- found in eight passes of eight bits
// over the candidates, histogram in shared memory with warp-aggregated atomics (values of one column share their leading
// bytes, so un-aggregated atomics would serialise) -- and, for an even count, the next larger element, found in one more pass.
#pragma once
#include "pe_common.cuh"

namespace pe {

// monotone map of a double's bits to an unsigned key (negative values included; -0.0 == 0.0)
__device__ __forceinline__ unsigned long long order_key(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v + 0.0);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double order_value(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

constexpr unsigned long long kNoKey = 0ull;  // order_key never yields 0 for a non-NaN value (0 would be the bits of a negative NaN)

struct SelectShared {
    unsigned int hist[256];
    unsigned long long wmin[32];
    unsigned int wcnt[32];
    unsigned long long prefix;
    unsigned int rank;
};

// Median of the keys key(i), i in [0, len), skipping entries whose key is kNoKey.  All threads of the block call it with the
// same arguments; returns NaN when there is no key.  key(i) must be cheap: it is evaluated nine or ten times per element.
template <class K>
__device__ double block_nanmedian(int len, K key, SelectShared &sh) {
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
    // number of keys
    unsigned int cnt = 0;
    for (int i = tid; i < len; i += nthr) cnt += key(i) != kNoKey ? 1u : 0u;
    cnt = (unsigned int)warp_sum((int)cnt);
    __syncthreads();
    if (lane == 0) sh.wcnt[warp] = cnt;
    __syncthreads();
    unsigned int n = 0;
    for (int w = 0; w < nwarps; ++w) n += sh.wcnt[w];
    if (n == 0) return nan("");
    unsigned int want = (n - 1u) / 2u;
    unsigned long long prefix = 0ull;
    for (int shift = 56; shift >= 0; shift -= 8) {
        for (int k = tid; k < 256; k += nthr) sh.hist[k] = 0u;
        __syncthreads();
        const unsigned long long himask = shift == 56 ? 0ull : (~0ull << (shift + 8));
        for (int i = tid; i < len; i += nthr) {
            const unsigned long long k = key(i);
            if (k == kNoKey || (k & himask) != prefix) continue;
            const unsigned int bin = (unsigned int)(k >> shift) & 0xffu;
            const unsigned int peers = __match_any_sync(__activemask(), bin);
            if (lane == __ffs(peers) - 1) atomicAdd(&sh.hist[bin], (unsigned int)__popc(peers));
        }
        __syncthreads();
        if (warp == 0) {
            unsigned int mine = 0;  // lane l owns bins 8l .. 8l+7
#pragma unroll
            for (int k = 0; k < 8; ++k) mine += sh.hist[8 * lane + k];
            const unsigned int before = (unsigned int)warp_excl_scan((int)mine, lane);
            if (want >= before && want < before + mine) {  // exactly one lane
                unsigned int acc = before;
                int k = 0;
                for (; k < 7; ++k) {
                    if (acc + sh.hist[8 * lane + k] > want) break;
                    acc += sh.hist[8 * lane + k];
                }
                sh.prefix = prefix | ((unsigned long long)(8 * lane + k) << shift);
                sh.rank = want - acc;
            }
        }
        __syncthreads();
        prefix = sh.prefix;
        want = sh.rank;
    }
    const double lo = order_value(prefix);
    if (n & 1u) return lo;
    // even count: the element of rank n/2 is `prefix` again when enough copies of it exist, else the smallest larger key
    unsigned int le = 0;
    unsigned long long mn = ~0ull;
    for (int i = tid; i < len; i += nthr) {
        const unsigned long long k = key(i);
        if (k == kNoKey) continue;
        if (k <= prefix)
            ++le;
        else
            mn = k < mn ? k : mn;
    }
    le = (unsigned int)warp_sum((int)le);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(kFull, mn, o);
        mn = other < mn ? other : mn;
    }
    __syncthreads();
    if (lane == 0) {
        sh.wcnt[warp] = le;
        sh.wmin[warp] = mn;
    }
    __syncthreads();
    unsigned int le_all = 0;
    unsigned long long mn_all = ~0ull;
    for (int w = 0; w < nwarps; ++w) {
        le_all += sh.wcnt[w];
        mn_all = sh.wmin[w] < mn_all ? sh.wmin[w] : mn_all;
    }
    const double hi = le_all > n / 2u ? lo : order_value(mn_all);
    return (lo + hi) / 2.0;
}

// Fixed-order block sum (deterministic); scratch: one double per warp.
__device__ __forceinline__ double block_sum_fixed(double v, double *scratch) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    double tot = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) tot += scratch[k];
    return tot;
}

}  // namespace pe

// pe_scan.cu -- error state, device info and the exclusive prefix sum used by the compaction steps.
#include <stdarg.h>
#include <string.h>
#include <atomic>
#include <mutex>
#include <vector>
#include <nvtx3/nvToolsExt.h>  // header-only NVTX 3: a no-op unless a profiler injects its library
#include "pe_common.cuh"

namespace pe {

static thread_local char g_error[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

// ------------------------------------------------------------------------------------------------ launch accounting
static std::atomic<long long> g_launches{0};
static std::atomic<int> g_profiling{0};
struct ProfEvent {
    const char *tag;
    cudaEvent_t start, stop;
};
static std::mutex g_prof_mutex;
static std::vector<ProfEvent> g_prof_events;   // recorded, not yet folded
static std::vector<ProfEvent> g_prof_free;     // reusable event pairs
struct ProfTotal {
    const char *tag;
    long long count;
    double ms;
};
static std::vector<ProfTotal> g_prof_totals;

ProfScope::ProfScope(const char *tag, cudaStream_t st) : slot(-1), stream(st) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    nvtxRangePushA(tag);  // one NVTX range per launch, named like the kernel (tracing)
    if (!g_profiling.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    ProfEvent ev;
    if (!g_prof_free.empty()) {
        ev = g_prof_free.back();
        g_prof_free.pop_back();
    } else {
        if (cudaEventCreate(&ev.start) != cudaSuccess || cudaEventCreate(&ev.stop) != cudaSuccess) return;
    }
    ev.tag = tag;
    cudaEventRecord(ev.start, st);
    g_prof_events.push_back(ev);
    slot = (int)g_prof_events.size() - 1;
}

ProfScope::~ProfScope() {
    nvtxRangePop();
    if (slot < 0) return;
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    if (slot < (int)g_prof_events.size()) cudaEventRecord(g_prof_events[slot].stop, stream);
}

// Folds all recorded event pairs into the per-tag totals (synchronises on each stop event).
static void prof_fold() {
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    for (ProfEvent &ev : g_prof_events) {
        float ms = 0.f;
        if (cudaEventSynchronize(ev.stop) == cudaSuccess && cudaEventElapsedTime(&ms, ev.start, ev.stop) == cudaSuccess) {
            bool found = false;
            for (ProfTotal &t : g_prof_totals)
                if (strcmp(t.tag, ev.tag) == 0) {
                    t.count += 1;
                    t.ms += ms;
                    found = true;
                    break;
                }
            if (!found) g_prof_totals.push_back({ev.tag, 1, (double)ms});
        }
        g_prof_free.push_back(ev);
    }
    g_prof_events.clear();
}

int sm_count() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            return 148;
    }
    return cached;
}

// ------------------------------------------------------------------------------------------------ scan
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

template <bool POPC>
__device__ __forceinline__ uint32_t scan_load(const uint32_t *in, int64_t i, int64_t n) {
    if (i >= n) return 0u;
    const uint32_t v = in[i];
    return POPC ? (uint32_t)__popc(v) : v;
}

// Block-wide exclusive scan of one value per thread; returns the exclusive prefix, *total = block sum.
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *total) {
    __shared__ uint32_t warp_tot[kScanThreads / 32];
    __shared__ uint32_t block_tot;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(kFull, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[w] = x;
    __syncthreads();
    if (w == 0) {
        uint32_t t = lane < kScanThreads / 32 ? warp_tot[lane] : 0u;
        uint32_t s = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(kFull, s, o);
            if (lane >= o) s += y;
        }
        if (lane < kScanThreads / 32) warp_tot[lane] = s - t;
        if (lane == kScanThreads / 32 - 1) block_tot = s;
    }
    __syncthreads();
    const uint32_t res = x - v + warp_tot[w];
    *total = block_tot;
    __syncthreads();
    return res;
}

template <bool POPC>
__global__ void __launch_bounds__(kScanThreads) scan_tile_sums(const uint32_t *__restrict__ in, int64_t n_host,
                                                                const int64_t *__restrict__ d_n,
                                                                uint32_t *__restrict__ block_sums) {
    const int64_t n = d_n ? min(*d_n, n_host) : n_host;  // never past the buffer capacity
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    uint32_t s = 0;
    if (base < n) {
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) s += scan_load<POPC>(in, base + k, n);
    }
    uint32_t tot;
    block_excl_scan(s, &tot);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024) scan_block_offsets(uint32_t *__restrict__ block_sums, int nblocks,
                                                            int64_t *__restrict__ d_total) {
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry_s;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int start = 0; start < nblocks; start += 1024) {
        const int i = start + threadIdx.x;
        const uint32_t v = i < nblocks ? block_sums[i] : 0u;
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(kFull, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_tot[w] = x;
        __syncthreads();
        if (w == 0) {
            const uint32_t t = warp_tot[lane];
            uint32_t s = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(kFull, s, o);
                if (lane >= o) s += y;
            }
            warp_tot[lane] = s - t;
        }
        __syncthreads();
        const uint32_t carry = carry_s;
        const uint32_t excl = x - v + warp_tot[w] + carry;
        if (i < nblocks) block_sums[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0 && d_total) *d_total = (int64_t)carry_s;
}

template <bool POPC>
__global__ void __launch_bounds__(kScanThreads) scan_apply(const uint32_t *__restrict__ in, uint32_t *__restrict__ out,
                                                            int64_t n_host, const int64_t *__restrict__ d_n,
                                                            const uint32_t *__restrict__ block_offsets) {
    const int64_t n = d_n ? min(*d_n, n_host) : n_host;  // never past the buffer capacity
    const int64_t tile0 = (int64_t)blockIdx.x * kScanTile;
    if (tile0 >= n) return;
    const int64_t base = tile0 + (int64_t)threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = scan_load<POPC>(in, base + k, n);
        s += v[k];
    }
    uint32_t tot;
    uint32_t run = block_excl_scan(s, &tot) + block_offsets[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
}

int64_t scan_ws_bytes(int64_t capacity) {
    const int64_t nblocks = (capacity + kScanTile - 1) / kScanTile;
    return align_up((nblocks + 1) * (int64_t)sizeof(uint32_t), 256);
}

int exclusive_scan_u32(const uint32_t *d_in, uint32_t *d_out, int64_t capacity, const int64_t *d_n, int64_t *d_total,
                       void *d_block_ws, cudaStream_t stream, bool popcount_input) {
    if (capacity <= 0) {
        if (d_total) PE_CUDA(cudaMemsetAsync(d_total, 0, sizeof(int64_t), stream));
        return PE_OK;
    }
    const int64_t nblocks64 = (capacity + kScanTile - 1) / kScanTile;
    PE_CHECK_ARG(nblocks64 < (1ll << 31), "scan: capacity %lld too large", (long long)capacity);
    const int nblocks = (int)nblocks64;
    uint32_t *block_sums = (uint32_t *)d_block_ws;
    if (popcount_input) {
        PE_LAUNCH("scan_tile_sums", stream, scan_tile_sums<true><<<nblocks, kScanThreads, 0, stream>>>(d_in, capacity, d_n, block_sums));
        PE_LAUNCH("scan_block_offsets", stream, scan_block_offsets<<<1, 1024, 0, stream>>>(block_sums, nblocks, d_total));
        PE_LAUNCH("scan_apply", stream, scan_apply<true><<<nblocks, kScanThreads, 0, stream>>>(d_in, d_out, capacity, d_n, block_sums));
    } else {
        PE_LAUNCH("scan_tile_sums", stream, scan_tile_sums<false><<<nblocks, kScanThreads, 0, stream>>>(d_in, capacity, d_n, block_sums));
        PE_LAUNCH("scan_block_offsets", stream, scan_block_offsets<<<1, 1024, 0, stream>>>(block_sums, nblocks, d_total));
        PE_LAUNCH("scan_apply", stream, scan_apply<false><<<nblocks, kScanThreads, 0, stream>>>(d_in, d_out, capacity, d_n, block_sums));
    }
    PE_LAUNCH_CHECK();
    return PE_OK;
}

}  // namespace pe

extern "C" {

int pe_abi_version(void) { return PE_ABI_VERSION; }

const char *pe_last_error(void) { return pe::g_error; }

long long pe_launch_count(void) { return pe::g_launches.load(); }

void pe_profile_enable(int on) {
    pe::prof_fold();
    pe::g_profiling.store(on ? 1 : 0);
}

void pe_profile_reset(void) {
    pe::prof_fold();
    std::lock_guard<std::mutex> lock(pe::g_prof_mutex);
    pe::g_prof_totals.clear();
}

int pe_profile_entries(void) {
    pe::prof_fold();
    std::lock_guard<std::mutex> lock(pe::g_prof_mutex);
    return (int)pe::g_prof_totals.size();
}

int pe_profile_get(int index, const char **name, long long *count, double *total_ms) {
    std::lock_guard<std::mutex> lock(pe::g_prof_mutex);
    if (index < 0 || index >= (int)pe::g_prof_totals.size()) return PE_ERR_ARG;
    if (name) *name = pe::g_prof_totals[index].tag;
    if (count) *count = pe::g_prof_totals[index].count;
    if (total_ms) *total_ms = pe::g_prof_totals[index].ms;
    return PE_OK;
}

int pe_device_info(int32_t *sm_count, int32_t *cc_major, int32_t *cc_minor) {
    int dev = 0, ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        pe::set_error("no CUDA device is visible");
        return PE_ERR_NO_DEVICE;
    }
    PE_CUDA(cudaGetDevice(&dev));
    int sms = 0, major = 0, minor = 0;
    PE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    PE_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    PE_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    if (sm_count) *sm_count = sms;
    if (cc_major) *cc_major = major;
    if (cc_minor) *cc_minor = minor;
    return PE_OK;
}

}  // extern "C"

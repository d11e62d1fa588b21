// sort.cu -- hand-written onesweep LSD radix sort of (key, value) pairs and the chained-scan
// instance-offset pass.  Replaces the three CUB calls of the reference's binning stage
// (/root/reference/.../rasterization/src/forward.cu:104-127,195-202): same results (stable ascending
// sort / exclusive prefix sum, both integer-exact), new implementation:
//   * one histogram kernel for all digit places, then one kernel per digit place that reads and
//     writes the data exactly once; cross-block digit offsets come from a decoupled look-back chain
//     (status word = 2 flag bits + 30-bit count, so no fences are needed);
//   * ranking inside a block is warp-cooperative: __match_any_sync groups equal digits, the lowest
//     lane of each group advances the warp's private digit counter in shared memory;
//   * element counts are read from device memory, so the host never waits for V before sorting.
#include "fgs_device.cuh"
#include "fgs_host.h"

namespace fgs {

namespace {

constexpr uint32_t kFlagLocal = 1u << 30;
constexpr uint32_t kFlagIncl = 2u << 30;
constexpr uint32_t kValueMask = (1u << 30) - 1u;
constexpr int kSortWarps = kSortBlock / 32;
#ifndef FGS_SORT_WINDOW
#define FGS_SORT_WINDOW 8
#endif
constexpr int kLookWindow = FGS_SORT_WINDOW;

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// exclusive scan of one value per thread over a 256-thread block; also returns the block total
__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t v, uint32_t* s_tmp /*>= 9 words*/, uint32_t* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_tmp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < kSortWarps ? s_tmp[lane] : 0u;
        uint32_t wi = w;
#pragma unroll
        for (int o = 1; o < kSortWarps; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        if (lane < kSortWarps) s_tmp[lane] = wi - w;
        if (lane == kSortWarps - 1) s_tmp[kSortWarps] = wi;
    }
    __syncthreads();
    const uint32_t r = s_tmp[warp] + incl - v;
    if (total) *total = s_tmp[kSortWarps];
    __syncthreads();
    return r;
}

struct PassPlan {
    int passes;
    int shift[4];
    int nbits[4];
};

inline PassPlan make_plan(int end_bit) {
    PassPlan p{};
    if (end_bit < 1) end_bit = 1;
    p.passes = (end_bit + 7) / 8;
    const int bpp = (end_bit + p.passes - 1) / p.passes;
    for (int i = 0; i < p.passes; ++i) {
        p.shift[i] = i * bpp;
        p.nbits[i] = end_bit - p.shift[i] < bpp ? end_bit - p.shift[i] : bpp;
    }
    return p;
}

template <typename KeyT>
__global__ void __launch_bounds__(kSortBlock) radix_hist_kernel(const KeyT* __restrict__ keys,
                                                                const uint32_t* __restrict__ n_dev,
                                                                uint32_t* __restrict__ hist, const PassPlan plan) {
    __shared__ uint32_t s_hist[4 * kRadixMax];
    for (int i = threadIdx.x; i < plan.passes * kRadixMax; i += kSortBlock) s_hist[i] = 0;
    __syncthreads();
    const uint32_t n = *n_dev;
    for (uint32_t i = blockIdx.x * kSortBlock + threadIdx.x; i < n; i += gridDim.x * kSortBlock) {
        const uint32_t k = (uint32_t)keys[i];
        for (int p = 0; p < plan.passes; ++p)
            atomicAdd(&s_hist[p * kRadixMax + ((k >> plan.shift[p]) & ((1u << plan.nbits[p]) - 1u))], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < plan.passes * kRadixMax; i += kSortBlock) {
        const uint32_t c = s_hist[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

#ifndef FGS_SORT_MINBLOCKS
#define FGS_SORT_MINBLOCKS 3
#endif
template <typename KeyT>
__global__ void __launch_bounds__(kSortBlock, FGS_SORT_MINBLOCKS) onesweep_kernel(
    const KeyT* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, KeyT* __restrict__ keys_out,
    uint32_t* __restrict__ vals_out, const uint32_t* __restrict__ n_dev, const uint32_t* __restrict__ hist,
    uint32_t* lookback, uint32_t* ticket, const int shift, const int nbits) {
    __shared__ uint32_t s_warp_hist[kSortWarps][kRadixMax];
    __shared__ uint32_t s_keys[kSortTile];
    __shared__ uint32_t s_vals[kSortTile];
    __shared__ uint32_t s_base[kRadixMax];
    __shared__ uint32_t s_lstart[kRadixMax];
    __shared__ uint32_t s_tmp[kSortWarps + 1];
    __shared__ uint32_t s_cnt[kRadixMax];
    __shared__ uint32_t s_bid;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_bid = atomicAdd(ticket, 1u);
    for (int i = tid; i < kSortWarps * kRadixMax; i += kSortBlock) (&s_warp_hist[0][0])[i] = 0;
    s_cnt[tid] = 0;
    __syncthreads();
    const uint32_t bid = s_bid;
    const uint32_t n = *n_dev;
    const uint32_t tile_start = bid * (uint32_t)kSortTile;
    if (tile_start >= n) return;
    const uint32_t tile_count = min((uint32_t)kSortTile, n - tile_start);
    const uint32_t radix = 1u << nbits, dmask = radix - 1u;
    const uint32_t lt_mask = (1u << lane) - 1u;

    // ---- load keys (warp-striped) and rank them inside the warp ------------------------------------
    const uint32_t warp_base = tile_start + (uint32_t)warp * 32u * kSortItems;
    uint32_t key[kSortItems];
    uint32_t rank[kSortItems];
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        const uint32_t idx = warp_base + i * 32u + lane;
        key[i] = idx < n ? (uint32_t)keys_in[idx] : 0xffffffffu;
    }
    // Early publication: a cheap shared-memory histogram of the block's digits goes out before the
    // (much longer) ranking, so that by the time successors look back the chain of partial counts is
    // already complete and only the inclusive wave has to propagate.
#pragma unroll
    for (int i = 0; i < kSortItems; ++i)
        if (warp_base + i * 32u + lane < n) atomicAdd(&s_cnt[(key[i] >> shift) & dmask], 1u);
    __syncthreads();
    if ((uint32_t)tid < radix)
        st_volatile_u32(lookback + (size_t)bid * kRadixMax + tid, s_cnt[tid] | (bid == 0 ? kFlagIncl : kFlagLocal));
    // all MATCH.ANY operations first: they do not depend on the shared-memory counters, so their (long)
    // latencies overlap instead of serialising with the counter updates below
    uint32_t peers[kSortItems];
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        const bool valid = warp_base + i * 32u + lane < n;
        peers[i] = __match_any_sync(0xffffffffu, valid ? (key[i] >> shift) & dmask : 0xffffffffu);
    }
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        const bool valid = warp_base + i * 32u + lane < n;
        const uint32_t d = (key[i] >> shift) & dmask;
        const uint32_t prior = valid ? s_warp_hist[warp][d] : 0u;
        rank[i] = prior + __popc(peers[i] & lt_mask);
        __syncwarp();
        if (valid && lane == __ffs((int)peers[i]) - 1) s_warp_hist[warp][d] = prior + __popc(peers[i]);
        __syncwarp();
    }
    __syncthreads();

    // ---- per digit: offsets of the warps inside the block, block count, look-back -----------------
    const uint32_t d = (uint32_t)tid;
    uint32_t cnt = 0;
    if (d < radix) {
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
            const uint32_t t = s_warp_hist[w][d];
            s_warp_hist[w][d] = cnt;
            cnt += t;
        }
    }
    // Windowed decoupled look-back: kLookWindow predecessors are read with independent loads per
    // step (one L2 latency per window instead of one per predecessor), then consumed in order until
    // the first inclusive entry; an entry that is not published yet restarts the window there.
    uint32_t prefix = 0;
    if (d < radix && bid > 0) {
        int j = (int)bid - 1;
        bool found = false;
        while (!found) {
            uint32_t v[kLookWindow];
#pragma unroll
            for (int w = 0; w < kLookWindow; ++w)
                v[w] = j - w >= 0 ? ld_volatile_u32(lookback + (size_t)(j - w) * kRadixMax + d) : kFlagIncl;
            int used = 0;
#pragma unroll
            for (int w = 0; w < kLookWindow; ++w) {
                if (!found && used == w) {
                    if ((v[w] >> 30) != 0u) {
                        prefix += v[w] & kValueMask;
                        found = (v[w] & kFlagIncl) != 0u;
                        used = w + 1;
                    }
                }
            }
            j -= used;
        }
        st_volatile_u32(lookback + (size_t)bid * kRadixMax + d, ((prefix + cnt) & kValueMask) | kFlagIncl);
    }
    const uint32_t g = d < radix ? hist[d] : 0u;
    const uint32_t gex = block_excl_scan_256(g, s_tmp, nullptr);
    const uint32_t lex = block_excl_scan_256(cnt, s_tmp, nullptr);
    s_base[tid] = gex + prefix - lex;
    s_lstart[tid] = lex;
    __syncthreads();

    // ---- scatter to shared memory in block-sorted order, then stream out coalesced runs ------------
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        const uint32_t idx = warp_base + i * 32u + lane;
        if (idx < n) {
            const uint32_t dd = (key[i] >> shift) & dmask;
            const uint32_t lp = s_lstart[dd] + s_warp_hist[warp][dd] + rank[i];
            s_keys[lp] = key[i];
            s_vals[lp] = vals_in[idx];
        }
    }
    __syncthreads();
    for (uint32_t lp = tid; lp < tile_count; lp += kSortBlock) {
        const uint32_t k = s_keys[lp];
        const uint32_t pos = s_base[(k >> shift) & dmask] + lp;
        if (keys_out) keys_out[pos] = (KeyT)k;
        vals_out[pos] = s_vals[lp];
    }
}

// ---- chained-scan exclusive prefix sum of n_touched in depth order -------------------------------
constexpr int kScanItems = 8;
constexpr int kScanTile = kSortBlock * kScanItems;

__global__ void __launch_bounds__(kSortBlock) offsets_scan_kernel(const uint32_t* __restrict__ sorted_idx,
                                                                  const uint32_t* __restrict__ n_touched,
                                                                  uint32_t* __restrict__ offsets,
                                                                  const uint32_t* __restrict__ n_dev,
                                                                  uint32_t* lookback, uint32_t* ticket) {
    __shared__ uint32_t s_tmp[kSortWarps + 1];
    __shared__ uint32_t s_bid;
    __shared__ uint32_t s_prefix;
    const int tid = threadIdx.x;
    if (tid == 0) s_bid = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t bid = s_bid;
    const uint32_t n = *n_dev;
    const uint32_t tile_start = bid * (uint32_t)kScanTile;
    if (tile_start >= n) return;
    uint32_t v[kScanItems];
    uint32_t sum = 0;
    const uint32_t first = tile_start + (uint32_t)tid * kScanItems;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const uint32_t i = first + k;
        v[k] = i < n ? __ldg(n_touched + __ldg(sorted_idx + i)) : 0u;
        sum += v[k];
    }
    uint32_t total;
    uint32_t excl = block_excl_scan_256(sum, s_tmp, &total);
    if (tid == 0) {
        uint32_t prefix = 0;
        if (bid == 0) {
            st_volatile_u32(lookback, total | kFlagIncl);
        } else {
            st_volatile_u32(lookback + bid, total | kFlagLocal);
            for (int j = (int)bid - 1; j >= 0; --j) {
                uint32_t w;
                do {
                    w = ld_volatile_u32(lookback + j);
                } while ((w >> 30) == 0u);
                prefix += w & kValueMask;
                if (w & kFlagIncl) break;
            }
            st_volatile_u32(lookback + bid, ((prefix + total) & kValueMask) | kFlagIncl);
        }
        s_prefix = prefix;
    }
    __syncthreads();
    excl += s_prefix;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const uint32_t i = first + k;
        if (i < n) offsets[i] = excl;
        excl += v[k];
    }
}

}  // namespace

int radix_sort_passes(int end_bit) { return make_plan(end_bit).passes; }

size_t radix_sort_lookback_words(int n_max, int end_bit) {
    return (size_t)make_plan(end_bit).passes * (size_t)div_up(n_max > 0 ? n_max : 1, kSortTile) * kRadixMax;
}

template <typename KeyT>
cudaError_t radix_sort_pairs(KeyT* keys, KeyT* keys_alt, uint32_t* vals, uint32_t* vals_alt, const uint32_t* n_dev,
                             int n_max, int end_bit, uint32_t* hist, uint32_t* lookback, uint32_t* tickets,
                             bool* result_in_alt, bool zero_scratch, cudaStream_t stream) {
    *result_in_alt = false;
    if (n_max <= 0) return cudaSuccess;
    const PassPlan plan = make_plan(end_bit);
    const int nblocks = div_up(n_max, kSortTile);
    if (zero_scratch) {
        cudaError_t e;
        if (lookback == hist + 4 * kRadixMax) {  // adjacent: one memset
            e = cudaMemsetAsync(hist, 0, sizeof(uint32_t) * (4 * kRadixMax + radix_sort_lookback_words(n_max, end_bit)), stream);
        } else {
            e = cudaMemsetAsync(hist, 0, sizeof(uint32_t) * plan.passes * kRadixMax, stream);
            if (e != cudaSuccess) return e;
            e = cudaMemsetAsync(lookback, 0, sizeof(uint32_t) * radix_sort_lookback_words(n_max, end_bit), stream);
        }
        if (e != cudaSuccess) return e;
    }
    const int hist_grid = nblocks < 148 * 8 ? nblocks : 148 * 8;
    radix_hist_kernel<KeyT><<<hist_grid, kSortBlock, 0, stream>>>(keys, n_dev, hist, plan);
    KeyT* kin = keys;
    KeyT* kout = keys_alt;
    uint32_t* vin = vals;
    uint32_t* vout = vals_alt;
    for (int p = 0; p < plan.passes; ++p) {
        onesweep_kernel<KeyT><<<nblocks, kSortBlock, 0, stream>>>(
            kin, vin, kout, vout, n_dev, hist + p * kRadixMax, lookback + (size_t)p * nblocks * kRadixMax,
            tickets + p, plan.shift[p], plan.nbits[p]);
        KeyT* tk = kin;
        kin = kout;
        kout = tk;
        uint32_t* tv = vin;
        vin = vout;
        vout = tv;
    }
    *result_in_alt = (plan.passes & 1) != 0;
    return cudaGetLastError();
}

template cudaError_t radix_sort_pairs<uint16_t>(uint16_t*, uint16_t*, uint32_t*, uint32_t*, const uint32_t*, int, int,
                                                uint32_t*, uint32_t*, uint32_t*, bool*, bool, cudaStream_t);
template cudaError_t radix_sort_pairs<uint32_t>(uint32_t*, uint32_t*, uint32_t*, uint32_t*, const uint32_t*, int, int,
                                                uint32_t*, uint32_t*, uint32_t*, bool*, bool, cudaStream_t);

size_t offsets_scan_lookback_words(int n_max) { return (size_t)div_up(n_max > 0 ? n_max : 1, kScanTile); }

cudaError_t launch_offsets_scan(const uint32_t* sorted_idx, const uint32_t* n_touched, uint32_t* offsets,
                                const uint32_t* n_dev, int n_max, uint32_t* lookback, uint32_t* ticket,
                                bool zero_scratch, cudaStream_t stream) {
    if (n_max <= 0) return cudaSuccess;
    if (zero_scratch) {
        cudaError_t e = cudaMemsetAsync(lookback, 0, sizeof(uint32_t) * offsets_scan_lookback_words(n_max), stream);
        if (e != cudaSuccess) return e;
    }
    offsets_scan_kernel<<<div_up(n_max, kScanTile), kSortBlock, 0, stream>>>(sorted_idx, n_touched, offsets, n_dev,
                                                                           lookback, ticket);
    return cudaGetLastError();
}

}  // namespace fgs

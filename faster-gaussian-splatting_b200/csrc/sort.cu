// sort.cu -- hand-written onesweep LSD radix sort of (key, value) pairs and the chained-scan
// instance-offset pass.  Replaces the three CUB calls of the reference's binning stage
// (/root/reference/.../rasterization/src/forward.cu:104-127,195-202): same results (stable ascending
// sort / exclusive prefix sum, both integer-exact), new implementation:
//   * one kernel per digit place.  The grid is at most one wave of co-scheduled blocks; a block owns a
//     CONTIGUOUS chunk of 2048-key tiles, histograms the whole chunk first (16-byte key loads), publishes
//     that histogram and resolves its cross-block digit offsets with ONE decoupled look-back (status
//     word = 2 flag bits + 30-bit count, so no fences are needed), then ranks and scatters its tiles one
//     after the other from a running per-digit base.  Measured on B200: with one look-back per 2048-key
//     tile the pass time was the length of the look-back chain (tiles / 16 L2 round trips: 44 us for
//     2.15 M pairs); per chunk the chain has <= 148 x occupancy links;
//   * the global digit histogram of pass p+1 is accumulated by pass p while it histograms its chunk (the
//     keys are in registers anyway); pass 0's comes from the kernel that produced the keys
//     (create_instances) or from a one-digit histogram kernel;
//   * ranking inside a tile is warp-cooperative: equal digits are grouped warp-wide, the lowest lane of
//     each group advances the warp's private digit counter in shared memory;
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
#define FGS_SORT_WINDOW 4
#endif
constexpr int kLookWindow = FGS_SORT_WINDOW;

#ifdef FGS_CUDA_EMU  // test-only CPU emulator (tests/cuda_emu): blocks run one after the other
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
#else
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
#endif

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

// global histogram of ONE digit place (pass 0 when the producer of the keys did not supply it)
template <typename KeyT>
__global__ void __launch_bounds__(kSortBlock) radix_hist_kernel(const KeyT* __restrict__ keys,
                                                                const uint32_t* __restrict__ n_dev,
                                                                uint32_t* __restrict__ hist, const int shift,
                                                                const int nbits, const uint32_t n_cap) {
    __shared__ uint32_t s_hist[kRadixMax];
    s_hist[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t n = bounded_count(n_dev, n_cap);
    const uint32_t dmask = (1u << nbits) - 1u;
    for (uint32_t i = blockIdx.x * kSortBlock + threadIdx.x; i < n; i += gridDim.x * kSortBlock)
        atomicAdd(&s_hist[((uint32_t)keys[i] >> shift) & dmask], 1u);
    __syncthreads();
    const uint32_t c = s_hist[threadIdx.x];
    if (c) atomicAdd(&hist[threadIdx.x], c);
}

#ifndef FGS_SORT_MINBLOCKS
#define FGS_SORT_MINBLOCKS 3
#endif
#ifndef FGS_SORT_WAVE_CAP
#define FGS_SORT_WAVE_CAP 2
#endif

// groups the lanes of a warp by digit: bit l of the result is set iff lane l holds the same digit
// (invalid lanes never match a valid one).  One ballot per digit bit; the MATCH.ANY instruction does the
// same in one go but measured far slower on B200 (ncu: ~40 % of a pass stalled waiting for its result).
__device__ __forceinline__ uint32_t match_digit(const uint32_t d, const bool valid, const int nbits) {
#ifdef FGS_SORT_MATCH_ANY
    return __match_any_sync(0xffffffffu, valid ? d : 0xffffffffu);
#else
    uint32_t m = __ballot_sync(0xffffffffu, valid);
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        if (b < nbits) {  // warp-uniform
            const bool bit = ((d >> b) & 1u) != 0u;
            const uint32_t bal = __ballot_sync(0xffffffffu, bit);
            m &= bit ? bal : ~bal;
        }
    }
    return m;
#endif
}

// One digit place of the LSD sort.  Block `bid` (ticket order, so every predecessor has started) owns
// tiles [bid * tpb, (bid + 1) * tpb) with tpb = ceil(tiles / gridDim.x).
template <typename KeyT>
__global__ void __launch_bounds__(kSortBlock, FGS_SORT_MINBLOCKS) onesweep_kernel(
    const KeyT* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, KeyT* __restrict__ keys_out,
    uint32_t* __restrict__ vals_out, const uint32_t* __restrict__ n_dev, const uint32_t* __restrict__ hist,
    uint32_t* __restrict__ hist_next, uint32_t* lookback, uint32_t* ticket, const int shift, const int nbits,
    const int shift_next, const int nbits_next, const int hist_replicas, const uint32_t n_cap) {
    static_assert(kSortBlock == kRadixMax, "one thread per digit");
    __shared__ uint32_t s_warp_hist[kSortWarps][kRadixMax];
    __shared__ uint32_t s_keys[kSortTile];
    __shared__ uint32_t s_vals[kSortTile];
    __shared__ uint32_t s_gpos[kRadixMax];    // per tile and digit: global position of the digit's run minus its local start
    __shared__ uint32_t s_lstart[kRadixMax];  // per tile and digit: local start of the digit's run
    __shared__ uint32_t s_tmp[kSortWarps + 1];
    __shared__ uint32_t s_cnt[kRadixMax];
    __shared__ uint32_t s_cnt_next[kRadixMax];
    __shared__ uint32_t s_bid;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_bid = atomicAdd(ticket, 1u);
    const uint32_t n = bounded_count(n_dev, n_cap);  // n_cap: the host-side bound the buffers were sized for
    // this pass's global digit counts (complete: written by earlier kernels); `hist_replicas` > 1: the
    // producer spread its atomics over that many copies (preprocess.cu)
    uint32_t g = 0;
    if (tid < (1 << nbits))
        for (int r = 0; r < hist_replicas; ++r) g += hist[r * kRadixMax + tid];
    for (int i = tid; i < kSortWarps * kRadixMax; i += kSortBlock) (&s_warp_hist[0][0])[i] = 0;
    s_cnt[tid] = 0;
    s_cnt_next[tid] = 0;
    __syncthreads();
    const uint32_t bid = s_bid;
    const uint32_t total_tiles = (n + (uint32_t)kSortTile - 1u) / (uint32_t)kSortTile;
    const uint32_t tpb = (total_tiles + gridDim.x - 1u) / gridDim.x;
    const uint32_t first_tile = bid * tpb;
    if (first_tile >= total_tiles) return;  // blocks behind the data: nobody looks back at them
    const uint32_t my_tiles = min(tpb, total_tiles - first_tile);
    const uint32_t chunk_start = first_tile * (uint32_t)kSortTile;
    const uint32_t chunk_end = min(n, chunk_start + my_tiles * (uint32_t)kSortTile);
    const uint32_t radix = 1u << nbits, dmask = radix - 1u;
    const bool has_next = hist_next != nullptr;
    const uint32_t dmask_next = (1u << nbits_next) - 1u;
    const uint32_t lt_mask = (1u << lane) - 1u;

    // ---- phase 1: digit histogram of the whole chunk (this pass and the next one) ---------------------
    auto count_key = [&](const uint32_t k) {
        atomicAdd(&s_cnt[(k >> shift) & dmask], 1u);
        if (has_next) atomicAdd(&s_cnt_next[(k >> shift_next) & dmask_next], 1u);
    };
    // the first tile goes to registers (warp-striped) and stays there for phase 3
    uint32_t key[kSortItems];
    uint32_t val[kSortItems];
    {
        const uint32_t warp_base = chunk_start + (uint32_t)warp * 32u * kSortItems;
#pragma unroll
        for (int i = 0; i < kSortItems; ++i) {
            const uint32_t idx = warp_base + i * 32u + lane;
            key[i] = idx < n ? (uint32_t)keys_in[idx] : 0xffffffffu;
        }
#pragma unroll
        for (int i = 0; i < kSortItems; ++i) {
            const uint32_t idx = warp_base + i * 32u + lane;
            val[i] = idx < n ? vals_in[idx] : 0u;
        }
#pragma unroll
        for (int i = 0; i < kSortItems; ++i)
            if (warp_base + i * 32u + lane < n) count_key(key[i]);
    }
    if (my_tiles > 1) {  // the other tiles of the chunk: 16-byte loads (tile starts are 16-byte aligned)
        constexpr uint32_t kVec = 16u / (uint32_t)sizeof(KeyT);
        const uint32_t rest_start = chunk_start + (uint32_t)kSortTile;
        const uint32_t nvec = (chunk_end - rest_start) / kVec;
        const uint4* src = reinterpret_cast<const uint4*>(keys_in + rest_start);
        for (uint32_t v = tid; v < nvec; v += kSortBlock) {
            const uint4 q = __ldg(src + v);
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (sizeof(KeyT) == 2) {
                    count_key(w[j] & 0xffffu);
                    count_key(w[j] >> 16);
                } else {
                    count_key(w[j]);
                }
            }
        }
        for (uint32_t i = rest_start + nvec * kVec + tid; i < chunk_end; i += kSortBlock) count_key((uint32_t)keys_in[i]);
    }
    __syncthreads();

    // ---- phase 2: publish, look back once, derive the chunk's global digit bases ------------------------
    const uint32_t d = (uint32_t)tid;
    const uint32_t chunk_cnt = s_cnt[tid];
    if (d < radix)
        st_volatile_u32(lookback + (size_t)bid * kRadixMax + d, chunk_cnt | (bid == 0 ? kFlagIncl : kFlagLocal));
    if (has_next) {
        const uint32_t c = s_cnt_next[tid];
        if (c) atomicAdd(hist_next + tid, c);
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
        st_volatile_u32(lookback + (size_t)bid * kRadixMax + d, ((prefix + chunk_cnt) & kValueMask) | kFlagIncl);
    }
    // global position of the chunk's first key with digit d; advanced tile by tile (thread d owns digit d)
    uint32_t base = block_excl_scan_256(g, s_tmp, nullptr) + prefix;

    // ---- phase 3: rank and scatter the chunk's tiles one after the other --------------------------------
    for (uint32_t t = 0; t < my_tiles; ++t) {
        const uint32_t tile_start = chunk_start + t * (uint32_t)kSortTile;
        const uint32_t tile_count = min((uint32_t)kSortTile, n - tile_start);
        const uint32_t warp_base = tile_start + (uint32_t)warp * 32u * kSortItems;
        if (t > 0) {
            for (int i = tid; i < kSortWarps * kRadixMax; i += kSortBlock) (&s_warp_hist[0][0])[i] = 0;
#pragma unroll
            for (int i = 0; i < kSortItems; ++i) {
                const uint32_t idx = warp_base + i * 32u + lane;
                key[i] = idx < n ? (uint32_t)keys_in[idx] : 0xffffffffu;
            }
#pragma unroll
            for (int i = 0; i < kSortItems; ++i) {
                const uint32_t idx = warp_base + i * 32u + lane;
                val[i] = idx < n ? vals_in[idx] : 0u;
            }
            __syncthreads();
        }
        // all matches first: they do not depend on the shared-memory counters, so their (long)
        // latencies overlap instead of serialising with the counter updates below
        uint32_t peers[kSortItems];
        uint32_t rank[kSortItems];
#pragma unroll
        for (int i = 0; i < kSortItems; ++i)
            peers[i] = match_digit((key[i] >> shift) & dmask, warp_base + i * 32u + lane < n, nbits);
        // the lowest lane of every digit group advances the warp's private counter with one shared-memory
        // atomic (same-warp shared accesses complete in program order, so item i+1 sees item i); the old
        // value reaches the group by shuffle: the items' atomics are independent and pipeline
#pragma unroll
        for (int i = 0; i < kSortItems; ++i) {
            const bool valid = warp_base + i * 32u + lane < n;
            rank[i] = 0;
            if (valid && lane == __ffs((int)peers[i]) - 1)
                rank[i] = atomicAdd(&s_warp_hist[warp][(key[i] >> shift) & dmask], (uint32_t)__popc(peers[i]));
            __syncwarp();
        }
#pragma unroll
        for (int i = 0; i < kSortItems; ++i) {
            const int leader = (__ffs((int)peers[i]) - 1) & 31;
            rank[i] = __shfl_sync(0xffffffffu, rank[i], leader) + __popc(peers[i] & lt_mask);
        }
        __syncthreads();

        // per digit: offsets of the warps inside the tile, tile count, local start, global position
        uint32_t cnt = 0;
        if (d < radix) {
#pragma unroll
            for (int w = 0; w < kSortWarps; ++w) {
                const uint32_t c = s_warp_hist[w][d];
                s_warp_hist[w][d] = cnt;
                cnt += c;
            }
        }
        const uint32_t lex = block_excl_scan_256(cnt, s_tmp, nullptr);
        s_lstart[tid] = lex;
        s_gpos[tid] = base - lex;
        base += cnt;
        __syncthreads();

        // scatter to shared memory in tile-sorted order, then stream out coalesced runs
#pragma unroll
        for (int i = 0; i < kSortItems; ++i) {
            if (warp_base + i * 32u + lane < n) {
                const uint32_t dd = (key[i] >> shift) & dmask;
                const uint32_t lp = s_lstart[dd] + s_warp_hist[warp][dd] + rank[i];
                s_keys[lp] = key[i];
                s_vals[lp] = val[i];
            }
        }
        __syncthreads();
        for (uint32_t lp = tid; lp < tile_count; lp += kSortBlock) {
            const uint32_t k = s_keys[lp];
            const uint32_t pos = s_gpos[(k >> shift) & dmask] + lp;
            if (keys_out) keys_out[pos] = (KeyT)k;
            vals_out[pos] = s_vals[lp];
        }
        // the next tile's first __syncthreads orders these reads before its writes to s_keys / s_gpos
    }
}

// ---- chained-scan exclusive prefix sum of n_touched in depth order -------------------------------
// 4096 elements per block (16 consecutive per thread, the sorted indices arrive with 16-byte loads);
// warp 0 resolves the block's prefix with a warp-wide decoupled look-back: 32 predecessors per L2 round trip.
#ifndef FGS_SCAN_ITEMS
#define FGS_SCAN_ITEMS 16
#endif
constexpr int kScanItems = FGS_SCAN_ITEMS;
constexpr int kScanTile = kSortBlock * kScanItems;
static_assert(kScanItems % 4 == 0, "16-byte index loads");

__global__ void __launch_bounds__(kSortBlock) offsets_scan_kernel(const uint32_t* __restrict__ sorted_idx,
                                                                  const uint2* __restrict__ tile_info,
                                                                  uint32_t* __restrict__ offsets,
                                                                  uint2* __restrict__ sorted_info,
                                                                  const uint32_t* __restrict__ n_dev,
                                                                  uint32_t* lookback, uint32_t* ticket) {
    __shared__ uint32_t s_tmp[kSortWarps + 1];
    __shared__ uint32_t s_bid;
    __shared__ uint32_t s_prefix;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_bid = atomicAdd(ticket, 1u);
    const uint32_t n = *n_dev;
    __syncthreads();
    const uint32_t bid = s_bid;
    const uint32_t tile_start = bid * (uint32_t)kScanTile;
    if (tile_start >= n) return;
    uint32_t v[kScanItems];
    uint32_t sum = 0;
    const uint32_t first = tile_start + (uint32_t)tid * kScanItems;
    const bool full = first + kScanItems <= n;
    if (full) {
        const uint4* src = reinterpret_cast<const uint4*>(sorted_idx + first);
#pragma unroll
        for (int k4 = 0; k4 < kScanItems / 4; ++k4) {
            const uint4 q = __ldg(src + k4);
            v[4 * k4 + 0] = q.x;
            v[4 * k4 + 1] = q.y;
            v[4 * k4 + 2] = q.z;
            v[4 * k4 + 3] = q.w;
        }
        uint4* dst = reinterpret_cast<uint4*>(sorted_info + first);
#pragma unroll
        for (int k2 = 0; k2 < kScanItems / 2; ++k2) {
            const uint2 a = __ldg(tile_info + v[2 * k2]);
            const uint2 b = __ldg(tile_info + v[2 * k2 + 1]);
            dst[k2] = make_uint4(a.x, a.y, b.x, b.y);
            v[2 * k2] = tile_info_count(a);
            v[2 * k2 + 1] = tile_info_count(b);
        }
    } else {
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            const uint32_t i = first + k;
            v[k] = 0u;
            if (i < n) {
                const uint2 a = __ldg(tile_info + __ldg(sorted_idx + i));
                sorted_info[i] = a;
                v[k] = tile_info_count(a);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) sum += v[k];
    uint32_t total;
    uint32_t excl = block_excl_scan_256(sum, s_tmp, &total);
    if (warp == 0) {
        uint32_t prefix = 0;
        if (lane == 0) st_volatile_u32(lookback + bid, total | (bid == 0 ? kFlagIncl : kFlagLocal));
        int j = (int)bid - 1;  // newest predecessor not consumed yet
        while (j >= 0) {       // warp-uniform
            const int idx = j - lane;
            const uint32_t w = idx >= 0 ? ld_volatile_u32(lookback + idx) : kFlagIncl;  // before block 0: inclusive 0
            const uint32_t not_ready = ~__ballot_sync(0xffffffffu, (w >> 30) != 0u);
            const uint32_t incl = __ballot_sync(0xffffffffu, (w & kFlagIncl) != 0u);
            const int n_ready = not_ready ? __ffs((int)not_ready) - 1 : 32;  // leading published entries
            const uint32_t ready_mask = n_ready == 32 ? 0xffffffffu : (1u << n_ready) - 1u;
            const uint32_t incl_in = incl & ready_mask;
            const int take = incl_in ? __ffs((int)incl_in) : n_ready;  // entries consumed, the inclusive one included
            prefix += __reduce_add_sync(0xffffffffu, lane < take ? (w & kValueMask) : 0u);
            if (incl_in) break;
            j -= take;
        }
        if (lane == 0) {
            if (bid > 0) st_volatile_u32(lookback + bid, ((prefix + total) & kValueMask) | kFlagIncl);
            s_prefix = prefix;
        }
    }
    __syncthreads();
    excl += s_prefix;
    if (full) {
        uint4* dst = reinterpret_cast<uint4*>(offsets + first);
#pragma unroll
        for (int k4 = 0; k4 < kScanItems / 4; ++k4) {
            uint4 q;
            q.x = excl;
            q.y = q.x + v[4 * k4 + 0];
            q.z = q.y + v[4 * k4 + 1];
            q.w = q.z + v[4 * k4 + 2];
            excl = q.w + v[4 * k4 + 3];
            dst[k4] = q;
        }
    } else {
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            const uint32_t i = first + k;
            if (i < n) offsets[i] = excl;
            excl += v[k];
        }
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
                             bool* result_in_alt, bool zero_scratch, const uint32_t* pass0_hist, int pass0_hist_replicas,
                             cudaStream_t stream) {
    *result_in_alt = false;
    if (n_max <= 0) return cudaSuccess;
    const PassPlan plan = make_plan(end_bit);
    const int tiles = div_up(n_max, kSortTile);
    static PerDevice<int> s_wave_dev;  // blocks of onesweep_kernel<KeyT> one wave holds
    int& s_wave = s_wave_dev.here();
    if (s_wave == 0) {
        int dev = 0, sms = 0, occ = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, onesweep_kernel<KeyT>, kSortBlock, 0);
        if (occ < 1) occ = 1;
        // two resident blocks per SM measured best on B200 (A/B in profiles/r1b_results.md): fewer links in the
        // look-back chain at the cost of more tiles per block
        if (occ > FGS_SORT_WAVE_CAP) occ = FGS_SORT_WAVE_CAP;
        s_wave = sms * occ;
    }
    const int grid = tiles < s_wave ? tiles : s_wave;
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
    if (!pass0_hist) {
        const int hist_grid = tiles < 148 * 4 ? tiles : 148 * 4;
        FGS_LAUNCH(radix_hist_kernel<KeyT>, hist_grid, kSortBlock, stream, keys, n_dev, hist, plan.shift[0], plan.nbits[0],
                                                                      (uint32_t)n_max);
    }
    KeyT* kin = keys;
    KeyT* kout = keys_alt;
    uint32_t* vin = vals;
    uint32_t* vout = vals_alt;
    for (int p = 0; p < plan.passes; ++p) {
        const bool last = p + 1 == plan.passes;
        FGS_LAUNCH(onesweep_kernel<KeyT>, grid, kSortBlock, stream,
            kin, vin, kout, vout, n_dev, p == 0 && pass0_hist ? pass0_hist : hist + p * kRadixMax,
            last ? nullptr : hist + (p + 1) * kRadixMax, lookback + (size_t)p * tiles * kRadixMax, tickets + p,
            plan.shift[p], plan.nbits[p], last ? 0 : plan.shift[p + 1], last ? 1 : plan.nbits[p + 1],
            p == 0 && pass0_hist ? pass0_hist_replicas : 1, (uint32_t)n_max);
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

int radix_sort_pass0_bits(int end_bit) { return make_plan(end_bit).nbits[0]; }

template cudaError_t radix_sort_pairs<uint16_t>(uint16_t*, uint16_t*, uint32_t*, uint32_t*, const uint32_t*, int, int,
                                                uint32_t*, uint32_t*, uint32_t*, bool*, bool, const uint32_t*, int, cudaStream_t);
template cudaError_t radix_sort_pairs<uint32_t>(uint32_t*, uint32_t*, uint32_t*, uint32_t*, const uint32_t*, int, int,
                                                uint32_t*, uint32_t*, uint32_t*, bool*, bool, const uint32_t*, int, cudaStream_t);

size_t offsets_scan_lookback_words(int n_max) { return (size_t)div_up(n_max > 0 ? n_max : 1, kScanTile); }

cudaError_t launch_offsets_scan(const uint32_t* sorted_idx, const uint2* tile_info, uint32_t* offsets, uint2* sorted_info,
                                const uint32_t* n_dev, int n_max, uint32_t* lookback, uint32_t* ticket,
                                bool zero_scratch, cudaStream_t stream) {
    if (n_max <= 0) return cudaSuccess;
    if (zero_scratch) {
        cudaError_t e = cudaMemsetAsync(lookback, 0, sizeof(uint32_t) * offsets_scan_lookback_words(n_max), stream);
        if (e != cudaSuccess) return e;
    }
    FGS_LAUNCH(offsets_scan_kernel, div_up(n_max, kScanTile), kSortBlock, stream, sorted_idx, tile_info, offsets, sorted_info,
                                                                           n_dev, lookback, ticket);
    return cudaGetLastError();
}

}  // namespace fgs

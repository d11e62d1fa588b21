// cuda_emu.h -- TEST INFRASTRUCTURE ONLY.  A thread-per-CUDA-thread emulator that lets g++ compile and run the
// kernel SOURCE of csrc/*.cu on host memory, so that indexing, barrier placement, warp collectives and the
// C-ABI launch logic of a kernel can be checked in a container without a GPU (and under -fsanitize=thread,
// which turns a missing __syncthreads() into a reported data race).  It is never part of the product: nothing
// under faster-gaussian-splatting_b200/ includes it, and the emulated libraries are built into tests/cuda_emu/_build/.
//
// Model: blocks run one after another; every CUDA thread of the running block is a std::thread.
//   __shared__      -> static (one block at a time, so one copy is the block's copy)
//   __syncthreads() -> std::barrier over the live threads of the block (a thread that returns drops out, as on
//                      hardware since Volta)
//   warp collectives (__shfl_*_sync, __ballot_sync, __any_sync, __all_sync, __syncwarp) -> per-warp barrier +
//                      exchange slots; every live lane of the warp must take part (full-mask use only)
//   atomics         -> GCC __atomic builtins / CAS loops
// Not modelled: memory-model subtleties between blocks (blocks are sequential), fast-math rounding, PTX.
#pragma once
#ifndef FGS_CUDA_EMU
#define FGS_CUDA_EMU 1
#endif

#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __noinline__ __attribute__((noinline))
#define __maxnreg__(n)
#define __shared__ static
#define __constant__ static

struct uint3 {
    unsigned x, y, z;
};
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(8) float2 {
    float x, y;
};
struct float3 {
    float x, y, z;
};
struct alignas(16) float4 {
    float x, y, z, w;
};
struct alignas(8) uint2 {
    unsigned x, y;
};
struct alignas(16) uint4 {
    unsigned x, y, z, w;
};
struct alignas(8) int2 {
    int x, y;
};
struct alignas(4) uchar4 {
    unsigned char x, y, z, w;
};
struct alignas(8) ushort4 {
    unsigned short x, y, z, w;
};
struct alignas(4) ushort2 {
    unsigned short x, y;
};
struct alignas(16) int4 {
    int x, y, z, w;
};
struct uint3_unused;
inline ushort4 make_ushort4(unsigned short x, unsigned short y, unsigned short z, unsigned short w) { return {x, y, z, w}; }
inline int4 make_int4(int x, int y, int z, int w) { return {x, y, z, w}; }
inline float2 make_float2(float x, float y) { return {x, y}; }
inline float3 make_float3(float x, float y, float z) { return {x, y, z}; }
inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
inline uint2 make_uint2(unsigned x, unsigned y) { return {x, y}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return {x, y, z, w}; }
inline int2 make_int2(int x, int y) { return {x, y}; }

typedef void* cudaStream_t;
typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0, cudaErrorInvalidValue = 1;
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) {
    std::memset(p, v, n);
    return cudaSuccess;
}
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) {
    std::memcpy(d, s, n);
    return cudaSuccess;
}
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
typedef struct cuda_emu_event* cudaEvent_t;
struct cuda_emu_event {
    int dummy;
};
constexpr unsigned cudaEventDisableTiming = 2, cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0;
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new cuda_emu_event{0}; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.0f; return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = 0; return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { *p = std::calloc(1, n); return cudaSuccess; }
template <typename T>
inline cudaError_t cudaMemcpyToSymbolAsync(T& symbol, const void* src, size_t n, size_t offset, cudaMemcpyKind, cudaStream_t) {
    std::memcpy(reinterpret_cast<char*>(&symbol) + offset, src, n);
    return cudaSuccess;
}
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize, cudaFuncAttributePreferredSharedMemoryCarveout };
template <typename K>
inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return cudaSuccess; }
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount };
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 3; return cudaSuccess; }  // a 3-SM "GPU"
template <typename K>
inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, K, int, size_t) { *n = 2; return cudaSuccess; }

namespace cuda_emu {

using Barrier = std::barrier<>;  // (a spin-then-yield barrier was tried: slower with 256 host threads on 8 cores)

struct WarpCtx {
    Barrier bar;
    uint64_t slot[32];
    std::atomic<unsigned> live;  // lanes that have not returned yet
    explicit WarpCtx(int lanes) : bar(lanes), live(lanes >= 32 ? 0xffffffffu : ((1u << lanes) - 1u)) {
        std::memset(slot, 0, sizeof(slot));
    }
};
struct BlockCtx {
    Barrier bar;
    std::vector<std::unique_ptr<WarpCtx>> warps;
    explicit BlockCtx(int threads) : bar(threads) {
        for (int w = 0; w * 32 < threads; ++w) warps.emplace_back(new WarpCtx(std::min(32, threads - w * 32)));
    }
};

inline thread_local uint3 t_threadIdx, t_blockIdx;
inline thread_local dim3 t_blockDim, t_gridDim;
inline thread_local BlockCtx* t_block = nullptr;
inline thread_local WarpCtx* t_warp = nullptr;
inline thread_local int t_lane = 0;
inline thread_local unsigned char* t_dyn_smem = nullptr;  // the running block's dynamic shared memory

// Order in which the blocks of a grid run.  Hardware promises none, so a kernel must not depend on it (kernels with
// inter-block dependencies take their logical block id from an atomic ticket).  FGS_EMU_BLOCK_ORDER=reverse|shuffle
// makes the emulator run the blocks in another order than 0, 1, 2, ... to check exactly that.
inline std::vector<unsigned> block_order(unsigned n) {
    std::vector<unsigned> order(n);
    for (unsigned i = 0; i < n; ++i) order[i] = i;
    const char* mode = std::getenv("FGS_EMU_BLOCK_ORDER");
    if (mode && std::strcmp(mode, "reverse") == 0) std::reverse(order.begin(), order.end());
    if (mode && std::strcmp(mode, "shuffle") == 0) {
        static uint64_t state = 0x9e3779b97f4a7c15ull;
        for (unsigned i = n; i > 1; --i) {
            state = state * 6364136223846793005ull + 1442695040888963407ull;
            std::swap(order[i - 1], order[(state >> 33) % i]);
        }
    }
    return order;
}

template <typename F>
void launch(dim3 grid, dim3 block, F&& body, size_t dyn_smem_bytes = 0) {
    const int threads = (int)(block.x * block.y * block.z);
    std::vector<unsigned char> dyn(dyn_smem_bytes + 256);
    unsigned char* dyn_base = dyn.data() + (256 - reinterpret_cast<uintptr_t>(dyn.data()) % 256) % 256;
    for (unsigned linear : block_order(grid.x * grid.y * grid.z)) {
        const unsigned bx = linear % grid.x, by = (linear / grid.x) % grid.y, bz = linear / (grid.x * grid.y);
        BlockCtx ctx(threads);
        std::vector<std::thread> pool;
        pool.reserve(threads);
        for (int t = 0; t < threads; ++t)
            pool.emplace_back([&, t] {
                t_threadIdx = {(unsigned)t % block.x, ((unsigned)t / block.x) % block.y, (unsigned)t / (block.x * block.y)};
                t_blockIdx = {bx, by, bz};
                t_blockDim = block;
                t_gridDim = grid;
                t_block = &ctx;
                t_warp = ctx.warps[t / 32].get();
                t_lane = t % 32;
                t_dyn_smem = dyn_base;
                body();
                t_warp->live.fetch_and(~(1u << t_lane));
                t_warp->bar.arrive_and_drop();  // a finished thread no longer takes part in barriers
                ctx.bar.arrive_and_drop();
            });
        for (auto& th : pool) th.join();
    }
}

// Kernels without barriers or warp collectives (declared so at the launch site with FGS_LAUNCH_SERIAL): every CUDA
// thread runs to completion on the calling host thread, one after the other.  A barrier inside such a kernel
// dereferences the null block context: the declaration is checked, not trusted.
template <typename F>
void launch_serial(dim3 grid, dim3 block, F&& body) {
    t_block = nullptr;
    t_warp = nullptr;
    t_blockDim = block;
    t_gridDim = grid;
    for (unsigned linear : block_order(grid.x * grid.y * grid.z)) {
        t_blockIdx = {linear % grid.x, (linear / grid.x) % grid.y, linear / (grid.x * grid.y)};
        for (unsigned tz = 0; tz < block.z; ++tz)
            for (unsigned ty = 0; ty < block.y; ++ty)
                for (unsigned tx = 0; tx < block.x; ++tx) {
                    t_threadIdx = {tx, ty, tz};
                    t_lane = (int)((tx + block.x * (ty + block.y * tz)) % 32);
                    body();
                }
    }
}

template <typename T>
inline T exchange(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "exchange: at most 8 bytes");
    WarpCtx& w = *t_warp;
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    w.slot[t_lane] = raw;
    w.bar.arrive_and_wait();
    raw = w.slot[src_lane & 31];
    w.bar.arrive_and_wait();
    T r;
    std::memcpy(&r, &raw, sizeof(T));
    return r;
}

// mbarrier state machine in the 64-bit word the kernel reserves: bit 0 phase | 15 bits arrivals left | 15 bits
// arrival count | 32 bits pending transaction bytes (two's complement).  Copies complete at issue.
inline std::mutex g_mbar_mutex;
inline void mbar_settle(uint64_t& w) {
    const uint64_t left = (w >> 1) & 0x7fff, count = (w >> 16) & 0x7fff;
    const int32_t tx = (int32_t)(w >> 32);
    if (left == 0 && tx == 0) w = ((w & 1) ^ 1) | (count << 1) | (count << 16);
}
inline void mbar_init(uint64_t* bar, uint32_t count) {
    std::lock_guard<std::mutex> g(g_mbar_mutex);
    *bar = ((uint64_t)count << 1) | ((uint64_t)count << 16);
}
inline void mbar_arrive(uint64_t* bar, uint32_t expect_bytes) {
    std::lock_guard<std::mutex> g(g_mbar_mutex);
    uint64_t w = *bar;
    const int32_t tx = (int32_t)(w >> 32) + (int32_t)expect_bytes;
    const uint64_t left = ((w >> 1) & 0x7fff) - 1;
    w = (w & 0x7fff0001ull) | (left << 1) | ((uint64_t)(uint32_t)tx << 32);
    mbar_settle(w);
    *bar = w;
}
inline bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    std::lock_guard<std::mutex> g(g_mbar_mutex);
    return (*bar & 1) != (parity & 1);
}
inline void bulk_copy(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    std::memcpy(dst, src, bytes);
    std::lock_guard<std::mutex> g(g_mbar_mutex);
    uint64_t w = *bar;
    const int32_t tx = (int32_t)(w >> 32) - (int32_t)bytes;
    w = (w & 0xffffffffull) | ((uint64_t)(uint32_t)tx << 32);
    mbar_settle(w);
    *bar = w;
}

}  // namespace cuda_emu

#define threadIdx (cuda_emu::t_threadIdx)
#define blockIdx (cuda_emu::t_blockIdx)
#define blockDim (cuda_emu::t_blockDim)
#define gridDim (cuda_emu::t_gridDim)
constexpr int warpSize = 32;

inline void __syncthreads() { cuda_emu::t_block->bar.arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { cuda_emu::t_warp->bar.arrive_and_wait(); }

template <typename T>
inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
    const int base = cuda_emu::t_lane & ~(width - 1);
    return cuda_emu::exchange(v, base + (src & (width - 1)));
}
template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int lane_mask, int width = 32) {
    (void)width;
    return cuda_emu::exchange(v, cuda_emu::t_lane ^ lane_mask);
}
template <typename T>
inline T __shfl_down_sync(unsigned, T v, unsigned delta, int width = 32) {
    const int src = cuda_emu::t_lane + (int)delta;
    const bool in = (src & ~(width - 1)) == (cuda_emu::t_lane & ~(width - 1));
    const T r = cuda_emu::exchange(v, in ? src : cuda_emu::t_lane);
    return in ? r : v;
}
template <typename T>
inline T __shfl_up_sync(unsigned, T v, unsigned delta, int width = 32) {
    const int src = cuda_emu::t_lane - (int)delta;
    const bool in = src >= 0 && (src & ~(width - 1)) == (cuda_emu::t_lane & ~(width - 1));
    const T r = cuda_emu::exchange(v, in ? src : cuda_emu::t_lane);
    return in ? r : v;
}
inline unsigned __ballot_sync(unsigned, int pred) {
    cuda_emu::WarpCtx& w = *cuda_emu::t_warp;
    w.slot[cuda_emu::t_lane] = pred ? 1u : 0u;
    w.bar.arrive_and_wait();
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) r |= (unsigned)(w.slot[l] & 1u) << l;
    r &= w.live.load();  // a lane that has returned does not vote
    w.bar.arrive_and_wait();
    return r;
}
inline unsigned __reduce_add_sync(unsigned, unsigned v) {
    cuda_emu::WarpCtx& w = *cuda_emu::t_warp;
    w.slot[cuda_emu::t_lane] = v;
    w.bar.arrive_and_wait();
    const unsigned live = w.live.load();
    unsigned r = 0;
    for (int l = 0; l < 32; ++l)
        if (live >> l & 1u) r += (unsigned)w.slot[l];
    w.bar.arrive_and_wait();
    return r;
}
inline unsigned __reduce_max_sync(unsigned, unsigned v) {
    cuda_emu::WarpCtx& w = *cuda_emu::t_warp;
    w.slot[cuda_emu::t_lane] = v;
    w.bar.arrive_and_wait();
    const unsigned live = w.live.load();
    unsigned r = 0;
    for (int l = 0; l < 32; ++l)
        if (live >> l & 1u) r = std::max(r, (unsigned)w.slot[l]);
    w.bar.arrive_and_wait();
    return r;
}
inline unsigned __reduce_min_sync(unsigned, unsigned v) {
    cuda_emu::WarpCtx& w = *cuda_emu::t_warp;
    w.slot[cuda_emu::t_lane] = v;
    w.bar.arrive_and_wait();
    const unsigned live = w.live.load();
    unsigned r = 0xffffffffu;
    for (int l = 0; l < 32; ++l)
        if (live >> l & 1u) r = std::min(r, (unsigned)w.slot[l]);
    w.bar.arrive_and_wait();
    return r;
}
inline unsigned __reduce_or_sync(unsigned, unsigned v) {
    cuda_emu::WarpCtx& w = *cuda_emu::t_warp;
    w.slot[cuda_emu::t_lane] = v;
    w.bar.arrive_and_wait();
    const unsigned live = w.live.load();
    unsigned r = 0;
    for (int l = 0; l < 32; ++l)
        if (live >> l & 1u) r |= (unsigned)w.slot[l];
    w.bar.arrive_and_wait();
    return r;
}
inline unsigned __match_any_sync(unsigned, unsigned v) {
    cuda_emu::WarpCtx& w = *cuda_emu::t_warp;
    w.slot[cuda_emu::t_lane] = v;
    w.bar.arrive_and_wait();
    const unsigned live = w.live.load();
    unsigned r = 0;
    for (int l = 0; l < 32; ++l)
        if ((live >> l & 1u) && (unsigned)w.slot[l] == v) r |= 1u << l;
    w.bar.arrive_and_wait();
    return r;
}
inline unsigned __activemask() { return cuda_emu::t_warp->live.load(); }
inline void __nanosleep(unsigned) { std::this_thread::yield(); }
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, !pred) == 0; }

// ---- atomics --------------------------------------------------------------------------------------------------
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline float atomicAdd(float* p, float v) {
    uint32_t* u = reinterpret_cast<uint32_t*>(p);
    uint32_t old = __atomic_load_n(u, __ATOMIC_RELAXED), want;
    float f;
    do {
        std::memcpy(&f, &old, 4);
        f += v;
        std::memcpy(&want, &f, 4);
    } while (!__atomic_compare_exchange_n(u, &old, want, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
    std::memcpy(&f, &old, 4);
    return f;
}
inline double atomicAdd(double* p, double v) {
    uint64_t* u = reinterpret_cast<uint64_t*>(p);
    uint64_t old = __atomic_load_n(u, __ATOMIC_RELAXED), want;
    double f;
    do {
        std::memcpy(&f, &old, 8);
        f += v;
        std::memcpy(&want, &f, 8);
    } while (!__atomic_compare_exchange_n(u, &old, want, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
    std::memcpy(&f, &old, 8);
    return f;
}
inline unsigned atomicMin(unsigned* p, unsigned v) {
    unsigned old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
    }
    return old;
}
inline unsigned atomicMax(unsigned* p, unsigned v) {
    unsigned old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v > old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
    }
    return old;
}
inline int atomicMin(int* p, int v) {
    int old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
    }
    return old;
}
inline int atomicMax(int* p, int v) {
    int old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v > old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
    }
    return old;
}
inline float4 atomicAdd(float4* p, float4 v) {  // sm_90+: one 16-byte reduction
    float4 old;
    old.x = atomicAdd(&p->x, v.x);
    old.y = atomicAdd(&p->y, v.y);
    old.z = atomicAdd(&p->z, v.z);
    old.w = atomicAdd(&p->w, v.w);
    return old;
}
inline float2 atomicAdd(float2* p, float2 v) {
    float2 old;
    old.x = atomicAdd(&p->x, v.x);
    old.y = atomicAdd(&p->y, v.y);
    return old;
}
inline unsigned atomicExch(unsigned* p, unsigned v) { return __atomic_exchange_n(p, v, __ATOMIC_RELAXED); }
inline unsigned atomicCAS(unsigned* p, unsigned cmp, unsigned v) {
    __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED);
    return cmp;
}
inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

// ---- intrinsics -----------------------------------------------------------------------------------------------
template <typename T>
inline T __ldg(const T* p) { return *p; }
template <typename T>
inline T __ldcg(const T* p) { return *p; }
inline float __expf(float x) { return std::exp(x); }
inline float __logf(float x) { return std::log(x); }
inline float __log2f(float x) { return std::log2(x); }
inline float __powf(float a, float b) { return std::pow(a, b); }
inline float __fdividef(float a, float b) { return a / b; }
inline float __frcp_rn(float a) { return 1.0f / a; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fsqrt_rn(float a) { return std::sqrt(a); }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float rsqrtf(float a) { return 1.0f / std::sqrt(a); }
inline float __fmaf_rn(float a, float b, float c) { return std::fma(a, b, c); }
inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return {std::fma(a.x, b.x, c.x), std::fma(a.y, b.y, c.y)}; }
inline float2 __fmul2_rn(float2 a, float2 b) { return {a.x * b.x, a.y * b.y}; }
inline float2 __fadd2_rn(float2 a, float2 b) { return {a.x + b.x, a.y + b.y}; }
inline float __frsqrt_rn(float a) { return 1.0f / std::sqrt(a); }
inline float __fsub_rn(float a, float b) { return a - b; }
inline int __float2int_rd(float a) { return (int)std::floor(a); }
inline int __float2int_ru(float a) { return (int)std::ceil(a); }
inline int __float2int_rn(float a) { return (int)std::nearbyint(a); }
inline int __float2int_rz(float a) { return (int)a; }
inline unsigned __float2uint_rd(float a) { return (unsigned)std::floor(a); }
inline unsigned __float2uint_ru(float a) { return (unsigned)std::ceil(a); }
inline unsigned __float2uint_rz(float a) { return (unsigned)a; }
inline float __saturatef(float a) { return a < 0.0f ? 0.0f : (a > 1.0f ? 1.0f : a); }
inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
inline int __float_as_int(float f) { int u; std::memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
inline float __int_as_float(int u) { float f; std::memcpy(&f, &u, 4); return f; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
// fabsf / fmaxf / fminf / fmaf / expf ...: the C library's global-namespace functions (<cmath>)
template <typename T>
inline T min(T a, T b) { return b < a ? b : a; }
template <typename T>
inline T max(T a, T b) { return a < b ? b : a; }
inline unsigned min(unsigned a, int b) { return min(a, (unsigned)b); }
inline unsigned min(int a, unsigned b) { return min((unsigned)a, b); }

// kernel launch through a macro shared with the nvcc build (csrc/fgs_launch.h)
#define FGS_LAUNCH(kernel, grid, block, stream, ...) cuda_emu::launch((grid), (block), [&] { kernel(__VA_ARGS__); })
#define FGS_LAUNCH_SERIAL(kernel, grid, block, stream, ...) cuda_emu::launch_serial((grid), (block), [&] { kernel(__VA_ARGS__); })
#define FGS_LAUNCH_SMEM(kernel, grid, block, smem, stream, ...) cuda_emu::launch((grid), (block), [&] { kernel(__VA_ARGS__); }, (smem))
#define FGS_DYNAMIC_SMEM(type, name, alignment) type* name = reinterpret_cast<type*>(cuda_emu::t_dyn_smem)

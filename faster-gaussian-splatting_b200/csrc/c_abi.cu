// c_abi.cu -- host drivers and the extern "C" surface declared in include/fgs_b200.h.
//
// Mirrors the reference's torch-free host layer (rasterization/src/forward.cu:11-259,
// backward.cu:8-125, inference.cu:11-226, adam/src/adam.cu:36-71) with a different pipeline:
//   preprocess -> [async read-back of (V, I)] -> onesweep depth sort -> chained-scan offsets
//   -> (wait for I) -> instances -> onesweep tile sort -> ranges + record gather -> persistent blend.
// One host wait per forward (the reference blocks three times), hidden behind the depth sort; all
// work goes to the caller's stream; no bucket buffers.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "../../include/fgs_b200.h"
#include "fgs_device.cuh"
#include "fgs_host.h"
#include "fgs_launch.h"

namespace fgs {
namespace {

thread_local char g_err[512] = "";
}  // namespace

// error sink of the self-contained operator files (loss.cu, densify.cu): same text buffer as everything else
int report_error(int code, const char* where, const char* detail) {
    std::snprintf(g_err, sizeof(g_err), "%s: %s", where, detail);
    return code;
}

namespace {

int fail(cudaError_t e, const char* where) {
    std::snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return (int)e;
}
int fail_arg(const char* msg) {
    std::snprintf(g_err, sizeof(g_err), "%s", msg);
    return -1;
}
#define FGS_TRY(expr, where)                           \
    do {                                               \
        cudaError_t e_ = (expr);                       \
        if (e_ != cudaSuccess) return fail(e_, where); \
    } while (0)

// ---- optional per-kernel timing (CUDA events on the launching stream), used by bench.py -----------
struct Profiler {
    static constexpr int kMaxRecords = 512;
    bool enabled = false;
    int n = 0;
    const char* names[kMaxRecords];
    cudaEvent_t start[kMaxRecords];
    cudaEvent_t stop[kMaxRecords];
    int created = 0;
};
Profiler& prof() {
    static Profiler p;
    return p;
}
struct ProfScope {
    int slot = -1;
    cudaStream_t stream;
    ProfScope(const char* name, cudaStream_t s) : stream(s) {
        Profiler& p = prof();
        if (!p.enabled || p.n >= Profiler::kMaxRecords) return;
        if (p.n >= p.created) {
            if (cudaEventCreate(&p.start[p.created]) != cudaSuccess || cudaEventCreate(&p.stop[p.created]) != cudaSuccess) return;
            ++p.created;
        }
        slot = p.n++;
        p.names[slot] = name;
        cudaEventRecord(p.start[slot], stream);
    }
    ~ProfScope() {
        if (slot >= 0) cudaEventRecord(prof().stop[slot], stream);
    }
};

#define FGS_KERNEL(expr, where)                        \
    do {                                               \
        ProfScope ps_(where, stream);                  \
        cudaError_t e_ = (expr);                       \
        if (e_ != cudaSuccess) return fail(e_, where); \
    } while (0)

int bits_for(int max_value) {  // number of bits needed to represent max_value (>= 1)
    int b = 1;
    while ((max_value >> b) != 0 && b < 32) ++b;
    return b;
}

// ---- scratch layouts (recomputed from sizes only, like buffer_utils.h:30-43) --------------------
struct Carver {
    char* base;
    size_t off = 0;
    explicit Carver(char* b) : base(b) {}
    template <typename T>
    T* take(size_t count) {
        off = align_up(off, 128);
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += count * sizeof(T);
        return p;
    }
    size_t bytes() { return align_up(off, 128); }
};

struct PrimitiveLayout {
    uint32_t* counters;
    Rec* recs;
    uint32_t* n_touched;
    uint2* tile_info;    // per Gaussian: tile footprint (fgs_device.cuh)
    uint2* sorted_info;  // the footprints of the visible Gaussians in depth order
    float4* colors;      // deferred-colour forward: colour per Gaussian (color_kernel), else unused
    uint32_t* keys[2];
    uint32_t* idx[2];
    uint32_t* offsets;
    uint32_t* hist;
    uint32_t* depth_hist;  // kDepthHistReplicas copies of the depth sort's first-digit histogram (filled by preprocess)
    uint32_t* lookback;
    uint32_t* scan_lookback;
    size_t zero_bytes;  // counters .. scan_lookback are contiguous: one memset clears them all
    size_t bytes;
    PrimitiveLayout(char* blob, int n) {
        Carver c(blob);
        const size_t nn = (size_t)(n > 0 ? n : 1);
        counters = c.take<uint32_t>(kCtrCount);
        hist = c.take<uint32_t>(4 * kRadixMax);
        depth_hist = c.take<uint32_t>(kDepthHistReplicas * kRadixMax);
        lookback = c.take<uint32_t>(radix_sort_lookback_words(n, 32));
        scan_lookback = c.take<uint32_t>(offsets_scan_lookback_words(n));
        zero_bytes = c.bytes();
        recs = c.take<Rec>(nn);
        n_touched = c.take<uint32_t>(nn);
        tile_info = c.take<uint2>(nn);
        sorted_info = c.take<uint2>(nn);
        keys[0] = c.take<uint32_t>(nn);
        keys[1] = c.take<uint32_t>(nn);
        idx[0] = c.take<uint32_t>(nn);
        idx[1] = c.take<uint32_t>(nn);
        offsets = c.take<uint32_t>(nn);
        colors = c.take<float4>(nn);
        bytes = c.bytes();
    }
};

struct TileLayout {
    uint2* ranges;
    uint32_t* tile_max_n;
    float* final_T;
    uint32_t* n_processed;
    size_t bytes;
    TileLayout(char* blob, int n_tiles, int n_pixels, bool training) {
        Carver c(blob);
        ranges = c.take<uint2>((size_t)n_tiles);
        tile_max_n = c.take<uint32_t>((size_t)n_tiles * kBlendWarps);  // last contributor per (tile, sub-tile) item
        final_T = training ? c.take<float>((size_t)n_pixels) : nullptr;
        n_processed = training ? c.take<uint32_t>((size_t)n_pixels) : nullptr;
        bytes = c.bytes();
    }
};

template <typename KeyT>
struct InstanceLayout {
    Rec* inst_recs;
    KeyT* keys[2];
    uint32_t* vals[2];
    uint32_t* hist;
    uint32_t* lookback;
    size_t bytes;
    InstanceLayout(char* blob, int n_inst, int end_bit) {
        Carver c(blob);
        const size_t ni = (size_t)(n_inst > 0 ? n_inst : 1);
        inst_recs = c.take<Rec>(ni);  // first: the only part the backward reads
        keys[0] = c.take<KeyT>(ni);
        keys[1] = c.take<KeyT>(ni);
        vals[0] = c.take<uint32_t>(ni);
        vals[1] = c.take<uint32_t>(ni);
        hist = c.take<uint32_t>(4 * kRadixMax);
        lookback = c.take<uint32_t>(radix_sort_lookback_words(n_inst, end_bit));
        bytes = c.bytes();
    }
};

Cam make_cam(const fgs_settings& s, const float* w2c, const float* cam_position, const float* bg_color) {
    Cam c{};
    c.w2c = w2c;
    c.cam_pos = cam_position;
    c.bg = bg_color;
    c.width = (float)s.width;
    c.height = (float)s.height;
    c.focal_x = s.focal_x;
    c.focal_y = s.focal_y;
    c.center_x = s.center_x;
    c.center_y = s.center_y;
    c.near_plane = s.near_plane;
    c.far_plane = s.far_plane;
    c.grid_w = div_up(s.width, kTileW);
    c.grid_h = div_up(s.height, kTileH);
    c.img_w = s.width;
    c.img_h = s.height;
    c.active_sh_bases = s.active_sh_bases;
    c.total_sh_rest = s.total_sh_bases_rest;
    c.proper_aa = s.proper_antialiasing;
    return c;
}

// pinned read-back slot, side stream and events, created once PER DEVICE (streams and events belong to the device
// that was current when they were created).  The 8-byte copy of (V, I) runs on its own stream so that the launching
// stream goes from preprocess straight into the depth sort (a copy queued between two kernels of one stream costs a
// hand-over to the copy engine and back).  `lock` serialises the host-side launch sequences of one device: the slot,
// the side stream and the capacity hints of a device are used by one forward / backward at a time, whichever host
// thread or stream it comes from (calls on different devices do not contend).
struct Readback {
    uint32_t* host = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t counted = nullptr;  // preprocess done (launching stream)
    cudaEvent_t event = nullptr;    // copy done (side stream)
    std::mutex lock;
};
Readback& readback() {
    static Readback rb[kMaxDevices];
    return rb[device_slot()];
}
cudaError_t ensure_readback() {
    Readback& rb = readback();
    if (rb.host) return cudaSuccess;
    cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&rb.host), 64, cudaHostAllocDefault);
    if (e != cudaSuccess) return e;
    e = cudaStreamCreateWithFlags(&rb.stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) return e;
    e = cudaEventCreateWithFlags(&rb.counted, cudaEventDisableTiming);
    if (e != cudaSuccess) return e;
    return cudaEventCreateWithFlags(&rb.event, cudaEventDisableTiming);
}

// Instance-buffer capacity the next forward is launched with before its own instance count has come back
// from the device (0: none yet).  Sizing speculatively keeps the host ahead of the GPU: the one read-back of
// (V, I) is then only *checked* after everything has been queued, instead of stalling the launch sequence
// between the offsets scan and create_instances.  Kept per device AND per shape class (tile count x training/inference):
// a process that alternates a 1080p training view with a 4K evaluation view, or drives two GPUs, keeps one hint for each
// instead of thrashing a single one.  A small direct-mapped table; a collision only costs the cold path once.
struct CapacityHints {
    static constexpr int kEntries = 8;
    struct Entry {
        int n_tiles = 0, n_primitives = 0, capacity = 0;
        bool training = false;
    } e[kEntries];
    int forced = -1;  // fgs_debug_set_capacity_hint: overrides the table for the next forward (tests)
    Entry& entry(int n_tiles, bool training) { return e[((unsigned)n_tiles * 2u + (training ? 1u : 0u)) % kEntries]; }
    int get(int n_tiles, int n_primitives, bool training) {
        if (forced >= 0) return forced;
        const Entry& x = entry(n_tiles, training);
        // a different scene size is a different workload: start cold rather than with a stale size
        return (x.n_tiles == n_tiles && x.training == training && x.n_primitives == n_primitives) ? x.capacity : 0;
    }
    void put(int n_tiles, int n_primitives, bool training, int capacity) {
        Entry& x = entry(n_tiles, training);
        x.n_tiles = n_tiles;
        x.n_primitives = n_primitives;
        x.training = training;
        x.capacity = capacity;
        forced = -1;
        last = capacity;
    }
    int last = 0;  // most recent suggestion on this device (fgs_debug_get_capacity_hint)
    int reruns = 0;  // forwards whose speculative tail had to be queued again (fgs_debug_get_forward_reruns)
};
CapacityHints& capacity_hints() {
    static CapacityHints h[kMaxDevices];
    return h[device_slot()];
}

// Opt-in overlapped schedule of the fused backward + Adam (see backward_impl); FGS_OVERLAP_ADAM=1 in the environment or
// fgs_debug_set_overlap_adam(1).  Off by default until it has been measured on hardware.
bool g_overlap_adam = [] {
    const char* e = std::getenv("FGS_OVERLAP_ADAM");
    return e != nullptr && e[0] == '1';
}();

// Deferred sh_rest update (fgs_backward_adam_ex with FGS_DEFER_SH_REST): per device, a second side stream and the event
// that marks the end of the pending update.  Whoever reads or writes the SH coefficients next makes its stream wait for
// that event first: the deferred-colour forward just before its colour kernel (everything in front of it overlaps the
// update), every other entry point at its very beginning (join_deferred).
struct Deferred {
    cudaStream_t stream = nullptr;
    cudaEvent_t start = nullptr;  // preprocess backward done (launching stream)
    cudaEvent_t done = nullptr;   // adam_rest_kernel done (side stream)
    bool pending = false;
    int ctas_per_sm = 2;
};
Deferred& deferred() {
    static Deferred d[kMaxDevices];
    return d[device_slot()];
}
cudaError_t ensure_deferred() {
    Deferred& d = deferred();
    if (d.stream) return cudaSuccess;
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);  // lowest priority: the update yields to the forward it runs under
    cudaError_t e = cudaStreamCreateWithPriority(&d.stream, cudaStreamNonBlocking, lo);
    if (e != cudaSuccess) return e;
    e = cudaEventCreateWithFlags(&d.start, cudaEventDisableTiming);
    if (e != cudaSuccess) return e;
    return cudaEventCreateWithFlags(&d.done, cudaEventDisableTiming);
}
// make `stream` wait for a pending deferred update of the current device (no-op if there is none)
cudaError_t join_deferred(cudaStream_t stream) {
    Deferred& d = deferred();
    if (!d.pending) return cudaSuccess;
    d.pending = false;
    return cudaStreamWaitEvent(stream, d.done, 0);
}

// instrumentation build only: device counters of the blend kernels (fgs_device.cuh, FGS_BLEND_STATS)
unsigned long long* blend_stats_buffer() {
#ifdef FGS_BLEND_STATS
    static PerDevice<unsigned long long*> buf;
    unsigned long long*& p = buf.here();
    if (!p && cudaMalloc(reinterpret_cast<void**>(&p), 16 * sizeof(unsigned long long)) == cudaSuccess)
        cudaMemset(p, 0, 16 * sizeof(unsigned long long));
    return p;
#else
    return nullptr;
#endif
}

int next_capacity_hint(int n_instances, int current) {
    // grow with 25 % head-room, shrink only when more than half of the buffer is unused: sizes stay stable
    const long long want = (long long)n_instances + n_instances / 4 + 65536;
    if (n_instances <= current && (long long)current <= 2 * want) return current;
    return want > 0x7fffffffLL ? 0x7fffffff : (int)want;
}

// deferred-colour forward (fgs_forward_ex): what the colour kernel needs, and the event it additionally waits for
struct DeferredColor {
    bool on;
    const float* means;
    const float* sh0;
    const float* sh_rest;
    cudaEvent_t wait;
    int n;
};

// everything after the offsets scan: instances, tile sort, ranges + record gather, blend.  `capacity` is the
// number of instances the instance buffers are sized for (>= the true count, or the caller re-runs this).
template <typename KeyT>
int finish_forward(fgs_resize_fn resize, void* user_instance, const PrimitiveLayout& pl, const TileLayout& tl,
                   const uint32_t* sorted_idx, int n, int capacity, int n_tiles, int end_bit, const Cam& cam,
                   float* image, bool training, bool to_chw, bool clamp_output, cudaStream_t stream,
                   const DeferredColor& dc) {
    InstanceLayout<KeyT> sizing(nullptr, capacity, end_bit);
    char* blob = resize(user_instance, sizing.bytes);
    if (!blob) return fail_arg("instance buffer resize callback returned NULL");
    InstanceLayout<KeyT> il(blob, capacity, end_bit);
    if (capacity > 0) {
        // sort scratch (histograms, then look-back words: adjacent) is cleared before create_instances,
        // which accumulates the histogram of the tile sort's first digit while it writes the keys
        FGS_TRY(cudaMemsetAsync(il.hist, 0, sizeof(uint32_t) * (4 * kRadixMax + radix_sort_lookback_words(capacity, end_bit)), stream),
                "memset tile sort scratch");
        FGS_KERNEL(launch_create_instances<KeyT>(sorted_idx, pl.offsets, pl.recs, pl.sorted_info, il.keys[0], il.vals[0],
                                              pl.counters + kCtrVisible, n, cam.grid_w, il.hist,
                                              radix_sort_pass0_bits(end_bit), pl.counters + kCtrBigList,
                                              /*work list: the depth sort's key buffers are dead by now*/ pl.keys[0],
                                              pl.keys[1], capacity, stream),
                "create_instances");
        bool in_alt = false;
        FGS_KERNEL(radix_sort_pairs<KeyT>(il.keys[0], il.keys[1], il.vals[0], il.vals[1], pl.counters + kCtrInstances,
                                       capacity, end_bit, il.hist, il.lookback, pl.counters + kCtrTicket0 + 5,
                                       &in_alt, /*zero_scratch=*/false, /*pass0_hist=*/il.hist, 1, stream),
                "tile sort");
        const float4* colors = nullptr;
        if (dc.on) {
            // the SH coefficients are only needed from here on: wait for whoever still writes them (the deferred Adam
            // update of the previous step on this device; `dc.wait`: e.g. their broadcast from a peer GPU), then colour
            // the visible Gaussians
            FGS_TRY(join_deferred(stream), "join deferred update");
            if (dc.wait) FGS_TRY(cudaStreamWaitEvent(stream, dc.wait, 0), "wait for the colour event");
            FGS_KERNEL(launch_color(dc.means, dc.sh0, dc.sh_rest, pl.n_touched, pl.colors, dc.n, cam, stream), "color");
            colors = pl.colors;
        }
        FGS_KERNEL(launch_ranges_and_gather<KeyT>(il.keys[in_alt ? 1 : 0], il.vals[in_alt ? 1 : 0], pl.recs, il.inst_recs,
                                               tl.ranges, pl.counters + kCtrInstances, capacity, cam.grid_w, stream, colors),
                "ranges_and_gather");
    }
    BlendFwdArgs a{};
    a.inst_recs = il.inst_recs;
    a.tile_ranges = tl.ranges;
    a.image = image;
    a.final_T = tl.final_T;
    a.n_processed = tl.n_processed;
    a.tile_max_n = tl.tile_max_n;
    a.queue = pl.counters + kCtrTileQueue;
    a.stats = blend_stats_buffer();
    a.n_tiles = n_tiles;
    a.training = training;
    a.to_chw = to_chw;
    a.clamp_output = clamp_output;
    FGS_KERNEL(launch_blend_forward(a, cam, stream), "blend_forward");
    return 0;
}

int finish_forward_any(bool keys16, fgs_resize_fn resize, void* user_instance, const PrimitiveLayout& pl,
                       const TileLayout& tl, const uint32_t* sorted_idx, int n, int capacity, int n_tiles, int end_bit,
                       const Cam& cam, float* image, bool training, bool to_chw, bool clamp_output, cudaStream_t stream,
                       const DeferredColor& dc) {
    return keys16 ? finish_forward<uint16_t>(resize, user_instance, pl, tl, sorted_idx, n, capacity, n_tiles, end_bit, cam,
                                             image, training, to_chw, clamp_output, stream, dc)
                  : finish_forward<uint32_t>(resize, user_instance, pl, tl, sorted_idx, n, capacity, n_tiles, end_bit, cam,
                                             image, training, to_chw, clamp_output, stream, dc);
}

int forward_impl(fgs_resize_fn resize, void* user_primitive, void* user_tile, void* user_instance, const float* means,
                 const float* scales, const float* rotations, const float* opacities, const float* sh0,
                 const float* sh_rest, const float* w2c, const float* cam_position, const float* bg_color, float* image,
                 int n, const fgs_settings* s, bool training, bool to_chw, bool clamp_output, cudaStream_t stream,
                 fgs_forward_state* state_out, bool defer_color = false, cudaEvent_t color_wait = nullptr) {
    if (!resize || !s || !image) return fail_arg("fgs_forward: NULL argument");
    if (defer_color && rotations == nullptr) return fail_arg("fgs_forward_ex: no deferred colour in covariance-input mode");
    // a pending deferred update of the SH coefficients: the deferred-colour forward only needs it before its colour
    // kernel (below); everything else reads the coefficients right away
    if (!defer_color) FGS_TRY(join_deferred(stream), "join deferred update");
    if (s->width <= 0 || s->height <= 0) return fail_arg("fgs_forward: empty image");
    if (s->total_sh_bases_rest < 0 || s->total_sh_bases_rest > 15) return fail_arg("fgs_forward: sh_rest must have <= 15 bases");
    const Cam cam = make_cam(*s, w2c, cam_position, bg_color);
    const int n_tiles = cam.grid_w * cam.grid_h;
    const int n_pixels = s->width * s->height;
    const int end_bit = bits_for(n_tiles - 1);
    if (cam.grid_w >= kMaxGridDim || cam.grid_h >= kMaxGridDim) return fail_arg("fgs_forward: image larger than 4095 tiles per side");
    FGS_TRY(ensure_readback(), "readback alloc");
    Readback& rb = readback();
    std::lock_guard<std::mutex> device_lock(rb.lock);
    CapacityHints& hints = capacity_hints();

    TileLayout tsz(nullptr, n_tiles, n_pixels, training);
    char* tile_blob = resize(user_tile, tsz.bytes);
    if (!tile_blob) return fail_arg("tile buffer resize callback returned NULL");
    TileLayout tl(tile_blob, n_tiles, n_pixels, training);
    PrimitiveLayout psz(nullptr, n);
    char* prim_blob = resize(user_primitive, psz.bytes);
    if (!prim_blob) return fail_arg("primitive buffer resize callback returned NULL");
    PrimitiveLayout pl(prim_blob, n);

    FGS_TRY(cudaMemsetAsync(pl.counters, 0, pl.zero_bytes, stream), "memset counters + sort scratch");
    // empty tiles keep the range (0, 0); tile_max_n is written for every item by the blend forward
    FGS_TRY(cudaMemsetAsync(tl.ranges, 0, sizeof(uint2) * n_tiles, stream), "memset ranges");

    // rotations == NULL: covariance-input mode, `scales` holds (N,6) 3D covariances (include/fgs_b200.h)
    FGS_KERNEL(launch_preprocess(means, scales, rotations, opacities, sh0, sh_rest, pl.recs, pl.n_touched, pl.tile_info, pl.keys[0],
                              pl.idx[0], pl.counters, pl.depth_hist, n, cam, !training, stream, rotations ? nullptr : scales,
                              defer_color),
            "preprocess");
    FGS_TRY(cudaEventRecord(rb.counted, stream), "event record");
    FGS_TRY(cudaStreamWaitEvent(rb.stream, rb.counted, 0), "read-back stream wait");
    FGS_TRY(cudaMemcpyAsync(rb.host, pl.counters, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, rb.stream), "count read-back");
    FGS_TRY(cudaEventRecord(rb.event, rb.stream), "event record");

    // depth order: bits [0, 32) of the positive float depth (sign bit is always 0 after the near cull)
    bool in_alt = false;
    FGS_KERNEL(radix_sort_pairs<uint32_t>(pl.keys[0], pl.keys[1], pl.idx[0], pl.idx[1], pl.counters + kCtrVisible, n, 32,
                                       pl.hist, pl.lookback, pl.counters + kCtrTicket0, &in_alt, /*zero_scratch=*/false,
                                       /*pass0_hist=*/pl.depth_hist, kDepthHistReplicas, stream),
            "depth sort");
    const uint32_t* sorted_idx = pl.idx[in_alt ? 1 : 0];
    FGS_KERNEL(launch_offsets_scan(sorted_idx, pl.tile_info, pl.offsets, pl.sorted_info, pl.counters + kCtrVisible, n, pl.scan_lookback,
                                pl.counters + kCtrTicket0 + 4, /*zero_scratch=*/false, stream),
            "offsets scan");

    // Launch the rest with the capacity the previous forward suggested, and only then look at this forward's
    // own count (the read-back has long completed by then).  No suggestion yet: wait for the count first.
    const bool keys16 = end_bit <= 16;
    // deferred colour: evaluated inside finish_forward, between the tile sort and the record gather -- as late as possible
    DeferredColor dc{};
    dc.on = defer_color;
    dc.means = means;
    dc.sh0 = sh0;
    dc.sh_rest = sh_rest;
    dc.wait = color_wait;
    dc.n = n;
    const int hinted = hints.get(n_tiles, n, training);
    int capacity = hinted;
    if (capacity > 0) {
        const int rc = finish_forward_any(keys16, resize, user_instance, pl, tl, sorted_idx, n, capacity, n_tiles, end_bit, cam,
                                          image, training, to_chw, clamp_output, stream, dc);
        if (rc != 0) return rc;
    }
    FGS_TRY(cudaEventSynchronize(rb.event), "count read-back wait");
    const int n_visible = (int)rb.host[0];
    const int n_instances = (int)rb.host[1];
    if (n_instances < 0) return fail_arg("fgs_forward: more than 2^31 instances");
    if (capacity == 0 || n_instances > capacity) {
        if (capacity > 0) {
            ++hints.reruns;
            // the speculative pass ran out of room: create_instances skipped every write that did not fit and
            // the stages behind it saw an empty instance list (bounded_count), so nothing was overrun and nothing
            // uninitialised was read.  Reset what it consumed -- blend queues, sort tickets, work-list length,
            // tile ranges -- and redo it with the exact size
            FGS_TRY(cudaMemsetAsync(pl.counters + kCtrTileQueue, 0, sizeof(uint32_t) * (kCtrCount - kCtrTileQueue), stream),
                    "memset counters (redo)");
            FGS_TRY(cudaMemsetAsync(tl.ranges, 0, sizeof(uint2) * n_tiles, stream), "memset ranges (redo)");
        }
        capacity = n_instances;
        const int rc = finish_forward_any(keys16, resize, user_instance, pl, tl, sorted_idx, n, capacity, n_tiles, end_bit, cam,
                                          image, training, to_chw, clamp_output, stream, dc);
        if (rc != 0) return rc;
    }
    hints.put(n_tiles, n, training, next_capacity_hint(n_instances, hinted));
    if (state_out) {
        state_out->n_instances = n_instances;
        state_out->n_buckets = 0;
        state_out->selector = 0;
        state_out->n_visible = n_visible;
    }
    return 0;
}

int backward_impl(const float* grad_image, const float* image, const float* means, const float* scales,
                  const float* rotations, const float* opacities, const float* sh_rest, const float* w2c,
                  const float* cam_position, const float* bg_color, const char* primitive_buffers,
                  const char* tile_buffers, const char* instance_buffers, PreBwdArgs pa, void* grad2d_scratch,
                  float* densification_info, int n, const fgs_settings* s, const fgs_forward_state* state,
                  cudaStream_t stream, bool defer_rest = false) {
    if (!s || !state) return fail_arg("fgs_backward: NULL argument");
    if (n <= 0) return 0;
    FGS_TRY(join_deferred(stream), "join deferred update");  // the backward reads (and, fused, writes) the SH coefficients
    const Cam cam = make_cam(*s, w2c, cam_position, bg_color);
    const int n_tiles = cam.grid_w * cam.grid_h;
    const int n_pixels = s->width * s->height;
    const int end_bit = bits_for(n_tiles - 1);
    PrimitiveLayout pl(const_cast<char*>(primitive_buffers), n);
    TileLayout tl(const_cast<char*>(tile_buffers), n_tiles, n_pixels, true);
    const Rec* inst_recs = end_bit <= 16
                               ? InstanceLayout<uint16_t>(const_cast<char*>(instance_buffers), state->n_instances, end_bit).inst_recs
                               : InstanceLayout<uint32_t>(const_cast<char*>(instance_buffers), state->n_instances, end_bit).inst_recs;
    Grad2D* grad2d = reinterpret_cast<Grad2D*>(grad2d_scratch);
    pa.n_touched = pl.n_touched;
    pa.n = n;
    // Overlapped schedule (opt-in): the zero-gradient Adam update of the sh_rest rows of invisible Gaussians -- pure HBM
    // traffic that depends on nothing the blend backward produces -- runs on the side stream while the issue-bound blend
    // backward leaves the memory system idle; the preprocess backward then skips those rows.
    const bool overlap = g_overlap_adam && state->n_instances > 0 && preprocess_backward_can_overlap(pa, cam);
    std::unique_lock<std::mutex> device_lock;
    if (overlap) {
        FGS_TRY(ensure_readback(), "side stream");
        Readback& rb = readback();
        device_lock = std::unique_lock<std::mutex>(rb.lock);
        FGS_TRY(cudaEventRecord(rb.counted, stream), "event record (overlap)");
        FGS_TRY(cudaStreamWaitEvent(rb.stream, rb.counted, 0), "side stream wait (overlap)");
        FGS_TRY(launch_adam_rest_invisible(pa, cam, rb.stream), "adam_rest_invisible");
        FGS_TRY(cudaEventRecord(rb.event, rb.stream), "event record (overlap)");
    }
    FGS_TRY(cudaMemsetAsync(grad2d, 0, sizeof(Grad2D) * (size_t)n, stream), "memset grad2d");
    if (state->n_instances > 0) {
        FGS_TRY(cudaMemsetAsync(pl.counters + kCtrTileQueueBwd, 0, sizeof(uint32_t), stream), "memset queue");
        BlendBwdArgs ba{};
        ba.inst_recs = inst_recs;
        ba.tile_ranges = tl.ranges;
        ba.image = image;
        ba.grad_image = grad_image;
        ba.final_T = tl.final_T;
        ba.n_processed = tl.n_processed;
        ba.tile_max_n = tl.tile_max_n;
        ba.grad2d = grad2d;
        ba.queue = pl.counters + kCtrTileQueueBwd;
        ba.stats = blend_stats_buffer();
        ba.n_tiles = n_tiles;
        FGS_KERNEL(launch_blend_backward(ba, cam, stream), "blend_backward");
    }
    pa.means = means;
    pa.scales = scales;
    pa.cov3d = rotations == nullptr;  // covariance-input mode: scales holds (N,6), g_scales receives (N,6)
    pa.rotations = rotations;
    pa.opacities = opacities;
    pa.sh_rest = sh_rest;
    pa.grad2d = grad2d;
    pa.densification_info = densification_info;
    if (overlap) FGS_TRY(cudaStreamWaitEvent(stream, readback().event, 0), "join side stream (overlap)");
    const bool defer = defer_rest && !overlap && preprocess_backward_can_defer_rest(pa, cam);
    pa.defer_rest = defer;
    FGS_KERNEL(launch_preprocess_backward(pa, cam, stream, overlap), "preprocess_backward");
    if (defer) {
        // the sh_rest part of the step (three quarters of the fused kernel's HBM traffic) goes to the low-priority side
        // stream; the caller's stream is free to start the next forward's geometry pass and sorts
        FGS_TRY(ensure_deferred(), "deferred stream");
        Deferred& d = deferred();
        FGS_TRY(cudaEventRecord(d.start, stream), "event record (deferred)");
        FGS_TRY(cudaStreamWaitEvent(d.stream, d.start, 0), "deferred stream wait");
        {
            ProfScope ps_("adam_rest (deferred, side stream)", d.stream);
            FGS_TRY(launch_adam_rest(pa, cam, d.stream, d.ctas_per_sm), "adam_rest");
        }
        FGS_TRY(cudaEventRecord(d.done, d.stream), "event record (deferred)");
        d.pending = true;
    }
    return 0;
}

}  // namespace
}  // namespace fgs

using namespace fgs;

extern "C" size_t fgs_sort_scratch_bytes(int32_t n_max, int32_t end_bit);

template <typename KeyT>
static int sort_pairs_entry(KeyT* keys, KeyT* keys_alt, uint32_t* vals, uint32_t* vals_alt, const uint32_t* n_dev,
                            int32_t n_max, int32_t end_bit, void* scratch, size_t scratch_bytes, int32_t* result_in_alt,
                            void* stream) {
    if (end_bit < 1 || end_bit > 8 * (int)sizeof(KeyT)) return fail_arg("fgs_sort_pairs: end_bit out of range for the key type");
    if (scratch_bytes < fgs_sort_scratch_bytes(n_max, end_bit)) return fail_arg("fgs_sort_pairs: scratch too small");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    Carver c(reinterpret_cast<char*>(scratch));
    uint32_t* tickets = c.take<uint32_t>(8);
    uint32_t* hist = c.take<uint32_t>(4 * kRadixMax);
    uint32_t* lookback = c.take<uint32_t>(radix_sort_lookback_words(n_max, end_bit));
    FGS_TRY(cudaMemsetAsync(tickets, 0, 8 * sizeof(uint32_t), st), "memset tickets");
    bool in_alt = false;
    FGS_TRY(radix_sort_pairs<KeyT>(keys, keys_alt, vals, vals_alt, n_dev, n_max, end_bit, hist, lookback, tickets, &in_alt,
                                   /*zero_scratch=*/true, /*pass0_hist=*/nullptr, 1, st),
            "radix_sort_pairs");
    if (result_in_alt) *result_in_alt = in_alt ? 1 : 0;
    return 0;
}

extern "C" {

int fgs_abi_version(void) { return FGS_B200_ABI_VERSION; }
const char* fgs_last_error(void) { return g_err; }

int fgs_forward(fgs_resize_fn resize, void* user_primitive, void* user_tile, void* user_instance, const float* means,
                const float* scales, const float* rotations, const float* opacities, const float* sh_coefficients_0,
                const float* sh_coefficients_rest, const float* w2c, const float* cam_position, const float* bg_color,
                float* image, int32_t n_primitives, const fgs_settings* settings, void* stream,
                fgs_forward_state* state_out) {
    return forward_impl(resize, user_primitive, user_tile, user_instance, means, scales, rotations, opacities,
                        sh_coefficients_0, sh_coefficients_rest, w2c, cam_position, bg_color, image, n_primitives,
                        settings, true, true, false, reinterpret_cast<cudaStream_t>(stream), state_out);
}

int fgs_inference(fgs_resize_fn resize, void* user_primitive, void* user_tile, void* user_instance, const float* means,
                  const float* scales, const float* rotations, const float* opacities, const float* sh_coefficients_0,
                  const float* sh_coefficients_rest, const float* w2c, const float* cam_position, const float* bg_color,
                  float* image, int32_t n_primitives, const fgs_settings* settings, int32_t to_chw, int32_t clamp_output,
                  void* stream) {
    return forward_impl(resize, user_primitive, user_tile, user_instance, means, scales, rotations, opacities,
                        sh_coefficients_0, sh_coefficients_rest, w2c, cam_position, bg_color, image, n_primitives,
                        settings, false, to_chw != 0, clamp_output != 0, reinterpret_cast<cudaStream_t>(stream), nullptr);
}

size_t fgs_grad2d_scratch_bytes(int32_t n_primitives) {
    return sizeof(Grad2D) * (size_t)(n_primitives > 0 ? n_primitives : 1);
}

int fgs_backward(const float* grad_image, const float* image, const float* means, const float* scales,
                 const float* rotations, const float* opacities, const float* sh_coefficients_rest, const float* w2c,
                 const float* cam_position, const float* bg_color, const char* primitive_buffers,
                 const char* tile_buffers, const char* instance_buffers, float* grad_means, float* grad_scales,
                 float* grad_rotations, float* grad_opacities, float* grad_sh_coefficients_0,
                 float* grad_sh_coefficients_rest, void* grad2d_scratch, float* densification_info, int32_t n_primitives,
                 const fgs_settings* settings, const fgs_forward_state* state, void* stream) {
    PreBwdArgs pa{};
    pa.g_means = grad_means;
    pa.g_scales = grad_scales;
    pa.g_rotations = grad_rotations;
    pa.g_opacities = grad_opacities;
    pa.g_sh0 = grad_sh_coefficients_0;
    pa.g_sh_rest = grad_sh_coefficients_rest;
    pa.fused = false;
    return backward_impl(grad_image, image, means, scales, rotations, opacities, sh_coefficients_rest, w2c, cam_position,
                         bg_color, primitive_buffers, tile_buffers, instance_buffers, pa, grad2d_scratch,
                         densification_info, n_primitives, settings, state, reinterpret_cast<cudaStream_t>(stream));
}

int fgs_backward_adam(const float* grad_image, const float* image, float* const params[6], float* const exp_avg[6],
                      float* const exp_avg_sq[6], const int32_t step[6], const double lr[6], double beta1, double beta2,
                      double eps, const float* w2c, const float* cam_position, const float* bg_color,
                      const char* primitive_buffers, const char* tile_buffers, const char* instance_buffers,
                      void* grad2d_scratch, float* densification_info, int32_t n_primitives, const fgs_settings* settings,
                      const fgs_forward_state* state, void* stream) {
    PreBwdArgs pa{};
    pa.fused = true;
    for (int t = 0; t < 6; ++t) {
        pa.p[t] = params[t];
        pa.m[t] = exp_avg[t];
        pa.v[t] = exp_avg_sq[t];
        // bias corrections in double on the host, adam.cu:54-56
        const double bc1_rcp = 1.0 / (1.0 - std::pow(beta1, (double)step[t]));
        const double bc2_sqrt_rcp = 1.0 / std::sqrt(1.0 - std::pow(beta2, (double)step[t]));
        pa.hyper.step_size[t] = (float)(lr[t] * bc1_rcp);
        pa.hyper.bias2_sqrt_rcp[t] = (float)bc2_sqrt_rcp;
    }
    pa.hyper.beta1 = (float)beta1;
    pa.hyper.beta2 = (float)beta2;
    pa.hyper.eps = (float)eps;
    // the kernel reads the pre-update parameters: means/scales/rotations/opacities/sh_rest ARE params[]
    return backward_impl(grad_image, image, params[0], params[1], params[2], params[3], params[5], w2c, cam_position,
                         bg_color, primitive_buffers, tile_buffers, instance_buffers, pa, grad2d_scratch,
                         densification_info, n_primitives, settings, state, reinterpret_cast<cudaStream_t>(stream));
}

int fgs_forward_ex(fgs_resize_fn resize, void* user_primitive, void* user_tile, void* user_instance, const float* means,
                   const float* scales, const float* rotations, const float* opacities, const float* sh_coefficients_0,
                   const float* sh_coefficients_rest, const float* w2c, const float* cam_position, const float* bg_color,
                   float* image, int32_t n_primitives, const fgs_settings* settings, void* stream,
                   fgs_forward_state* state_out, uint32_t flags, void* color_wait_event) {
    return forward_impl(resize, user_primitive, user_tile, user_instance, means, scales, rotations, opacities,
                        sh_coefficients_0, sh_coefficients_rest, w2c, cam_position, bg_color, image, n_primitives,
                        settings, true, true, false, reinterpret_cast<cudaStream_t>(stream), state_out,
                        (flags & FGS_DEFER_COLOR) != 0, reinterpret_cast<cudaEvent_t>(color_wait_event));
}

int fgs_backward_adam_ex(const float* grad_image, const float* image, float* const params[6], float* const exp_avg[6],
                         float* const exp_avg_sq[6], const int32_t step[6], const double lr[6], double beta1, double beta2,
                         double eps, const float* w2c, const float* cam_position, const float* bg_color,
                         const char* primitive_buffers, const char* tile_buffers, const char* instance_buffers,
                         void* grad2d_scratch, float* densification_info, int32_t n_primitives,
                         const fgs_settings* settings, const fgs_forward_state* state, void* stream, uint32_t flags) {
    PreBwdArgs pa{};
    pa.fused = true;
    for (int t = 0; t < 6; ++t) {
        pa.p[t] = params[t];
        pa.m[t] = exp_avg[t];
        pa.v[t] = exp_avg_sq[t];
        const double bc1_rcp = 1.0 / (1.0 - std::pow(beta1, (double)step[t]));
        const double bc2_sqrt_rcp = 1.0 / std::sqrt(1.0 - std::pow(beta2, (double)step[t]));
        pa.hyper.step_size[t] = (float)(lr[t] * bc1_rcp);
        pa.hyper.bias2_sqrt_rcp[t] = (float)bc2_sqrt_rcp;
    }
    pa.hyper.beta1 = (float)beta1;
    pa.hyper.beta2 = (float)beta2;
    pa.hyper.eps = (float)eps;
    return backward_impl(grad_image, image, params[0], params[1], params[2], params[3], params[5], w2c, cam_position,
                         bg_color, primitive_buffers, tile_buffers, instance_buffers, pa, grad2d_scratch,
                         densification_info, n_primitives, settings, state, reinterpret_cast<cudaStream_t>(stream),
                         (flags & FGS_DEFER_SH_REST) != 0);
}

int fgs_join_deferred(void* stream) {
    FGS_TRY(join_deferred(reinterpret_cast<cudaStream_t>(stream)), "fgs_join_deferred");
    return 0;
}

void fgs_debug_set_deferred_ctas_per_sm(int32_t ctas) { deferred().ctas_per_sm = ctas > 0 ? ctas : 2; }

int fgs_adam_step(const float* param_grad, float* param, float* exp_avg, float* exp_avg_sq, size_t n_elements,
                  int32_t step_count, double learning_rate, double beta1, double beta2, double epsilon, void* stream) {
    const double bc1_rcp = 1.0 / (1.0 - std::pow(beta1, (double)step_count));
    const double bc2_sqrt_rcp = 1.0 / std::sqrt(1.0 - std::pow(beta2, (double)step_count));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    FGS_TRY(join_deferred(st), "join deferred update");  // `param` may be the tensor a deferred update is still writing
    {
        ProfScope ps_("adam_step", st);
        FGS_TRY(launch_adam(param_grad, param, exp_avg, exp_avg_sq, n_elements, (float)(learning_rate * bc1_rcp),
                            (float)beta1, (float)beta2, (float)epsilon, (float)bc2_sqrt_rcp, st),
                "adam_step");
    }
    return 0;
}

int fgs_pruning_scores(fgs_resize_fn resize, void* user_primitive, void* user_tile, void* user_instance,
                       const float* means, const float* scales, const float* rotations, const float* opacities,
                       const float* sh_coefficients_0, const float* sh_coefficients_rest, const float* w2c,
                       const float* cam_position, const float* bg_color, float* image_scratch, float* scores,
                       int32_t n_primitives, const fgs_settings* settings, void* stream) {
    // forward with the per-pixel bookkeeping of the training path, then the score variant of the blend backward
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    fgs_forward_state state{};
    const int rc = forward_impl(resize, user_primitive, user_tile, user_instance, means, scales, rotations, opacities,
                                sh_coefficients_0, sh_coefficients_rest, w2c, cam_position, bg_color, image_scratch,
                                n_primitives, settings, true, true, false, st, &state);
    if (rc != 0 || state.n_instances <= 0) return rc;
    const Cam cam = make_cam(*settings, w2c, cam_position, bg_color);
    const int n_tiles = cam.grid_w * cam.grid_h;
    const int end_bit = bits_for(n_tiles - 1);
    // the blobs were just (re)allocated by the callbacks: ask for them again with the same sizes
    PrimitiveLayout psz(nullptr, n_primitives);
    PrimitiveLayout pl(resize(user_primitive, 0), n_primitives);
    TileLayout tl(resize(user_tile, 0), n_tiles, settings->width * settings->height, true);
    char* inst_blob = resize(user_instance, 0);
    if (!pl.counters || !tl.ranges || !inst_blob) return fail_arg("fgs_pruning_scores: resize(user, 0) must return the current buffer");
    const Rec* inst_recs = end_bit <= 16 ? InstanceLayout<uint16_t>(inst_blob, state.n_instances, end_bit).inst_recs
                                         : InstanceLayout<uint32_t>(inst_blob, state.n_instances, end_bit).inst_recs;
    FGS_TRY(cudaMemsetAsync(pl.counters + kCtrTileQueueBwd, 0, sizeof(uint32_t), st), "memset queue");
    BlendBwdArgs ba{};
    ba.inst_recs = inst_recs;
    ba.tile_ranges = tl.ranges;
    ba.image = image_scratch;
    ba.grad_image = nullptr;
    ba.final_T = tl.final_T;
    ba.n_processed = tl.n_processed;
    ba.tile_max_n = tl.tile_max_n;
    ba.grad2d = nullptr;
    ba.scores = scores;
    ba.queue = pl.counters + kCtrTileQueueBwd;
    ba.n_tiles = n_tiles;
    {
        cudaStream_t stream_ = st;
        (void)stream_;
        ProfScope ps_("pruning_scores", st);
        FGS_TRY(launch_blend_backward(ba, cam, st), "pruning_scores");
    }
    (void)psz;
    return 0;
}

int fgs_update_3d_filter(const float* positions, const float* w2c, float* filter_3d, uint8_t* visibility_mask,
                         int32_t n_points, int32_t width, int32_t height, float focal_x, float focal_y, float center_x,
                         float center_y, float near_plane, float clipping_tolerance, float distance2filter, void* stream) {
    FGS_TRY(launch_update_3d_filter(positions, w2c, filter_3d, visibility_mask, n_points, width, height, focal_x, focal_y,
                                    center_x, center_y, near_plane, clipping_tolerance, distance2filter,
                                    reinterpret_cast<cudaStream_t>(stream)),
            "update_3d_filter");
    return 0;
}

int fgs_relocation_adjustment(const float* old_opacities, const float* old_scales, const int64_t* n_samples_per_primitive,
                              float* new_opacities, float* new_scales, int32_t n_primitives, void* stream) {
    FGS_TRY(launch_relocation(old_opacities, old_scales, reinterpret_cast<const long long*>(n_samples_per_primitive),
                              new_opacities, new_scales, n_primitives, reinterpret_cast<cudaStream_t>(stream)),
            "relocation_adjustment");
    return 0;
}

int fgs_add_noise(const float* raw_scales, const float* raw_rotations, const float* raw_opacities,
                  const float* random_samples, float* means, int32_t n_primitives, float current_lr, void* stream) {
    FGS_TRY(launch_add_noise(raw_scales, raw_rotations, raw_opacities, random_samples, means, n_primitives, current_lr,
                             reinterpret_cast<cudaStream_t>(stream)),
            "add_noise");
    return 0;
}

void fgs_debug_set_overlap_adam(int32_t on) { g_overlap_adam = on != 0; }
int32_t fgs_debug_get_overlap_adam(void) { return g_overlap_adam ? 1 : 0; }

void fgs_debug_set_capacity_hint(int32_t capacity) { capacity_hints().forced = capacity > 0 ? capacity : 0; }
int32_t fgs_debug_get_capacity_hint(void) { return capacity_hints().last; }
int32_t fgs_debug_get_forward_reruns(void) { return capacity_hints().reruns; }

int fgs_debug_blend_stats(uint64_t out[16], int32_t reset) {
#ifdef FGS_BLEND_STATS
    unsigned long long* p = blend_stats_buffer();
    if (!p) return fail_arg("fgs_debug_blend_stats: no counter buffer");
    FGS_TRY(cudaDeviceSynchronize(), "blend stats sync");
    if (out) FGS_TRY(cudaMemcpy(out, p, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost), "blend stats copy");
    if (reset) FGS_TRY(cudaMemset(p, 0, 16 * sizeof(unsigned long long)), "blend stats reset");
    return 0;
#else
    (void)out;
    (void)reset;
    return fail_arg("fgs_debug_blend_stats: the library was not built with -DFGS_BLEND_STATS=1");
#endif
}

void fgs_profile_enable(int32_t on) {
    prof().enabled = on != 0;
    prof().n = 0;
}

int fgs_profile_read(char* out, size_t out_bytes) {
    Profiler& p = prof();
    if (!out || out_bytes == 0) return fail_arg("fgs_profile_read: NULL buffer");
    size_t off = 0;
    out[0] = 0;
    for (int i = 0; i < p.n; ++i) {
        FGS_TRY(cudaEventSynchronize(p.stop[i]), "profile sync");
        float ms = 0.f, t0 = 0.f;
        FGS_TRY(cudaEventElapsedTime(&ms, p.start[i], p.stop[i]), "profile elapsed");
        FGS_TRY(cudaEventElapsedTime(&t0, p.start[0], p.start[i]), "profile elapsed");
        const int w = std::snprintf(out + off, out_bytes - off, "%s=%.6f@%.6f\n", p.names[i], ms, t0);
        if (w < 0 || (size_t)w >= out_bytes - off) break;
        off += (size_t)w;
    }
    p.n = 0;
    return 0;
}

size_t fgs_sort_scratch_bytes(int32_t n_max, int32_t end_bit) {
    return sizeof(uint32_t) * (4 * kRadixMax + radix_sort_lookback_words(n_max, end_bit) + 8) + 512;
}

int fgs_sort_pairs_u32(uint32_t* keys, uint32_t* keys_alt, uint32_t* vals, uint32_t* vals_alt, const uint32_t* n_dev,
                       int32_t n_max, int32_t end_bit, void* scratch, size_t scratch_bytes, int32_t* result_in_alt,
                       void* stream) {
    return sort_pairs_entry<uint32_t>(keys, keys_alt, vals, vals_alt, n_dev, n_max, end_bit, scratch, scratch_bytes,
                                      result_in_alt, stream);
}

int fgs_sort_pairs_u16(uint16_t* keys, uint16_t* keys_alt, uint32_t* vals, uint32_t* vals_alt, const uint32_t* n_dev,
                       int32_t n_max, int32_t end_bit, void* scratch, size_t scratch_bytes, int32_t* result_in_alt,
                       void* stream) {
    return sort_pairs_entry<uint16_t>(keys, keys_alt, vals, vals_alt, n_dev, n_max, end_bit, scratch, scratch_bytes,
                                      result_in_alt, stream);
}

}  // extern "C"

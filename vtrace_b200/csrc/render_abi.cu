// render_abi.cu — the C ABI of librender (include/vtrace_abi.h): the reference's seven FFI
// symbols (src/render.rs:110-128; lib/entry.c, lib/memory.c) implemented on CUDA, plus the
// vt_* headless extensions.  Global singleton state like the reference's `renderer glbl`
// (lib/entry.c:21).  No CPU fallback: every path either runs the CUDA kernels or fails.
#include "../../include/vtrace_abi.h"
#include "kernels.h"

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>

namespace {

using namespace vt;

constexpr uint32_t kMaxTextures = 65536; // MAX_TEXTURES, lib/common.h:35
constexpr size_t kSmemMaskBudget = 160 * 1024; // masks up to this size are staged in shared memory
constexpr uint32_t kBinMinInstances = 16;       // below this the per-pixel loop over all instances is cheaper than binning

struct State {
    bool inited = false;
    int device = 0;
    int sm_count = 148;
    int max_smem_optin = 48 * 1024;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    vt_config cfg{};

    // volumes ("textures")
    std::vector<VolumeDesc> vols;
    VolumeDesc* d_vols = nullptr;
    size_t d_vols_cap = 0;
    bool vols_dirty = false;
    uint32_t* d_arena = nullptr; // all stop masks, contiguous
    uint32_t arena_words = 0, arena_cap = 0;
    struct BrickAlloc { uint32_t* codes; uint32_t* base; uint32_t* pool; float* heights; uchar4* colors; BrickVolume* d_desc; };
    std::vector<BrickAlloc> brick_allocs; // procedural volumes (extension)
    bool any_bricks = false;
    uint32_t max_idx_bits = 0; // widest stop-mask index of any dense volume (the wavefront kernel packs it into 30 bits)
    uint8_t* h_tex_staging = nullptr; // pinned; grows to the next power of two (lib/memory.c:297-302)
    size_t tex_staging_size = 0;

    // instances
    // pinned staging the engine writes into (lib/memory.c:245-247): a ring of kInstRing buffers, so that the host can fill the
    // next frame's instances while earlier uploads are still queued (no host wait unless it runs kInstRing frames ahead)
    static constexpr int kInstRing = 4;
    float* h_inst = nullptr;       // kInstRing x inst_cap x 16 floats
    int inst_slot = 0;             // the buffer handed out last
    const float* inst_src = nullptr; // != nullptr: the instance-setup kernel reads the staging buffer itself (few instances)
    cudaEvent_t ev_inst[kInstRing] = {nullptr, nullptr, nullptr, nullptr};
    bool inst_busy[kInstRing] = {false, false, false, false};
    float* d_inst = nullptr;
    InstUniforms* d_iu = nullptr;
    uint32_t inst_cap = 0, inst_count = 1;

    // screen-space instance bins (used when the scene has more than kBinMinInstances instances)
    uint32_t* d_bin_offset = nullptr;
    uint32_t* d_bin_count = nullptr;
    uint32_t* d_bin_list = nullptr;
    uint32_t* d_bin_cursor = nullptr;
    uint32_t* h_bin_cursor = nullptr; // pinned
    uint32_t bin_cap_bins = 0, bin_cap_list = 0;
    bool bins_used = false;
    // world-space instance grid for the path tracer's bounce rays (world_grid.cuh)
    float* d_world_aabb = nullptr;
    uint32_t world_aabb_cap = 0;
    WorldGrid* d_world_hdr = nullptr;
    WorldGrid* h_world_hdr = nullptr; // pinned copy of the last frame's header
    uint32_t* d_world_cells = nullptr; // offset | count | cursor, kWorldGridMaxCells each
    uint32_t* d_world_list = nullptr;
    uint32_t world_list_cap = 0;
    bool world_used = false;

    // tables
    uint32_t clear_rgba = 0;
    float* d_dec = nullptr;
    float* d_thr = nullptr;

    // framebuffer
    uint32_t fb_w = 0, fb_h = 0;
    HitRecord* d_rec = nullptr;
    uchar4* d_color = nullptr;
    float* d_depth = nullptr;
    unsigned long long* d_accum = nullptr;     // in use (own or caller-provided)
    unsigned long long* d_accum_own = nullptr; // the library's allocation
    // fused cross-GPU accumulation (vt_fused_reduce_*): 0 off, 1 root (owns the double buffer), 2 peer (maps it)
    int fused_mode = 0;
    uint4* fused_base = nullptr;     // root memory: [2 buffers][world ranks][pixels][2] partial sums (32 B per pixel)
    size_t fused_pixels = 0;         // pixels per slot
    uint32_t fused_index = 0, fused_rank = 0, fused_world = 1;
    // flag synchronisation of the fused accumulation (kernels.cu, flag_*_kernel): frame sequence number, the
    // flags in the root's memory (arrive[2][kFusedMaxWorld], consumed[2]), a local error word
    uint32_t fused_seq = 0;
    uint32_t* fused_flags = nullptr;
    uint32_t* d_fused_err = nullptr;
    uint32_t* h_fused_err = nullptr; // pinned
    bool fused_sync = true;
    bool fused_rows = false;         // vt_fused_reduce_partition(1, ...): the frame is shared out by rows of tiles, not by samples
    uint32_t fused_relief = 0, fused_relief_den = 8; // ... of which the root owns (den - relief) for every den of another rank (RowShare)
    // root: the last frame's own sums are still in d_accum_own (the root does not push; its first vt_resolve / vt_read_accum
    // of the frame takes them from there) and its counters in d_stats (published by the same kernel), ring slot below
    bool fused_root_live = false;
    uint32_t fused_root_slot = 0;
    uint32_t fused_seq_rendered = 0; // sequence number of the last frame rendered: every frame needs its own (vt_fused_reduce_next_frame)
    unsigned long long* fused_sum = nullptr; // root, on demand: materialised sums for vt_read_accum
    // asynchronous colour read-back (vt_read_color_async): a second colour buffer, a copy stream, one event per buffer
    uchar4* d_color_alt = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_color_ready = nullptr, ev_copy_done[2] = {nullptr, nullptr}; // [0] belongs to d_color, [1] to d_color_alt
    bool copy_pending[2] = {false, false};
    // A read-back is not started when it is asked for but when the NEXT frame's trace kernel starts (or when somebody waits
    // for it): while a device->host copy saturates PCIe, the GPU's command fetches queue behind it, and the small kernels at
    // the start of a frame (set-up, clear) were measured 30-70 us slower.  Under the long trace kernel nothing is fetched.
    struct { bool armed = false; uint8_t* dst = nullptr; const uchar4* src = nullptr; size_t bytes = 0; cudaEvent_t done = nullptr; } deferred_copy;
    bool defer_readback = true; // VT_DEFER_READBACK
    uint32_t sky_missing_spp = 0;            // != 0: the last frame left the accumulators outside the screen rectangle untouched (spp x sky missing)
    unsigned long long* d_stats = nullptr;
    unsigned long long* h_stats = nullptr; // pinned: kRing slots of 4 counters (rays, iterations, work-claim counter, analytic rays)
    void* h_readback = nullptr;            // pinned staging for vt_read_*
    size_t readback_size = 0;

    // Frames can be enqueued back to back without a host synchronisation (vt_render_async); their
    // timing events and counters live in a ring and are folded into `stats` at the next synchronisation.
    static constexpr uint32_t kRing = 64;
    cudaEvent_t ev_begin[kRing] = {}, ev_trace0[kRing] = {}, ev_trace1[kRing] = {}, ev_end[kRing] = {};
    uint64_t ring_head = 0, ring_done = 0; // frames enqueued / frames folded into stats
    bool frame_pending = false;
    uint32_t refill_threshold = 12; // tuning knob of the wavefront / persistent-lane kernels (VT_REFILL)
    uint32_t refill_batch = 10;     // wavefront kernel: stopped lanes wait until this many can be refilled together (VT_REFILL_BATCH)
    uint32_t item_spp = 16;         // wavefront kernel: most samples per work item (VT_ITEM_SPP)
    uint32_t items_per_warp = 6;    // wavefront kernel: work items wanted per resident warp (VT_ITEMS_PER_WARP)
    uint32_t item_order = 0;        // wavefront kernel: order of the work items (VT_ITEM_ORDER, FrameParams::item_order)

    vt_stats stats{};
    user_input input{};
    char err[512] = {0};
};

State g;

int fail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g.err, sizeof g.err, fmt, ap);
    va_end(ap);
    fprintf(stderr, "ERROR: %s\n", g.err); // the reference logs the same way, e.g. lib/memory.c:288
    return -1;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            (void)cudaGetLastError(); /* a recoverable failure must not poison the next launch's cudaGetLastError() */ \
            return fail("%s -> %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);    \
        }                                                                                          \
    } while (0)

uint32_t env_u32(const char* name, uint32_t dflt) {
    const char* v = getenv(name);
    return v && *v ? (uint32_t)strtoul(v, nullptr, 0) : dflt;
}

uint32_t ceil_log2(uint32_t x) {
    uint32_t b = 0;
    while ((1u << b) < x) ++b;
    return b;
}

size_t round_up_p2(size_t x) { // lib/memory.c round_up_p2
    size_t p = 1;
    while (p < x) p <<= 1;
    return p;
}

// host-side mat4 helpers, same operation order as the oracle / the device code
void mat4_mul(const float* a, const float* b, float* r) {
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i)
            r[j * 4 + i] = ((a[0 * 4 + i] * b[j * 4 + 0] + a[1 * 4 + i] * b[j * 4 + 1]) + a[2 * 4 + i] * b[j * 4 + 2]) + a[3 * 4 + i] * b[j * 4 + 3];
}

double srgb_to_linear(double c) { return c <= 0.04045 ? c / 12.92 : pow((c + 0.055) / 1.055, 2.4); }

int issue_deferred_copy(cudaEvent_t after);

int alloc_framebuffer() {
    const uint32_t w = g.cfg.width, h = g.cfg.height;
    if (w == g.fb_w && h == g.fb_h && g.d_color) return 0;
    if (issue_deferred_copy(nullptr)) return -1; // (a read-back of the old buffers that no frame has started yet)
    CK(cudaStreamSynchronize(g.stream));
    if (g.copy_stream) CK(cudaStreamSynchronize(g.copy_stream));
    cudaFree(g.d_rec); cudaFree(g.d_color); cudaFree(g.d_color_alt); cudaFree(g.d_depth); cudaFree(g.d_accum_own);
    g.d_rec = nullptr; g.d_color = nullptr; g.d_color_alt = nullptr; g.d_depth = nullptr; g.d_accum = nullptr; g.d_accum_own = nullptr;
    g.copy_pending[0] = g.copy_pending[1] = false;
    const size_t n = (size_t)w * h;
    CK(cudaMalloc(&g.d_rec, n * sizeof(HitRecord)));
    CK(cudaMalloc(&g.d_color, n * sizeof(uchar4)));
    CK(cudaMalloc(&g.d_depth, n * sizeof(float)));
    CK(cudaMalloc(&g.d_accum_own, n * 3 * sizeof(unsigned long long)));
    g.d_accum = g.d_accum_own; // a caller-provided buffer does not survive a resize
    CK(cudaMemsetAsync(g.d_rec, 0xFF, n * sizeof(HitRecord), g.stream));
    CK(cudaMemsetAsync(g.d_color, 0, n * sizeof(uchar4), g.stream));
    CK(cudaMemsetAsync(g.d_depth, 0, n * sizeof(float), g.stream));
    CK(cudaMemsetAsync(g.d_accum, 0, n * 3 * sizeof(unsigned long long), g.stream));
    g.sky_missing_spp = 0;
    g.fb_w = w; g.fb_h = h;
    return 0;
}

int ensure_readback(size_t bytes) {
    if (bytes <= g.readback_size) return 0;
    if (g.h_readback) cudaFreeHost(g.h_readback);
    g.h_readback = nullptr;
    g.readback_size = 0;
    CK(cudaMallocHost(&g.h_readback, bytes));
    g.readback_size = bytes;
    return 0;
}

// up to this many instances are read by the setup kernel straight from the staging buffer (VT_DIRECT_INSTANCES)
static uint32_t direct_instances() { static const uint32_t n = env_u32("VT_DIRECT_INSTANCES", 64); return n; }

float* inst_staging(int slot) { return g.h_inst + (size_t)slot * g.inst_cap * 16; }

int ensure_instances(uint32_t n) {
    if (n <= g.inst_cap) return 0;
    // lib/memory.c:239-243: the instance buffers are re-created at the next power of two
    const uint32_t cap = (uint32_t)round_up_p2(n);
    CK(cudaStreamSynchronize(g.stream));
    float* h_new = nullptr;
    CK(cudaMallocHost(&h_new, (size_t)State::kInstRing * cap * 64));
    memset(h_new, 0, (size_t)State::kInstRing * cap * 64);
    if (g.h_inst) {
        memcpy(h_new, inst_staging(g.inst_slot), (size_t)g.inst_cap * 64); // the current contents survive, in slot 0
        cudaFreeHost(g.h_inst);
    }
    g.h_inst = h_new;
    g.inst_slot = 0;
    for (int i = 0; i < State::kInstRing; ++i) {
        g.inst_busy[i] = false;
        if (!g.ev_inst[i]) CK(cudaEventCreateWithFlags(&g.ev_inst[i], cudaEventDisableTiming));
    }
    cudaFree(g.d_inst); cudaFree(g.d_iu);
    g.d_inst = nullptr; g.d_iu = nullptr;
    CK(cudaMalloc(&g.d_inst, (size_t)cap * 64));
    CK(cudaMalloc(&g.d_iu, (size_t)cap * sizeof(InstUniforms)));
    g.inst_cap = cap;
    CK(cudaMemcpyAsync(g.d_inst, g.h_inst, (size_t)cap * 64, cudaMemcpyHostToDevice, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    if (g.inst_src) g.inst_src = inst_staging(0);
    return 0;
}

int render_async(const float* P, const float* V, bool clear_accum, bool resolve);

// starts the read-back that vt_read_color_async armed; `after` (optional): not before this event of the render stream
int issue_deferred_copy(cudaEvent_t after) {
    if (!g.deferred_copy.armed) return 0;
    g.deferred_copy.armed = false;
    CK(cudaStreamWaitEvent(g.copy_stream, g.ev_color_ready, 0)); // its frame is complete
    if (after) CK(cudaStreamWaitEvent(g.copy_stream, after, 0));
    CK(cudaMemcpyAsync(g.deferred_copy.dst, g.deferred_copy.src, g.deferred_copy.bytes, cudaMemcpyDeviceToHost, g.copy_stream));
    CK(cudaEventRecord(g.deferred_copy.done, g.copy_stream));
    return 0;
}

// Called before anything writes the colour buffer: a buffer whose asynchronous read-back (vt_read_color_async) may still be
// in flight is left alone — the frame goes to the other one, after (device-side) waiting for that one's own read-back.
int prepare_color_target() {
    if (!g.copy_pending[0]) return 0;
    if (!g.d_color_alt) CK(cudaMalloc(&g.d_color_alt, (size_t)g.cfg.width * g.cfg.height * sizeof(uchar4)));
    std::swap(g.d_color, g.d_color_alt);
    std::swap(g.ev_copy_done[0], g.ev_copy_done[1]);
    std::swap(g.copy_pending[0], g.copy_pending[1]);
    if (g.copy_pending[0]) {
        CK(cudaStreamWaitEvent(g.stream, g.ev_copy_done[0], 0));
        g.copy_pending[0] = false;
    }
    return 0;
}

// After a lean frame (render_async) the accumulators outside the instance's screen rectangle hold stale data: write
// spp x sky there before anybody reads or adds to them.
int complete_accum() {
    if (!g.sky_missing_spp) return 0;
    SrgbTables lut{g.d_dec, g.d_thr};
    CK(launch_resolve_rect(g.d_iu, g.d_accum, g.cfg.width, g.cfg.height, g.sky_missing_spp, 1u, lut, g.d_color, true, nullptr, nullptr, g.sm_count, g.stream));
    g.sky_missing_spp = 0;
    return 0;
}

static constexpr uint32_t kFusedMaxWorld = 64;
static uint32_t* fused_arrive(uint32_t half) { return g.fused_flags + half * kFusedMaxWorld; }
static uint32_t* fused_consumed(uint32_t half) { return g.fused_flags + 2 * kFusedMaxWorld + half; }
// partial sums travel as 3 x u32 when no rank can exceed 32 bits per channel: every rank renders at most
// total_spp samples of at most 2^24 each (total_spp is the same on all ranks, so they agree on the layout)
// the root's summation first waits for every rank's "sums of this frame are in place" flag
static FusedSync fused_wait_all() {
    FusedSync fs{};
    if (g.fused_sync && g.fused_world > 1) { // (the root's own sums do not travel: ranks 1 .. world - 1)
        fs.wait_flags = fused_arrive(g.fused_index) + 1; fs.wait_count = g.fused_world - 1; fs.wait_target = g.fused_seq;
        fs.err = g.d_fused_err;
    }
    return fs;
}
// rows of tiles per rank (vt_fused_reduce_partition); world 1 = every rank holds every row (frames shared by samples)
static RowShare fused_row_share() {
    RowShare rs{0u, 1u, g.fused_relief_den, 0u};
    if (g.fused_rows) {
        rs.rank = g.fused_rank; rs.world = g.fused_world; rs.c = g.fused_world > 1 ? g.fused_relief : 0u;
        // (profiling aid: one rank of an N-rank job on its own — VT_FUSED_ROWS_AS_WORLD=N traces and moves rank 0's rows only)
        static const uint32_t as_world = env_u32("VT_FUSED_ROWS_AS_WORLD", 0);
        if (as_world && g.fused_world == 1) { rs.world = as_world; rs.c = g.fused_relief; }
    }
    return rs;
}
static bool fused_compact() { return (g.cfg.total_spp ? g.cfg.total_spp : g.cfg.spp) <= 255u; }
// The root's summation (vt_resolve / vt_read_accum).  The first one of a frame also collects the root's own sums from its
// local accumulators and publishes the frame's counters, so the frame's end event is recorded again behind it.
static cudaError_t fused_root_sum(unsigned long long* accum_out) {
    const uint32_t total = g.cfg.total_spp ? g.cfg.total_spp : g.cfg.spp;
    SrgbTables lut{g.d_dec, g.d_thr};
    uint4* half = g.fused_base + (size_t)g.fused_index * g.fused_world * g.fused_pixels * 2;
    FusedSync fs = fused_wait_all();
    if (g.fused_root_live) { fs.stats = g.d_stats; fs.host_stats = g.h_stats + 4 * g.fused_root_slot; }
    cudaError_t e = launch_resolve_partials(g.d_iu, half, g.fused_world, g.cfg.width, g.cfg.height, total ? total : 1, lut, g.d_color, accum_out,
                                            fused_compact(), fused_row_share(), g.fused_root_live ? g.d_accum_own : nullptr, half,
                                            fs, g.sm_count, g.stream);
    if (e == cudaSuccess && g.fused_root_live) {
        g.fused_root_live = false;
        e = cudaEventRecord(g.ev_end[g.fused_root_slot], g.stream);
    }
    return e;
}

int finish_frame() {
    if (!g.frame_pending) return 0;
    CK(cudaStreamSynchronize(g.stream));
    g.frame_pending = false;
    if (g.fused_mode == 1 && g.fused_root_live) // counters of a root frame that was not summed yet (normally vt_resolve's kernel publishes them)
        CK(cudaMemcpy(g.h_stats + 4 * g.fused_root_slot, g.d_stats, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    // the binner reports how many list entries it needed.  Bins whose segment did not fit made their pixels visit
    // every instance (kernels.cu, bin_range) — exact, only slower, and free of side effects whoever owns the
    // accumulator clear — so nothing is rendered again: the next frame simply gets a larger list.
    if (g.bins_used && *g.h_bin_cursor > g.bin_cap_list) {
        const uint32_t need = *g.h_bin_cursor + *g.h_bin_cursor / 2;
        cudaFree(g.d_bin_list);
        g.d_bin_list = nullptr;
        g.bin_cap_list = 0;
        CK(cudaMalloc(&g.d_bin_list, (size_t)need * 4));
        g.bin_cap_list = need;
        g.stats.bin_list_grown += 1;
    }
    if (g.fused_mode && g.d_fused_err) { // did a flag wait give up?  (the stream is idle here: a blocking 4-byte copy)
        CK(cudaMemcpy(g.h_fused_err, g.d_fused_err, 4, cudaMemcpyDeviceToHost));
        if (*g.h_fused_err) {
            *g.h_fused_err = 0;
            CK(cudaMemset(g.d_fused_err, 0, 4));
            return fail("fused cross-GPU accumulation: timed out waiting for another rank");
        }
    }
    // the world grid only reports that its lists did not fit (that frame looped over all instances, which is
    // exact): give the next frame room
    if (g.world_used && g.h_world_hdr->total > g.world_list_cap) {
        const uint32_t need = g.h_world_hdr->total + g.h_world_hdr->total / 2;
        cudaFree(g.d_world_list);
        g.d_world_list = nullptr;
        g.world_list_cap = 0;
        CK(cudaMalloc(&g.d_world_list, (size_t)need * 4));
        g.world_list_cap = need;
    }
    for (; g.ring_done < g.ring_head; ++g.ring_done) {
        const uint32_t slot = (uint32_t)(g.ring_done % State::kRing);
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, g.ev_trace0[slot], g.ev_trace1[slot]) == cudaSuccess) {
            g.stats.last_trace_ms = ms;
            g.stats.trace_ms_sum += ms;
            g.stats.trace_frames += 1;
        }
        if (cudaEventElapsedTime(&ms, g.ev_begin[slot], g.ev_end[slot]) == cudaSuccess) g.stats.last_frame_ms = ms;
        const unsigned long long* hs = g.h_stats + 4 * slot;
        if (g.cfg.mode == VT_MODE_PRIMARY || g.cfg.mode == VT_MODE_RAYS) {
            g.stats.rays = (uint64_t)g.cfg.width * g.cfg.height + hs[0]; // + shadow rays
            g.stats.iterations = hs[1];
            g.stats.analytic_rays = 0;
        } else {
            g.stats.rays = hs[0];
            g.stats.iterations = hs[1];
            g.stats.analytic_rays = hs[3];
        }
        g.stats.rays_sum += g.stats.rays;
        g.stats.iterations_sum += g.stats.iterations;
        g.stats.analytic_rays_sum += g.stats.analytic_rays;
    }
    return 0;
}

int render_async(const float* P, const float* V, bool clear_accum, bool resolve) {
    if (!g.inited) return fail("render before entry()");
    CK(cudaSetDevice(g.device));
    if (alloc_framebuffer()) return -1;
    // A frame that used the instance bins must be checked for list overflow before the next one; otherwise
    // frames may queue up (their events/counters live in a ring of kRing slots).
    if (g.frame_pending && (g.bins_used || g.ring_head - g.ring_done >= State::kRing) && finish_frame()) return -1;
    const uint32_t slot = (uint32_t)(g.ring_head % State::kRing);

    FrameParams fp{};
    float Pi[16], Vi[16], Vc[16], Vci[16];
    mat4_inverse(P, Pi);                 // trace.frag:48
    mat4_inverse(V, Vi);                 // trace.frag:49
    memcpy(Vc, V, sizeof Vc);            // trace.frag:54-57
    Vc[12] = 0.0f; Vc[13] = 0.0f; Vc[14] = 0.0f;
    mat4_inverse(Vc, Vci);               // trace.frag:59
    mat4_mul(Vci, Pi, fp.RD);            // (inverse(Vc) * inverse(P)) * sp
    mat4_mul(P, V, fp.PV);               // trace.vert:45
    const float origin[4] = {0.0f, 0.0f, 0.0f, 1.0f};
    for (int i = 0; i < 3; ++i)          // trace.frag:51 cam_pos
        fp.eye[i] = ((Vi[0 * 4 + i] * origin[0] + Vi[1 * 4 + i] * origin[1]) + Vi[2 * 4 + i] * origin[2]) + Vi[3 * 4 + i] * origin[3];
    fp.width = (int32_t)g.cfg.width;
    fp.height = (int32_t)g.cfg.height;
    fp.vw = (float)g.cfg.width;                                                              // lib/command.c:80
    fp.vh = (g.cfg.flags & VT_FLAG_VIEWPORT_H_IS_W) ? (float)g.cfg.width : (float)g.cfg.height; // lib/command.c:81
    fp.sxn = 2.0f / fp.vw;
    fp.syn = 2.0f / fp.vh;
    fp.n_inst = g.inst_count;
    fp.n_volumes = (uint32_t)g.vols.size();
    fp.flags = g.cfg.flags;
    fp.spp = g.cfg.spp;
    fp.bounces = g.cfg.bounces;
    fp.seed = g.cfg.seed;
    fp.sample_first = g.cfg.sample_first;
    fp.sample_stride = g.cfg.sample_stride ? g.cfg.sample_stride : 1;
    fp.refill_threshold = g.refill_threshold;
    fp.refill_batch = g.refill_batch;
    fp.item_spp = g.item_spp;
    fp.items_per_warp = g.items_per_warp;
    fp.item_order = g.item_order;
    {   // sun direction (0.4, -0.8, 0.45), normalised (same float operations as the oracle)
        const float sx = 0.4f, sy = -0.8f, sz = 0.45f;
        const float l = sqrtf((sx * sx + sy * sy) + sz * sz);
        fp.sun[0] = sx / l; fp.sun[1] = sy / l; fp.sun[2] = sz / l;
    }
    fp.any_bricks = g.any_bricks ? 1u : 0u;
    fp.max_idx_bits = g.max_idx_bits;
    fp.clear_rgba = g.clear_rgba;
    fp.sky_spp = g.cfg.spp;
    if (g.fused_mode && g.cfg.mode == VT_MODE_PATHS) {
        if (g.inst_count != 1 || g.any_bricks || (g.cfg.flags & VT_FLAG_PER_PIXEL_PATHS) || g.max_idx_bits > 29)
            return fail("fused cross-GPU accumulation needs the single-instance wavefront kernel");
        if (g.fused_pixels != (size_t)g.cfg.width * g.cfg.height) return fail("fused accumulation buffer does not match the framebuffer size");
        if (g.fused_mode == 1 && g.fused_root_live)
            return fail("fused cross-GPU accumulation: the root must call vt_resolve (or vt_read_accum) for every frame before it starts the next");
        if (g.fused_seq != 0 && g.fused_seq == g.fused_seq_rendered) // (the flags of this sequence number are already up: the root would not wait)
            return fail("fused cross-GPU accumulation: vt_fused_reduce_next_frame must be called before every frame");
        fp.sky_spp = 0u; // pixels outside the screen rectangle are resolved analytically on the root
        if (g.fused_rows) {
            if (g.cfg.sample_first != 0 || g.cfg.sample_stride > 1 || (g.cfg.total_spp && g.cfg.total_spp != g.cfg.spp))
                return fail("fused accumulation by tile rows: every rank traces all samples (sample_first 0, sample_stride 1, spp = total_spp)");
            fp.rows = fused_row_share();
        }
    }

    if (g.vols_dirty) {
        if (g.vols.size() > g.d_vols_cap) {
            cudaFree(g.d_vols);
            g.d_vols = nullptr;
            g.d_vols_cap = round_up_p2(g.vols.size());
            CK(cudaMalloc(&g.d_vols, g.d_vols_cap * sizeof(VolumeDesc)));
        }
        CK(cudaMemcpyAsync(g.d_vols, g.vols.data(), g.vols.size() * sizeof(VolumeDesc), cudaMemcpyHostToDevice, g.stream));
        CK(cudaStreamSynchronize(g.stream)); // g.vols may reallocate before the copy would run
        g.vols_dirty = false;
    }
    if (ensure_instances(g.inst_count)) return -1;

    // render_tick on a single-instance scene: nothing outside the instance's screen rectangle is cleared, added to or
    // read (kernels.cu, clear_rect / resolve_rect); the sums there are materialised on demand (complete_accum) — at the
    // latest before another kind of frame replaces the instance uniforms the rectangle lives in.
    static const bool lean_enabled = env_u32("VT_LEAN_FRAME", 1) != 0;
    const bool lean = g.cfg.mode == VT_MODE_PATHS && clear_accum && resolve && !g.fused_mode && g.d_accum == g.d_accum_own &&
                      paths_use_wave_kernel(fp) && lean_enabled;
    if (lean) fp.sky_spp = 0u;
    else if (complete_accum()) return -1;

    CK(cudaEventRecord(g.ev_begin[slot], g.stream));
    uint32_t* consumed_flag = nullptr;
    uint32_t consumed_value = 0;
    if (g.fused_mode && g.cfg.mode == VT_MODE_PATHS) {
        if (g.fused_seq == 0) { g.fused_seq = 1; g.fused_index = 1; } // (vt_fused_reduce_next_frame was not called yet)
        g.fused_seq_rendered = g.fused_seq;
        // the root is done with the previous frame once it starts this one: its half may be refilled
        // (raised by the instance-setup kernel below, the first kernel of the frame)
        if (g.fused_sync && g.fused_mode == 1 && g.fused_seq > 1) {
            consumed_flag = fused_consumed((g.fused_seq - 1) & 1u);
            consumed_value = g.fused_seq - 1;
        }
    }
    // (the setup kernel also zeroes the frame's counters; a few instances are read straight from the page-locked staging
    // buffer the engine wrote them into — neither costs the frame a copy-engine operation)
    CK(launch_instance_setup(g.inst_src ? g.inst_src : g.d_inst, g.inst_count, g.d_vols, fp, g.d_iu, consumed_flag, consumed_value,
                             g.d_stats, g.stream));
    if (g.inst_src) CK(cudaEventRecord(g.ev_inst[g.inst_slot], g.stream)); // the staging buffer is in use until here
    g.stats.launches += 1;

    // many instances: bin their screen rectangles (16x16-pixel bins) so a pixel only visits its own
    BinTable bins{};
    g.bins_used = g.inst_count > kBinMinInstances && !(g.cfg.flags & VT_FLAG_NO_BINNING);
    if (g.bins_used) {
        const uint32_t bx = (g.cfg.width + (1u << kBinShift) - 1) >> kBinShift, by = (g.cfg.height + (1u << kBinShift) - 1) >> kBinShift;
        if (bx * by > g.bin_cap_bins) {
            CK(cudaStreamSynchronize(g.stream));
            cudaFree(g.d_bin_offset); cudaFree(g.d_bin_count);
            g.d_bin_offset = g.d_bin_count = nullptr;
            CK(cudaMalloc(&g.d_bin_offset, (size_t)bx * by * 4));
            CK(cudaMalloc(&g.d_bin_count, (size_t)bx * by * 4));
            g.bin_cap_bins = bx * by;
        }
        if (!g.d_bin_list) {
            size_t cap = (size_t)g.inst_count * 64 > (1u << 20) ? (size_t)g.inst_count * 64 : (1u << 20);
            cap = env_u32("VT_BIN_CAP", (uint32_t)cap); // (tests shrink it to exercise the grow-and-retry path)
            CK(cudaMalloc(&g.d_bin_list, cap * 4));
            g.bin_cap_list = (uint32_t)cap;
        }
        CK(cudaMemsetAsync(g.d_bin_cursor, 0, 4, g.stream));
        CK(launch_bin_instances(g.d_iu, g.inst_count, bx, by, g.d_bin_offset, g.d_bin_count, g.d_bin_list, g.bin_cap_list,
                                g.d_bin_cursor, g.stream));
        CK(cudaMemcpyAsync(g.h_bin_cursor, g.d_bin_cursor, 4, cudaMemcpyDeviceToHost, g.stream));
        g.stats.launches += 1;
        bins.offset = g.d_bin_offset; bins.count = g.d_bin_count; bins.list = g.d_bin_list;
        bins.bins_x = bx; bins.bins_y = by; bins.enabled = 1;
    }

    // path tracing over many instances: a world-space grid so that bounce rays only test the instances near them
    WorldGridTable wg{};
    g.world_used = g.cfg.mode == VT_MODE_PATHS && g.inst_count >= kWorldGridMinInstances && !(g.cfg.flags & VT_FLAG_NO_BINNING);
    if (g.world_used) {
        if (g.inst_count > g.world_aabb_cap) {
            CK(cudaStreamSynchronize(g.stream));
            cudaFree(g.d_world_aabb);
            g.d_world_aabb = nullptr;
            g.world_aabb_cap = 0;
            CK(cudaMalloc(&g.d_world_aabb, (size_t)g.inst_count * 6 * sizeof(float)));
            g.world_aabb_cap = g.inst_count;
        }
        if (!g.d_world_cells) CK(cudaMalloc(&g.d_world_cells, (size_t)kWorldGridMaxCells * 3 * 4));
        if (!g.d_world_list) {
            const uint32_t cap = env_u32("VT_WORLD_CAP", g.inst_count * 16 > (1u << 16) ? g.inst_count * 16 : (1u << 16));
            CK(cudaMalloc(&g.d_world_list, (size_t)cap * 4));
            g.world_list_cap = cap;
        }
        uint32_t* offset = g.d_world_cells;
        uint32_t* count = g.d_world_cells + kWorldGridMaxCells;
        uint32_t* cursor = g.d_world_cells + 2 * kWorldGridMaxCells;
        CK(launch_world_grid(g.d_iu, g.inst_count, g.d_world_aabb, g.d_world_hdr, offset, count, cursor, g.d_world_list,
                             g.world_list_cap, g.stream));
        CK(cudaMemcpyAsync(g.h_world_hdr, g.d_world_hdr, sizeof(WorldGrid), cudaMemcpyDeviceToHost, g.stream));
        g.stats.launches += 4;
        wg.hdr = g.d_world_hdr; wg.offset = offset; wg.count = count; wg.list = g.d_world_list; wg.enabled = 1;
    }

    const bool in_smem = !(g.cfg.flags & VT_FLAG_FORCE_GLOBAL_MASKS) && g.arena_words > 0 &&
                         (size_t)g.arena_words * 4 <= kSmemMaskBudget &&
                         trace_smem_bytes(g.arena_words, true) <= (size_t)g.max_smem_optin;
    g.stats.masks_in_smem = in_smem ? 1u : 0u;
    FrameBuffers fb{};
    fb.records = (g.cfg.flags & VT_FLAG_NO_HIT_RECORDS) ? nullptr : g.d_rec;
    if (prepare_color_target()) return -1;
    fb.color = g.d_color;
    fb.depth = g.d_depth;
    const bool fused = g.fused_mode && g.cfg.mode == VT_MODE_PATHS;
    fb.accum = fused ? g.d_accum_own : g.d_accum;
    fb.stats = g.d_stats;
    SrgbTables lut{g.d_dec, g.d_thr};

    if (g.cfg.mode == VT_MODE_RAYS) {
        const unsigned long long n = (unsigned long long)g.cfg.width * g.cfg.height;
        CK(cudaEventRecord(g.ev_trace0[slot], g.stream));
        if (issue_deferred_copy(g.ev_trace0[slot])) return -1;
        CK(launch_trace_rays(fp, g.d_iu, g.d_arena, n, (unsigned long long)g.cfg.sample_first * n, fb, g.sm_count, g.stream));
        CK(cudaEventRecord(g.ev_trace1[slot], g.stream));
        g.stats.launches += 1;
    } else if (g.cfg.mode == VT_MODE_PRIMARY) {
        CK(cudaEventRecord(g.ev_trace0[slot], g.stream));
        if (issue_deferred_copy(g.ev_trace0[slot])) return -1;
        CK(launch_trace_primary(fp, g.d_iu, bins, g.d_arena, g.arena_words, in_smem, lut, fb, g.sm_count, g.stream));
        CK(cudaEventRecord(g.ev_trace1[slot], g.stream));
        g.stats.launches += 1;
    } else {
        if (lean) {
            CK(launch_clear_rect(g.d_iu, g.d_accum, g.cfg.width, g.cfg.height, g.sm_count, g.stream));
            g.stats.launches += 1;
            g.sky_missing_spp = g.cfg.spp;
        } else if (clear_accum) {
            CK(cudaMemsetAsync(g.d_accum, 0, (size_t)g.cfg.width * g.cfg.height * 3 * sizeof(unsigned long long), g.stream));
        }
        CK(cudaEventRecord(g.ev_trace0[slot], g.stream));
        if (issue_deferred_copy(g.ev_trace0[slot])) return -1;
        CK(launch_trace_paths(fp, g.d_iu, bins, wg, g.d_arena, g.arena_words, in_smem, lut, fb, g.sm_count, g.stream));
        CK(cudaEventRecord(g.ev_trace1[slot], g.stream));
        g.stats.launches += 1;
        if (fused && g.fused_mode == 1) { // the root's sums stay where they are until its summation kernel collects them
            g.fused_root_live = true;
            g.fused_root_slot = slot;
        } else if (fused) { // stream this rank's sums of the covered rectangle into its slot in the root's memory
            const uint32_t half = g.fused_index, seq = g.fused_seq;
            uint4* peer_slot = g.fused_base + ((size_t)half * g.fused_world + g.fused_rank) * g.fused_pixels * 2;
            FusedSync fs{};
            if (g.fused_sync) {
                if (g.fused_rank != 0 && seq > 2) { // the root must be done with the frame that used this half before
                    fs.wait_flags = fused_consumed(half); fs.wait_count = 1; fs.wait_target = seq - 2;
                }
                fs.signal_flag = fused_arrive(half) + g.fused_rank; fs.signal_value = seq; // "my sums of frame seq are in place"
                fs.done_counter = g.d_fused_err + 1;
                fs.err = g.d_fused_err;
            }
            fs.stats = g.d_stats; fs.host_stats = g.h_stats + 4 * slot; // (the frame's counters travel with this kernel)
            CK(launch_push_partial(g.d_iu, g.d_accum_own, peer_slot, g.cfg.width, g.cfg.height, fused_compact(), fused_row_share(), fs,
                                   g.sm_count, g.stream));
            g.stats.launches += 1;
        }
        if (resolve && !fused) {
            const uint32_t total = g.cfg.total_spp ? g.cfg.total_spp : g.cfg.spp;
            if (lean) CK(launch_resolve_rect(g.d_iu, g.d_accum, g.cfg.width, g.cfg.height, g.cfg.spp, total ? total : 1, lut, g.d_color, false,
                                             g.d_stats, g.h_stats + 4 * slot, g.sm_count, g.stream));
            else CK(launch_resolve(g.d_accum, g.cfg.width * g.cfg.height, total ? total : 1, lut, g.d_color, g.stream));
            g.stats.launches += 1;
        }
    }
    if (!lean && !(g.fused_mode && g.cfg.mode == VT_MODE_PATHS))
        CK(cudaMemcpyAsync(g.h_stats + 4 * slot, g.d_stats, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, g.stream));
    CK(cudaEventRecord(g.ev_end[slot], g.stream));
    g.ring_head += 1;
    g.frame_pending = true;
    g.stats.frames += 1;
    return 0;
}

int64_t read_back(const void* d_src, size_t bytes, void* out, size_t capacity) {
    if (!g.inited) return fail("read before entry()");
    if (!out || capacity < bytes) return fail("read-back buffer too small: need %zu bytes, have %zu", bytes, capacity);
    if (cudaSetDevice(g.device) != cudaSuccess) return fail("cudaSetDevice failed");
    // a page-locked destination (cudaHostAlloc / cudaHostRegister / a pinned torch tensor) is written by the
    // copy engine directly; pageable memory goes through the library's pinned staging buffer
    cudaPointerAttributes attr{};
    const bool pinned = cudaPointerGetAttributes(&attr, out) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    if (!pinned) (void)cudaGetLastError();
    if (!pinned && ensure_readback(bytes)) return -1;
    void* dst = pinned ? out : g.h_readback;
    if (finish_frame()) return -1; // (first: whatever finishing a frame may enqueue must precede the copy)
    if (cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, g.stream) != cudaSuccess) return fail("read-back copy failed");
    if (cudaStreamSynchronize(g.stream) != cudaSuccess) return fail("read-back sync failed");
    if (!pinned) memcpy(out, g.h_readback, bytes);
    return (int64_t)bytes;
}

} // namespace

// ============================================================================================
// Part 1 — the reference's FFI

extern "C" uint64_t entry(void) {
    if (g.inited) return 0;
    int dev = (int)env_u32("VT_DEVICE", env_u32("LOCAL_RANK", 0));
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        fail("no CUDA device: %s", cudaGetErrorString(e));
        return ((uint64_t)(uint32_t)e << 32) | 1u;
    }
    if (dev >= count) { // (silently falling back to device 0 would stack several ranks on one GPU)
        fail("device %d requested (VT_DEVICE / LOCAL_RANK) but only %d CUDA device(s) are visible", dev, count);
        return 6u;
    }
    g.device = dev;
#define CKE(call)                                                                     \
    do {                                                                              \
        cudaError_t e_ = (call);                                                      \
        if (e_ != cudaSuccess) {                                                      \
            (void)cudaGetLastError();                                                 \
            fail("%s -> %s", #call, cudaGetErrorString(e_));                          \
            return ((uint64_t)(uint32_t)e_ << 32) | 2u;                               \
        }                                                                             \
    } while (0)
    CKE(cudaSetDevice(dev));
    cudaDeviceProp prop{};
    CKE(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10) {
        fail("librender is built for sm_100a only; device %d is sm_%d%d", dev, prop.major, prop.minor);
        return 3u;
    }
    g.sm_count = prop.multiProcessorCount;
    g.max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    CKE(cudaStreamCreateWithFlags(&g.own_stream, cudaStreamNonBlocking));
    g.stream = g.own_stream;
    for (uint32_t k = 0; k < State::kRing; ++k) {
        CKE(cudaEventCreate(&g.ev_begin[k]));
        CKE(cudaEventCreate(&g.ev_trace0[k]));
        CKE(cudaEventCreate(&g.ev_trace1[k]));
        CKE(cudaEventCreate(&g.ev_end[k]));
    }
    CKE(configure_kernels(g.max_smem_optin));

    // sRGB tables (VK_FORMAT_R8G8B8A8_SRGB textures, lib/memory.c:317; B8G8R8A8_SRGB target, lib/swapchain.c:88)
    float dec[256], thr[256];
    for (int k = 0; k < 256; ++k) {
        dec[k] = (float)srgb_to_linear((double)k / 255.0);
        thr[k] = k == 0 ? 0.0f : (float)srgb_to_linear(((double)k - 0.5) / 255.0);
    }
    {   // clear values (lib/command.c:56-61) pushed through the sRGB target's encoding once, here
        const float clear[3] = {53.0f / 100.0f, 81.0f / 100.0f, 92.0f / 100.0f};
        g.clear_rgba = 255u << 24;
        for (int c = 0; c < 3; ++c) {
            uint32_t k = 0;
            for (uint32_t bit = 128; bit; bit >>= 1)
                if (clear[c] >= thr[k | bit]) k |= bit;
            g.clear_rgba |= k << (8 * c);
        }
    }
    CKE(cudaMalloc(&g.d_dec, sizeof dec));
    CKE(cudaMalloc(&g.d_thr, sizeof thr));
    CKE(cudaMemcpy(g.d_dec, dec, sizeof dec, cudaMemcpyHostToDevice));
    CKE(cudaMemcpy(g.d_thr, thr, sizeof thr, cudaMemcpyHostToDevice));
    CKE(cudaMalloc(&g.d_stats, 4 * sizeof(unsigned long long)));
    CKE(cudaMalloc(&g.d_world_hdr, sizeof(WorldGrid)));
    CKE(cudaMallocHost(&g.h_world_hdr, sizeof(WorldGrid)));
    memset(g.h_world_hdr, 0, sizeof(WorldGrid));
    CKE(cudaMalloc(&g.d_bin_cursor, 4));
    CKE(cudaMallocHost(&g.h_bin_cursor, 4));
    *g.h_bin_cursor = 0;
    CKE(cudaMallocHost(&g.h_stats, 4 * State::kRing * sizeof(unsigned long long)));
    memset(g.h_stats, 0, 4 * State::kRing * sizeof(unsigned long long));

    // defaults: the reference's fixed 1000x1000 window (lib/entry.c:62), overridable from the
    // environment so the unmodified Rust engine can be configured without new calls
    g.cfg = vt_config{};
    g.cfg.width = env_u32("VT_WIDTH", 1000);
    g.cfg.height = env_u32("VT_HEIGHT", 1000);
    g.cfg.mode = env_u32("VT_MODE", VT_MODE_PRIMARY);
    g.cfg.flags = env_u32("VT_FLAGS", 0);
    g.cfg.spp = env_u32("VT_SPP", 1);
    g.cfg.bounces = env_u32("VT_BOUNCES", 4);
    g.cfg.seed = env_u32("VT_SEED", 0x5EED);
    g.cfg.sample_first = env_u32("VT_SAMPLE_FIRST", 0);
    g.cfg.sample_stride = env_u32("VT_SAMPLE_STRIDE", 1);
    g.cfg.total_spp = env_u32("VT_TOTAL_SPP", 0);
    g.cfg.max_frames = (int32_t)env_u32("VT_MAX_FRAMES", 0);
    g.cfg.device = dev;
    g.refill_threshold = env_u32("VT_REFILL", 12);
    g.refill_batch = env_u32("VT_REFILL_BATCH", 10);
    g.item_spp = env_u32("VT_ITEM_SPP", 16);
    if (g.item_spp < 1) g.item_spp = 1;
    if (g.item_spp > 255) g.item_spp = 255; // an item's per-pixel sums live in 32-bit shared counters: 255 samples of < 2^24 each
    g.defer_readback = env_u32("VT_DEFER_READBACK", 1) != 0;
    g.deferred_copy.armed = false;
    g.item_order = env_u32("VT_ITEM_ORDER", 0); // (1 and 2 measured: whole frame +6 % / +1 %; a rank's tile rows -4 % .. +8 %, box to box)
    if (g.item_order > 2u) g.item_order = 0;
    g.items_per_warp = env_u32("VT_ITEMS_PER_WARP", 6);
    if (g.items_per_warp < 1) g.items_per_warp = 1;
    if (g.refill_batch < 1) g.refill_batch = 1;
    if (g.refill_batch > 32) g.refill_batch = 32;
    if (g.refill_threshold < 1) g.refill_threshold = 1;
    if (g.refill_threshold > 32) g.refill_threshold = 32;

    g.inited = true;
    g.inst_count = 1; // lib/memory.c:236,251: an empty scene still draws one (stale, zeroed) instance
    if (ensure_instances(1)) { g.inited = false; return 4u; }
    if (alloc_framebuffer()) { g.inited = false; return 5u; }
    CKE(cudaStreamSynchronize(g.stream));
#undef CKE
    return 0;
}

extern "C" int32_t render_tick(int32_t* window_width, int32_t* window_height, const render_tick_info* info) {
    if (!g.inited) return -1;
    // lib/entry.c:151-153: the window-close check comes first; headless = frame budget
    if (g.cfg.max_frames > 0 && g.stats.frames >= (uint64_t)g.cfg.max_frames) return -1;
    if (!info || !info->perspective || !info->camera) return -1;
    if (render_async((const float*)info->perspective, (const float*)info->camera, true, true)) return -1;
    if (finish_frame()) return -1;
    if (window_width) *window_width = (int32_t)g.cfg.width;   // lib/entry.c:244
    if (window_height) *window_height = (int32_t)g.cfg.height; // lib/entry.c:245
    return 0;
}

extern "C" user_input* get_input_data_pointer(void) { return &g.input; }

extern "C" int32_t add_texture(const uint8_t* data, uint32_t width, uint32_t height, uint32_t depth) {
    if (!g.inited) return fail("add_texture before entry()");
    if (g.vols.size() >= kMaxTextures) return fail("Tried allocating too many textures"); // lib/memory.c:287-290
    if (!data || !width || !height || !depth) return fail("add_texture: empty volume");
    CK(cudaSetDevice(g.device));
    const uint32_t xb = ceil_log2(width + 2) < 5 ? 5 : ceil_log2(width + 2);
    const uint32_t yb = ceil_log2(height + 2) < 2 ? 2 : ceil_log2(height + 2);
    const uint32_t zbits = ceil_log2(depth + 2);
    if (xb + yb + zbits > 31) return fail("add_texture: %ux%ux%u does not fit the 31-bit stop-mask index", width, height, depth);
    const size_t bytes = (size_t)4 * width * height * depth;
    const uint32_t mask_words = ((1u << (xb - 5)) << yb) * (depth + 2);
    const uint32_t mask_words_padded = (mask_words + 3u) & ~3u;

    // staging grows to the next power of two and the bytes are copied before returning
    // (lib/memory.c:297-307: `data` is only borrowed for the call)
    if (bytes > g.tex_staging_size) {
        CK(cudaStreamSynchronize(g.stream));
        if (g.h_tex_staging) cudaFreeHost(g.h_tex_staging);
        g.h_tex_staging = nullptr;
        g.tex_staging_size = round_up_p2(bytes);
        CK(cudaMallocHost(&g.h_tex_staging, g.tex_staging_size));
    }
    CK(cudaStreamSynchronize(g.stream)); // the previous upload may still read the staging buffer
    memcpy(g.h_tex_staging, data, bytes);

    uint8_t* d_rgba = nullptr;
    CK(cudaMalloc(&d_rgba, bytes));
    CK(cudaMemcpyAsync(d_rgba, g.h_tex_staging, bytes, cudaMemcpyHostToDevice, g.stream));

    // the arena doubles like the reference's texture memory blocks (lib/memory.c:278-284)
    if (g.arena_words + mask_words_padded > g.arena_cap) {
        uint32_t cap = g.arena_cap ? g.arena_cap : 8192;
        while (cap < g.arena_words + mask_words_padded) cap *= 2;
        uint32_t* d_new = nullptr;
        CK(cudaMalloc(&d_new, (size_t)cap * 4));
        if (g.arena_words) CK(cudaMemcpyAsync(d_new, g.d_arena, (size_t)g.arena_words * 4, cudaMemcpyDeviceToDevice, g.stream));
        CK(cudaStreamSynchronize(g.stream));
        cudaFree(g.d_arena);
        g.d_arena = d_new;
        g.arena_cap = cap;
    }
    VolumeDesc v{};
    v.rgba = d_rgba;
    v.w = width; v.h = height; v.d = depth;
    v.xb = xb; v.yb = yb;
    v.mask_off = g.arena_words;
    v.mask_words = mask_words_padded;
    v.bricks = nullptr;
    // texel == voxel when the reference's coordinate round trip floor(fl(v/s)*s) is the identity
    // (true for 16, 40, 50, 64, ...; false e.g. for 22 or 41) — lets the kernels skip three divisions per hit
    v.remap_identity = 1;
    {
        const uint32_t dims[3] = {width, height, depth};
        for (int a = 0; a < 3 && v.remap_identity; ++a) {
            const float sz = (float)(int32_t)dims[a];
            for (uint32_t x = 0; x < dims[a]; ++x) {
                const float u = (float)(int32_t)x / sz;
                if ((int32_t)floorf(u * sz) != (int32_t)x) { v.remap_identity = 0; break; }
            }
        }
    }
    if (mask_words_padded > mask_words)
        CK(cudaMemsetAsync(g.d_arena + g.arena_words + mask_words, 0xFF, (size_t)(mask_words_padded - mask_words) * 4, g.stream));
    CK(launch_build_mask(d_rgba, width, height, depth, xb, yb, g.d_arena + g.arena_words, mask_words, g.stream));
    g.stats.launches += 1;
    g.arena_words += mask_words_padded;
    g.vols.push_back(v);
    if (xb + yb + zbits > g.max_idx_bits) g.max_idx_bits = xb + yb + zbits;
    g.vols_dirty = true;
    return (int32_t)(g.vols.size() - 1); // lib/memory.c:292,384
}

// extension: a large procedural volume generated on the device as sparse bricks.
extern "C" int32_t vt_add_volume_procedural(uint32_t kind, uint32_t width, uint32_t height, uint32_t depth, uint32_t seed) {
    if (!g.inited) return fail("vt_add_volume_procedural before entry()");
    if (g.vols.size() >= kMaxTextures) return fail("Tried allocating too many textures");
    if (kind != kVolumeHeightmap && kind != kVolumeSparseBricks) return fail("vt_add_volume_procedural: unknown kind %u", kind);
    if (!width || !height || !depth || (width | height | depth) & 7u || width > 8192 || height > 8192 || depth > 8192)
        return fail("vt_add_volume_procedural: dimensions must be multiples of 8, at most 8192");
    {   // brick volumes assume texel == voxel (true for powers of two; checked like add_texture does)
        const uint32_t dims[3] = {width, height, depth};
        for (int a = 0; a < 3; ++a) {
            const float sz = (float)(int32_t)dims[a];
            for (uint32_t x = 0; x < dims[a]; ++x)
                if ((int32_t)floorf(((float)(int32_t)x / sz) * sz) != (int32_t)x)
                    return fail("vt_add_volume_procedural: size %u does not map voxels to texels one to one", dims[a]);
        }
    }
    CK(cudaSetDevice(g.device));
    const size_t bricks = (size_t)(width >> 3) * (height >> 3) * (depth >> 3);
    if (bricks >= (1ull << 31)) return fail("vt_add_volume_procedural: too many bricks");
    State::BrickAlloc a{};
    uint32_t* d_counter = nullptr;
    uint32_t n = 0;
    CK(cudaMalloc(&d_counter, 4));
    if (kind == kVolumeHeightmap) {
        CK(cudaMalloc(&a.heights, (size_t)width * depth * sizeof(float)));
        CK(launch_heightmap(a.heights, width, height, depth, seed, g.stream));
        g.stats.launches += 1;
    }
    // the brick grid is stored with a one-brick border that marks "outside" (kernels.h, BrickVolume)
    const uint32_t pbx = (width >> 3) + 2, pby = (height >> 3) + 2, pbz = (depth >> 3) + 2;
    const size_t padded = (size_t)pbx * pby * pbz;
    const size_t entries = (padded + 15) / 16, scan_words = (entries + 1023) / 1024 + 2;
    uint32_t* scratch = nullptr;
    CK(cudaMalloc(&a.codes, entries * 4));
    CK(cudaMalloc(&a.base, entries * 4));
    CK(cudaMalloc(&scratch, scan_words * 4));
    CK(cudaMemsetAsync(a.codes, 0, entries * 4, g.stream));
    // pass 1 marks and counts the bricks with voxels, the directory is finished, pass 2 fills the pool in grid order
    CK(cudaMemsetAsync(d_counter, 0, 4, g.stream));
    CK(launch_brick_build(kind, seed, width, height, depth, a.heights, a.codes, a.base, nullptr, d_counter, g.stream));
    CK(cudaMemcpyAsync(&n, d_counter, 4, cudaMemcpyDeviceToHost, g.stream));
    CK(launch_brick_finalize(pbx, pby, pbz, a.codes, a.base, scratch, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    CK(cudaMalloc(&a.pool, (size_t)(n ? n : 1) * 64));
    CK(launch_brick_build(kind, seed, width, height, depth, a.heights, a.codes, a.base, a.pool, d_counter, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    cudaFree(scratch);
    g.stats.launches += 7;
    BrickVolume bv{};
    bv.codes = a.codes; bv.base = a.base; bv.pool = a.pool; bv.heights = a.heights; bv.colors = nullptr;
    bv.kind = kind; bv.seed = seed;
    bv.bx = pbx; bv.by = pby; bv.bz = pbz;
    bv.n_bricks = n;
    CK(cudaMalloc(&a.d_desc, sizeof(BrickVolume)));
    CK(cudaMemcpyAsync(a.d_desc, &bv, sizeof bv, cudaMemcpyHostToDevice, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    cudaFree(d_counter);
    g.brick_allocs.push_back(a);
    VolumeDesc v{};
    v.rgba = nullptr;
    v.w = width; v.h = height; v.d = depth;
    v.xb = 5; v.yb = 2; v.mask_off = 0; v.mask_words = 0;
    v.remap_identity = 1;
    v.bricks = a.d_desc;
    g.vols.push_back(v);
    g.vols_dirty = true;
    g.any_bricks = true;
    return (int32_t)(g.vols.size() - 1);
}

// extension: a caller-supplied sparse volume, as 8^3 occupancy bricks with one colour each.
extern "C" int32_t vt_add_volume_bricks(const uint32_t* brick_coords, const uint32_t* masks, const uint8_t* colors, size_t n_bricks,
                                        uint32_t width, uint32_t height, uint32_t depth) {
    if (!g.inited) return fail("vt_add_volume_bricks before entry()");
    if (g.vols.size() >= kMaxTextures) return fail("Tried allocating too many textures");
    if (!width || !height || !depth || (width | height | depth) & 7u || width > 8192 || height > 8192 || depth > 8192)
        return fail("vt_add_volume_bricks: dimensions must be multiples of 8, at most 8192");
    if (n_bricks && (!brick_coords || !masks || !colors)) return fail("vt_add_volume_bricks: null input");
    if (n_bricks >= (1ull << 31)) return fail("vt_add_volume_bricks: too many bricks");
    {
        const uint32_t dims[3] = {width, height, depth};
        for (int a = 0; a < 3; ++a) {
            const float sz = (float)(int32_t)dims[a];
            for (uint32_t x = 0; x < dims[a]; ++x)
                if ((int32_t)floorf(((float)(int32_t)x / sz) * sz) != (int32_t)x)
                    return fail("vt_add_volume_bricks: size %u does not map voxels to texels one to one", dims[a]);
        }
    }
    CK(cudaSetDevice(g.device));
    const size_t n1 = n_bricks ? n_bricks : 1;
    State::BrickAlloc a{};
    uint32_t* d_coords = nullptr;
    uint32_t* d_bad = nullptr;
    uint32_t bad = 0;
    const uint32_t pbx = (width >> 3) + 2, pby = (height >> 3) + 2, pbz = (depth >> 3) + 2;
    const size_t padded = (size_t)pbx * pby * pbz;
    const size_t entries = (padded + 15) / 16, scan_words = (entries + 1023) / 1024 + 2;
    uint32_t* scratch = nullptr;
    uint32_t* d_masks = nullptr;
    uchar4* d_colors = nullptr;
    CK(cudaMalloc(&a.codes, entries * 4));
    CK(cudaMalloc(&a.base, entries * 4));
    CK(cudaMalloc(&scratch, scan_words * 4));
    CK(cudaMalloc(&a.pool, n1 * 64));
    CK(cudaMalloc(&a.colors, n1 * 4));
    CK(cudaMalloc(&d_masks, n1 * 64));
    CK(cudaMalloc(&d_colors, n1 * 4));
    CK(cudaMalloc(&d_coords, n1 * 12));
    CK(cudaMalloc(&d_bad, 4));
    CK(cudaMemsetAsync(a.codes, 0, entries * 4, g.stream));
    CK(cudaMemsetAsync(d_bad, 0, 4, g.stream));
    // the inputs are only borrowed for the call (like add_texture's): synchronous copies
    CK(cudaMemcpy(d_masks, masks, n_bricks * 64, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_colors, colors, n_bricks * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_coords, brick_coords, n_bricks * 12, cudaMemcpyHostToDevice));
    // mark the bricks, finish the directory, then move every brick's words and colour to its slot (grid order)
    CK(launch_brick_index(d_coords, (uint32_t)n_bricks, width >> 3, height >> 3, depth >> 3, a.codes, d_bad, g.stream));
    CK(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    if (!bad) {
        CK(launch_brick_finalize(pbx, pby, pbz, a.codes, a.base, scratch, g.stream));
        CK(launch_brick_place(d_coords, (uint32_t)n_bricks, width >> 3, height >> 3, a.codes, a.base, d_masks, d_colors, a.pool, a.colors, g.stream));
        CK(cudaStreamSynchronize(g.stream));
    }
    cudaFree(d_coords); cudaFree(d_bad); cudaFree(d_masks); cudaFree(d_colors); cudaFree(scratch);
    g.stats.launches += 7;
    if (bad) {
        cudaFree(a.codes); cudaFree(a.base); cudaFree(a.pool); cudaFree(a.colors);
        return fail("vt_add_volume_bricks: %u brick coordinates outside the volume", bad);
    }
    BrickVolume bv{};
    bv.codes = a.codes; bv.base = a.base; bv.pool = a.pool; bv.heights = nullptr; bv.colors = a.colors;
    bv.kind = kVolumeUploadedBricks; bv.seed = 0;
    bv.bx = pbx; bv.by = pby; bv.bz = pbz;
    bv.n_bricks = (uint32_t)n_bricks;
    CK(cudaMalloc(&a.d_desc, sizeof(BrickVolume)));
    CK(cudaMemcpy(a.d_desc, &bv, sizeof bv, cudaMemcpyHostToDevice));
    g.brick_allocs.push_back(a);
    VolumeDesc v{};
    v.w = width; v.h = height; v.d = depth;
    v.xb = 5; v.yb = 2;
    v.remap_identity = 1;
    v.bricks = a.d_desc;
    g.vols.push_back(v);
    g.vols_dirty = true;
    g.any_bricks = true;
    return (int32_t)(g.vols.size() - 1);
}

extern "C" float* start_update_instances(uint32_t instance_count) {
    if (!g.inited) { fail("start_update_instances before entry()"); return nullptr; }
    if (instance_count == 0) instance_count = 1; // lib/memory.c:236
    if (cudaSetDevice(g.device) != cudaSuccess) return nullptr;
    const uint32_t keep = g.inst_count < g.inst_cap ? g.inst_count : g.inst_cap;
    if (ensure_instances(instance_count)) return nullptr;
    // the next staging buffer of the ring; its previous upload (kInstRing updates ago) must have been consumed
    const int prev = g.inst_slot, slot = (g.inst_slot + 1) % State::kInstRing;
    if (g.inst_busy[slot]) {
        if (cudaEventSynchronize(g.ev_inst[slot]) != cudaSuccess) return nullptr;
        g.inst_busy[slot] = false;
    }
    // the reference hands out one persistently mapped buffer (lib/memory.c:245-247): what was written before is still there
    if (keep) memcpy(inst_staging(slot), inst_staging(prev), (size_t)keep * 64);
    g.inst_slot = slot;
    g.inst_count = instance_count;
    return inst_staging(slot);
}
extern "C" int32_t end_update_instances(uint32_t instance_count) {
    if (!g.inited) return fail("end_update_instances before entry()");
    if (instance_count == 0) instance_count = 1; // lib/memory.c:251
    if (instance_count > g.inst_cap) return fail("end_update_instances: %u > capacity %u", instance_count, g.inst_cap);
    CK(cudaSetDevice(g.device));
    g.inst_count = instance_count;
    if (instance_count <= direct_instances()) {
        g.inst_src = inst_staging(g.inst_slot); // page-locked memory is device-accessible (unified addressing): 64 bytes per instance over PCIe
    } else {
        g.inst_src = nullptr;
        // lib/memory.c:257-264: staging -> device copy
        CK(cudaMemcpyAsync(g.d_inst, inst_staging(g.inst_slot), (size_t)instance_count * 64, cudaMemcpyHostToDevice, g.stream));
    }
    CK(cudaEventRecord(g.ev_inst[g.inst_slot], g.stream));
    g.inst_busy[g.inst_slot] = true;
    return 0;
}

extern "C" void cleanup(void) {
    if (!g.inited) return;
    cudaSetDevice(g.device);
    cudaDeviceSynchronize(); // vkDeviceWaitIdle, lib/entry.c:101
    if (g.fused_mode == 1) cudaFree(g.fused_base);
    if (g.fused_mode == 2) cudaIpcCloseMemHandle(g.fused_base);
    g.fused_base = nullptr;
    g.fused_flags = nullptr;
    g.fused_rows = false;
    g.fused_root_live = false;
    cudaFree(g.d_fused_err);
    if (g.h_fused_err) cudaFreeHost(g.h_fused_err);
    g.d_fused_err = g.h_fused_err = nullptr;
    g.fused_mode = 0;
    for (auto& v : g.vols) cudaFree(const_cast<uint8_t*>(v.rgba));
    g.vols.clear();
    for (auto& b : g.brick_allocs) { cudaFree(b.codes); cudaFree(b.base); cudaFree(b.pool); cudaFree(b.heights); cudaFree(b.colors); cudaFree(b.d_desc); }
    g.brick_allocs.clear();
    cudaFree(g.d_vols); cudaFree(g.d_arena); cudaFree(g.d_inst); cudaFree(g.d_iu); cudaFree(g.d_dec); cudaFree(g.d_thr);
    g.deferred_copy.armed = false;
    if (g.copy_stream) { cudaStreamSynchronize(g.copy_stream); cudaStreamDestroy(g.copy_stream); cudaEventDestroy(g.ev_color_ready); cudaEventDestroy(g.ev_copy_done[0]); cudaEventDestroy(g.ev_copy_done[1]); }
    cudaFree(g.d_color_alt);
    cudaFree(g.d_rec); cudaFree(g.d_color); cudaFree(g.d_depth); cudaFree(g.d_accum_own); cudaFree(g.d_stats);
    cudaFree(g.d_bin_offset); cudaFree(g.d_bin_count); cudaFree(g.d_bin_list); cudaFree(g.d_bin_cursor);
    if (g.h_bin_cursor) cudaFreeHost(g.h_bin_cursor);
    cudaFree(g.d_world_aabb); cudaFree(g.d_world_hdr); cudaFree(g.d_world_cells); cudaFree(g.d_world_list);
    if (g.h_world_hdr) cudaFreeHost(g.h_world_hdr);
    g.d_world_aabb = nullptr; g.d_world_hdr = nullptr; g.h_world_hdr = nullptr; g.d_world_cells = g.d_world_list = nullptr;
    g.world_aabb_cap = g.world_list_cap = 0; g.world_used = false;
    if (g.h_inst) cudaFreeHost(g.h_inst);
    for (int i = 0; i < State::kInstRing; ++i) if (g.ev_inst[i]) cudaEventDestroy(g.ev_inst[i]);
    if (g.h_tex_staging) cudaFreeHost(g.h_tex_staging);
    if (g.h_stats) cudaFreeHost(g.h_stats);
    if (g.h_readback) cudaFreeHost(g.h_readback);
    for (uint32_t k = 0; k < State::kRing; ++k) {
        cudaEventDestroy(g.ev_begin[k]); cudaEventDestroy(g.ev_trace0[k]); cudaEventDestroy(g.ev_trace1[k]); cudaEventDestroy(g.ev_end[k]);
    }
    if (g.own_stream) cudaStreamDestroy(g.own_stream);
    const user_input keep = g.input;
    g = State{};
    g.input = keep;
}

// ============================================================================================
// Part 2 — headless extensions

extern "C" int32_t vt_get_config(vt_config* out) {
    if (!g.inited || !out) return -1;
    *out = g.cfg;
    return 0;
}

extern "C" int32_t vt_configure(const vt_config* cfg) {
    if (!g.inited) return fail("vt_configure before entry()");
    if (!cfg || !cfg->width || !cfg->height) return fail("vt_configure: bad size");
    if (cfg->mode > VT_MODE_RAYS) return fail("vt_configure: unknown mode %u", cfg->mode);
    if ((uint64_t)cfg->width * cfg->height > (1ull << 28)) return fail("vt_configure: framebuffer too large");
    CK(cudaSetDevice(g.device));
    if (finish_frame()) return -1;
    const int32_t dev = g.cfg.device;
    g.cfg = *cfg;
    g.cfg.device = dev; // the device is fixed at entry()
    if (!g.cfg.sample_stride) g.cfg.sample_stride = 1;
    return alloc_framebuffer();
}

extern "C" int32_t vt_render_async(const float* projection, const float* camera) {
    if (!projection || !camera) return -1;
    // PATHS: adds this process's samples into the accumulators; clearing / resolving is the
    // launcher's job (vt_clear_accum / vt_resolve) because a cross-rank reduction sits between.
    return render_async(projection, camera, false, false);
}

extern "C" int32_t vt_render_frame_async(const float* projection, const float* camera) {
    if (!projection || !camera) return -1;
    return render_async(projection, camera, true, true);
}

extern "C" int32_t vt_synchronize(void) {
    if (!g.inited) return -1;
    CK(cudaSetDevice(g.device));
    if (finish_frame()) return -1;
    CK(cudaStreamSynchronize(g.stream));
    return 0;
}

extern "C" int64_t vt_read_hits(vt_hit_record* out, size_t capacity) {
    static_assert(sizeof(vt_hit_record) == sizeof(HitRecord), "hit record layout");
    return read_back(g.d_rec, (size_t)g.cfg.width * g.cfg.height * sizeof(HitRecord), out, capacity);
}
extern "C" int64_t vt_read_color(uint8_t* rgba8, size_t capacity) {
    return read_back(g.d_color, (size_t)g.cfg.width * g.cfg.height * 4, rgba8, capacity);
}
extern "C" int64_t vt_read_color_async(uint8_t* pinned_rgba8, size_t capacity) {
    const size_t bytes = (size_t)g.cfg.width * g.cfg.height * 4;
    if (!g.inited) return fail("read before entry()");
    if (!pinned_rgba8 || capacity < bytes) return fail("read-back buffer too small: need %zu bytes, have %zu", bytes, capacity);
    if (cudaSetDevice(g.device) != cudaSuccess) return fail("cudaSetDevice failed");
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, pinned_rgba8) != cudaSuccess || attr.type != cudaMemoryTypeHost) {
        (void)cudaGetLastError();
        return fail("vt_read_color_async: the destination must be page-locked host memory");
    }
    if (!g.copy_stream) {
        CK(cudaStreamCreateWithFlags(&g.copy_stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&g.ev_color_ready, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&g.ev_copy_done[0], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&g.ev_copy_done[1], cudaEventDisableTiming));
    }
    if (issue_deferred_copy(nullptr)) return -1;     // (an earlier read-back that no frame has started yet)
    CK(cudaEventRecord(g.ev_color_ready, g.stream)); // the frame enqueued last is complete
    g.deferred_copy.armed = true;
    g.deferred_copy.dst = pinned_rgba8; g.deferred_copy.src = g.d_color; g.deferred_copy.bytes = bytes; g.deferred_copy.done = g.ev_copy_done[0];
    g.copy_pending[0] = true; // (the next frame renders into the other colour buffer: prepare_color_target)
    if (!g.defer_readback && issue_deferred_copy(nullptr)) return -1;
    return (int64_t)bytes;
}

extern "C" int32_t vt_read_color_wait(void) {
    if (!g.inited) return -1;
    if (issue_deferred_copy(nullptr)) return -1;
    if (g.copy_stream) CK(cudaStreamSynchronize(g.copy_stream));
    return 0;
}

extern "C" int32_t vt_read_color_fence(void) {
    if (!g.inited) return -1;
    if (!g.copy_stream) return 0;
    if (issue_deferred_copy(nullptr)) return -1;
    for (int i = 0; i < 2; ++i)
        if (g.copy_pending[i]) CK(cudaStreamWaitEvent(g.stream, g.ev_copy_done[i], 0));
    return 0;
}

extern "C" int64_t vt_read_color_bgra(uint8_t* bgra8, size_t capacity) {
    const int64_t n = read_back(g.d_color, (size_t)g.cfg.width * g.cfg.height * 4, bgra8, capacity);
    for (int64_t p = 0; p + 3 < n; p += 4) { const uint8_t r = bgra8[p]; bgra8[p] = bgra8[p + 2]; bgra8[p + 2] = r; }
    return n;
}

extern "C" int32_t vt_write_ppm(const char* path) {
    if (!g.inited || !path) return -1;
    const size_t n = (size_t)g.cfg.width * g.cfg.height;
    std::vector<uint8_t> rgba(n * 4);
    if (read_back(g.d_color, n * 4, rgba.data(), rgba.size()) < 0) return -1;
    FILE* f = fopen(path, "wb");
    if (!f) return fail("vt_write_ppm: cannot open %s", path);
    fprintf(f, "P6\n%u %u\n255\n", g.cfg.width, g.cfg.height);
    std::vector<uint8_t> rgb(n * 3);
    for (size_t p = 0; p < n; ++p) { rgb[3 * p] = rgba[4 * p]; rgb[3 * p + 1] = rgba[4 * p + 1]; rgb[3 * p + 2] = rgba[4 * p + 2]; }
    const bool ok = fwrite(rgb.data(), 1, rgb.size(), f) == rgb.size();
    fclose(f);
    return ok ? 0 : fail("vt_write_ppm: short write to %s", path);
}

extern "C" int64_t vt_read_depth(float* depth, size_t capacity) {
    return read_back(g.d_depth, (size_t)g.cfg.width * g.cfg.height * 4, depth, capacity);
}
extern "C" int64_t vt_read_accum(uint64_t* accum, size_t capacity) {
    if (g.inited && g.fused_mode == 1) { // materialise the sum of all ranks' partial sums
        if (cudaSetDevice(g.device) != cudaSuccess) return -1;
        if (!g.fused_sum && cudaMalloc(&g.fused_sum, g.fused_pixels * 24) != cudaSuccess) return fail("vt_read_accum: out of memory");
        if (fused_root_sum(g.fused_sum) != cudaSuccess) return fail("vt_read_accum: resolve failed");
        return read_back(g.fused_sum, g.fused_pixels * 24, accum, capacity);
    }
    if (g.inited && g.sky_missing_spp) {
        if (cudaSetDevice(g.device) != cudaSuccess || finish_frame() || complete_accum()) return -1;
    }
    return read_back(g.d_accum, (size_t)g.cfg.width * g.cfg.height * 24, accum, capacity);
}

extern "C" void* vt_accum_device_ptr(void) {
    if (!g.inited) return nullptr;
    if (g.sky_missing_spp && (cudaSetDevice(g.device) != cudaSuccess || finish_frame() || complete_accum())) return nullptr;
    return (void*)g.d_accum;
}

extern "C" int32_t vt_set_accum_buffer(void* device_ptr) {
    if (!g.inited) return -1;
    CK(cudaSetDevice(g.device));
    if (finish_frame()) return -1;
    if (complete_accum()) return -1;
    if (g.fused_mode && device_ptr) return fail("vt_set_accum_buffer: a fused reduction accumulates in the library's own buffer (vt_fused_reduce_disable first)");
    CK(cudaStreamSynchronize(g.stream));
    g.d_accum = device_ptr ? (unsigned long long*)device_ptr : g.d_accum_own;
    return 0;
}

extern "C" int32_t vt_clear_accum(void) {
    if (!g.inited) return -1;
    CK(cudaSetDevice(g.device));
    CK(cudaMemsetAsync(g.d_accum, 0, (size_t)g.cfg.width * g.cfg.height * 24, g.stream));
    g.sky_missing_spp = 0;
    return 0;
}

extern "C" int32_t vt_resolve(void) {
    if (!g.inited) return -1;
    CK(cudaSetDevice(g.device));
    if (complete_accum()) return -1;
    const uint32_t total = g.cfg.total_spp ? g.cfg.total_spp : g.cfg.spp;
    SrgbTables lut{g.d_dec, g.d_thr};
    if (g.fused_mode == 2) return fail("vt_resolve: only the root of a fused reduction holds the sums");
    if (prepare_color_target()) return -1;
    if (g.fused_mode == 1) {
        CK(fused_root_sum(nullptr));
    } else {
        CK(launch_resolve(g.d_accum, g.cfg.width * g.cfg.height, total ? total : 1, lut, g.d_color, g.stream));
    }
    g.stats.launches += 1;
    return 0;
}

// ---- fused cross-GPU accumulation over NVLink peer memory (one process per GPU, one node) -----------
static int fused_common(uint32_t rank, uint32_t world) {
    if (world < 1 || rank >= world) return fail("fused reduction: bad rank %u of %u", rank, world);
    g.fused_pixels = (size_t)g.cfg.width * g.cfg.height;
    g.fused_rank = rank;
    g.fused_world = world;
    g.fused_index = 0;
    g.fused_seq = 0;
    g.fused_seq_rendered = 0;
    g.fused_root_live = false;
    g.fused_rows = false;
    g.fused_sync = env_u32("VT_FUSED_SYNC", 1) != 0; // 0: the caller orders the ranks itself (a stream barrier per frame)
    if (world > kFusedMaxWorld) return fail("fused reduction: at most %u ranks", kFusedMaxWorld);
    if (!g.d_fused_err) {
        CK(cudaMalloc(&g.d_fused_err, 8)); // [0] error word, [1] the push kernel's done-block counter
        CK(cudaMallocHost(&g.h_fused_err, 4));
    }
    *g.h_fused_err = 0;
    CK(cudaMemsetAsync(g.d_fused_err, 0, 8, g.stream));
    // the local accumulators must start (and, thanks to push_partial, stay) clear
    CK(cudaMemsetAsync(g.d_accum_own, 0, g.fused_pixels * 3 * sizeof(unsigned long long), g.stream));
    CK(cudaStreamSynchronize(g.stream));
    return 0;
}

extern "C" int32_t vt_fused_reduce_export(uint8_t handle[64], uint32_t world) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    if (!g.inited || !handle) return -1;
    CK(cudaSetDevice(g.device));
    if (finish_frame()) return -1;
    if (g.fused_mode) return fail("fused reduction already set up");
    if (fused_common(0, world)) return -1;
    const size_t data = 2 * (size_t)world * g.fused_pixels * 2 * sizeof(uint4);
    const size_t bytes = data + (2 * kFusedMaxWorld + 2) * sizeof(uint32_t); // partial sums, then the flags
    CK(cudaMalloc(&g.fused_base, bytes));
    CK(cudaMemset(g.fused_base, 0, bytes));
    g.fused_flags = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(g.fused_base) + data);
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, g.fused_base));
    memcpy(handle, &h, 64);
    g.fused_mode = 1;
    return 0;
}

extern "C" int32_t vt_fused_reduce_import(const uint8_t handle[64], uint32_t rank, uint32_t world) {
    if (!g.inited || !handle) return -1;
    CK(cudaSetDevice(g.device));
    if (finish_frame()) return -1;
    if (g.fused_mode) return fail("fused reduction already set up");
    if (rank == 0) return fail("vt_fused_reduce_import: rank 0 is the root (it exports)");
    if (fused_common(rank, world)) return -1;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    void* p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    g.fused_base = (uint4*)p;
    g.fused_flags = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(p) + 2 * (size_t)world * g.fused_pixels * 2 * sizeof(uint4));
    g.fused_mode = 2;
    return 0;
}

extern "C" int32_t vt_fused_reduce_partition(uint32_t by_tile_rows, uint32_t root_relief_num, uint32_t root_relief_den) {
    if (!g.inited || !g.fused_mode) return fail("vt_fused_reduce_partition: no fused reduction set up");
    if (root_relief_den < 1u || root_relief_den > 64u || root_relief_num >= root_relief_den)
        return fail("vt_fused_reduce_partition: the root's relief is num / den with den in 1 .. 64 and num < den");
    CK(cudaSetDevice(g.device));
    if (finish_frame()) return -1;
    g.fused_rows = by_tile_rows != 0;
    g.fused_relief = root_relief_num;
    g.fused_relief_den = root_relief_den;
    return 0;
}

extern "C" int32_t vt_fused_reduce_next_frame(void) {
    if (!g.inited || !g.fused_mode) return -1;
    g.fused_seq += 1;
    g.fused_index = g.fused_seq & 1u;
    return 0;
}

extern "C" int32_t vt_fused_reduce_disable(void) {
    if (!g.inited) return -1;
    CK(cudaSetDevice(g.device));
    if (finish_frame()) return -1;
    CK(cudaStreamSynchronize(g.stream));
    if (g.fused_mode == 1) cudaFree(g.fused_base);
    if (g.fused_mode == 2) cudaIpcCloseMemHandle(g.fused_base);
    cudaFree(g.fused_sum);
    g.fused_sum = nullptr;
    g.fused_base = nullptr;
    g.fused_mode = 0;
    return 0;
}

extern "C" int32_t vt_set_stream(void* cuda_stream) {
    if (!g.inited) return -1;
    CK(cudaSetDevice(g.device));
    if (finish_frame()) return -1;
    CK(cudaStreamSynchronize(g.stream));
    g.stream = cuda_stream ? (cudaStream_t)cuda_stream : g.own_stream;
    return 0;
}

extern "C" int32_t vt_get_stats(vt_stats* out) {
    if (!g.inited || !out) return -1;
    if (g.frame_pending) {
        if (cudaSetDevice(g.device) != cudaSuccess || finish_frame()) return -1;
    }
    *out = g.stats;
    g.stats.trace_ms_sum = 0.0f; // the sums cover the frames since the previous call
    g.stats.trace_frames = 0;
    g.stats.rays_sum = g.stats.iterations_sum = g.stats.analytic_rays_sum = 0;
    return 0;
}

extern "C" int32_t vt_set_user_input(const user_input* in) {
    if (!in) return -1;
    g.input = *in;
    return 0;
}

extern "C" const char* vt_last_error(void) { return g.err; }

static_assert(sizeof(user_input) == 40, "UserInput is 40 bytes (src/render.rs:37-51)");
static_assert(sizeof(render_tick_info) == 2 * sizeof(void*), "RenderTickInfo is two pointers (src/render.rs:177-181)");
static_assert(sizeof(vt_hit_record) == 16, "hit record is 16 bytes");

#ifdef VT_WAVE_STATS
// variant builds only (python -m vtrace_b200.build --variant stats -DVT_WAVE_STATS): per-phase counters of the wavefront kernel
namespace vt { cudaError_t read_wave_stats(unsigned long long* out16); }
extern "C" int vt_debug_wave_stats(uint64_t* out16) {
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    return vt::read_wave_stats(reinterpret_cast<unsigned long long*>(out16)) == cudaSuccess ? 0 : -1;
}
// histograms in 5 us buckets since each warp started: [0,64) work exhausted, [64,128) warp finished, [128,192) first item claimed
namespace vt { cudaError_t read_wave_times(unsigned int* out192); }
extern "C" int vt_debug_wave_times(uint32_t* out192) {
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    return vt::read_wave_times(out192) == cudaSuccess ? 0 : -1;
}
#endif

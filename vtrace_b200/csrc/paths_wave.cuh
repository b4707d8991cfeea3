// paths_wave.cuh — trace_paths_wave_kernel: the path tracer for single-instance scenes
// (BASELINE configs[2]) as a warp-local wavefront engine.  Included from kernels.cu (namespace vt).
//
// Same paths, same arithmetic, same integer sums as trace_path() — only the schedule differs.
// Rays of one warp need 0..150 DDA iterations (22 on average in the bench scene, a fifth of them at
// most 3), paths end after 1..5 segments, and regenerating a ray (shade + bounce, or a new camera
// ray) costs more than marching it.  Every warp owns a pool of kSlots paths in shared memory and
// alternates between two kinds of full-width work:
//
//   generate : 32 lanes shade and bounce 32 hits (taken from the warp's hit stack), or start 32 new
//              camera rays (jobs = covered pixel x sample, claimed per (tile, samples) item from a
//              global counter).  A lane that produced a ray keeps it in registers and walks it right
//              away — a fresh ray never travels through shared memory.
//   march    : the stepping loop is one PTX block, fully predicated (a lane whose ray stopped — or
//              that has no ray — keeps its "stopped" predicate and executes nothing): 16 SASS
//              instructions per voxel step, no branch inside a burst of 4 steps, one vote per burst.
//              When `refill_batch` lanes have stopped they decode hit / miss and write their exit
//              state: a hit goes on the hit stack, a ray that left the volume moves its throughput to
//              the miss list and frees its slot at once; idle lanes pull the rays that the previous
//              march parked.  When none are left and fewer than `refill_threshold` lanes still walk,
//              the walkers are parked and the warp generates again.
//   sky      : whenever the miss list holds 32 entries, 32 lanes add throughput x clear colour
//              (integer atomics).
//
// Slots, hits, parked rays and misses are kept in small per-warp stacks that are pushed with
// ballot / popc ranks — no state bytes, no compaction passes.  Everything is warp-local (__syncwarp
// only): no inter-warp queues, no block barriers after the prologue.  Radiance is 2^-24 fixed point
// added with integer atomics, RNG streams are keyed by (pixel, sample): the result is bit-identical to
// the per-pixel kernel and to the CPU oracle.
//
// Slot layout (96 bytes, 3 + 3 uint4; the 48-byte stride makes 128-bit accesses of consecutive
// slots bank-conflict-free):
//   ray  q0 = side.xyz, idx      q1 = signed delta.xyz (len / dir), step signs
//        q2 = prev idx, steps, rng key, -
//   path p0 = thr.rgb, pixel handle   p1 = pos.xyz, len          p2 = dir.xyz, meta
//
// Multi-GPU: every rank accumulates in its OWN buffer (remote atomics from inside this kernel were tried and
// dropped) and a separate kernel moves the sums to the root.  A rank traces either a subset of
// the samples of every pixel (fp.sample_first / sample_stride) or every sample of the tile rows fp.rows
// deals to it (RowShare); the per-tile flushes are integer adds, so the image is the same either way.
#pragma once

// Shape of a CTA: warps, paths in flight per warp, CTAs per SM.  (Compile-time knobs so that variants can be
// built side by side: python -m vtrace_b200.build --variant NAME -DVT_WAVE_WARPS=.. -DVT_WAVE_SLOTS=.. -DVT_WAVE_CTAS=..)
// The kernel is bound by latency below ~24 warps per SM (12: 1.51 ms, 16: 1.25, 20: 1.11, 24: 1.08 on the bench frame).
#ifndef VT_WAVE_WARPS
#define VT_WAVE_WARPS 24
#endif
#ifndef VT_WAVE_SLOTS
#define VT_WAVE_SLOTS 64
#endif
#ifndef VT_WAVE_CTAS
#define VT_WAVE_CTAS 1
#endif
static constexpr int kWaveWarps = VT_WAVE_WARPS;
static constexpr int kWaveThreads = kWaveWarps * 32;
static constexpr int kWaveCtas = VT_WAVE_CTAS;
static constexpr int kSlots = VT_WAVE_SLOTS; // paths in flight per warp
static constexpr int kSlotGroups = (kSlots + 31) / 32;
static constexpr int kMissCap = 64;          // miss list entries (drained in batches of 32 as soon as it holds 32)
static_assert(kSlots % 16 == 0 && kSlots >= 32 && kSlots <= 256, "slot ids travel as bytes; the pool is 16-byte granular");
// slots + miss list + tile accumulators + free stack + hit stack + parked list + covered-pixel table
static constexpr uint32_t kPoolBytes = kSlots * 96 + kMissCap * 16 + 32 * 3 * 4 + kSlots + kSlots + 32 + 32;
static_assert(kPoolBytes % 16 == 0, "pool alignment");
// -DVT_WAVE_STATS (variant builds only): per-phase counters, read with vt_debug_wave_stats()
#ifdef VT_WAVE_STATS
__device__ unsigned long long vt_wave_stats[32];
__device__ unsigned int vt_wave_times[3 * 64]; // histograms (5 us buckets since the warp started): work exhausted, finished, first item claimed
#define VT_STAT(i, v) (wstat[i] += (v))
__device__ __forceinline__ unsigned long long wave_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#else
#define VT_STAT(i, v) ((void)0)
#endif

size_t wave_smem_bytes(uint32_t arena_words, bool masks_in_smem) {
    return trace_smem_bytes(arena_words, masks_in_smem) + size_t(kWaveWarps) * kPoolBytes;
}

__device__ __forceinline__ uint32_t pack_signs(const int32_t step[3]) {
    return (uint32_t)(step[0] + 1) | ((uint32_t)(step[1] + 1) << 2) | ((uint32_t)(step[2] + 1) << 4);
}

// ---- the stepping loop -------------------------------------------------------------------------
// One voxel step (trace.frag:76-86) of a lane whose "stopped" predicate ps is false; a lane with ps
// set executes nothing.  %0-%2 side, %3 idx, %4 rec (previous idx | burst position << 30),
// %8-%10 delta, %11-%13 index increments, %14 mask base.  LD = the mask word load for word index t.
#define VT_WAVE_SUB(LD, REC)                                                                        \
    "shr.u32 t, %3, 5;\n" LD                                                                         \
    "shf.l.wrap.b32 b, 0, 1, %3;\n"                                                                  \
    "lop3.or.b32 b|ps, b, w, 0, 0xc0, ps;\n" /* :78 filled (or border): the walk ends here       */ \
    "min.f32 m, %0, %1;\n"                                                                           \
    "min.f32 m, m, %2;\n"                                                                            \
    "setp.eq.and.f32 px, %0, m, !ps;\n"   /* :83 mask = side <= min(other two)                   */ \
    "setp.eq.and.f32 py, %1, m, !ps;\n"                                                              \
    "setp.eq.and.f32 pz, %2, m, !ps;\n" REC                                                          \
    "@px add.rn.f32 %0, %0, %8;\n"        /* :84                                                 */ \
    "@py add.rn.f32 %1, %1, %9;\n"                                                                   \
    "@pz add.rn.f32 %2, %2, %10;\n"                                                                  \
    "@px add.s32 %3, %3, %11;\n"          /* :85                                                 */ \
    "@py add.s32 %3, %3, %12;\n"                                                                     \
    "@pz add.s32 %3, %3, %13;\n"
#define VT_WAVE_LD_SMEM "shl.b32 t, t, 2;\n" "add.u32 t, t, %14;\n" "ld.shared.u32 w, [t];\n"
#define VT_WAVE_LD_GLOBAL "mad.wide.u32 ta, t, 4, %14;\n" "ld.global.nc.u32 w, [ta];\n"
// Iterations per burst (between two votes): 4 or 8.  The position inside the burst lives in the top bits of rec.
#ifndef VT_WAVE_BURST
#define VT_WAVE_BURST 4
#endif
#if VT_WAVE_BURST == 8
#define VT_WAVE_SUBS(LD)                                                                             \
    VT_WAVE_SUB(LD, "@!ps mov.u32 %4, %3;\n")                                                        \
    VT_WAVE_SUB(LD, "@!ps add.u32 %4, %3, 0x20000000;\n")                                            \
    VT_WAVE_SUB(LD, "@!ps add.u32 %4, %3, 0x40000000;\n")                                            \
    VT_WAVE_SUB(LD, "@!ps add.u32 %4, %3, 0x60000000;\n")                                            \
    VT_WAVE_SUB(LD, "@!ps add.u32 %4, %3, 0x80000000;\n")                                            \
    VT_WAVE_SUB(LD, "@!ps add.u32 %4, %3, 0xa0000000;\n")                                            \
    VT_WAVE_SUB(LD, "@!ps add.u32 %4, %3, 0xc0000000;\n")                                            \
    VT_WAVE_SUB(LD, "@!ps add.u32 %4, %3, 0xe0000000;\n")                                            \
    "@!ps add.u32 %5, %5, 8;\n"          /* :86, eight at a time                                 */
static constexpr uint32_t kWaveIdxBits = 29, kWaveBurst = 8;
#else
#define VT_WAVE_SUBS(LD)                                                                             \
    VT_WAVE_SUB(LD, "@!ps mov.u32 %4, %3;\n")                                                        \
    VT_WAVE_SUB(LD, "@!ps add.u32 %4, %3, 0x40000000;\n")                                            \
    VT_WAVE_SUB(LD, "@!ps add.u32 %4, %3, 0x80000000;\n")                                            \
    VT_WAVE_SUB(LD, "@!ps add.u32 %4, %3, 0xc0000000;\n")                                            \
    "@!ps add.u32 %5, %5, 4;\n"          /* :86, four at a time                                  */
static constexpr uint32_t kWaveIdxBits = 30, kWaveBurst = 4;
#endif
// The march keeps "previous index" and the step count modulo the burst length in one register (rec): the top bits are
// the position inside the burst, so a volume's stop-mask index must fit kWaveIdxBits bits (launch_trace_paths checks).
static constexpr uint32_t kRecIdxMask = (1u << kWaveIdxBits) - 1u, kRecFresh = (kWaveBurst - 1u) << kWaveIdxBits;
#define VT_WAVE_LOOP(LD)                                                                            \
    "{\n"                                                                                            \
    ".reg .pred ps, px, py, pz, pc;\n"                                                               \
    ".reg .u32 t, w, b, n;\n"                                                                        \
    ".reg .u64 ta;\n"                                                                                \
    ".reg .f32 m;\n"                                                                                 \
    "setp.ne.u32 ps, %6, 0;\n"                                                                       \
    "WAVE_LOOP:\n"                                                                                   \
    VT_WAVE_SUBS(LD)                                                                                 \
    "vote.sync.ballot.b32 %7, ps, 0xffffffff;\n"                                                     \
    "popc.b32 n, %7;\n"                                                                              \
    "setp.lt.u32 pc, n, %15;\n"                                                                      \
    "@pc bra.uni WAVE_LOOP;\n"                                                                       \
    "selp.u32 %6, 1, 0, ps;\n"                                                                       \
    "}\n"

// Steps the warp's rays in bursts of four iterations until at least `k_stop` lanes are stopped (lanes
// without a ray count as stopped and sit on the border bit idx 0).  `rec` must enter with its position bits
// set (kRecFresh); afterwards a lane's ray has taken steps + ((rec >> kWaveIdxBits) + 1 & kWaveBurst - 1) iterations
// and the index before its last iteration is rec & kRecIdxMask.  Returns the ballot of stopped lanes.
template <bool kSmem>
__device__ __forceinline__ uint32_t wave_walk(const Vol& vol, float& sx, float& sy, float& sz, float dx, float dy, float dz,
                                              uint32_t& idx, uint32_t& rec, uint32_t& steps, uint32_t& stopped, uint32_t ix,
                                              uint32_t iy, uint32_t iz, uint32_t k_stop, uint32_t smem0) {
    uint32_t v;
    if (kSmem) {
        const uint32_t base = smem0 + kSmemMaskOff + vol.mask_off * 4u;
        asm volatile(VT_WAVE_LOOP(VT_WAVE_LD_SMEM)
                     : "+f"(sx), "+f"(sy), "+f"(sz), "+r"(idx), "+r"(rec), "+r"(steps), "+r"(stopped), "=r"(v)
                     : "f"(dx), "f"(dy), "f"(dz), "r"(ix), "r"(iy), "r"(iz), "r"(base), "r"(k_stop));
    } else {
        const uint32_t* base = vol.arena + vol.mask_off;
        asm volatile(VT_WAVE_LOOP(VT_WAVE_LD_GLOBAL)
                     : "+f"(sx), "+f"(sy), "+f"(sz), "+r"(idx), "+r"(rec), "+r"(steps), "+r"(stopped), "=r"(v)
                     : "f"(dx), "f"(dy), "f"(dz), "r"(ix), "r"(iy), "r"(iz), "l"(base), "r"(k_stop));
    }
    return v;
}

template <bool kSmem>
__global__ void __launch_bounds__(kWaveThreads, kWaveCtas) trace_paths_wave_kernel(const __grid_constant__ FrameParams fp,
                                                                          const InstUniforms* __restrict__ inst,
                                                                          const uint32_t* __restrict__ mask_arena,
                                                                          uint32_t arena_words, SrgbTables lut, FrameBuffers fb) {
    stage_tables<kSmem>(mask_arena, arena_words, lut.decode);
    // The shared-window address of the CTA's dynamic shared memory is pinned in one register (and every pool pointer
    // derived from it): left alone, the compiler rebuilds it from %cluster_ctarank wherever a pointer is needed.
    uint32_t smem0 = smem_u32(vt_smem);
    asm volatile("" : "+r"(smem0));
    unsigned char* const smem_base = static_cast<unsigned char*>(__cvta_shared_to_generic(smem0));
    const float* dec = reinterpret_cast<const float*>(smem_base + kSmemLutOff);
    const InstUniforms* Ip = inst; // the one instance
    const Vol vol{Ip->w, Ip->h, Ip->d, Ip->xb, Ip->yb, Ip->mask_off, mask_arena};
    const float size[3] = {(float)(int32_t)Ip->w, (float)(int32_t)Ip->h, (float)(int32_t)Ip->d};
    const uint32_t xb = vol.xb, zb = vol.xb + vol.yb;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t pool_off = kSmemMaskOff + (kSmem ? arena_words * 4u : 0u);
    uint4* ray = reinterpret_cast<uint4*>(smem_base + pool_off + warp * kPoolBytes);
    uint4* path = ray + kSlots * 3;
    uint4* miss_list = path + kSlots * 3;                                 // (thr.rgb, pixel handle) of rays that left the volume
    uint32_t* wacc = reinterpret_cast<uint32_t*>(miss_list + kMissCap);   // radiance sums of the current item's tile (2^-24 fixed point)
    uint8_t* free_stack = reinterpret_cast<uint8_t*>(wacc + 32 * 3);      // slots without a path
    uint8_t* hit_stack = free_stack + kSlots;                             // slots whose ray stopped on a filled voxel
    uint8_t* parked = hit_stack + kSlots;                                 // slots whose ray is still walking (written when a march ends)
    uint8_t* cov_pix = parked + 32; // cov_pix[c] = tile-local index of the item's c-th covered pixel

    const int tiles_x = (fp.width + kTileW - 1) / kTileW;
    const int tiles_y = (fp.height + kTileH - 1) / kTileH;
    const int n_tiles = tiles_x * tiles_y;
    // Only tiles that touch the instance's screen rectangle become work items; every other pixel sees
    // nothing but sky and is written by the prologue below.
    const bool any_cov = Ip->bounds[0] <= Ip->bounds[1] && Ip->bounds[2] <= Ip->bounds[3];
    const int ctx0 = any_cov ? Ip->bounds[0] / kTileW : 0, ctx1 = any_cov ? Ip->bounds[1] / kTileW : -1;
    const int cty0 = any_cov ? Ip->bounds[2] / kTileH : 0, cty1 = any_cov ? Ip->bounds[3] / kTileH : -1;
    // multi-GPU, frame shared out by rows of tiles: this rank's rows (RowShare; a few of the enumerated rows may lie outside
    // the rectangle: their tiles have no covered pixel and cost one claim)
    const RowShare rows = fp.rows;
    const int own_rows = any_cov ? (int)rows.rows_in((uint32_t)cty0, (uint32_t)cty1) : 0;
    const int cov_w = ctx1 - ctx0 + 1, cov_tiles = cov_w * own_rows;
    // Work item = one tile x `item_spp` samples.  Items are sized so that there are several per resident
    // warp even when a rank only has a few samples per pixel (multi-GPU), otherwise the tail dominates.
    const uint32_t kItemSpp = fp.item_spp; // most samples per work item
    uint32_t item_spp = kItemSpp;
    {
        const long long want_items = (long long)fp.items_per_warp * gridDim.x * kWaveWarps;
        long long chunks = cov_tiles > 0 ? (want_items + cov_tiles - 1) / cov_tiles : 1;
        if (chunks < (long long)((fp.spp + kItemSpp - 1) / kItemSpp)) chunks = (fp.spp + kItemSpp - 1) / kItemSpp;
        if (chunks > (long long)fp.spp) chunks = fp.spp;
        if (chunks < 1) chunks = 1;
        item_spp = (fp.spp + (uint32_t)chunks - 1) / (uint32_t)chunks;
        if (item_spp < 1) item_spp = 1;
    }
    // Chunks are handed out in order (chunk-major), so the LAST ones decide how long the slowest warp runs
    // after the counter is exhausted: the final chunk's samples are split again and again in halves
    // (8 -> 4, 2, 1, 1), guided-self-scheduling style.
    const uint32_t n_full = fp.spp ? (fp.spp + item_spp - 1) / item_spp - 1 : 0; // chunks of exactly item_spp samples
    const uint32_t tail_spp = fp.spp - n_full * item_spp;                        // 1 .. item_spp samples left for the tail chunks
    uint32_t n_tail = 0;
    for (uint32_t rem = tail_spp; rem; rem -= (rem + 1) / 2) ++n_tail;
    const int n_chunks = (int)(n_full + n_tail);
    const int n_items = cov_tiles * n_chunks;

    const float sky[3] = {53.0f / 100.0f, 81.0f / 100.0f, 92.0f / 100.0f}; // lib/command.c:57-59
    unsigned long long sky_q[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) sky_q[c] = __float2ull_rz((1.0f * sky[c]) * 16777216.0f);
    unsigned long long rays = 0, iters = 0, analytic = 0;
#ifdef VT_WAVE_STATS
    const unsigned long long t_start = wave_timer_ns();
    unsigned long long t_exhausted = 0, t_first = 0;
    uint32_t wstat[32];
    for (int i = 0; i < 32; ++i) wstat[i] = 0;
#endif

    // ---- prologue: pixels outside the screen rectangle see only sky, for every sample.  No other
    // warp ever touches them, so a plain read-modify-write is enough (they still count as rays).
    {
        const int gw = blockIdx.x * kWaveWarps + warp, nw = gridDim.x * kWaveWarps;
        for (int tile = gw; tile < n_tiles; tile += nw) {
            const int tx = tile % tiles_x, ty = tile / tiles_x;
            const int px = tx * kTileW + (lane & 7), py = ty * kTileH + (lane >> 3);
            const bool in_frame = px < fp.width && py < fp.height && rows.owner((uint32_t)ty) == rows.rank;
            const bool may_hit = !(px < Ip->bounds[0] || px > Ip->bounds[1] || py < Ip->bounds[2] || py > Ip->bounds[3]);
            if (in_frame && !may_hit) {
                const size_t p = (size_t)py * (size_t)fp.width + (size_t)px;
                if (fp.sky_spp) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) fb.accum[3 * p + c] += sky_q[c] * fp.sky_spp;
                }
                rays += fp.spp;
                analytic += fp.spp;
            }
        }
    }

#pragma unroll
    for (int g = 0; g < kSlotGroups; ++g)
        if (lane + 32 * g < kSlots) free_stack[lane + 32 * g] = (uint8_t)(lane + 32 * g);
    wacc[lane * 3 + 0] = 0u; wacc[lane * 3 + 1] = 0u; wacc[lane * 3 + 2] = 0u;
    __syncwarp();
    int n_free = kSlots, n_hit = 0, n_miss = 0, n_parked = 0; // warp-uniform stack heights

    // current work item (warp-uniform)
    bool work_left = true;
    int it_x0 = 0, it_y0 = 0;
    uint32_t it_tile = 0xFFFFFFFFu, it_ncov = 1, it_magic = 0, it_next = 0, it_njobs = 0, it_s0 = 0;

    // the lane's walking ray (registers): side, delta, stop-mask index and its per-axis increments, rec = previous
    // index | burst position << 30, steps, slot (-1: none); stopped != 0: nothing to step
    float sx = 0.0f, sy = 0.0f, sz = 0.0f, dx = 0.0f, dy = 0.0f, dz = 0.0f;
    uint32_t idx = 0, ix = 0, iy = 0, iz = 0, rec = kRecFresh, steps = 0, stopped = 1u; // idx 0 = border bit, where lanes without a ray sit
    int my_slot = -1;

    // A path is identified by its pixel handle = tile << 5 | tile-local pixel.  Radiance of paths that
    // belong to the warp's current tile goes to shared-memory accumulators (flushed once per item);
    // stragglers of earlier items go straight to the frame accumulators.  Integer adds: order-free.
    auto add_sky = [&](uint32_t handle, float t0, float t1, float t2) {
        const float thr[3] = {t0, t1, t2};
        if ((handle >> 5) == it_tile) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float q = (thr[c] * sky[c]) * 16777216.0f;
                if (q == q && q > 0.0f) atomicAdd(&wacc[(handle & 31u) * 3 + c], (uint32_t)__float2ull_rz(q));
            }
        } else {
            const uint32_t tile = handle >> 5, pix = handle & 31u;
            const uint32_t ty = tile / (uint32_t)tiles_x, tx = tile - ty * (uint32_t)tiles_x;
            const size_t p = (size_t)(ty * kTileH + (pix >> 3)) * (size_t)fp.width + (size_t)(tx * kTileW + (pix & 7u));
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float q = (thr[c] * sky[c]) * 16777216.0f;
                if (q == q && q > 0.0f) atomicAdd(fb.accum + 3 * p + c, __float2ull_rz(q));
            }
        }
    };
    // moves the current tile's sums to the frame accumulators (lane = tile-local pixel)
    auto flush_tile = [&]() {
        __syncwarp();
        const int my_px = it_x0 + (lane & 7), my_py = it_y0 + (lane >> 3);
        if (it_tile != 0xFFFFFFFFu && my_px < fp.width && my_py < fp.height) {
            const size_t p = (size_t)my_py * (size_t)fp.width + (size_t)my_px;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const uint32_t v = wacc[lane * 3 + c];
                if (v) atomicAdd(fb.accum + 3 * p + c, (unsigned long long)v);
                wacc[lane * 3 + c] = 0u;
            }
        }
        __syncwarp();
    };
    // ---- sky batch: up to 32 entries of the miss list add throughput x clear colour
    auto sky_batch = [&]() {
        __syncwarp();
        const int n = n_miss < 32 ? n_miss : 32;
        if (lane < n) {
            const uint4 e = miss_list[n_miss - n + lane];
            add_sky(e.w, __uint_as_float(e.x), __uint_as_float(e.y), __uint_as_float(e.z));
        }
        VT_STAT(4, 1); VT_STAT(5, n);
        n_miss -= n;
        __syncwarp();
    };
    // stack pushes (all lanes call; `mine` lanes push).  The miss list has room for 32 more entries whenever
    // it holds fewer than 32, which every caller ensures (see sky_batch calls).
    auto push_hit = [&](bool mine, uint32_t slot) {
        const uint32_t m = __ballot_sync(0xffffffffu, mine);
        if (mine) hit_stack[n_hit + __popc(m & lt_mask)] = (uint8_t)slot;
        n_hit += __popc(m);
    };
    auto push_free = [&](bool mine, uint32_t slot) {
        const uint32_t m = __ballot_sync(0xffffffffu, mine);
        if (mine) free_stack[n_free + __popc(m & lt_mask)] = (uint8_t)slot;
        n_free += __popc(m);
    };
    auto push_miss = [&](bool mine, uint32_t slot) { // the ray left the volume: its throughput waits in the miss list, the slot is free
        const uint32_t m = __ballot_sync(0xffffffffu, mine);
        if (mine) {
            const int r = __popc(m & lt_mask);
            miss_list[n_miss + r] = path[slot * 3 + 0];
            free_stack[n_free + r] = (uint8_t)slot;
        }
        n_miss += __popc(m);
        n_free += __popc(m);
    };
    // Starts the walk of a fresh ray (pos, dir[, start voxel]) that belongs to `slot`: the path part goes to the slot,
    // the ray itself stays in this lane's registers.  Returns 1: walking, 2: the (rare) slow path ran to a hit,
    // 0: it left the volume (or never entered the padded volume).
    auto launch_ray = [&](uint32_t slot, uint32_t handle, const float pos[3], const float dir[3], bool has_start, const int32_t sv[3],
                          const float thr[3], uint32_t key, uint32_t meta) -> uint32_t {
        Dda r;
        uint32_t i0;
        const DdaMode mode = dda_init(vol, pos, dir, has_start, sv, r, i0);
        // the sign of len / dir is the step direction (the march needs nothing else); step == 0 travels in the packed signs
        const float raw[3] = {r.step[0] < 0 ? -r.delta[0] : r.delta[0], r.step[1] < 0 ? -r.delta[1] : r.delta[1],
                              r.step[2] < 0 ? -r.delta[2] : r.delta[2]};
        path[slot * 3 + 0] = make_uint4(__float_as_uint(thr[0]), __float_as_uint(thr[1]), __float_as_uint(thr[2]), handle);
        path[slot * 3 + 1] = make_uint4(__float_as_uint(pos[0]), __float_as_uint(pos[1]), __float_as_uint(pos[2]), __float_as_uint(r.len));
        path[slot * 3 + 2] = make_uint4(__float_as_uint(dir[0]), __float_as_uint(dir[1]), __float_as_uint(dir[2]), meta);
        ray[slot * 3 + 1] = make_uint4(__float_as_uint(raw[0]), __float_as_uint(raw[1]), __float_as_uint(raw[2]), pack_signs(r.step));
        ray[slot * 3 + 2] = make_uint4(i0, 0u, key, 0u);
        if (mode == kDdaFast) {
            sx = r.side[0]; sy = r.side[1]; sz = r.side[2];
            dx = r.delta[0]; dy = r.delta[1]; dz = r.delta[2];
            ix = (uint32_t)r.step[0]; iy = (uint32_t)r.step[1] << xb; iz = (uint32_t)r.step[2] << zb;
            idx = i0; rec = i0 | kRecFresh; steps = 0u;
            my_slot = (int)slot;
            stopped = 0u;
            return 1u;
        }
        if (mode == kDdaSlow) {
            dda_slow<kSmem>(vol, r); // rare: a direction component is exactly 0; runs to completion here
            iters += r.steps;
            if (r.hit) {
                // re-express the result in the fast walk's exit format (bit index, previous index, steps)
                const uint32_t hidx = (uint32_t)(r.v[0] + 1) | ((uint32_t)(r.v[1] + 1) << xb) | ((uint32_t)(r.v[2] + 1) << zb);
                const uint32_t jx = (uint32_t)r.step[0], jy = (uint32_t)r.step[1] << xb, jz = (uint32_t)r.step[2] << zb;
                const uint32_t lm = r.steps ? r.last_mask : 0u;
                const uint32_t hprev = hidx - ((lm & 1u) ? jx : 0u) - ((lm & 2u) ? jy : 0u) - ((lm & 4u) ? jz : 0u);
                ray[slot * 3 + 0] = make_uint4(__float_as_uint(r.side[0]), __float_as_uint(r.side[1]), __float_as_uint(r.side[2]), hidx);
                *reinterpret_cast<uint2*>(&ray[slot * 3 + 2]) = make_uint2(hprev, r.steps);
                return 2u;
            }
        }
        return 0u;
    };

    for (;;) {
        // ---- make sure there is a work item with jobs (or learn that the frame is exhausted) ------
        while (work_left && it_next >= it_njobs) {
            const int item = claim_tiles(fb.stats + 2, lane, 1);
#ifdef VT_WAVE_STATS
            if (item >= n_items && !t_exhausted) t_exhausted = wave_timer_ns();
            if (item < n_items && !t_first) t_first = wave_timer_ns();
#endif
            if (item >= n_items) { work_left = false; break; }
            flush_tile();
            int chunk, cy, cx;
            if (fp.item_order == 0u) { // chunk-major over the tiles in row order: a tile's chunks are spread in time
                chunk = item / cov_tiles;
                const int ct = item - chunk * cov_tiles;
                cy = ct / cov_w; cx = ct - cy * cov_w;
            } else {
                // Tile-major, from the centre of the rectangle outwards: the instance fills the middle of its screen rectangle,
                // so the tiles whose paths bounce longest are started first and the frame ends on border tiles, whose paths
                // mostly leave through the sky after one segment — the drain of the last paths is short.  Rings are counted from
                // the outside (ring r = tiles r away from the nearest edge); q counts tiles from the END of the order.
                int ct;
                if (fp.item_order == 1u) { ct = item / n_chunks; chunk = item - ct * n_chunks; }
                else { chunk = item / cov_tiles; ct = item - chunk * cov_tiles; } // (2: chunk-major, every pass from the centre outwards)
                const int q = cov_tiles - 1 - ct;
                int r = 0, before = 0; // tiles in the rings outside ring r
                for (;;) {
                    const int w_in = cov_w - 2 * (r + 1), h_in = own_rows - 2 * (r + 1);
                    const int upto = cov_tiles - (w_in > 0 && h_in > 0 ? w_in * h_in : 0); // tiles in rings 0 .. r
                    if (q < upto) break;
                    before = upto;
                    ++r;
                }
                const int w_r = cov_w - 2 * r, h_r = own_rows - 2 * r; // ring r is the border of a w_r x h_r box at (r, r)
                int k = q - before;
                if (w_r == 1) { cy = r + k; cx = r; }                                          // a single column
                else if (h_r == 1 || k < w_r) { cy = r; cx = r + k; }                          // top edge (or the single row)
                else if (k < 2 * w_r) { cy = r + h_r - 1; cx = r + (k - w_r); }                // bottom edge
                else { k -= 2 * w_r; cy = r + 1 + (k >> 1); cx = (k & 1) ? r + w_r - 1 : r; }  // left / right columns in between
            }
            const int ty = (int)rows.row((uint32_t)cty0, (uint32_t)cy);
            const int tile = ty * tiles_x + (ctx0 + cx);
            it_tile = (uint32_t)tile;
            it_x0 = (ctx0 + cx) * kTileW;
            it_y0 = ty * kTileH;
            const int my_px = it_x0 + (lane & 7), my_py = it_y0 + (lane >> 3);
            const bool in_frame = my_px < fp.width && my_py < fp.height;
            const bool may_hit = in_frame && !(my_px < Ip->bounds[0] || my_px > Ip->bounds[1] || my_py < Ip->bounds[2] || my_py > Ip->bounds[3]);
            const uint32_t cov = __ballot_sync(0xffffffffu, may_hit);
            it_ncov = __popc(cov);
            if (may_hit) cov_pix[__popc(cov & lt_mask)] = (uint8_t)lane;
            uint32_t ns = item_spp;
            it_s0 = (uint32_t)chunk * item_spp;
            if ((uint32_t)chunk >= n_full) { // tail chunk t: what is left after t halvings
                uint32_t rem = tail_spp;
                it_s0 = n_full * item_spp;
                for (uint32_t t = (uint32_t)chunk - n_full; t; --t) { const uint32_t take = (rem + 1) / 2; it_s0 += take; rem -= take; }
                ns = (rem + 1) / 2;
            }
            it_njobs = it_ncov * ns;
            it_next = 0;
            if (it_ncov == 0) it_ncov = 1; // (no jobs; keeps the division below defined)
            it_magic = 0xFFFFFFFFu / it_ncov + 1u; // floor(job / ncov) == umulhi(job, magic) for job * ncov < 2^32
            __syncwarp();
        }
        const bool jobs = work_left && it_next < it_njobs;
        if (n_miss >= 32) sky_batch(); // (keeps room for 32 more entries)

        // ---- what to generate: a full batch if there is one; else march what is parked; else the better partial batch
        const int low = (int)fp.refill_threshold;
        const int avail = jobs ? (int)(it_njobs - it_next) : 0;
        const int can_primary = n_free < avail ? n_free : avail;
        int gen; // 0 nothing, 1 bounce, 2 primary
        if (n_hit >= 32) gen = 1;
        else if (can_primary >= 32 || (can_primary > 0 && can_primary == avail && n_free >= 32)) gen = 2; // (an item's last jobs count as a full batch)
        else if (n_parked >= low) gen = 0;
        else if (n_hit > 0 && n_hit >= can_primary) gen = 1;
        else if (can_primary > 0) gen = 2;
        else if (n_parked > 0) gen = 0;
        else if (n_miss > 0) { sky_batch(); continue; }
        else break;

        if (gen == 2) {
            // ---- primary batch: new camera rays --------------------------------------------------
            const uint32_t n = can_primary < 32 ? (uint32_t)can_primary : 32u;
            bool enters = false;
            uint32_t handle = 0, key = 0, meta = 0;
            float d[3] = {0.0f, 0.0f, 0.0f}, pos[3] = {0.0f, 0.0f, 0.0f};
            if ((uint32_t)lane < n) {
                const uint32_t job = it_next + lane; // sample-major: neighbouring lanes get neighbouring pixels
                const uint32_t si = it_ncov == 1u ? job : __umulhi(job, it_magic), ci = job - si * it_ncov; // (the magic number of 1 is 2^32)
                const uint32_t pix = cov_pix[ci];
                const int px = it_x0 + (int)(pix & 7u), py = it_y0 + (int)(pix >> 3);
                const uint32_t pixel = (uint32_t)py * (uint32_t)fp.width + (uint32_t)px;
                Rng rng;
                rng_init(rng, fp.seed, pixel, fp.sample_first + (it_s0 + si) * fp.sample_stride);
                const float jx = rng_u01(rng), jy = rng_u01(rng);
                const float fx = (float)px + jx, fy = (float)py + jy;
                handle = it_tile << 5 | pix;
                rays += 1;
                // camera ray in the instance's model space: o = eye, d = dirm * (x_ndc, y_ndc, 1)
                const float x_ndc = fx * fp.sxn - 1.0f;
                const float y_ndc = fy * fp.syn - 1.0f;
                float o[3], lo3[3], hi3[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    d[k] = (Ip->dirm[0 * 3 + k] * x_ndc + Ip->dirm[1 * 3 + k] * y_ndc) + Ip->dirm[3 * 3 + k];
                    o[k] = Ip->eye_m[k];
                    lo3[k] = Ip->slab_lo[k];
                    hi3[k] = Ip->slab_hi[k];
                }
                float tn;
                int axis;
                if (!slab_unit_cube(o, lo3, hi3, d, tn, axis)) { // leaves through the sky: no slot needed
#pragma unroll
                    for (int c = 0; c < 3; ++c) atomicAdd(&wacc[pix * 3 + c], (uint32_t)sky_q[c]);
                } else {
                    float mp[3];
                    entry_point(o, d, tn, axis, mp);
#pragma unroll
                    for (int k = 0; k < 3; ++k) pos[k] = (mp[k] + 0.5f) * size[k];
                    enters = true;
                    key = rng.key;
                    meta = (uint32_t)axis << 4 | rng.ctr << 8;
                }
            }
            it_next += n;
            VT_STAT(0, 1); VT_STAT(1, n);
            // the rays that enter the cube pop a slot each
            const uint32_t em = __ballot_sync(0xffffffffu, enters);
            uint32_t res = 1u;
            uint32_t slot = 0;
            if (enters) {
                slot = free_stack[n_free - 1 - __popc(em & lt_mask)];
                const float one[3] = {1.0f, 1.0f, 1.0f};
                const int32_t none[3] = {0, 0, 0};
                res = launch_ray(slot, handle, pos, d, false, none, one, key, meta);
            }
            n_free -= __popc(em);
            __syncwarp();
            push_hit(enters && res == 2u, slot);
            push_miss(enters && res == 0u, slot);
        } else if (gen == 1) {
            // ---- bounce batch: shade the hits, start the next segment ---------------------------
            const int n = n_hit < 32 ? n_hit : 32;
            uint32_t res = 1u; // 0 left the volume, 1 nothing to report (walking / no work), 2 slow-path hit, 3 path ended
            uint32_t slot = 0;
            if (lane < n) {
                slot = hit_stack[n_hit - n + lane];
                const uint4 q0 = ray[slot * 3 + 0], q1 = ray[slot * 3 + 1], q2 = ray[slot * 3 + 2];
                const uint4 p0 = path[slot * 3 + 0], p1 = path[slot * 3 + 1], p2 = path[slot * 3 + 2];
                Dda r;
                r.side[0] = __uint_as_float(q0.x); r.side[1] = __uint_as_float(q0.y); r.side[2] = __uint_as_float(q0.z);
                r.delta[0] = fabsf(__uint_as_float(q1.x)); r.delta[1] = fabsf(__uint_as_float(q1.y)); r.delta[2] = fabsf(__uint_as_float(q1.z));
#pragma unroll
                for (int k = 0; k < 3; ++k) r.step[k] = (int32_t)((q1.w >> (2 * k)) & 3u) - 1;
                r.pos[0] = __uint_as_float(p1.x); r.pos[1] = __uint_as_float(p1.y); r.pos[2] = __uint_as_float(p1.z);
                r.len = __uint_as_float(p1.w);
                r.dir[0] = __uint_as_float(p2.x); r.dir[1] = __uint_as_float(p2.y); r.dir[2] = __uint_as_float(p2.z);
                dda_finish_fast(vol, r, q0.w, q2.x, q2.y);
                const uint32_t handle = p0.w;
                uint32_t bounce = p2.w & 15u;
                const int entry_axis = (int)((p2.w >> 4) & 3u);
                Rng rng{q2.z, p2.w >> 8};
                float thr[3] = {__uint_as_float(p0.x), __uint_as_float(p0.y), __uint_as_float(p0.z)};
                const uchar4 s = fetch_texel(Ip->rgba, Ip->w, Ip->h, Ip->d, Ip->remap_identity != 0, r.v);
                thr[0] = thr[0] * dec[s.x];
                thr[1] = thr[1] * dec[s.y];
                thr[2] = thr[2] * dec[s.z];
                if (bounce == fp.bounces) {
                    res = 3u; // path length exhausted: contributes nothing
                } else {
                    ++bounce;
                    const uint32_t lm = r.steps ? r.last_mask : (1u << entry_axis);
                    const int a = (lm & 1u) ? 0 : ((lm & 2u) ? 1 : 2);
                    float npos[3], ndir[3];
                    int32_t nsv[3];
                    rng_sphere(rng, ndir);
                    int nsign = 0;
                    const float t = r.steps ? ((a == 0 ? r.side[0] : (a == 1 ? r.side[1] : r.side[2])) -
                                               (a == 0 ? r.delta[0] : (a == 1 ? r.delta[1] : r.delta[2])))
                                            : 0.0f;
                    const float tl = t / r.len;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        float p = r.pos[k] + r.dir[k] * tl;
                        const float lo = (float)r.v[k], hi = (float)(r.v[k] + 1);
                        p = p < lo ? lo : p;
                        p = p > hi ? hi : p;
                        npos[k] = p;
                        nsv[k] = r.v[k];
                        if (k == a) {
                            nsign = r.step[k] != 0 ? -r.step[k] : (r.pos[k] <= 0.5f * size[k] ? -1 : 1);
                            npos[k] = (float)(r.v[k] + (nsign > 0 ? 1 : 0));
                            nsv[k] += nsign;
                            ndir[k] += (float)nsign;
                        }
                    }
                    const float l2 = (ndir[0] * ndir[0] + ndir[1] * ndir[1]) + ndir[2] * ndir[2];
                    if (l2 < 1e-6f) {
#pragma unroll
                        for (int k = 0; k < 3; ++k) ndir[k] = (k == a) ? (float)nsign : 0.0f;
                    } else {
                        const float rl = 1.0f / sqrtf(l2);
                        ndir[0] *= rl; ndir[1] *= rl; ndir[2] *= rl;
                    }
                    rays += 1;
                    res = launch_ray(slot, handle, npos, ndir, true, nsv, thr, rng.key, bounce | (uint32_t)a << 4 | rng.ctr << 8);
                }
            }
            VT_STAT(2, 1); VT_STAT(3, n);
            n_hit -= n;
            __syncwarp(); // (the batch's reads of the hit stack precede the pushes below)
            push_hit(res == 2u, slot);
            push_free(res == 3u, slot);
            push_miss(res == 0u, slot);
        }

        // ---- march: step the fresh rays, pull the parked ones into idle lanes, write exits back ------
        {
            if (n_miss >= 32) sky_batch(); // (the generation above may have pushed misses)
            const int total = n_parked;
            int rc = 0;
            uint32_t idle_mask = __ballot_sync(0xffffffffu, my_slot < 0); // lanes without a ray
            uint32_t it32 = 0;
            int thresh = -1;
            VT_STAT(6, 1); VT_STAT(7, total + 32 - __popc(idle_mask));
            for (;;) {
                if (rc < total && idle_mask) { // hand the parked rays to the idle lanes
                    const int rank = __popc(idle_mask & lt_mask);
                    if (my_slot < 0 && rc + rank < total) {
                        my_slot = parked[rc + rank];
                        const uint4 q0 = ray[my_slot * 3 + 0], q1 = ray[my_slot * 3 + 1];
                        const uint2 q2 = *reinterpret_cast<const uint2*>(&ray[my_slot * 3 + 2]);
                        sx = __uint_as_float(q0.x); sy = __uint_as_float(q0.y); sz = __uint_as_float(q0.z); idx = q0.w;
                        dx = fabsf(__uint_as_float(q1.x)); dy = fabsf(__uint_as_float(q1.y)); dz = fabsf(__uint_as_float(q1.z));
                        // step direction = sign of len / dir (never 0 on the fast path): +-1 in the axis' field of the index
                        ix = (uint32_t)(((int32_t)q1.x >> 31) * 2 + 1);
                        iy = (uint32_t)(((int32_t)q1.y >> 31) * 2 + 1) << xb;
                        iz = (uint32_t)(((int32_t)q1.z >> 31) * 2 + 1) << zb;
                        rec = q2.x | kRecFresh;
                        steps = q2.y;
                        stopped = 0u;
                    }
                    const int want = __popc(idle_mask);
                    rc += want < total - rc ? want : total - rc;
                    idle_mask = __ballot_sync(0xffffffffu, my_slot < 0);
                }
                const int nact = 32 - __popc(idle_mask);
                if (!nact) { n_parked = 0; break; }
                if (thresh < 0) thresh = nact < low ? nact : low;
                if (rc >= total && nact < thresh) {
                    // too few lanes left: park them and go generate more rays
                    __syncwarp(); // (every read of the parked list precedes its rewrite)
                    if (my_slot >= 0) {
                        ray[my_slot * 3 + 0] = make_uint4(__float_as_uint(sx), __float_as_uint(sy), __float_as_uint(sz), idx);
                        *reinterpret_cast<uint2*>(&ray[my_slot * 3 + 2]) = make_uint2(rec & kRecIdxMask, steps);
                        parked[__popc(~idle_mask & lt_mask)] = (uint8_t)my_slot;
                        my_slot = -1;
                        idx = 0;
                        stopped = 1u;
                    }
                    n_parked = nact;
                    VT_STAT(10, nact);
                    break;
                }
                // Step until it pays to look at the stopped lanes: while parked rays remain, once `refill_batch`
                // lanes can be refilled together; afterwards, once fewer than `thresh` lanes still walk (the rest
                // is then parked).  Lanes without a ray count as stopped.
                uint32_t k_stop = rc < total ? (uint32_t)(32 - nact) + fp.refill_batch : (uint32_t)(33 - thresh);
                k_stop = k_stop > 32u ? 32u : k_stop;
                const uint32_t v = wave_walk<kSmem>(vol, sx, sy, sz, dx, dy, dz, idx, rec, steps, stopped, ix, iy, iz, k_stop, smem0);
                VT_STAT(8, 1); VT_STAT(9, __popc(v & ~idle_mask)); VT_STAT(11, nact);
                // the rays that ended: on a filled voxel (hit) or on the border (left the volume)
                const bool fin = stopped && my_slot >= 0;
                bool hit = false;
                if (fin) {
                    const uint32_t st = steps + (((rec >> kWaveIdxBits) + 1u) & (kWaveBurst - 1u));
                    const uint32_t vx = (idx & ((1u << xb) - 1u)) - 1u, vy = ((idx >> xb) & ((1u << vol.yb) - 1u)) - 1u, vz = (idx >> zb) - 1u;
                    hit = vx < vol.w && vy < vol.h && vz < vol.d;
                    if (hit) {
                        ray[my_slot * 3 + 0] = make_uint4(__float_as_uint(sx), __float_as_uint(sy), __float_as_uint(sz), idx);
                        *reinterpret_cast<uint2*>(&ray[my_slot * 3 + 2]) = make_uint2(rec & kRecIdxMask, st);
                    }
                    it32 += st;
#ifdef VT_WAVE_STATS
                    { // histogram of ray lengths: 0, 1, 2, 3, 4-7, 8-15, 16-31, 32-63, 64+
                        const int b = st < 4 ? (int)st : (st < 8 ? 4 : (st < 16 ? 5 : (st < 32 ? 6 : (st < 64 ? 7 : 8))));
                        for (int k = 0; k < 9; ++k) {
                            const uint32_t m = __ballot_sync(__activemask(), b == k);
                            wstat[12 + k] += (lane == (__ffs(__activemask()) - 1)) ? __popc(m) : 0;
                        }
                    }
#endif
                }
                push_hit(fin && hit, (uint32_t)my_slot);
                push_miss(fin && !hit, (uint32_t)my_slot);
                if (fin) { my_slot = -1; idx = 0; }
                idle_mask = v; // every stopped lane is idle now
                if (n_miss >= 32) sky_batch();
            }
            iters += it32;
            __syncwarp();
        }
    }

    flush_tile();
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        rays += __shfl_xor_sync(0xffffffffu, rays, o);
        iters += __shfl_xor_sync(0xffffffffu, iters, o);
        analytic += __shfl_xor_sync(0xffffffffu, analytic, o);
    }
#ifdef VT_WAVE_STATS
    if (lane == 0) {
        for (int i = 0; i < 32; ++i) atomicAdd(&vt_wave_stats[i], (unsigned long long)wstat[i]);
        const unsigned long long t_end = wave_timer_ns();
        auto bucket = [&](unsigned long long t) { const unsigned long long b = (t - t_start) / 5000ull; return (int)(b > 63 ? 63 : b); };
        atomicAdd(&vt_wave_times[bucket(t_exhausted ? t_exhausted : t_end)], 1u);
        atomicAdd(&vt_wave_times[64 + bucket(t_end)], 1u);
        atomicAdd(&vt_wave_times[128 + bucket(t_first ? t_first : t_end)], 1u);
    }
#endif
    if (lane == 0) {
        if (rays) atomicAdd(fb.stats + 0, rays);
        if (iters) atomicAdd(fb.stats + 1, iters);
        if (analytic) atomicAdd(fb.stats + 3, analytic);
    }
}

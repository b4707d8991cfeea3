/*
 * vtrace_abi.h — C ABI of librender (B200 / CUDA backend for vtrace's voxel tracing path).
 *
 * Part 1 is EXACTLY the surface the reference's Rust engine binds in
 * src/render.rs:110-128 and today links from lib/entry.c + lib/memory.c (static librender.a,
 * build.rs:71-72,96-97).  Same symbol names, argument meaning, return conventions, ownership
 * and threading rules, so this library is a drop-in for that path.  The Vulkan window,
 * swapchain and present are replaced by a headless framebuffer that Part 2 reads back.
 *
 * Part 2 (vt_*) are extension symbols the Rust engine does not know about: headless
 * configuration, framebuffer read-back, statistics and the hooks a multi-process launcher
 * (one process per GPU) needs to reduce the accumulation buffers with NCCL.
 *
 * Plain C: fixed-width integers, raw pointers and sizes only.  No CUDA, torch or C++ types.
 * Every entry point is callable from any OS thread (the reference calls from two,
 * src/main.rs:40-50), never concurrently; each one selects its CUDA device itself.
 * Nothing unwinds across the boundary; errors are return codes.
 */
#ifndef VTRACE_ABI_H
#define VTRACE_ABI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ======================================================================================== */
/* Part 1 — the reference's FFI, unchanged                                                   */

/* lib/common.h:89-92  <->  RenderTickInfo, src/render.rs:177-181.
 * Both point at 16 floats, column-major mat4 (glm-rs layout): projection then camera,
 * the two push constants of lib/command.c:97-98.  Valid only during the call. */
typedef struct render_tick_info {
    void* perspective;
    void* camera;
} render_tick_info;

/* lib/common.h:145-151  <->  UserInput, src/render.rs:37-51 (40 bytes). */
typedef struct user_input {
    uint8_t keys[6]; /* w a s d space lshift */
    double mouse_x;
    double mouse_y;
    double last_mouse_x;
    double last_mouse_y;
} user_input;

/* replaces lib/entry.c:54-57 (+ init, :59-98).  One-time initialisation: selects the CUDA
 * device, creates the stream, framebuffers and tables, reads VT_* environment overrides
 * (see vt_config).  Returns 0 on success; non-zero = (cudaError << 32) | custom code, which
 * the Rust side turns into a panic (src/render.rs:185-188). */
uint64_t entry(void);

/* replaces lib/entry.c:150-248.  Renders one frame with the given projection / camera into
 * the headless framebuffer, then writes the framebuffer size to `*window_width` and `*window_height`
 * (lib/entry.c:244-245; the engine derives its aspect ratio from them, src/render.rs:282-292).
 * Returns 0 = keep going, -1 = stop (configured frame budget exhausted) or error. */
int32_t render_tick(int32_t* window_width, int32_t* window_height, const render_tick_info* info);

/* replaces lib/entry.c:146-148.  Pointer into library-owned static storage, valid for the
 * life of the process.  Headless: all zero unless scripted through vt_set_user_input. */
user_input* get_input_data_pointer(void);

/* replaces lib/memory.c:286-385.  Copies 4*width*height*depth bytes of RGBA8 (`Color`,
 * src/voxel/common.rs:65-72; x fastest as the Vulkan copy reads it, lib/memory.c:353-366)
 * before returning, uploads them and builds the volume's traversal masks (the slot that
 * lib/raytrace.c's never-consumed BLAS build occupies).  Returns the dense texture id
 * (0, 1, 2, ... lib/memory.c:292,384) or -1 (MAX_TEXTURES reached, lib/memory.c:287-290,
 * or device error). */
int32_t add_texture(const uint8_t* data, uint32_t width, uint32_t height, uint32_t depth);

/* Extension (no reference counterpart): volumes too large for add_texture's
 * dense RGBA8 upload (its u32 byte count overflows at 1024^3, lib/memory.c:297) are generated on the
 * device from a procedural definition and stored as sparse 8^3 occupancy bricks.  Returns a texture
 * id usable in instances exactly like add_texture's, or -1.  Dimensions: multiples of 8, <= 8192. */
#define VT_VOLUME_HEIGHTMAP 1u      /* terrain: altitude <= H/2 + H/4 * fbm(value noise, 5 octaves)          */
#define VT_VOLUME_SPARSE_BRICKS 2u  /* 2 % of the 8^3 bricks non-empty (hash), half of their voxels filled */
int32_t vt_add_volume_procedural(uint32_t kind, uint32_t width, uint32_t height, uint32_t depth, uint32_t seed);

/* Extension: a caller-supplied sparse volume.  n_bricks bricks of 8^3 voxels: brick_coords = n x 3
 * uint32 (voxel coordinate / 8), masks = n x 16 uint32 (voxel (x,y,z) of a brick is bit
 * (x | (y&3) << 3) of word ((z&7) << 1 | (y&7) >> 2)), colors = n x RGBA8 (one colour per brick; filled
 * voxels are opaque).  The arrays are only borrowed for the call.  Returns a texture id or -1. */
int32_t vt_add_volume_bricks(const uint32_t* brick_coords, const uint32_t* masks, const uint8_t* colors, size_t n_bricks,
                             uint32_t width, uint32_t height, uint32_t depth);

/* replaces lib/memory.c:235-248.  Returns a write pointer with room for max(instance_count,1)
 * 64-byte column-major mat4s (texture id bit-cast into element [3][3], src/render.rs:74-78),
 * valid until end_update_instances; NULL on failure. */
float* start_update_instances(uint32_t instance_count);

/* replaces lib/memory.c:250-267.  Commits the first instance_count instances (0 is coerced
 * to 1, lib/memory.c:251).  0 = ok, -1 = error. */
int32_t end_update_instances(uint32_t instance_count);

/* replaces lib/entry.c:100-144.  Waits for the device, frees everything. */
void cleanup(void);

/* ======================================================================================== */
/* Part 2 — headless extensions (new symbols; the Rust engine never calls them)              */

#define VT_MISS 0xFFFFFFFFu

/* render modes */
#define VT_MODE_PRIMARY 0u /* the reference's pass: one DDA per covered fragment (trace.frag)  */
#define VT_MODE_PATHS 1u   /* extension: spp jittered paths per pixel, diffuse bounces, sky     */
#define VT_MODE_RAYS 2u    /* extension: width*height incoherent rays through instance 0's volume (config 4) */

/* flags */
#define VT_FLAG_VIEWPORT_H_IS_W 1u /* reproduce lib/command.c:80-81 (viewport height = width)  */
#define VT_FLAG_NO_HIT_RECORDS 2u  /* do not write the per-pixel hit records                    */
#define VT_FLAG_FORCE_GLOBAL_MASKS 4u /* keep traversal masks in global memory (no smem staging) */
/* (bit 8u was the round-1 persistent-lane schedule, removed: the wavefront engine superseded it) */
#define VT_FLAG_SHADOW_RAYS 64u     /* PRIMARY: one shadow ray towards the sun (0.4,-0.8,0.45) per winning fragment  */
#define VT_FLAG_NO_BINNING 32u      /* visit every instance per pixel instead of the screen-space bins (testing) */
#define VT_FLAG_PER_PIXEL_PATHS 16u /* PATHS, one instance: the general per-pixel kernel instead of the wavefront engine */

/* Per-pixel derived hit record (not an output of the reference). 16 bytes. */
typedef struct vt_hit_record {
    uint32_t hit_voxel; /* X + W*(Y + H*Z) of model_ray_voxel (trace.frag:68,85) at the hit; VT_MISS  */
    uint32_t packed;    /* bits 0-15: `steps` (trace.frag:73,86) at the hit; bits 16-18: axes advanced */
                        /* by the last executed iteration (`mask`, :83), or the proxy-entry axis when  */
                        /* steps == 0; bits 19-21: model_ray_step < 0 per axis (:69); 0 on miss        */
    uint32_t instance;  /* gl_InstanceIndex (trace.vert:37) of the winning fragment; VT_MISS          */
    uint32_t iters;     /* DDA iterations executed by all fragments covering the pixel (late-Z)        */
} vt_hit_record;

typedef struct vt_config {
    uint32_t width, height;   /* framebuffer; the reference hard-codes 1000x1000, lib/entry.c:62 */
    uint32_t mode;            /* VT_MODE_*                                                        */
    uint32_t flags;           /* VT_FLAG_*                                                        */
    uint32_t spp;             /* PATHS: samples per pixel rendered BY THIS PROCESS per frame      */
    uint32_t bounces;         /* PATHS: diffuse bounces after the primary segment                 */
    uint32_t seed;            /* PATHS: RNG seed                                                  */
    uint32_t sample_first;    /* PATHS: global index of this process's first sample (= rank);     */
                              /* RAYS: this process traces rays [sample_first*w*h, (sample_first+1)*w*h) */
    uint32_t sample_stride;   /* PATHS: distance between its samples (= number of ranks)          */
    uint32_t total_spp;       /* PATHS: spp over all ranks, the divisor used by vt_resolve        */
    int32_t max_frames;       /* render_tick returns -1 after this many frames; <= 0: never       */
    int32_t device;           /* read-only: CUDA device chosen by entry() (VT_DEVICE, LOCAL_RANK, 0) */
} vt_config;

typedef struct vt_stats {
    uint64_t frames;          /* frames rendered since entry()                                    */
    uint64_t rays;            /* last frame: ray segments traced (primary fragments + bounces)    */
    uint64_t iterations;      /* last frame: DDA loop iterations (= sum of `steps`)               */
    uint64_t launches;        /* kernels launched since entry()                                   */
    float last_trace_ms;      /* last frame: device time of the trace kernel (CUDA events)        */
    float last_frame_ms;      /* last frame: device time of everything render_tick enqueued       */
    uint32_t masks_in_smem;   /* 1 when the traversal masks were staged in shared memory          */
    uint32_t trace_frames;    /* frames folded into trace_ms_sum since the previous vt_get_stats  */
    float trace_ms_sum;       /* sum of the trace kernel's device time over those frames          */
    uint32_t bin_list_grown;  /* times the instance-bin list was too small (such a frame is exact but slower) and was grown */
    uint64_t analytic_rays;   /* last frame, PATHS: the part of `rays` that was never marched — camera samples of  */
                              /* pixels outside every instance's screen rectangle, resolved as spp x sky           */
    uint64_t rays_sum;        /* `rays`, `iterations`, `analytic_rays` summed over the frames folded since the     */
    uint64_t iterations_sum;  /* previous vt_get_stats (frames may differ: the sums are not last-frame x frames)    */
    uint64_t analytic_rays_sum;
} vt_stats;

/* Current configuration (defaults: 1000x1000, PRIMARY, as lib/entry.c:62). */
int32_t vt_get_config(vt_config* out);
/* Applies a configuration; (re)allocates framebuffers.  0 = ok, -1 = error. */
int32_t vt_configure(const vt_config* cfg);

/* Enqueue one frame without host synchronisation or read-back (device-resident measurement);
 * projection / camera as in render_tick_info.  Up to 64 frames may be in flight; their counters and
 * kernel timings are folded into vt_stats at the next synchronising call.  0 = ok. */
int32_t vt_render_async(const float* projection, const float* camera);
/* render_tick without the wait: a whole frame (clear, trace, resolve into the colour buffer) enqueued on the
 * library's stream, so that a host that pipelines its frames never stalls the device between them.  0 = ok. */
int32_t vt_render_frame_async(const float* projection, const float* camera);
/* Block until everything enqueued so far has finished.  0 = ok. */
int32_t vt_synchronize(void);

/* Read back the last frame.  `capacity` in bytes; returns bytes written or -1. */
int64_t vt_read_hits(vt_hit_record* out, size_t capacity);
int64_t vt_read_color(uint8_t* rgba8, size_t capacity); /* R,G,B,A bytes, sRGB-encoded         */
/* Pipelined read-back: enqueue the copy of the frame enqueued last (vt_render_frame_async / render_tick / vt_resolve) into
 * PAGE-LOCKED host memory on the library's copy stream and return at once; the next frame renders into a second colour
 * buffer, so tracing frame k+1 overlaps the PCIe transfer of frame k.  The destination is valid after vt_read_color_wait
 * (host blocks until every read-back issued so far has landed).  vt_read_color_fence makes the library's STREAM wait for
 * them instead (no host wait), e.g. to time a pipelined loop with events.  The transfer itself starts when the next
 * frame's trace kernel starts, or at the next vt_read_color_wait / _fence / _async, whichever comes first: while a
 * device-to-host copy saturates PCIe the GPU's command fetches queue behind it, so a copy that runs under the set-up
 * kernels of the next frame delays them, and one that runs under its long trace kernel delays nothing
 * (VT_DEFER_READBACK=0 starts it at once).  Returns the byte count, -1 on error. */
int64_t vt_read_color_async(uint8_t* pinned_rgba8, size_t capacity);
int32_t vt_read_color_wait(void);
int32_t vt_read_color_fence(void);

/* The same frame in the reference swapchain's byte order, VK_FORMAT_B8G8R8A8_SRGB (lib/swapchain.c:88). */
int64_t vt_read_color_bgra(uint8_t* bgra8, size_t capacity);
/* Writes the last frame as a binary PPM (P6, RGB); 0 = ok.  For eyeballing / diffing against a real run of the reference. */
int32_t vt_write_ppm(const char* path);
int64_t vt_read_depth(float* depth, size_t capacity);   /* D32 depth buffer (proxy-face depth) */
int64_t vt_read_accum(uint64_t* accum, size_t capacity); /* PATHS: 3 x u64 per pixel, 2^-24 fixed point */

/* PATHS multi-process plumbing: the accumulation buffer lives in device memory; a launcher
 * reduces it across ranks (NCCL sum over 3*w*h uint64) and then resolves it to colour. */
void* vt_accum_device_ptr(void);
/* Use caller-owned device memory (3 * width * height uint64, e.g. a launcher's tensor that it
 * hands to NCCL) as the accumulation buffer; NULL returns to the library's own. */
int32_t vt_set_accum_buffer(void* device_ptr);
int32_t vt_clear_accum(void);
int32_t vt_resolve(void);
/* Fused cross-GPU accumulation (one process per GPU, one node) — replaces the all-reduce of the
 * accumulation buffers.  Rank 0 (the root) allocates a partial-sum buffer with one slot per rank and
 * exports it as a 64-byte CUDA IPC handle; the other ranks import it.  Each frame every rank traces
 * its samples into its own accumulators, then a small kernel streams the pixels of the instance's
 * screen rectangle — the only ones that can differ from "spp x sky" — into the rank's slot in the
 * root's memory with vector stores over NVLink (16 bytes per pixel while a rank's sums fit 32 bits,
 * i.e. up to 255 samples; 32 bytes otherwise) and clears them locally, and raises the rank's
 * sequence-number flag in the root's memory.  The root's vt_resolve (or vt_read_accum) waits for every
 * rank's flag, sums the slots (integers: bit-identical to an all-reduce) and encodes the frame; a rank
 * reuses a half of the double buffer only after the root has moved on from the frame that used it.
 * No collective and no host synchronisation between the ranks.  Per frame, on every rank:
 * vt_fused_reduce_next_frame, vt_render_async; then vt_resolve on the root — for every frame, and before
 * the root enqueues its next one (starting a frame is what tells the other ranks, which run at most two
 * frames ahead, that the previous frame's half of the buffer may be refilled).  A rank that does not
 * arrive within 10 s makes the next synchronising call fail instead of hanging the GPU.  Environment VT_FUSED_SYNC=0 (read at export /
 * import) turns the flags off; the launcher must then put a stream-ordered barrier between
 * vt_render_async and the root's vt_resolve.  Single-instance PATHS scenes only. */
int32_t vt_fused_reduce_export(uint8_t handle[64], uint32_t world);
int32_t vt_fused_reduce_import(const uint8_t handle[64], uint32_t rank, uint32_t world);
int32_t vt_fused_reduce_next_frame(void);
int32_t vt_fused_reduce_disable(void);
/* How the ranks of a fused reduction share a frame.  0 (default): by samples — rank r traces samples
 * sample_first + k * sample_stride of every pixel, and the root adds the slots up.  1: by rows of 8x4-pixel tiles —
 * rank r traces EVERY sample (configure sample_first 0, sample_stride 1, spp = total_spp) of the tile rows
 * ty = r (mod world), pushes only those rows, and the root takes each pixel from its owner's slot: per-pixel work
 * (camera set-up, accumulator traffic, NVLink bytes, the root's summation) is divided by the number of ranks instead
 * of being repeated on each.  The image is the same, bit for bit.  root_relief_num / root_relief_den (den 1 .. 64,
 * num < den, the same on every rank): the root, which also sums the slots and encodes the frame, is dealt den - num tile
 * rows for every den of another rank.  Call on every rank after export / import. */
int32_t vt_fused_reduce_partition(uint32_t by_tile_rows, uint32_t root_relief_num, uint32_t root_relief_den);

/* Run on a caller-provided cudaStream_t (e.g. the launcher's current stream); NULL = own. */
int32_t vt_set_stream(void* cuda_stream);

int32_t vt_get_stats(vt_stats* out);
/* Measured roofline denominators on the device entry() chose (streaming kernels, best of 4, CUDA events):
 * kind 0 = L2 read GB/s (32 MiB resident buffer, L1 bypassed), 1 = shared-memory read GB/s (LDS.128),
 * 2 = HBM read GB/s (2 GiB buffer).  The .vox scenes are cache-resident: their algorithmic
 * bandwidth is reported against these next to the HBM figure.  0 = ok. */
int32_t vt_measure_peak(uint32_t kind, double* gb_per_s);
int32_t vt_set_user_input(const user_input* in);
/* Last error as text (static storage). */
const char* vt_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* VTRACE_ABI_H */

/*
 * vtrace_oracle.h — CPU ORACLE for the vtrace voxel ray-traversal hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the shipped product (vtrace_b200/,
 * include/, librender) may include, link or call this.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * use it, and only as the checker / CPU baseline.
 *
 * What it is: a C restatement of the reference's GPU programs
 *   shaders/trace.vert:32-47, shaders/trace.frag:41-90
 * plus the fixed-function state that decides which fragment wins
 *   lib/pipeline.c:114-152 (cull/depth/blend), lib/command.c:56-102 (clear,
 *   viewport h = w, scissor), lib/descriptor.c:97-117 (NEAREST, clamp),
 *   lib/memory.c:317 (R8G8B8A8_SRGB), lib/memory.c:22-40 (unit cube proxy).
 *
 * Pinning status: the reference ships no golden vectors for this path
 *  and cannot be built or run here (no Vulkan/GLFW/rustc).
 * The traversal core (vo_frag_main) is pinned instead against the reference's
 * own compiled shader binary shaders/trace.frag.spv, executed by
 * tools/spirv_interp.py; the resulting vectors live in tests/golden/.  The
 * .vox loader restatement is pinned by the SHA-256s of SURVEY.md §A.4.
 * The rasteriser (pixel -> proxy-face point) and everything path-tracing
 * (bounces, RNG, accumulation) have no reference counterpart: "parity
 * unpinned" for those parts; they are DEFINED here.
 */
#ifndef VTRACE_ORACLE_H
#define VTRACE_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VO_MISS 0xFFFFFFFFu

/* flags */
#define VO_FLAG_VIEWPORT_H_IS_W 1u /* reference-faithful viewport (lib/command.c:80-81) */
#define VO_FLAG_SHADOW_RAYS 64u    /* extension: one shadow ray towards the sun per primary hit */
#define VO_FLAG_SUBSET_8 0x10000u  /* oracle only: compute every 8th pixel in x and y, leave the rest untouched */

/* procedural volume kinds of the large-scene extension */
#define VO_VOLUME_HEIGHTMAP 1u
#define VO_VOLUME_SPARSE_BRICKS 2u
#define VO_VOLUME_BRICKS 3u /* caller-supplied bricks */

/* Per-pixel derived hit record. 16 bytes. */
typedef struct vo_hit_record {
    uint32_t hit_voxel; /* X + W*(Y + H*Z) of model_ray_voxel at the hit, VO_MISS otherwise   */
    uint32_t packed;    /* bits 0-15 steps at hit, bits 16-18 mask of last executed step,     */
                        /* bits 19-21 (step<0) per axis; 0 on miss                            */
    uint32_t instance;  /* gl_InstanceIndex of the winning fragment, VO_MISS otherwise        */
    uint32_t iters;     /* DDA loop iterations executed by ALL fragments covering this pixel  */
} vo_hit_record;

typedef struct vo_scene vo_scene;

vo_scene* vo_scene_create(void);
void vo_scene_destroy(vo_scene*);
/* mirrors add_texture (lib/memory.c:286): copies 4*w*h*d bytes, returns id or -1 */
int32_t vo_add_texture(vo_scene*, const uint8_t* rgba, uint32_t w, uint32_t h, uint32_t d);
/* extension: a procedural volume (no texel storage); returns its texture id or -1 */
int32_t vo_add_volume_procedural(vo_scene*, uint32_t kind, uint32_t w, uint32_t h, uint32_t d, uint32_t seed);
/* extension: caller-supplied sparse volume: n bricks with coordinates (voxel / 8), 16 occupancy words and one RGBA colour each */
int32_t vo_add_volume_bricks(vo_scene*, const uint32_t* coords, const uint32_t* masks, const uint8_t* colors, uint64_t n,
                             uint32_t w, uint32_t h, uint32_t d);
/* mirrors start/end_update_instances (lib/memory.c:235-267): n x 16 floats, column-major,
 * texture id bit-cast into element [3][3] (src/render.rs:74-78). n == 0 is coerced to 1. */
void vo_set_instances(vo_scene*, const float* mats, uint32_t n);

/* One frame of primary rays = one vkCmdDrawIndexed (lib/command.c:102).
 * P, V: 16 floats column-major (push constants, lib/command.c:97-98).
 * Any output pointer may be NULL.  rgba8 is R,G,B,A byte order, sRGB-encoded.
 * Returns total DDA iterations (sum of records[].iters). */
uint64_t vo_render_primary(const vo_scene*, const float* P, const float* V, int width, int height,
                           uint32_t flags, vo_hit_record* records, uint8_t* rgba8, float* depth,
                           int num_threads);

/* Path-tracing extension (not in the reference; defined in DESIGN.md §3).
 * Renders samples s = sample_first + k*sample_stride, k in [0, sample_count).
 * accum: 3 x uint64 per pixel, fixed-point 2^-24 radiance sums (added to, not cleared).
 * stats[0] += ray segments traced, stats[1] += DDA iterations. */
void vo_render_paths(const vo_scene*, const float* P, const float* V, int width, int height,
                     uint32_t flags, uint32_t bounces, uint32_t seed, uint32_t sample_first,
                     uint32_t sample_stride, uint32_t sample_count, uint64_t* accum,
                     uint64_t* stats, int num_threads);

/* shadow rays traced by the most recent vo_render_primary (VO_FLAG_SHADOW_RAYS) */
uint64_t vo_last_shadow_rays(void);

/* Incoherent-ray extension (config 4): rays `first` .. `first + n` through instance 0's volume.
 * Returns the total DDA iterations. */
uint64_t vo_render_rays(const vo_scene*, uint64_t n, uint64_t first, uint32_t seed, vo_hit_record* records, uint8_t* rgba8,
                        int num_threads);

/* accum -> RGBA8 (sRGB-encoded), dividing by total_spp */
void vo_resolve(const uint64_t* accum, int width, int height, uint32_t total_spp, uint8_t* rgba8);

/* Run ONLY trace.frag's main() (shaders/trace.frag:41-90) for one fragment with explicit
 * varyings; used to pin the restatement against the SPIR-V interpreter's vectors.
 * out[0]=hit(0/1) out[1..3]=voxel out[4]=steps out[5]=last mask bits; color: 4 floats. */
void vo_frag_main(const float* P, const float* V, const float* M, const float* screen_position,
                  const float* model_position, const uint8_t* rgba, uint32_t w, uint32_t h,
                  uint32_t d, int32_t* out, float* color, float* frag_depth);

/* Uniform matrices derived per frame / per instance; exposed so tests can compare the
 * product's host-side derivation bit for bit. out: 16 floats. */
void vo_mat4_inverse(const float* m, float* out);
void vo_mat4_mul(const float* a, const float* b, float* out);

/* .vox loader restatement (src/voxel/magica_voxel.rs:18-44 over dot_vox 4.1.0).
 * Writes the first model as the reference's RawDynamicChunk<Color> bytes (z fastest,
 * src/voxel/rawchunk.rs:292) = exactly what add_texture receives.  Returns 0 on success.
 * Call with out == NULL to query dims only. */
int vo_load_vox(const char* path, uint32_t dims[3], uint8_t* out, uint64_t out_capacity);
/* the same for model `model` of a multi-model file (one per SIZE / XYZI pair, file order; -6: no such model).  A file
 * without RGBA chunk gets MagicaVoxel's published default palette (unpinned: dot_vox's own copy is not available offline). */
int vo_load_vox_model(const char* path, uint32_t model, uint32_t dims[3], uint8_t* out, uint64_t out_capacity);

/* sRGB helpers shared by resolve and blend (exposed for tests) */
float vo_srgb_decode(uint8_t c);
uint8_t vo_srgb_encode(float linear);

/* texels of volume `tex` (any kind) in a box, x fastest, 4 bytes each (tests: procedural / brick volume -> dense texture) */
int vo_read_texels(const vo_scene*, uint32_t tex, uint32_t x0, uint32_t y0, uint32_t z0, uint32_t nx, uint32_t ny, uint32_t nz,
                   uint8_t* rgba);

int vo_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif

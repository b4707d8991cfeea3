// kernels.h — device-side data layout and the launch wrappers render_abi.cu calls.
//
// Data layout in HBM:
//   * volume texels : dense RGBA8, x fastest, exactly the bytes add_texture received
//                     (lib/memory.c:353-366).  Read ONCE per ray, at the hit.
//   * stop masks    : one bit per voxel of the volume padded by a one-voxel border,
//                     bit = 1 when the DDA must stop there (texel alpha > 0, or border =
//                     the ray left the volume).  Row / plane strides are powers of two so
//                     a voxel's bit index is x' | y' << xb | z' << (xb + yb).  All volumes
//                     live in ONE arena so a single TMA bulk copy stages every mask of a
//                     small scene into shared memory.
//   * instance table: per-instance uniforms produced by the instance-setup kernel (the
//                     trace.vert replacement).
//   * framebuffer   : 16-byte hit records, RGBA8 colour, D32 depth, 3 x u64 accumulators.
#pragma once

#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace vt {

struct BrickVolume;

struct VolumeDesc {
    const uint8_t* rgba; // dense W*H*D RGBA8
    uint32_t w, h, d;
    uint32_t xb, yb;     // log2(row bits), log2(rows per plane) of the padded stop mask
    uint32_t mask_off;   // word offset of this volume's mask inside the arena
    uint32_t mask_words; // multiple of 4 (16-byte granules for the bulk copy)
    uint32_t remap_identity; // 1 when floor(fl(v/s)*s) == v on all three axes (texel == voxel)
    const BrickVolume* bricks; // non-null: a procedural brick volume (rgba / mask fields unused)
};

// Large procedural volumes (extension; §8f rank 2): occupancy only, as a
// two-level sparse structure; colours are a function of the voxel position.
//   codes : one word per 16 consecutive bricks of the padded brick grid = 16 two-bit codes (3 = the brick has at least one
//           filled voxel, 2 = it has none but a brick of its 3x3x3 neighbourhood has or is outside, 1 = the brick lies in the
//           one-brick border = outside the volume, 0 = nothing within one brick: the walk needs no lookup for a whole burst)
//   base  : per such entry, the number of 3-codes in all earlier entries
//   pool  : 16 words per non-empty brick, in grid order (slot = base + number of 3-codes before the brick in its entry); voxel (x,y,z) of a brick is bit (x | (y&3) << 3) of word ((z&7) << 1 | (y&7) >> 2)
// The brick grid is stored padded by one brick on every side.
struct BrickVolume {
    const uint32_t* codes;
    const uint32_t* base;
    const uint32_t* pool;
    const float* heights;  // heightmap kind: w*d column heights
    const uchar4* colors;  // uploaded bricks: one RGBA colour per pool slot
    uint32_t kind, seed;   // VT_VOLUME_*
    uint32_t bx, by, bz;   // PADDED brick grid dimensions (bricks per axis + 2)
    uint32_t n_bricks;     // non-empty bricks in the pool
};
static constexpr uint32_t kVolumeDense = 0, kVolumeHeightmap = 1, kVolumeSparseBricks = 2, kVolumeUploadedBricks = 3;

// Per-frame uniforms, passed by value as a kernel parameter (constant bank, no loads).
// How the ranks of a fused multi-GPU reduction share a frame by rows of 8x4-pixel tiles.  Tile rows are dealt in cycles of
// c * (world - 1) + (k - c) * world rows: the first c rounds of a cycle go to ranks 1 .. world - 1 only, the remaining k - c
// rounds to every rank in turn, so the root (rank 0), which also sums the slots and encodes the frame, owns (k - c) rows of a
// cycle and everybody else k.  world <= 1: one rank owns everything.
struct RowShare {
    uint32_t rank, world, k, c;
#ifdef __CUDACC__
    __host__ __device__ uint32_t cycle() const { return c * (world - 1u) + (k - c) * world; }
    __host__ __device__ uint32_t own_per_cycle() const { return rank == 0u ? k - c : k; }
    __host__ __device__ uint32_t owner(uint32_t ty) const {
        if (world <= 1u) return 0u;
        const uint32_t q = ty % cycle(), skip = c * (world - 1u);
        return q < skip ? 1u + q % (world - 1u) : (q - skip) % world;
    }
    // position inside a cycle of this rank's j-th row of the cycle (j < own_per_cycle())
    __host__ __device__ uint32_t own_pos(uint32_t j) const {
        const uint32_t skip = c * (world - 1u);
        if (rank == 0u) return skip + j * world;
        return j < c ? (rank - 1u) + j * (world - 1u) : skip + rank + (j - c) * world;
    }
    // this rank's rows among the tile rows [t0, t1], enumerated by cycle (a superset: rows_in() entries, of which row(i) may
    // fall outside [t0, t1] — callers skip those)
    __host__ __device__ uint32_t rows_in(uint32_t t0, uint32_t t1) const {
        if (world <= 1u) return t1 - t0 + 1u;
        return (t1 / cycle() - t0 / cycle() + 1u) * own_per_cycle();
    }
    __host__ __device__ uint32_t row(uint32_t t0, uint32_t i) const {
        if (world <= 1u) return t0 + i;
        const uint32_t n = own_per_cycle();
        return (t0 / cycle() + i / n) * cycle() + own_pos(i % n);
    }
#endif
};

struct FrameParams {
    float RD[16]; // inverse(centered camera) * inverse(projection), trace.frag:59
    float PV[16]; // projection * camera, trace.vert:45
    float eye[3]; // inverse(camera) * (0,0,0,1), trace.frag:51
    float vw, vh; // viewport, lib/command.c:80-81
    float sxn, syn; // 2/vw, 2/vh: pixel -> NDC scale
    int32_t width, height;
    uint32_t n_inst;
    uint32_t n_volumes;
    uint32_t flags;
    // path-tracing extension
    uint32_t spp, bounces, seed, sample_first, sample_stride;
    uint32_t refill_threshold; // wavefront / persistent-lane kernels: park the walkers when fewer than this many are left
    uint32_t refill_batch;     // wavefront kernel: hand out new rays once this many lanes have stopped
    uint32_t item_spp;        // wavefront kernel: most samples per work item (tile x samples), VT_ITEM_SPP
    uint32_t items_per_warp;  // wavefront kernel: work items wanted per resident warp, VT_ITEMS_PER_WARP
    float sun[3];              // unit vector towards the sun, world space (shadow-ray extension)
    uint32_t any_bricks;       // the scene contains a procedural brick volume
    uint32_t max_idx_bits;     // widest stop-mask bit index of the scene's dense volumes
    uint32_t clear_rgba;       // clear colour (lib/command.c:56-61) as stored by the sRGB target: r | g<<8 | b<<16 | a<<24
    uint32_t sky_spp;          // samples whose sky radiance this rank adds for pixels outside every screen
                               // rectangle (= spp normally; fused multi-GPU reduction: total on the root, 0 elsewhere)
    uint32_t item_order;       // wavefront kernel: 0 = chunk-major over tiles in row order, 1 = tile-major from the centre of the
                               // rectangle outwards, 2 = chunk-major, every pass from the centre outwards (VT_ITEM_ORDER)
    RowShare rows;             // wavefront kernel: the tile rows this rank traces (world <= 1: all; vt_fused_reduce_partition)
};

// Per-instance uniforms (trace.vert outputs that are flat per instance + derived matrices).
struct __align__(16) InstUniforms {
    float MVP[16];   // (P V) M           column-major
    float Mi[12];    // inverse(M), rows 0-2 of columns 0-3: Mi[c*3 + r]
    float M[12];     // M,          rows 0-2 of columns 0-3
    float dirm[12];  // inverse(M)3x3 * RD rows 0-2: clip point -> model-space ray direction
    float eye_m[3];  // camera position in model space
    uint32_t valid;  // texture id in range
    float sun_m[3];  // inverse(M)3x3 * sun: shadow-ray direction in model space
    uint32_t pad0;
    int32_t bounds[4]; // conservative screen rectangle of the proxy cube: x0, x1, y0, y1 (inclusive); 16-byte aligned
    float slab_lo[3]; // -0.5 - eye_m
    float slab_hi[3]; //  0.5 - eye_m
    uint32_t tex;
    uint32_t w, h, d;
    uint32_t xb, yb;
    uint32_t mask_off;
    uint32_t remap_identity;
    uint32_t pad;
    const uint8_t* rgba;
    const BrickVolume* bricks; // non-null: procedural brick volume
    // lin[k] != 0: inverse(M)'s 3x3 part is diagonal with (signed) powers of two, lin = its diagonal, ilin = 1 / lin.  The
    // instance's ray direction is then lin * (world direction) EXACTLY, and 1 / d = ilin * (1 / world direction) exactly:
    // the path tracer divides once per ray instead of once per instance it tests (every instance of the reference's
    // default world — translations, and chunks scaled by 2 — is of this kind).
    float lin[3], ilin[3];
    uint32_t pad1[2];
};

static_assert(offsetof(InstUniforms, bounds) % 16 == 0, "bounds are loaded as one int4");
static_assert(sizeof(InstUniforms) % 16 == 0, "instance table stride");

struct HitRecord { uint32_t hit_voxel, packed, instance, iters; };

struct FrameBuffers {
    HitRecord* records;   // may be null
    uchar4* color;
    float* depth;         // may be null
    unsigned long long* accum; // 3 per pixel
    unsigned long long* stats; // [0] rays, [1] iterations, [2] tile-scheduler counter, [3] spare
};

// Screen-space instance bins (what a tiling rasteriser's binner produces): for every 16x16-pixel
// bin the instances whose conservative screen rectangle touches it, ascending = draw order.
struct BinTable {
    const uint32_t* offset; // per bin: first entry in `list`
    const uint32_t* count;  // per bin: number of entries
    const uint32_t* list;   // instance indices
    uint32_t bins_x, bins_y;
    uint32_t enabled;       // 0: loop over all instances
    uint32_t pad;
};
static constexpr uint32_t kBinShift = 4; // 16x16 pixels

// World-space uniform grid over the instances' bounding boxes: narrows the instance loop of the path
// tracer's bounce rays (world_grid.cuh).  The header is written on the device every frame.
struct WorldGrid {
    float lo[3], cell[3], inv_cell[3];
    uint32_t res[3];
    uint32_t n_cells;
    uint32_t total;    // list entries needed
    uint32_t overflow; // 1: the lists did not fit (or a model matrix is not finite) -> full instance loop this frame
    uint32_t pad;
};
struct WorldGridTable {
    const WorldGrid* hdr;
    const uint32_t* offset; // per cell: first entry in `list`
    const uint32_t* count;  // per cell: number of entries
    const uint32_t* list;   // instance indices
    uint32_t enabled;       // 0: loop over all instances
    uint32_t pad;
};
static constexpr uint32_t kWorldGridMaxRes = 64, kWorldGridMaxCells = 64 * 64 * 64;
static constexpr uint32_t kWorldGridMinInstances = 8;

struct SrgbTables {
    const float* decode;    // 256
    const float* threshold; // 256
};

// GLSL inverse(mat4) (cofactor expansion, fixed operation order); 16 floats column-major.
// One source for host and device so the per-frame (host) and per-instance (device) inverses
// follow the same operation order as the oracle.
__host__ __device__ void mat4_inverse(const float* m, float* out);

// launch wrappers (all asynchronous on `stream`); return cudaError_t
cudaError_t launch_build_mask(const uint8_t* rgba, uint32_t w, uint32_t h, uint32_t d, uint32_t xb, uint32_t yb,
                              uint32_t* mask, uint32_t mask_words, cudaStream_t stream);
cudaError_t launch_instance_setup(const float* instances, uint32_t n, const VolumeDesc* volumes, FrameParams fp,
                                  InstUniforms* out, uint32_t* flag, uint32_t flag_value, unsigned long long* zero_stats,
                                  cudaStream_t stream);
// bins the instances' screen rectangles; *cursor (device) ends up holding the number of list entries
// needed — if it exceeds `capacity` the lists are incomplete and the caller must retry with more room
cudaError_t launch_bin_instances(const InstUniforms* inst, uint32_t n_inst, uint32_t bins_x, uint32_t bins_y, uint32_t* offset,
                                 uint32_t* count, uint32_t* list, uint32_t capacity, uint32_t* cursor, cudaStream_t stream);
cudaError_t launch_trace_primary(const FrameParams& fp, const InstUniforms* inst, BinTable bins, const uint32_t* mask_arena,
                                 uint32_t arena_words, bool masks_in_smem, SrgbTables lut, FrameBuffers fb,
                                 int sm_count, cudaStream_t stream);
cudaError_t launch_world_grid(const InstUniforms* inst, uint32_t n_inst, float* aabb, WorldGrid* hdr, uint32_t* offset, uint32_t* count,
                              uint32_t* cursor, uint32_t* list, uint32_t capacity, cudaStream_t stream);
cudaError_t launch_trace_paths(const FrameParams& fp, const InstUniforms* inst, BinTable bins, WorldGridTable wg, const uint32_t* mask_arena,
                               uint32_t arena_words, bool masks_in_smem, SrgbTables lut, FrameBuffers fb,
                               int sm_count, cudaStream_t stream);
// procedural brick volumes: column heights; mark + count the bricks with voxels (pool == nullptr); finish the directory
// (launch_brick_finalize: neighbourhood codes, border, slot bases; scratch = entries / 1024 + 2 words, its last used word
// receives the number of bricks with voxels); then fill the pool
cudaError_t launch_heightmap(float* heights, uint32_t w, uint32_t h, uint32_t d, uint32_t seed, cudaStream_t stream);
cudaError_t launch_brick_build(uint32_t kind, uint32_t seed, uint32_t w, uint32_t h, uint32_t d, const float* heights, uint32_t* codes,
                               const uint32_t* base, uint32_t* pool, uint32_t* counter, cudaStream_t stream);
cudaError_t launch_brick_finalize(uint32_t pbx, uint32_t pby, uint32_t pbz, uint32_t* codes, uint32_t* base, uint32_t* scratch,
                                  cudaStream_t stream);
// caller-supplied bricks: coords (n x 3) -> codes; after launch_brick_finalize the masks / colours move to their slots
cudaError_t launch_brick_index(const uint32_t* coords, uint32_t n, uint32_t bx, uint32_t by, uint32_t bz, uint32_t* codes, uint32_t* bad,
                               cudaStream_t stream);
cudaError_t launch_brick_place(const uint32_t* coords, uint32_t n, uint32_t bx, uint32_t by, const uint32_t* codes, const uint32_t* base,
                               const uint32_t* masks, const uchar4* colors, uint32_t* pool, uchar4* colors_out, cudaStream_t stream);
// incoherent-ray mode: rays [first, first + n) through instance 0's volume
cudaError_t launch_trace_rays(const FrameParams& fp, const InstUniforms* inst, const uint32_t* mask_arena, unsigned long long n,
                              unsigned long long first, FrameBuffers fb, int sm_count, cudaStream_t stream);
// fused multi-GPU reduction: move this rank's sums of the covered rectangle into its slot of the root's
// partial buffer (peer memory) and clear them locally; then, on the root, sum all slots and resolve
// What the push and summation kernels do about the flags (all optional: nullptr = nothing).
struct FusedSync {
    const uint32_t* wait_flags; // wait until wait_flags[0 .. wait_count) >= wait_target before touching the slots
    uint32_t wait_count, wait_target;
    uint32_t* signal_flag;      // once every block is done: *signal_flag = signal_value
    uint32_t signal_value;
    uint32_t* done_counter;     // device word, 0 between launches
    uint32_t* err;
    // the frame's counters (4 x u64, device) are published to mapped page-locked host memory by block 0 (no copy-engine
    // operation on the frame's stream: it would queue behind the previous frame's colour read-back)
    const unsigned long long* stats;
    unsigned long long* host_stats;
};

// rows.world > 1: the frame is shared out by rows of tiles (RowShare): a rank pushes only its own rows, and the root takes a
// pixel from its owner's slot instead of summing all
cudaError_t launch_push_partial(const InstUniforms* inst, unsigned long long* local_accum, uint4* slot, uint32_t width, uint32_t height,
                                bool compact, RowShare rows, FusedSync fs, int sm_count, cudaStream_t stream);
// root_local != nullptr: the root's own sums are still in its local accumulators (it does not push): they are taken from
// there, cleared, and parked in the root's slot (root_slot) for later calls on the same frame
cudaError_t launch_resolve_partials(const InstUniforms* inst, const uint4* partials, uint32_t world, uint32_t width, uint32_t height,
                                    uint32_t total_spp, SrgbTables lut, uchar4* color, unsigned long long* accum_out, bool compact,
                                    RowShare rows, unsigned long long* root_local, uint4* root_slot, FusedSync fs, int sm_count,
                                    cudaStream_t stream);
// single-instance frames: zero / resolve only the instance's screen rectangle; sky_only: write spp x sky into the
// accumulators outside it instead (they are not touched by such a frame otherwise)
cudaError_t launch_clear_rect(const InstUniforms* inst, unsigned long long* accum, uint32_t width, uint32_t height, int sm_count,
                              cudaStream_t stream);
cudaError_t launch_resolve_rect(const InstUniforms* inst, unsigned long long* accum, uint32_t width, uint32_t height, uint32_t spp,
                                uint32_t total_spp, SrgbTables lut, uchar4* color, bool sky_only, const unsigned long long* stats,
                                unsigned long long* host_stats, int sm_count, cudaStream_t stream);
bool paths_use_wave_kernel(const FrameParams& fp);
cudaError_t launch_resolve(const unsigned long long* accum, uint32_t n_pixels, uint32_t total_spp, SrgbTables lut,
                           uchar4* color, cudaStream_t stream);
// one-time: opt in to large dynamic shared memory
cudaError_t configure_kernels(int max_smem_optin);
// bytes of dynamic shared memory the trace kernels need for a given arena size
size_t trace_smem_bytes(uint32_t arena_words, bool masks_in_smem);

} // namespace vt

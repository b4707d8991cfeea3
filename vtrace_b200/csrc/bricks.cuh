// bricks.cuh — large procedural volumes (extension).
// Included from kernels.cu inside namespace vt.
//
// A 1024^3 or 4096^3 volume cannot be a dense RGBA8 texture (the reference's own add_texture
// overflows its u32 byte count at 1024^3, lib/memory.c:297).  Such volumes are generated on the
// device from a procedural definition and stored as occupancy only:
//   directory: per 16 consecutive bricks of the padded brick grid one word of two-bit codes (3: the brick has voxels,
//   2: it has none but a brick of its 3x3x3 neighbourhood has some or is outside, 1: the brick lies in the one-brick border =
//   outside the volume, 0: nothing within one brick) and one word with the number of bricks with voxels in all earlier
//   entries.  The pool holds the 16 words (512 bits) of the bricks with voxels in grid order, so a brick's slot is that
//   base + the number of 3-codes before it in its entry (no per-brick slot table: at 4096^3 that table was 543 MB, codes and
//   bases are 34 MB each and stay in L2).
// The brick grid carries a one-brick border on every side whose codes say "outside", so
// the walk needs no coordinate compares: leaving the volume is found by the same lookup as entering
// a brick.
// Traversal keeps the reference's per-voxel float DDA (trace.frag:73-87) bit for bit — same steps,
// same ties — but touches memory only when the ray enters a new brick (its two-bit code, then the slot base) and,
// inside non-empty bricks, one word per step; empty bricks are walked with arithmetic alone.
// Colour is a function of the voxel position, evaluated at the hit.
#pragma once

static constexpr uint32_t kSlotEmpty = 0xFFFFFFFFu; // brick without voxels
static constexpr uint32_t kSlotFree = 0xFFFFFFFEu;  // brick without voxels whose 26 neighbours have none either (and are inside).  Anything below is a pool slot.

// directory entries: 16 bricks each.  Codes and slot bases live in two arrays of one word per entry: the base is only
// needed for a brick with voxels, and the codes alone (34 MB at 4096^3) stay L2-resident (an interleaved 8-byte entry
// was measured 9 % slower on configs[4]).
__host__ __device__ __forceinline__ size_t dir_entries(size_t padded_bricks) { return (padded_bricks + 15) / 16; }
__device__ __forceinline__ void dir_set(uint32_t* __restrict__ codes, size_t bi, uint32_t bit) {
    atomicOr(codes + (bi >> 4), (1u << bit) << ((bi & 15u) * 2u));
}
// code of brick bi: 0 free, 1 outside, 2 empty, 3 has voxels
__device__ __forceinline__ uint32_t dir_code(const uint32_t* __restrict__ codes, uint32_t bi, uint32_t& entry) {
    entry = __ldg(codes + (bi >> 4));
    return (entry >> ((bi & 15u) * 2u)) & 3u;
}
// pool slot of a brick with voxels: its entry's base + the 3-codes before it in the entry
__device__ __forceinline__ uint32_t dir_slot(const uint32_t* __restrict__ base, uint32_t bi, uint32_t entry) {
    return __ldg(base + (bi >> 4)) + __popc(entry & (entry >> 1) & 0x55555555u & ((1u << ((bi & 15u) * 2u)) - 1u));
}

// index of brick (x, y, z) in the padded grid (pbx, pby = brick counts + 2)
__host__ __device__ __forceinline__ uint32_t brick_index(uint32_t pbx, uint32_t pby, uint32_t x, uint32_t y, uint32_t z) {
    return ((z + 1u) * pby + (y + 1u)) * pbx + (x + 1u);
}

__host__ __device__ __forceinline__ uint32_t vt_mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
__host__ __device__ __forceinline__ uint32_t vt_hash3(uint32_t x, uint32_t y, uint32_t z, uint32_t seed) {
    uint32_t h = vt_mix32(x * 0x9E3779B1u + seed);
    h = vt_mix32(h ^ (y * 0x85EBCA77u));
    h = vt_mix32(h ^ (z * 0xC2B2AE3Du));
    return h;
}

// heightmap kind: height(x,z) = H/2 + H/4 * fbm (5 octaves of value noise, lattices 128..8 voxels)
__device__ float heightmap_height(uint32_t x, uint32_t z, uint32_t H, uint32_t seed) {
    float n = 0.0f, amp = 0.5f;
    for (uint32_t o = 0; o < 5; ++o) {
        const uint32_t cell = 128u >> o;
        const uint32_t ix = x / cell, iz = z / cell;
        const float fx = (float)(x % cell) / (float)cell, fz = (float)(z % cell) / (float)cell;
        const float ux = (fx * fx) * (3.0f - 2.0f * fx), uz = (fz * fz) * (3.0f - 2.0f * fz);
        const float v00 = (float)(vt_hash3(ix, iz, o, seed) >> 8) * (1.0f / 16777216.0f);
        const float v10 = (float)(vt_hash3(ix + 1, iz, o, seed) >> 8) * (1.0f / 16777216.0f);
        const float v01 = (float)(vt_hash3(ix, iz + 1, o, seed) >> 8) * (1.0f / 16777216.0f);
        const float v11 = (float)(vt_hash3(ix + 1, iz + 1, o, seed) >> 8) * (1.0f / 16777216.0f);
        const float a = v00 + ux * (v10 - v00);
        const float b = v01 + ux * (v11 - v01);
        const float v = a + uz * (b - a);
        n = n + amp * (2.0f * v - 1.0f);
        amp = amp * 0.5f;
    }
    return 0.5f * (float)H + (0.25f * (float)H) * n;
}

__global__ void heightmap_kernel(float* __restrict__ heights, uint32_t w, uint32_t h, uint32_t d, uint32_t seed) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, z = blockIdx.y;
    if (x < w && z < d) heights[(size_t)z * w + x] = heightmap_height(x, z, h, seed);
}

cudaError_t launch_heightmap(float* heights, uint32_t w, uint32_t h, uint32_t d, uint32_t seed, cudaStream_t stream) {
    heightmap_kernel<<<dim3((w + 127) / 128, d), 128, 0, stream>>>(heights, w, h, d, seed);
    return cudaGetLastError();
}

// is voxel (x,y,z) filled?  (world +Y points down on screen, so the ground fills the high-y side)
__device__ __forceinline__ bool proc_filled(uint32_t kind, uint32_t seed, const float* __restrict__ heights, uint32_t w, uint32_t h,
                                            uint32_t x, uint32_t y, uint32_t z) {
    if (kind == kVolumeHeightmap) {
        const uint32_t alt = h - 1u - y;
        return (float)alt <= __ldg(heights + (size_t)z * w + x);
    }
    const bool brick = (vt_hash3(x >> 3, y >> 3, z >> 3, seed) & 0xFFFFu) < 1311u; // 2 % of the bricks
    return brick && (vt_hash3(x, y, z, seed ^ 0x5bd1e995u) & 1u);                    // half of their voxels
}

__device__ __forceinline__ uchar4 proc_color(const BrickVolume* __restrict__ bv, uint32_t h, uint32_t x, uint32_t y, uint32_t z) {
    const uint32_t kind = bv->kind, seed = bv->seed;
    if (kind == kVolumeUploadedBricks) { // one colour per brick, found through the table again (once per ray)
        const uint32_t bi = brick_index(bv->bx, bv->by, x >> 3, y >> 3, z >> 3);
        uint32_t entry;
        dir_code(bv->codes, bi, entry);
        const uchar4 c = __ldg(bv->colors + dir_slot(bv->base, bi, entry));
        return make_uchar4(c.x, c.y, c.z, 255);
    }
    if (kind == kVolumeHeightmap) {
        const uint32_t band = ((h - 1u - y) * 4u) / h;
        const uint32_t pal[4] = {0x323c48u, 0x388060u, 0x787878u, 0xf5f0f0u}; // b<<16 | g<<8 | r
        const uint32_t c = pal[band];
        return make_uchar4((unsigned char)c, (unsigned char)(c >> 8), (unsigned char)(c >> 16), 255);
    }
    const uint32_t c = vt_hash3(x, y, z, seed ^ 0x27d4eb2fu);
    return make_uchar4((unsigned char)(c | 0x40u), (unsigned char)((c >> 8) | 0x40u), (unsigned char)((c >> 16) | 0x40u), 255);
}

// One thread per brick.  pool == nullptr: mark the bricks with voxels (bit 0 of their code) and count them (*counter);
// otherwise (after brick_finalize) write each one's 16 words to its slot.
__global__ void brick_build_kernel(uint32_t kind, uint32_t seed, uint32_t w, uint32_t h, uint32_t d, const float* __restrict__ heights,
                                   uint32_t* __restrict__ codes, const uint32_t* __restrict__ base, uint32_t* __restrict__ pool,
                                   uint32_t* __restrict__ counter) {
    const uint32_t bxn = w >> 3, byn = h >> 3, bzn = d >> 3;
    const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= (size_t)bxn * byn * bzn) return;
    const uint32_t bx = (uint32_t)(b % bxn), by = (uint32_t)((b / bxn) % byn), bz = (uint32_t)(b / ((size_t)bxn * byn));
    // cheap rejection before touching 512 voxels
    if (kind == kVolumeSparseBricks) {
        if ((vt_hash3(bx, by, bz, seed) & 0xFFFFu) >= 1311u) return;
    } else {
        float hmax = -INFINITY;
        for (uint32_t k = 0; k < 64; ++k) hmax = fmaxf(hmax, __ldg(heights + (size_t)(bz * 8 + (k >> 3)) * w + bx * 8 + (k & 7)));
        const uint32_t alt_min = h - 1u - (by * 8 + 7);
        if ((float)alt_min > hmax) return; // the whole brick is above the terrain
    }
    uint32_t words[16];
    uint32_t any = 0;
#pragma unroll 1
    for (uint32_t wi = 0; wi < 16; ++wi) {
        uint32_t bits = 0;
        const uint32_t z = bz * 8 + (wi >> 1);
        for (uint32_t k = 0; k < 32; ++k) {
            const uint32_t x = bx * 8 + (k & 7), y = by * 8 + ((wi & 1) << 2) + (k >> 3);
            bits |= (proc_filled(kind, seed, heights, w, h, x, y, z) ? 1u : 0u) << k;
        }
        words[wi] = bits;
        any |= bits;
    }
    if (!any) return;
    const uint32_t pb = brick_index(bxn + 2, byn + 2, bx, by, bz);
    if (!pool) {
        dir_set(codes, pb, 0);
        atomicAdd(counter, 1u);
        return;
    }
    uint32_t entry;
    dir_code(codes, pb, entry);
    const uint32_t slot = dir_slot(base, pb, entry);
    for (uint32_t wi = 0; wi < 16; ++wi) pool[(size_t)slot * 16 + wi] = words[wi];
}

// The directory is finished in four passes after every brick with voxels has bit 0 of its code set:
// 1. scatter: bit 1 of every brick in the 3x3x3 neighbourhood of a brick with voxels or of a border brick (one thread per
//    padded brick; only the few per cent that qualify scatter);
__global__ void brick_dilate_kernel(uint32_t pbx, uint32_t pby, uint32_t pbz, uint32_t* __restrict__ codes) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)pbx * pby * pbz) return;
    const int x = (int)(i % pbx), y = (int)((i / pbx) % pby), z = (int)(i / ((size_t)pbx * pby));
    const bool border = x == 0 || y == 0 || z == 0 || x == (int)pbx - 1 || y == (int)pby - 1 || z == (int)pbz - 1;
    if (!border && !((codes[i >> 4] >> ((i & 15u) * 2u)) & 1u)) return;
    for (int dz = -1; dz <= 1; ++dz)
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                const int nx = x + dx, ny = y + dy, nz = z + dz;
                if (nx < 0 || ny < 0 || nz < 0 || nx >= (int)pbx || ny >= (int)pby || nz >= (int)pbz) continue;
                dir_set(codes, ((size_t)nz * pby + ny) * pbx + nx, 1);
            }
}
// 2. encode, one thread per entry: (bit 0, bit 1, border) -> 3 has voxels, 1 outside, 2 near something, 0 free; the entry's
//    count of 3-codes goes to base[] for the scan;
__global__ void brick_encode_kernel(uint32_t pbx, uint32_t pby, uint32_t pbz, size_t entries, uint32_t* __restrict__ codes,
                                    uint32_t* __restrict__ base) {
    const size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= entries) return;
    const uint32_t v = codes[w];
    const size_t n = (size_t)pbx * pby * pbz;
    uint32_t out = 0, cnt = 0;
    for (uint32_t k = 0; k < 16; ++k) {
        const size_t i = w * 16 + k;
        if (i >= n) break;
        const uint32_t x = (uint32_t)(i % pbx), y = (uint32_t)((i / pbx) % pby), z = (uint32_t)(i / ((size_t)pbx * pby));
        const bool border = x == 0 || y == 0 || z == 0 || x == pbx - 1 || y == pby - 1 || z == pbz - 1;
        const uint32_t p = (v >> (2 * k)) & 3u;
        const uint32_t q = border ? 1u : ((p & 1u) ? 3u : ((p & 2u) ? 2u : 0u));
        cnt += q == 3u ? 1u : 0u;
        out |= q << (2 * k);
    }
    codes[w] = out;
    base[w] = cnt;
}
// 3./4. exclusive scan of the counts (1024 entries per block: block sums, their scan by one block, then the entries).
__global__ void __launch_bounds__(256) brick_scan_sums_kernel(size_t entries, const uint32_t* __restrict__ base, uint32_t* __restrict__ sums) {
    __shared__ uint32_t part[8];
    const size_t first = (size_t)blockIdx.x * 1024 + threadIdx.x * 4;
    uint32_t s = 0;
    for (int k = 0; k < 4; ++k)
        if (first + k < entries) s += base[first + k];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int k = 0; k < 8; ++k) t += part[k];
        sums[blockIdx.x] = t;
    }
}
__global__ void brick_scan_top_kernel(uint32_t n_blocks, uint32_t* __restrict__ sums, uint32_t* __restrict__ total) {
    if (blockIdx.x || threadIdx.x) return;
    uint32_t run = 0;
    for (uint32_t b = 0; b < n_blocks; ++b) { const uint32_t v = sums[b]; sums[b] = run; run += v; }
    *total = run;
}
__global__ void __launch_bounds__(256) brick_scan_apply_kernel(size_t entries, uint32_t* __restrict__ base, const uint32_t* __restrict__ sums) {
    __shared__ uint32_t part[256];
    const size_t first = (size_t)blockIdx.x * 1024 + threadIdx.x * 4;
    uint32_t c[4] = {0, 0, 0, 0};
    for (int k = 0; k < 4; ++k)
        if (first + k < entries) c[k] = base[first + k];
    part[threadIdx.x] = c[0] + c[1] + c[2] + c[3];
    __syncthreads();
    if (threadIdx.x == 0) { // 256 partial sums: a serial exclusive scan by one thread is a few hundred cycles
        uint32_t run = sums[blockIdx.x];
        for (int k = 0; k < 256; ++k) { const uint32_t v = part[k]; part[k] = run; run += v; }
    }
    __syncthreads();
    uint32_t run = part[threadIdx.x];
    for (int k = 0; k < 4; ++k)
        if (first + k < entries) { base[first + k] = run; run += c[k]; }
}

// `scratch`: (entries + 1023) / 1024 + 1 words.  *total (device, scratch's last word) = number of bricks with voxels.
cudaError_t launch_brick_finalize(uint32_t pbx, uint32_t pby, uint32_t pbz, uint32_t* codes, uint32_t* base, uint32_t* scratch,
                                  cudaStream_t stream) {
    const size_t n = (size_t)pbx * pby * pbz;
    const size_t entries = dir_entries(n);
    const uint32_t n_blocks = (uint32_t)((entries + 1023) / 1024);
    brick_dilate_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(pbx, pby, pbz, codes);
    brick_encode_kernel<<<(unsigned)((entries + 255) / 256), 256, 0, stream>>>(pbx, pby, pbz, entries, codes, base);
    brick_scan_sums_kernel<<<n_blocks, 256, 0, stream>>>(entries, base, scratch);
    brick_scan_top_kernel<<<1, 32, 0, stream>>>(n_blocks, scratch, scratch + n_blocks);
    brick_scan_apply_kernel<<<n_blocks, 256, 0, stream>>>(entries, base, scratch);
    return cudaGetLastError();
}

cudaError_t launch_brick_build(uint32_t kind, uint32_t seed, uint32_t w, uint32_t h, uint32_t d, const float* heights, uint32_t* codes,
                               const uint32_t* base, uint32_t* pool, uint32_t* counter, cudaStream_t stream) {
    const size_t bricks = (size_t)(w >> 3) * (h >> 3) * (d >> 3);
    const int threads = 128;
    brick_build_kernel<<<(unsigned)((bricks + threads - 1) / threads), threads, 0, stream>>>(kind, seed, w, h, d, heights, codes, base, pool, counter);
    return cudaGetLastError();
}

// caller-supplied bricks, pass 1: one thread per brick marks its code (bit 0); coordinates outside the volume are counted in *bad
__global__ void brick_index_kernel(const uint32_t* __restrict__ coords, uint32_t n, uint32_t bx, uint32_t by, uint32_t bz,
                                   uint32_t* __restrict__ codes, uint32_t* __restrict__ bad) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t x = coords[3 * i], y = coords[3 * i + 1], z = coords[3 * i + 2];
    if (x >= bx || y >= by || z >= bz) { atomicAdd(bad, 1u); return; }
    dir_set(codes, brick_index(bx + 2, by + 2, x, y, z), 0);
}
// pass 2 (after brick_finalize): the caller's occupancy words and colour move to the brick's slot (grid order)
__global__ void brick_place_kernel(const uint32_t* __restrict__ coords, uint32_t n, uint32_t bx, uint32_t by,
                                   const uint32_t* __restrict__ codes, const uint32_t* __restrict__ base, const uint32_t* __restrict__ masks, const uchar4* __restrict__ colors, uint32_t* __restrict__ pool,
                                   uchar4* __restrict__ colors_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t bi = brick_index(bx + 2, by + 2, coords[3 * i], coords[3 * i + 1], coords[3 * i + 2]);
    uint32_t entry;
    dir_code(codes, bi, entry);
    const uint32_t slot = dir_slot(base, bi, entry);
    for (uint32_t wi = 0; wi < 16; ++wi) pool[(size_t)slot * 16 + wi] = masks[(size_t)i * 16 + wi];
    colors_out[slot] = colors[i];
}

cudaError_t launch_brick_index(const uint32_t* coords, uint32_t n, uint32_t bx, uint32_t by, uint32_t bz, uint32_t* codes, uint32_t* bad,
                               cudaStream_t stream) {
    if (n) brick_index_kernel<<<(n + 127) / 128, 128, 0, stream>>>(coords, n, bx, by, bz, codes, bad);
    return cudaGetLastError();
}
cudaError_t launch_brick_place(const uint32_t* coords, uint32_t n, uint32_t bx, uint32_t by, const uint32_t* codes, const uint32_t* base,
                               const uint32_t* masks, const uchar4* colors, uint32_t* pool, uchar4* colors_out, cudaStream_t stream) {
    if (n) brick_place_kernel<<<(n + 127) / 128, 128, 0, stream>>>(coords, n, bx, by, codes, base, masks, colors, pool, colors_out);
    return cudaGetLastError();
}

// The reference's DDA (trace.frag:63-89) over a brick volume, as a resumable walk: same state, same
// float operations in the same order as dda_init / dda_step / dda_slow_impl — only the occupancy
// test differs.  brick_walk_burst() runs up to 8 loop iterations; the persistent-lane ray kernel interleaves
// the walks of 32 lanes and refills lanes whose ray ended.
//
// Fast path (no NaN/inf in side/delta): the voxel is kept as a brick (padded brick coordinates, `cell`) plus its
// position INSIDE the brick, three 5-bit fields of one word (`loc` = fx | fy << 8 | fz << 16, field = local
// coordinate + 8).  A step is three predicated adds on `loc`; a field that leaves 8..15 has crossed a face of the
// brick, and bit 3 of a field is set exactly while it is inside, so "did this lane leave its brick?" is ONE
// instruction (LOP3 with predicate output: ~loc & stop != 0) instead of compares on absolute coordinates.
// `stop` = 0x080808 in an empty brick; 0x202020 (bits that are never set: the test always fires) in a brick with
// voxels, where every step needs its pool word; 0 in a brick whose neighbourhood is empty too (kSlotFree): such a
// lane walks a whole burst through brick faces — the fields then range over 0..23 and say how many bricks it moved.
// Only when the test fires is anything looked up: the brick's directory entry (code 1, the border = the ray left the
// volume; code 3 comes with the brick's pool slot), one pool word per step inside a brick with voxels.  Every fast-path iteration advances
// >= 1 voxel, so the steps < W+H+D bound of :74 cannot bind before the ray is outside.
struct BrickWalk {
    float sx, sy, sz;
    uint32_t loc;          // position inside the brick: (x & 7) + 8 | ((y & 7) + 8) << 8 | ((z & 7) + 8) << 16
    uint32_t ploc;         // loc before the last iteration | iterations run by the current burst << 24
    uint32_t cell;         // index of the brick in the padded brick grid, (cz * pby + cy) * pbx + cx
    uint32_t ix, iy, iz;   // per-axis increments of loc: +-1, +-1 << 8, +-1 << 16
    uint32_t stop;         // see above
    uint32_t slot;
    uint32_t steps, last;
};
static constexpr uint32_t kLocInside = 0x00080808u, kLocAlways = 0x00202020u, kLocFields = 0x001F1F1Fu, kLocLow = 0x00070707u;
#ifndef VT_MARCH_BURST
#define VT_MARCH_BURST 8 // DDA iterations between brick lookups, coherent rays (primary / shadow); 4: 1.99 ms, 6: 1.79, 8: 1.74 (configs[3])
#endif
#ifndef VT_RAY_BURST
#define VT_RAY_BURST 8   // same, incoherent rays (trace_rays_kernel); 6: 44.0 ms, 8: 42.2 (configs[4])
#endif

// literal transcription of the loop for rays with a zero direction component (0 * inf = NaN, :84), run to
// the end.  Out of line and fed by value, so the callers' ray state stays in registers.
struct BrickSlowState {
    int32_t v[3];
    float side[3];
    uint32_t steps, last;
};
__device__ __noinline__ int brick_walk_slow(const uint32_t* __restrict__ codes, const uint32_t* __restrict__ base,
                                            const uint32_t* __restrict__ pool, uint32_t pbx, uint32_t pby, uint32_t W, uint32_t H,
                                            uint32_t D, float d0, float d1, float d2, int32_t s0, int32_t s1, int32_t s2,
                                            BrickSlowState* io) {
    int32_t vx = io->v[0], vy = io->v[1], vz = io->v[2];
    float sx = io->side[0], sy = io->side[1], sz = io->side[2];
    uint32_t steps = 0, last = 0;
    int status = 2;
    while (steps < W + H + D && (uint32_t)vx < W && (uint32_t)vy < H && (uint32_t)vz < D) { // :74-75
        const uint32_t bi = brick_index(pbx, pby, (uint32_t)vx >> 3, (uint32_t)vy >> 3, (uint32_t)vz >> 3);
        uint32_t entry;
        if (dir_code(codes, bi, entry) == 3u) {
            const uint32_t wv = __ldg(pool + ((size_t)dir_slot(base, bi, entry) << 4) + ((((uint32_t)vz & 7u) << 1) | (((uint32_t)vy & 7u) >> 2)));
            if ((wv >> (((uint32_t)vx & 7u) | (((uint32_t)vy & 3u) << 3))) & 1u) { status = 1; break; } // :78-80
        }
        const bool m0 = sx <= vt_fmin(sy, sz); // :83
        const bool m1 = sy <= vt_fmin(sz, sx);
        const bool m2 = sz <= vt_fmin(sx, sy);
        sx += (m0 ? 1.0f : 0.0f) * d0; // :84
        sy += (m1 ? 1.0f : 0.0f) * d1;
        sz += (m2 ? 1.0f : 0.0f) * d2;
        vx += m0 ? s0 : 0; // :85
        vy += m1 ? s1 : 0;
        vz += m2 ? s2 : 0;
        last = (m0 ? 1u : 0u) | (m1 ? 2u : 0u) | (m2 ? 4u : 0u);
        ++steps; // :86
    }
    io->v[0] = vx; io->v[1] = vy; io->v[2] = vz;
    io->side[0] = sx; io->side[1] = sy; io->side[2] = sz;
    io->steps = steps; io->last = last;
    return status;
}

// The walk left the brick it knew to be empty (or stands in one that has voxels): look at the current
// voxel.  0 = empty, keep walking; 1 = filled (:78-80); 2 = outside the volume (:75).
__device__ __forceinline__ int brick_walk_lookup(const BrickVolume& bv, BrickWalk& k) {
    const uint32_t moved = (k.loc ^ k.ploc) & kLocFields; // fields the last iteration changed (:83), before loc is re-based
    if ((~k.loc & kLocInside) != 0u) { // crossed at least one face: field >> 3 = 0 / 1 / 2 -> one brick down / same / up
        const uint32_t row = bv.bx, plane = bv.bx * bv.by;
        k.cell = k.cell + ((k.loc >> 3) & 0x3u) + ((k.loc >> 11) & 0x3u) * row + ((k.loc >> 19) & 0x3u) * plane - (1u + row + plane);
        k.loc = (k.loc & kLocLow) | kLocInside;
        const uint32_t bi = k.cell;
        uint32_t entry;
        const uint32_t code = dir_code(bv.codes, bi, entry);
        k.slot = code == 3u ? dir_slot(bv.base, bi, entry) : (code == 0u ? kSlotFree : kSlotEmpty);
        k.stop = code == 3u ? kLocAlways : (code == 0u ? 0u : kLocInside);
        if (code == 1u) return 2; // the border: outside the volume
    }
    if (k.slot < kSlotFree) {
        const uint32_t lx = k.loc & 7u, ly = (k.loc >> 8) & 7u, lz = (k.loc >> 16) & 7u;
        const uint32_t wv = __ldg(bv.pool + ((size_t)k.slot << 4) + ((lz << 1) | (ly >> 2)));
        if ((wv >> (lx | ((ly & 3u) << 3))) & 1u) {
            k.last = ((moved & 0x1Fu) ? 1u : 0u) | ((moved & 0x1F00u) ? 2u : 0u) | ((moved & 0x1F0000u) ? 4u : 0u);
            return 1;
        }
    }
    return 0;
}

// ray setup (trace.frag:63-71) + the test of the start voxel; returns the walk status (0 = walking)
__device__ __forceinline__ int brick_walk_begin(const BrickVolume& bv, uint32_t W, uint32_t H, uint32_t D, const float pos[3],
                                                const float dir[3], bool has_start, const int32_t sv[3], Dda& r, BrickWalk& k) {
    const float size[3] = {(float)(int32_t)W, (float)(int32_t)H, (float)(int32_t)D};
    float sgn[3];
    r.len = sqrtf((dir[0] * dir[0] + dir[1] * dir[1]) + dir[2] * dir[2]); // length(), :70
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        r.pos[c] = pos[c];
        r.dir[c] = dir[c];
        r.v[c] = has_start ? sv[c] : __float2int_rz(floorf(vt_fmin(pos[c], size[c] - 1.0f))); // :68
        sgn[c] = dir[c] > 0.0f ? 1.0f : (dir[c] < 0.0f ? -1.0f : 0.0f);                       // :69
        r.step[c] = (int32_t)sgn[c];
        r.delta[c] = fabsf(r.len / dir[c]);                                                   // :70
        r.side[c] = ((sgn[c] * ((float)r.v[c] - pos[c]) + sgn[c] * 0.5f) + 0.5f) * r.delta[c]; // :71
    }
    r.steps = 0;
    r.last_mask = 0;
    r.hit = false;
    const bool finite = isfinite(r.delta[0]) && isfinite(r.delta[1]) && isfinite(r.delta[2]) && isfinite(r.side[0]) &&
                        isfinite(r.side[1]) && isfinite(r.side[2]);
    // the packed walk needs the start voxel within one brick of the volume (true for every caller:
    // rays start inside, shadow / bounce rays at most one voxel outside)
    const bool near = r.v[0] >= -8 && r.v[0] < (int32_t)W + 8 && r.v[1] >= -8 && r.v[1] < (int32_t)H + 8 && r.v[2] >= -8 &&
                      r.v[2] < (int32_t)D + 8;
    int status = 0;
    k.steps = 0; k.last = 0;
    if (!finite || !near) {
        BrickSlowState io;
#pragma unroll
        for (int c = 0; c < 3; ++c) { io.v[c] = r.v[c]; io.side[c] = r.side[c]; }
        status = brick_walk_slow(bv.codes, bv.base, bv.pool, bv.bx, bv.by, W, H, D, r.delta[0], r.delta[1], r.delta[2], r.step[0],
                                 r.step[1], r.step[2], &io);
        // park the finished ray in the packed state (clamped: the voxel of a miss is never read)
        r.v[0] = max(-8, min(io.v[0], (int32_t)W + 7)); r.v[1] = max(-8, min(io.v[1], (int32_t)H + 7)); r.v[2] = max(-8, min(io.v[2], (int32_t)D + 7));
        r.side[0] = io.side[0]; r.side[1] = io.side[1]; r.side[2] = io.side[2];
        k.steps = io.steps; k.last = io.last;
    }
    // padded brick = (v + 8) >> 3, position inside = (v + 8) & 7
    const uint32_t ux = (uint32_t)(r.v[0] + 8), uy = (uint32_t)(r.v[1] + 8), uz = (uint32_t)(r.v[2] + 8);
    k.cell = ((uz >> 3) * bv.by + (uy >> 3)) * bv.bx + (ux >> 3);
    k.loc = ((ux & 7u) | ((uy & 7u) << 8) | ((uz & 7u) << 16)) | kLocInside;
    k.ix = (uint32_t)r.step[0]; k.iy = (uint32_t)r.step[1] << 8; k.iz = (uint32_t)r.step[2] << 16;
    k.sx = r.side[0]; k.sy = r.side[1]; k.sz = r.side[2];
    k.ploc = k.loc;
    // no current brick yet: look the start brick up (the same code as after crossing a face, with no face crossed)
    if (status == 0) {
        const uint32_t bi = k.cell;
        uint32_t entry;
        const uint32_t code = dir_code(bv.codes, bi, entry);
        k.slot = code == 3u ? dir_slot(bv.base, bi, entry) : (code == 0u ? kSlotFree : kSlotEmpty);
        k.stop = code == 3u ? kLocAlways : (code == 0u ? 0u : kLocInside);
        status = code == 1u ? 2 : brick_walk_lookup(bv, k);
    } else {
        k.slot = kSlotEmpty;
        k.stop = kLocInside;
    }
    return status;
}

// Up to kBurst iterations of the while loop of trace.frag:75-87 through the current brick, then ONE lookup if
// the walk left it.  The lanes of a warp run the burst together and meet again for the lookup, so its loads are
// issued by many lanes at once instead of by whichever lane happens to cross a brick face in a given iteration.
// One iteration in PTX: 12 SASS instructions (FMNMX3, 3 FSETP, the record of the previous position, 3 + 3
// predicated adds, the face test).  lf = "this lane has left its brick"; such a lane idles through the rest of
// the burst, predicated off, instead of jumping ahead to a lookup of its own.
//   %0-2 side, %3 loc, %4 ploc (written with the iteration's number in bits 24+), %5 left (out),
//   %6-8 delta, %9-11 increments of loc, %12 stop bits
#define VT_BRICK_STEP_PTX(CODE)                  \
    "min.f32 m, %0, %1;\n"                       \
    "min.f32 m, m, %2;\n"                        \
    "setp.eq.and.f32 px, %0, m, !lf;\n"          \
    "setp.eq.and.f32 py, %1, m, !lf;\n"          \
    "setp.eq.and.f32 pz, %2, m, !lf;\n"          \
    "@!lf add.u32 %4, %3, " CODE ";\n"           \
    "@px add.rn.f32 %0, %0, %6;\n"               \
    "@py add.rn.f32 %1, %1, %7;\n"               \
    "@pz add.rn.f32 %2, %2, %8;\n"               \
    "@px add.u32 %3, %3, %9;\n"                  \
    "@py add.u32 %3, %3, %10;\n"                 \
    "@pz add.u32 %3, %3, %11;\n"                 \
    "lop3.or.b32 t|lf, %3, %12, 0, 0x0C, lf;\n"  /* ~loc & stop != 0: a field left 8..15 */
#define VT_BRICK_BURST_ASM(STEPS)                                                                                         \
    asm volatile("{\n"                                                                                                    \
                 ".reg .pred lf, px, py, pz;\n"                                                                           \
                 ".reg .f32 m;\n"                                                                                         \
                 ".reg .u32 t;\n"                                                                                         \
                 "setp.ne.u32 lf, 0, 0;\n" STEPS "selp.u32 %5, 1, 0, lf;\n"                                              \
                 "}\n"                                                                                                    \
                 : "+f"(k.sx), "+f"(k.sy), "+f"(k.sz), "+r"(k.loc), "+r"(k.ploc), "=r"(left)                                \
                 : "f"(r.delta[0]), "f"(r.delta[1]), "f"(r.delta[2]), "r"(k.ix), "r"(k.iy), "r"(k.iz), "r"(k.stop))
#define VT_BRICK_STEPS_4 VT_BRICK_STEP_PTX("0x01000000") VT_BRICK_STEP_PTX("0x02000000") VT_BRICK_STEP_PTX("0x03000000") VT_BRICK_STEP_PTX("0x04000000")
#define VT_BRICK_STEPS_6 VT_BRICK_STEPS_4 VT_BRICK_STEP_PTX("0x05000000") VT_BRICK_STEP_PTX("0x06000000")
#define VT_BRICK_STEPS_8 VT_BRICK_STEPS_6 VT_BRICK_STEP_PTX("0x07000000") VT_BRICK_STEP_PTX("0x08000000")

template <int kBurst>
__device__ __forceinline__ int brick_walk_burst(const BrickVolume& bv, const Dda& r, BrickWalk& k) {
    static_assert(kBurst == 4 || kBurst == 6 || kBurst == 8, "burst lengths with a PTX body (at most 8: see kSlotFree)");
    // Per iteration (trace.frag:83-86), no NaN: side <= min(other two) is side == min(all three); vec3(mask) * delta is
    // a predicated add.  The asm is volatile and `left` comes out of it, so the lookup is reached from ONE branch by all
    // lanes that need it together.
    uint32_t left;
    if (kBurst == 4) VT_BRICK_BURST_ASM(VT_BRICK_STEPS_4);
    if (kBurst == 6) VT_BRICK_BURST_ASM(VT_BRICK_STEPS_6);
    if (kBurst == 8) VT_BRICK_BURST_ASM(VT_BRICK_STEPS_8);
    k.steps += k.ploc >> 24; // :86, the iterations this burst ran
    k.ploc &= 0x00FFFFFFu;
    // a lane that walked on through the faces of a kSlotFree brick looks up where it is now (branch-free: every lane
    // that needs the lookup reaches it together)
    left |= (k.stop == 0u ? 1u : 0u) & ((~k.loc & kLocInside) != 0u ? 1u : 0u);
    return left ? brick_walk_lookup(bv, k) : 0;
}

__device__ __forceinline__ void brick_walk_finish(const BrickVolume& bv, const BrickWalk& k, bool hit, Dda& r) {
    const uint32_t cz = k.cell / (bv.bx * bv.by), rem = k.cell - cz * (bv.bx * bv.by), cy = rem / bv.bx, cx = rem - cy * bv.bx;
    r.v[0] = (int32_t)((cx << 3) + (k.loc & 7u)) - 8;
    r.v[1] = (int32_t)((cy << 3) + ((k.loc >> 8) & 7u)) - 8;
    r.v[2] = (int32_t)((cz << 3) + ((k.loc >> 16) & 7u)) - 8;
    r.side[0] = k.sx; r.side[1] = k.sy; r.side[2] = k.sz;
    r.steps = k.steps;
    r.last_mask = k.last;
    r.hit = hit;
}

__device__ __forceinline__ void dda_march_bricks(const BrickVolume& bv, uint32_t W, uint32_t H, uint32_t D, const float pos[3],
                                                 const float dir[3], bool has_start, const int32_t sv[3], Dda& r) {
    BrickWalk k;
    int status = brick_walk_begin(bv, W, H, D, pos, dir, has_start, sv, r, k);
    while (status == 0) status = brick_walk_burst<VT_MARCH_BURST>(bv, r, k);
    brick_walk_finish(bv, k, status == 1, r);
}

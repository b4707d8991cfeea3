// bricks.cuh — large procedural volumes (extension).
// Included from kernels.cu inside namespace vt.
//
// A 1024^3 or 4096^3 volume cannot be a dense RGBA8 texture (the reference's own add_texture
// overflows its u32 byte count at 1024^3, lib/memory.c:297).  Such volumes are generated on the
// device from a procedural definition and stored as occupancy only:
//   l1 bit per 8^3 brick  ->  table[brick] = pool slot  ->  16 words (512 bits) per non-empty brick.
// Traversal keeps the reference's per-voxel float DDA (trace.frag:73-87) bit for bit — same steps,
// same ties — but touches memory only when the ray enters a new brick (l1 bit, then the slot) and,
// inside non-empty bricks, one word per step; empty bricks are walked with arithmetic alone.
// Colour is a function of the voxel position, evaluated at the hit.
#pragma once

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
        const size_t b = ((size_t)(z >> 3) * bv->by + (y >> 3)) * bv->bx + (x >> 3);
        const uchar4 c = __ldg(bv->colors + __ldg(bv->table + b));
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

// One thread per brick.  pool == nullptr: count the non-empty bricks (*counter); otherwise fill.
__global__ void brick_build_kernel(uint32_t kind, uint32_t seed, uint32_t w, uint32_t h, uint32_t d, const float* __restrict__ heights,
                                   uint32_t* __restrict__ l1, uint32_t* __restrict__ table, uint32_t* __restrict__ pool,
                                   uint32_t pool_capacity, uint32_t* __restrict__ counter) {
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
    const uint32_t slot = atomicAdd(counter, 1u);
    if (!pool || slot >= pool_capacity) return;
    for (uint32_t wi = 0; wi < 16; ++wi) pool[(size_t)slot * 16 + wi] = words[wi];
    table[b] = slot;
    atomicOr(l1 + (b >> 5), 1u << (b & 31));
}

cudaError_t launch_brick_build(uint32_t kind, uint32_t seed, uint32_t w, uint32_t h, uint32_t d, const float* heights, uint32_t* l1,
                               uint32_t* table, uint32_t* pool, uint32_t pool_capacity, uint32_t* counter, cudaStream_t stream) {
    const size_t bricks = (size_t)(w >> 3) * (h >> 3) * (d >> 3);
    const int threads = 128;
    brick_build_kernel<<<(unsigned)((bricks + threads - 1) / threads), threads, 0, stream>>>(kind, seed, w, h, d, heights, l1, table, pool,
                                                                                          pool_capacity, counter);
    return cudaGetLastError();
}

// caller-supplied bricks: one thread per brick writes its table entry and l1 bit
__global__ void brick_index_kernel(const uint32_t* __restrict__ coords, uint32_t n, uint32_t bx, uint32_t by, uint32_t bz,
                                   uint32_t* __restrict__ l1, uint32_t* __restrict__ table, uint32_t* __restrict__ bad) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t x = coords[3 * i], y = coords[3 * i + 1], z = coords[3 * i + 2];
    if (x >= bx || y >= by || z >= bz) { atomicAdd(bad, 1u); return; }
    const size_t b = ((size_t)z * by + y) * bx + x;
    table[b] = i;
    atomicOr(l1 + (b >> 5), 1u << (b & 31));
}

cudaError_t launch_brick_index(const uint32_t* coords, uint32_t n, uint32_t bx, uint32_t by, uint32_t bz, uint32_t* l1, uint32_t* table,
                               uint32_t* bad, cudaStream_t stream) {
    if (n) brick_index_kernel<<<(n + 127) / 128, 128, 0, stream>>>(coords, n, bx, by, bz, l1, table, bad);
    return cudaGetLastError();
}

// The reference's DDA (trace.frag:63-89) over a brick volume, as a resumable walk: same state, same
// float operations in the same order as dda_init / dda_step / dda_slow_impl — only the occupancy
// test differs.  brick_walk_step() runs ONE loop iteration; the persistent-lane ray kernel interleaves
// the walks of 32 lanes and refills lanes whose ray ended.
struct BrickWalk {
    int32_t vx, vy, vz;
    float sx, sy, sz;
    uint32_t steps, last;
    uint32_t cur_key, cur_slot;
    bool finite;
};

__device__ __forceinline__ void brick_walk_init(uint32_t W, uint32_t H, uint32_t D, const float pos[3], const float dir[3],
                                                bool has_start, const int32_t sv[3], Dda& r, BrickWalk& k) {
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
    k.finite = isfinite(r.delta[0]) && isfinite(r.delta[1]) && isfinite(r.delta[2]) && isfinite(r.side[0]) &&
               isfinite(r.side[1]) && isfinite(r.side[2]);
    k.vx = r.v[0]; k.vy = r.v[1]; k.vz = r.v[2];
    k.sx = r.side[0]; k.sy = r.side[1]; k.sz = r.side[2];
    k.steps = 0; k.last = 0;
    k.cur_key = 0xFFFFFFFFu; k.cur_slot = 0xFFFFFFFFu;
}

// one iteration of the while loop of trace.frag:75-87; returns 0 = keep walking, 1 = hit, 2 = left / out of steps
__device__ __forceinline__ int brick_walk_step(const BrickVolume& bv, uint32_t W, uint32_t H, uint32_t D, const Dda& r, BrickWalk& k) {
    if (!(k.steps < W + H + D && (uint32_t)k.vx < W && (uint32_t)k.vy < H && (uint32_t)k.vz < D)) return 2; // :74-75
    // texture(tex, voxel / size): brick volumes only exist for sizes where the texel IS the voxel
    const uint32_t key = ((uint32_t)k.vx >> 3) | (((uint32_t)k.vy >> 3) << 10) | (((uint32_t)k.vz >> 3) << 20);
    if (key != k.cur_key) { // entered a new brick: one l1 bit, and the slot if it is set
        k.cur_key = key;
        const uint32_t b = (((uint32_t)k.vz >> 3) * bv.by + ((uint32_t)k.vy >> 3)) * bv.bx + ((uint32_t)k.vx >> 3); // < 2^31 bricks
        const uint32_t bit = (__ldg(bv.l1 + (b >> 5)) >> (b & 31)) & 1u;
        k.cur_slot = bit ? __ldg(bv.table + b) : 0xFFFFFFFFu;
    }
    if (k.cur_slot != 0xFFFFFFFFu) {
        const uint32_t wv = __ldg(bv.pool + ((size_t)k.cur_slot << 4) + ((((uint32_t)k.vz & 7u) << 1) | (((uint32_t)k.vy & 7u) >> 2)));
        if ((wv >> (((uint32_t)k.vx & 7u) | (((uint32_t)k.vy & 3u) << 3))) & 1u) return 1; // :78-80
    }
    bool m0, m1, m2;
    if (k.finite) { // no NaN: side <= min(other two) is side == min(all three); vec3(mask) * delta is a predicated add
        const float m = fminf(fminf(k.sx, k.sy), k.sz);
        m0 = k.sx == m; m1 = k.sy == m; m2 = k.sz == m;
        if (m0) k.sx += r.delta[0];
        if (m1) k.sy += r.delta[1];
        if (m2) k.sz += r.delta[2];
    } else {
        m0 = k.sx <= vt_fmin(k.sy, k.sz); // :83
        m1 = k.sy <= vt_fmin(k.sz, k.sx);
        m2 = k.sz <= vt_fmin(k.sx, k.sy);
        k.sx += (m0 ? 1.0f : 0.0f) * r.delta[0]; // :84
        k.sy += (m1 ? 1.0f : 0.0f) * r.delta[1];
        k.sz += (m2 ? 1.0f : 0.0f) * r.delta[2];
    }
    k.vx += m0 ? r.step[0] : 0; // :85
    k.vy += m1 ? r.step[1] : 0;
    k.vz += m2 ? r.step[2] : 0;
    k.last = (m0 ? 1u : 0u) | (m1 ? 2u : 0u) | (m2 ? 4u : 0u);
    ++k.steps; // :86
    return 0;
}

__device__ __forceinline__ void brick_walk_finish(const BrickWalk& k, bool hit, Dda& r) {
    r.v[0] = k.vx; r.v[1] = k.vy; r.v[2] = k.vz;
    r.side[0] = k.sx; r.side[1] = k.sy; r.side[2] = k.sz;
    r.steps = k.steps;
    r.last_mask = k.last;
    r.hit = hit;
}

__device__ __forceinline__ void dda_march_bricks(const BrickVolume& bv, uint32_t W, uint32_t H, uint32_t D, const float pos[3],
                                                 const float dir[3], bool has_start, const int32_t sv[3], Dda& r) {
    BrickWalk k;
    brick_walk_init(W, H, D, pos, dir, has_start, sv, r, k);
    int status;
    while ((status = brick_walk_step(bv, W, H, D, r, k)) == 0) {}
    brick_walk_finish(k, status == 1, r);
}

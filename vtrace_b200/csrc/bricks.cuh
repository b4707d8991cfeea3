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

__device__ __forceinline__ uchar4 proc_color(uint32_t kind, uint32_t seed, uint32_t h, uint32_t x, uint32_t y, uint32_t z) {
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

// The reference's DDA (trace.frag:63-89) over a brick volume.  Same state, same float operations in
// the same order as dda_init / dda_step / dda_slow_impl — only the occupancy test differs.
__device__ __forceinline__ void dda_march_bricks(const BrickVolume& bv, uint32_t W, uint32_t H, uint32_t D, const float pos[3],
                                                 const float dir[3], bool has_start, const int32_t sv[3], Dda& r) {
    const int32_t isz[3] = {(int32_t)W, (int32_t)H, (int32_t)D};
    const float size[3] = {(float)isz[0], (float)isz[1], (float)isz[2]};
    float sgn[3];
    r.len = sqrtf((dir[0] * dir[0] + dir[1] * dir[1]) + dir[2] * dir[2]); // length(), :70
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        r.pos[k] = pos[k];
        r.dir[k] = dir[k];
        r.v[k] = has_start ? sv[k] : __float2int_rz(floorf(vt_fmin(pos[k], size[k] - 1.0f))); // :68
        sgn[k] = dir[k] > 0.0f ? 1.0f : (dir[k] < 0.0f ? -1.0f : 0.0f);                       // :69
        r.step[k] = (int32_t)sgn[k];
        r.delta[k] = fabsf(r.len / dir[k]);                                                   // :70
        r.side[k] = ((sgn[k] * ((float)r.v[k] - pos[k]) + sgn[k] * 0.5f) + 0.5f) * r.delta[k]; // :71
    }
    r.steps = 0;
    r.last_mask = 0;
    r.hit = false;
    const bool finite = isfinite(r.delta[0]) && isfinite(r.delta[1]) && isfinite(r.delta[2]) && isfinite(r.side[0]) &&
                        isfinite(r.side[1]) && isfinite(r.side[2]);
    const uint32_t max_steps = W + H + D; // :74
    uint32_t cur_key = 0xFFFFFFFFu, cur_slot = 0xFFFFFFFFu;
    int32_t vx = r.v[0], vy = r.v[1], vz = r.v[2];
    float sx = r.side[0], sy = r.side[1], sz = r.side[2];
    uint32_t steps = 0, last = 0;
    while (steps < max_steps && (uint32_t)vx < W && (uint32_t)vy < H && (uint32_t)vz < D) { // :75
        // texture(tex, voxel / size): brick volumes only exist for sizes where the texel IS the voxel
        const uint32_t key = ((uint32_t)vx >> 3) | (((uint32_t)vy >> 3) << 10) | (((uint32_t)vz >> 3) << 20);
        if (key != cur_key) { // entered a new brick: one l1 bit, and the slot if it is set
            cur_key = key;
            const size_t b = ((size_t)((uint32_t)vz >> 3) * bv.by + ((uint32_t)vy >> 3)) * bv.bx + ((uint32_t)vx >> 3);
            const uint32_t bit = (__ldg(bv.l1 + (b >> 5)) >> (b & 31)) & 1u;
            cur_slot = bit ? __ldg(bv.table + b) : 0xFFFFFFFFu;
        }
        if (cur_slot != 0xFFFFFFFFu) {
            const uint32_t wv = __ldg(bv.pool + (size_t)cur_slot * 16 + ((((uint32_t)vz & 7u) << 1) | (((uint32_t)vy & 7u) >> 2)));
            if ((wv >> (((uint32_t)vx & 7u) | (((uint32_t)vy & 3u) << 3))) & 1u) { r.hit = true; break; } // :78-80
        }
        bool m0, m1, m2;
        if (finite) { // no NaN: side <= min(other two) is side == min(all three); vec3(mask) * delta is a predicated add
            const float m = fminf(fminf(sx, sy), sz);
            m0 = sx == m; m1 = sy == m; m2 = sz == m;
            if (m0) sx += r.delta[0];
            if (m1) sy += r.delta[1];
            if (m2) sz += r.delta[2];
        } else {
            m0 = sx <= vt_fmin(sy, sz); // :83
            m1 = sy <= vt_fmin(sz, sx);
            m2 = sz <= vt_fmin(sx, sy);
            sx += (m0 ? 1.0f : 0.0f) * r.delta[0]; // :84
            sy += (m1 ? 1.0f : 0.0f) * r.delta[1];
            sz += (m2 ? 1.0f : 0.0f) * r.delta[2];
        }
        vx += m0 ? r.step[0] : 0; // :85
        vy += m1 ? r.step[1] : 0;
        vz += m2 ? r.step[2] : 0;
        last = (m0 ? 1u : 0u) | (m1 ? 2u : 0u) | (m2 ? 4u : 0u);
        ++steps; // :86
    }
    r.v[0] = vx; r.v[1] = vy; r.v[2] = vz;
    r.side[0] = sx; r.side[1] = sy; r.side[2] = sz;
    r.steps = steps;
    r.last_mask = last;
}

// world_grid.cuh — uniform world-space grid over the instances' bounding boxes.
// Included from kernels.cu inside namespace vt.
//
// The path tracer's bounce rays must find, among ALL instances, the ones a world-space ray
// enters, in increasing (tn, index) order.  The reference has no such query (its rasteriser only ever
// shoots camera rays, lib/command.c:102); a loop over every instance is exact but O(n) per ray segment
// and the engine draws 121 to 9 261 instances per frame (src/world.rs:143-198).  The grid narrows the loop
// to the instances registered in the cells the ray passes, WITHOUT changing the answer:
//   * an instance is registered in every cell its (slightly enlarged) world bounding box touches;
//   * a ray visits its cells in order; in a cell it accepts candidates whose box-entry parameter
//     tn <= the parameter at which the ray leaves the cell, smallest (tn, index) first — exactly the
//     order of the full loop, because an instance entered at tn has its entry point, hence its box, in the
//     cell the ray is in at tn;
//   * tn itself comes from the same slab test, in the instance's model space, as in the full loop.
// Rebuilt every frame on the device (instances move every frame, src/world.rs:120-141): bounds +
// resolution (one block), count, scan (one block), fill.  If the instance lists do not fit the frame
// falls back to the full loop (hdr->overflow) and the host grows the list for the next frame.
#pragma once

static constexpr float kWorldPad = 1.0e-3f; // relative enlargement of every box; rounding in the cell walk is ~1e-6

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// One block.  aabb[j] = world bounding box of instance j's unit cube (lo xyz, hi xyz), empty (lo > hi) for
// instances that cannot be hit; hdr = bounds of all boxes, cell size ~ the average instance size.
__global__ void __launch_bounds__(1024) world_bounds_kernel(const InstUniforms* __restrict__ inst, uint32_t n, float* __restrict__ aabb,
                                                            WorldGrid* __restrict__ hdr) {
    __shared__ float s_red[8][32];
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    float size_sum = 0.0f, cnt = 0.0f, bad = 0.0f;
    for (uint32_t j = threadIdx.x; j < n; j += blockDim.x) {
        const InstUniforms* J = inst + j;
        float b[6] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY};
        if (J->valid) {
            float emax = 0.0f;
            bool finite = true;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float c = J->M[3 * 3 + k];
                float e = 0.5f * ((fabsf(J->M[0 * 3 + k]) + fabsf(J->M[1 * 3 + k])) + fabsf(J->M[2 * 3 + k]));
                e = e * (1.0f + kWorldPad) + 1.0e-6f * (fabsf(c) + 1.0f);
                b[k] = c - e;
                b[3 + k] = c + e;
                emax = fmaxf(emax, e);
                finite = finite && isfinite(c) && isfinite(e);
            }
            if (finite) {
#pragma unroll
                for (int k = 0; k < 3; ++k) { lo[k] = fminf(lo[k], b[k]); hi[k] = fmaxf(hi[k], b[3 + k]); }
                size_sum += 2.0f * emax;
                cnt += 1.0f;
            } else {
                bad = 1.0f; // a non-finite model matrix: no grid this frame
            }
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) aabb[(size_t)j * 6 + k] = b[k];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float v[9] = {lo[0], lo[1], lo[2], hi[0], hi[1], hi[2], size_sum, cnt, bad};
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        v[k] = k < 3 ? warp_min(v[k]) : (k < 6 ? warp_max(v[k]) : warp_sum(v[k]));
    }
    // 9 values through 8 rows: bad rides with cnt (cnt < 0 marks it)
    if (v[8] > 0.0f) v[7] = -1.0e30f;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) s_red[k][warp] = v[k];
    }
    __syncthreads();
    if (warp == 0) {
        const int nw = (blockDim.x + 31) >> 5;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float x = lane < nw ? s_red[k][lane] : (k < 3 ? INFINITY : (k < 6 ? -INFINITY : 0.0f));
            v[k] = k < 3 ? warp_min(x) : (k < 6 ? warp_max(x) : warp_sum(x));
        }
        if (lane == 0) {
            WorldGrid g;
            g.total = 0;
            g.overflow = v[7] < 0.0f ? 1u : 0u;
            g.pad = 0;
            const float count = v[7];
            if (!(count > 0.0f)) { // nothing to hit (or a bad matrix): one empty cell
#pragma unroll
                for (int k = 0; k < 3; ++k) { g.lo[k] = 0.0f; g.cell[k] = 1.0f; g.inv_cell[k] = 1.0f; g.res[k] = 1; }
            } else {
                const float s = v[6] / count; // average instance size
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float pad = kWorldPad * s;
                    const float l = v[k] - pad;
                    const float ext = (v[3 + k] + pad) - l;
                    float r = ceilf(ext / s);
                    r = r < 1.0f ? 1.0f : (r > (float)kWorldGridMaxRes ? (float)kWorldGridMaxRes : r);
                    g.lo[k] = l;
                    g.res[k] = (uint32_t)r;
                    g.cell[k] = ext / r;
                    g.inv_cell[k] = r / ext;
                }
            }
            g.n_cells = g.res[0] * g.res[1] * g.res[2];
            *hdr = g;
        }
    }
}

// cell range [c0, c1] (inclusive, per axis) a box touches; false for empty boxes
__device__ __forceinline__ bool world_cell_range(const WorldGrid& g, const float* __restrict__ b, int c0[3], int c1[3]) {
    if (!(b[0] <= b[3])) return false;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int r = (int)g.res[k];
        int a = (int)floorf((b[k] - g.lo[k]) * g.inv_cell[k]);
        int z = (int)floorf((b[3 + k] - g.lo[k]) * g.inv_cell[k]);
        c0[k] = a < 0 ? 0 : (a > r - 1 ? r - 1 : a);
        c1[k] = z < 0 ? 0 : (z > r - 1 ? r - 1 : z);
    }
    return true;
}

// pass 0: count entries per cell; pass 1: write them
template <int kPass>
__global__ void world_register_kernel(const float* __restrict__ aabb, uint32_t n, const WorldGrid* __restrict__ hdr,
                                      uint32_t* __restrict__ count, const uint32_t* __restrict__ offset, uint32_t* __restrict__ cursor,
                                      uint32_t* __restrict__ list) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const WorldGrid g = *hdr;
    if (g.overflow) return;
    int c0[3], c1[3];
    if (!world_cell_range(g, aabb + (size_t)j * 6, c0, c1)) return;
    for (int z = c0[2]; z <= c1[2]; ++z)
        for (int y = c0[1]; y <= c1[1]; ++y)
            for (int x = c0[0]; x <= c1[0]; ++x) {
                const uint32_t cell = ((uint32_t)z * g.res[1] + (uint32_t)y) * g.res[0] + (uint32_t)x;
                if (kPass == 0) atomicAdd(count + cell, 1u);
                else list[offset[cell] + atomicAdd(cursor + cell, 1u)] = j;
            }
}

// One block: exclusive scan of count[0 .. n_cells) into offset, cursor cleared, total / overflow into the header.
__global__ void __launch_bounds__(1024) world_scan_kernel(WorldGrid* __restrict__ hdr, const uint32_t* __restrict__ count,
                                                          uint32_t* __restrict__ offset, uint32_t* __restrict__ cursor, uint32_t capacity) {
    __shared__ uint32_t s_part[1024];
    const uint32_t n = hdr->n_cells;
    const uint32_t per = (n + blockDim.x - 1) / blockDim.x;
    const uint32_t begin = threadIdx.x * per, end = begin + per < n ? begin + per : n;
    uint32_t sum = 0;
    for (uint32_t i = begin; i < end; ++i) sum += count[i];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    // Hillis-Steele over the 1024 partial sums
    for (uint32_t d = 1; d < blockDim.x; d <<= 1) {
        const uint32_t v = threadIdx.x >= d ? s_part[threadIdx.x - d] : 0u;
        __syncthreads();
        s_part[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = s_part[threadIdx.x] - sum; // exclusive
    for (uint32_t i = begin; i < end; ++i) {
        offset[i] = run;
        cursor[i] = 0;
        run += count[i];
    }
    if (threadIdx.x == blockDim.x - 1) {
        hdr->total = s_part[threadIdx.x];
        if (s_part[threadIdx.x] > capacity) hdr->overflow = 1u;
    }
}

cudaError_t launch_world_grid(const InstUniforms* inst, uint32_t n_inst, float* aabb, WorldGrid* hdr, uint32_t* offset, uint32_t* count,
                              uint32_t* cursor, uint32_t* list, uint32_t capacity, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(count, 0, (size_t)kWorldGridMaxCells * 4, stream);
    if (e != cudaSuccess) return e;
    world_bounds_kernel<<<1, 1024, 0, stream>>>(inst, n_inst, aabb, hdr);
    const unsigned blocks = (n_inst + 127) / 128;
    world_register_kernel<0><<<blocks, 128, 0, stream>>>(aabb, n_inst, hdr, count, offset, cursor, list);
    world_scan_kernel<<<1, 1024, 0, stream>>>(hdr, count, offset, cursor, capacity);
    world_register_kernel<1><<<blocks, 128, 0, stream>>>(aabb, n_inst, hdr, count, offset, cursor, list);
    return cudaGetLastError();
}

// ---- traversal --------------------------------------------------------------------------------

// Cell walk of a world ray (Amanatides & Woo); the ray starts inside the grid (a bounce leaves from the
// surface of an instance, whose box is inside the bounds).
struct WorldWalk {
    int c[3], step[3];
    float tmax[3], tdelta[3];
    uint32_t res[3];
};

__device__ __forceinline__ void world_walk_begin(const WorldGrid& g, const float ow[3], const float dw[3], WorldWalk& w) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int r = (int)g.res[k];
        int c = (int)floorf((ow[k] - g.lo[k]) * g.inv_cell[k]);
        c = c < 0 ? 0 : (c > r - 1 ? r - 1 : c);
        w.c[k] = c;
        w.res[k] = g.res[k];
        if (dw[k] > 0.0f) {
            w.step[k] = 1;
            w.tmax[k] = ((g.lo[k] + (float)(c + 1) * g.cell[k]) - ow[k]) / dw[k];
            w.tdelta[k] = g.cell[k] / dw[k];
        } else if (dw[k] < 0.0f) {
            w.step[k] = -1;
            w.tmax[k] = ((g.lo[k] + (float)c * g.cell[k]) - ow[k]) / dw[k];
            w.tdelta[k] = -g.cell[k] / dw[k];
        } else {
            w.step[k] = 0;
            w.tmax[k] = INFINITY;
            w.tdelta[k] = INFINITY;
        }
        if (!(w.tmax[k] == w.tmax[k])) w.tmax[k] = INFINITY; // NaN direction components: never cross on this axis
    }
}

// index of the current cell and the ray parameter at which the ray leaves it (+inf for the last cell on the
// way out, so that rounding at the outer boundary cannot drop a candidate)
__device__ __forceinline__ uint32_t world_walk_cell(const WorldWalk& w, float& t_exit, bool& last_cell) {
    const float t = fminf(fminf(w.tmax[0], w.tmax[1]), w.tmax[2]);
    const int a = (w.tmax[0] <= w.tmax[1] && w.tmax[0] <= w.tmax[2]) ? 0 : (w.tmax[1] <= w.tmax[2] ? 1 : 2);
    const int next = (a == 0 ? w.c[0] + w.step[0] : (a == 1 ? w.c[1] + w.step[1] : w.c[2] + w.step[2]));
    const int r = (int)(a == 0 ? w.res[0] : (a == 1 ? w.res[1] : w.res[2]));
    const bool leaves = !(t < INFINITY) || next < 0 || next >= r;
    t_exit = leaves ? INFINITY : t;
    last_cell = leaves;
    return ((uint32_t)w.c[2] * w.res[1] + (uint32_t)w.c[1]) * w.res[0] + (uint32_t)w.c[0];
}

__device__ __forceinline__ void world_walk_next(WorldWalk& w) {
    const int a = (w.tmax[0] <= w.tmax[1] && w.tmax[0] <= w.tmax[2]) ? 0 : (w.tmax[1] <= w.tmax[2] ? 1 : 2);
    if (a == 0) { w.c[0] += w.step[0]; w.tmax[0] += w.tdelta[0]; }
    else if (a == 1) { w.c[1] += w.step[1]; w.tmax[1] += w.tdelta[1]; }
    else { w.c[2] += w.step[2]; w.tmax[2] += w.tdelta[2]; }
}

// kernels.cu — sm_100a kernels for vtrace's voxel ray-traversal path.
//
// What runs here replaces the GPU programs of the reference:
//   instance_setup_kernel  <- shaders/trace.vert:32-47 (per-instance part) + the uniform
//                             inverse() calls trace.frag repeats per fragment (:48,49,59,65)
//   trace_primary_kernel   <- rasteriser + shaders/trace.frag:41-90 + depth/blend state
//                             (lib/pipeline.c:114-152, lib/command.c:56-102)
//   trace_paths_kernel     <- extension: jittered paths with diffuse bounces
//   build_mask_kernel      <- takes the slot of lib/raytrace.c's acceleration-structure build
//
// Arithmetic contract: compiled with --fmad=false, IEEE div/sqrt, no fast-math.  Every
// float op that can influence a traversal decision is a separately rounded binary32 op in
// the same order as oracle/vtrace_oracle.c, so hit voxel / face / steps are bit-exact.
//
// The DDA keeps the reference's per-voxel float stepping (side += delta is a running sum,
// trace.frag:84 — closed-form skipping would change `steps`), but makes each step cheap:
// one bit test against a padded "stop mask" (filled OR outside), no coordinate compares,
// no texel fetch until the hit.  For scenes whose masks fit, the whole mask arena is
// staged into shared memory with one TMA bulk copy per CTA.
#include "kernels.h"

#include <cstdint>

namespace vt {

#define VT_FLAG_VIEWPORT_H_IS_W 1u
#define VT_FLAG_NO_HIT_RECORDS 2u
#define VT_FLAG_PER_PIXEL_PATHS 16u
#define VT_FLAG_SHADOW_RAYS 64u
#define VT_MISS 0xFFFFFFFFu

static constexpr int kBlockThreads = 256; // 8 warps
static constexpr int kTileW = 8, kTileH = 4; // one warp = one 8x4 pixel tile, so neighbouring lanes trace neighbouring rays

// -------------------------------------------------------------------------------------------
// dynamic shared memory layout:
//   [0,16) mbarrier | [16,1040) sRGB decode LUT | [1040,4112) per-warp radiance accumulators | masks
static constexpr uint32_t kSmemLutOff = 16;
static constexpr uint32_t kSmemAccOff = 16 + 1024;
static constexpr uint32_t kSmemMaskOff = kSmemAccOff + (kBlockThreads / 32) * 32 * 3 * 4;
static_assert(kSmemMaskOff % 16 == 0, "bulk copy destination must be 16-byte aligned");

size_t trace_smem_bytes(uint32_t arena_words, bool masks_in_smem) {
    return kSmemMaskOff + (masks_in_smem ? size_t(arena_words) * 4 : 0);
}

// -------------------------------------------------------------------------------------------
// small helpers

__device__ __forceinline__ float vt_fmin(float a, float b) { // IEEE minNum, as the oracle fixes it
    if (a != a) return b;
    if (b != b) return a;
    return a < b ? a : b;
}

__device__ __forceinline__ int32_t texel_of(int32_t v, float size, int32_t isize) {
    // texture(tex, voxel / size), NEAREST, clamp-to-edge (trace.frag:76, lib/descriptor.c:100-115)
    float u = (float)v / size;
    int32_t i = __float2int_rz(floorf(u * size));
    i = i < 0 ? 0 : i;
    i = i > isize - 1 ? isize - 1 : i;
    return i;
}

extern __shared__ __align__(128) unsigned char vt_smem[];

// mbarrier + TMA bulk copy (cp.async.bulk, SASS UBLKCP) -------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// -------------------------------------------------------------------------------------------
// build_mask_kernel: RGBA8 texels -> padded stop mask.  One thread per 32-bit mask word.

__global__ void build_mask_kernel(const uint8_t* __restrict__ rgba, uint32_t W, uint32_t H, uint32_t D, uint32_t xb,
                                  uint32_t yb, uint32_t* __restrict__ mask, uint32_t mask_words) {
    uint32_t word = blockIdx.x * blockDim.x + threadIdx.x;
    if (word >= mask_words) return;
    const uint32_t bit0 = word << 5;
    const uint32_t yq = (bit0 >> xb) & ((1u << yb) - 1u);
    const uint32_t zq = bit0 >> (xb + yb);
    const int32_t y = (int32_t)yq - 1, z = (int32_t)zq - 1;
    uint32_t bits = 0;
    const bool row_inside = y >= 0 && y < (int32_t)H && z >= 0 && z < (int32_t)D;
    int32_t ty = 0, tz = 0;
    if (row_inside) {
        ty = texel_of(y, (float)(int32_t)H, (int32_t)H);
        tz = texel_of(z, (float)(int32_t)D, (int32_t)D);
    }
    for (uint32_t b = 0; b < 32; ++b) {
        const uint32_t xq = (bit0 & ((1u << xb) - 1u)) + b;
        const int32_t x = (int32_t)xq - 1;
        bool stop = true; // border and unused padding both stop the walk
        if (row_inside && x >= 0 && x < (int32_t)W) {
            const int32_t tx = texel_of(x, (float)(int32_t)W, (int32_t)W);
            const size_t t = (size_t)tx + (size_t)W * ((size_t)ty + (size_t)H * (size_t)tz);
            stop = rgba[4 * t + 3] > 0; // texSample.w > 0.0, trace.frag:78
        }
        bits |= (stop ? 1u : 0u) << b;
    }
    mask[word] = bits;
}

cudaError_t launch_build_mask(const uint8_t* rgba, uint32_t w, uint32_t h, uint32_t d, uint32_t xb, uint32_t yb,
                              uint32_t* mask, uint32_t mask_words, cudaStream_t stream) {
    const int threads = 128;
    build_mask_kernel<<<(mask_words + threads - 1) / threads, threads, 0, stream>>>(rgba, w, h, d, xb, yb, mask, mask_words);
    return cudaGetLastError();
}

// -------------------------------------------------------------------------------------------
// instance_setup_kernel: one thread per instance.

// GLSL inverse(mat4) — same cofactor expansion, same operation order as the oracle.
__host__ __device__ void mat4_inverse(const float* m, float* r) {
#define A(r_, c_) m[(c_)*4 + (r_)]
#define B(r_, c_) r[(c_)*4 + (r_)]
    float s0 = A(0, 0) * A(1, 1) - A(1, 0) * A(0, 1);
    float s1 = A(0, 0) * A(1, 2) - A(1, 0) * A(0, 2);
    float s2 = A(0, 0) * A(1, 3) - A(1, 0) * A(0, 3);
    float s3 = A(0, 1) * A(1, 2) - A(1, 1) * A(0, 2);
    float s4 = A(0, 1) * A(1, 3) - A(1, 1) * A(0, 3);
    float s5 = A(0, 2) * A(1, 3) - A(1, 2) * A(0, 3);
    float c5 = A(2, 2) * A(3, 3) - A(3, 2) * A(2, 3);
    float c4 = A(2, 1) * A(3, 3) - A(3, 1) * A(2, 3);
    float c3 = A(2, 1) * A(3, 2) - A(3, 1) * A(2, 2);
    float c2 = A(2, 0) * A(3, 3) - A(3, 0) * A(2, 3);
    float c1 = A(2, 0) * A(3, 2) - A(3, 0) * A(2, 2);
    float c0 = A(2, 0) * A(3, 1) - A(3, 0) * A(2, 1);
    float det = ((((s0 * c5 - s1 * c4) + s2 * c3) + s3 * c2) - s4 * c1) + s5 * c0;
    float id = 1.0f / det;
    B(0, 0) = ((A(1, 1) * c5 - A(1, 2) * c4) + A(1, 3) * c3) * id;
    B(0, 1) = ((-A(0, 1) * c5 + A(0, 2) * c4) - A(0, 3) * c3) * id;
    B(0, 2) = ((A(3, 1) * s5 - A(3, 2) * s4) + A(3, 3) * s3) * id;
    B(0, 3) = ((-A(2, 1) * s5 + A(2, 2) * s4) - A(2, 3) * s3) * id;
    B(1, 0) = ((-A(1, 0) * c5 + A(1, 2) * c2) - A(1, 3) * c1) * id;
    B(1, 1) = ((A(0, 0) * c5 - A(0, 2) * c2) + A(0, 3) * c1) * id;
    B(1, 2) = ((-A(3, 0) * s5 + A(3, 2) * s2) - A(3, 3) * s1) * id;
    B(1, 3) = ((A(2, 0) * s5 - A(2, 2) * s2) + A(2, 3) * s1) * id;
    B(2, 0) = ((A(1, 0) * c4 - A(1, 1) * c2) + A(1, 3) * c0) * id;
    B(2, 1) = ((-A(0, 0) * c4 + A(0, 1) * c2) - A(0, 3) * c0) * id;
    B(2, 2) = ((A(3, 0) * s4 - A(3, 1) * s2) + A(3, 3) * s0) * id;
    B(2, 3) = ((-A(2, 0) * s4 + A(2, 1) * s2) - A(2, 3) * s0) * id;
    B(3, 0) = ((-A(1, 0) * c3 + A(1, 1) * c1) - A(1, 2) * c0) * id;
    B(3, 1) = ((A(0, 0) * c3 - A(0, 1) * c1) + A(0, 2) * c0) * id;
    B(3, 2) = ((-A(3, 0) * s3 + A(3, 1) * s1) - A(3, 2) * s0) * id;
    B(3, 3) = ((A(2, 0) * s3 - A(2, 1) * s1) + A(2, 2) * s0) * id;
#undef A
#undef B
}

__global__ void instance_setup_kernel(const float* __restrict__ instances, uint32_t n,
                                      const VolumeDesc* __restrict__ volumes, const FrameParams fp,
                                      InstUniforms* __restrict__ out, uint32_t* flag, uint32_t flag_value,
                                      unsigned long long* zero_stats) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && zero_stats) { // the frame's counters (rays, iterations, work-claim counter, analytic rays) start at 0
        zero_stats[0] = 0ull; zero_stats[1] = 0ull; zero_stats[2] = 0ull; zero_stats[3] = 0ull;
    }
    if (i == 0 && flag) { // fused multi-GPU accumulation, root: "the previous frame is consumed" (see flag_wait below)
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(flag_value) : "memory");
    }
    if (i >= n) return;
    float M[16], Mi[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) M[k] = instances[(size_t)i * 16 + k];
    InstUniforms I;
    I.tex = __float_as_uint(M[15]); // trace.vert:38 floatBitsToInt(model[3][3])
    M[15] = 1.0f;                   // trace.vert:39-40
    I.valid = I.tex < fp.n_volumes ? 1u : 0u;
    I.pad = 0;
    I.pad0 = 0;
    I.bricks = nullptr;
    I.sun_m[0] = I.sun_m[1] = I.sun_m[2] = 0.0f;
    if (!I.valid) {
        I.w = I.h = I.d = I.xb = I.yb = I.mask_off = I.remap_identity = 0;
        I.rgba = nullptr;
        for (int k = 0; k < 16; ++k) I.MVP[k] = 0.0f;
        for (int k = 0; k < 12; ++k) I.Mi[k] = I.M[k] = I.dirm[k] = 0.0f;
        for (int k = 0; k < 3; ++k) I.eye_m[k] = I.slab_lo[k] = I.slab_hi[k] = I.lin[k] = I.ilin[k] = 0.0f;
        I.pad1[0] = I.pad1[1] = 0;
        I.bounds[0] = 1; I.bounds[1] = 0; I.bounds[2] = 1; I.bounds[3] = 0; // empty
        out[i] = I;
        return;
    }
    const VolumeDesc v = volumes[I.tex];
    I.w = v.w; I.h = v.h; I.d = v.d; I.xb = v.xb; I.yb = v.yb; I.mask_off = v.mask_off; I.rgba = v.rgba;
    I.remap_identity = v.remap_identity;
    I.bricks = v.bricks;
    mat4_inverse(M, Mi); // trace.frag:65
    // MVP = (P V) * M
    for (int j = 0; j < 4; ++j)
        for (int r = 0; r < 4; ++r)
            I.MVP[j * 4 + r] = ((fp.PV[0 * 4 + r] * M[j * 4 + 0] + fp.PV[1 * 4 + r] * M[j * 4 + 1]) + fp.PV[2 * 4 + r] * M[j * 4 + 2]) +
                               fp.PV[3 * 4 + r] * M[j * 4 + 3];
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 3; ++r) {
            I.Mi[c * 3 + r] = Mi[c * 4 + r];
            I.M[c * 3 + r] = M[c * 4 + r];
            I.dirm[c * 3 + r] = (Mi[0 * 4 + r] * fp.RD[c * 4 + 0] + Mi[1 * 4 + r] * fp.RD[c * 4 + 1]) + Mi[2 * 4 + r] * fp.RD[c * 4 + 2];
        }
    for (int r = 0; r < 3; ++r) {
        I.eye_m[r] = ((Mi[0 * 4 + r] * fp.eye[0] + Mi[1 * 4 + r] * fp.eye[1]) + Mi[2 * 4 + r] * fp.eye[2]) + Mi[3 * 4 + r] * 1.0f;
        I.slab_lo[r] = -0.5f - I.eye_m[r];
        I.slab_hi[r] = 0.5f - I.eye_m[r];
        I.sun_m[r] = (Mi[0 * 4 + r] * fp.sun[0] + Mi[1 * 4 + r] * fp.sun[1]) + Mi[2 * 4 + r] * fp.sun[2];
    }
    {   // is inverse(M)'s linear part diagonal with powers of two?  (see InstUniforms::lin)
        bool ok = true;
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 3; ++r) {
                const uint32_t bits = __float_as_uint(Mi[c * 4 + r]);
                if (c != r) ok = ok && (bits << 1) == 0u; // +-0
                else ok = ok && (bits & 0x007FFFFFu) == 0u && ((bits >> 23) & 0xFFu) >= 64u && ((bits >> 23) & 0xFFu) <= 190u; // +-2^e, |e| <= 63
            }
        for (int k = 0; k < 3; ++k) {
            I.lin[k] = ok ? Mi[k * 4 + k] : 0.0f;
            I.ilin[k] = ok ? 1.0f / Mi[k * 4 + k] : 0.0f;
        }
        I.pad1[0] = I.pad1[1] = 0;
    }
    // Conservative screen rectangle of the proxy cube (what the rasteriser would bin): project the 8 corners, +-2 pixels of
    // slack.  A pixel outside the rectangle cannot be covered for any sample position inside it, so the trace kernels may
    // skip the instance (and, when no instance remains, the pixel) without changing any result.
    // A cube that reaches behind the eye plane (w <= 0) is first cut at w = eps: what lies in 0 < w < eps can only be seen
    // by an on-screen pixel if |x|, |y| <= w < eps in clip space, i.e. within |A^-1|_inf * eps of the centre of projection
    // e (A = rows x, y, w of MVP, A e = -t) — and eps is chosen as half of what the cube's distance from e allows.  Eye
    // inside or next to the cube, or no centre of projection (singular A): whole screen.
    float minx = INFINITY, maxx = -INFINITY, miny = INFINITY, maxy = -INFINITY;
    bool whole = false, cut = false;
    float cX[8], cY[8], cW[8], wmin = INFINITY;
    for (int c = 0; c < 8; ++c) {
        const float cx = (c & 1) ? 0.5f : -0.5f, cy = (c & 2) ? 0.5f : -0.5f, cz = (c & 4) ? 0.5f : -0.5f;
        cX[c] = ((I.MVP[0] * cx + I.MVP[4] * cy) + I.MVP[8] * cz) + I.MVP[12];
        cY[c] = ((I.MVP[1] * cx + I.MVP[5] * cy) + I.MVP[9] * cz) + I.MVP[13];
        cW[c] = ((I.MVP[3] * cx + I.MVP[7] * cy) + I.MVP[11] * cz) + I.MVP[15];
        wmin = fminf(wmin, cW[c]);
        if (!(cW[c] == cW[c])) whole = true;
    }
    float eps = 1e-6f, slack = 2.0f;
    if (!whole && !(wmin > 1e-6f)) {
        const float a00 = I.MVP[0], a01 = I.MVP[4], a02 = I.MVP[8], a10 = I.MVP[1], a11 = I.MVP[5], a12 = I.MVP[9];
        const float a20 = I.MVP[3], a21 = I.MVP[7], a22 = I.MVP[11];
        const float c00 = a11 * a22 - a12 * a21, c01 = a12 * a20 - a10 * a22, c02 = a10 * a21 - a11 * a20;
        const float det = (a00 * c00 + a01 * c01) + a02 * c02;
        const float id = 1.0f / det;
        // rows of A^-1
        const float r0[3] = {c00 * id, (a02 * a21 - a01 * a22) * id, (a01 * a12 - a02 * a11) * id};
        const float r1[3] = {c01 * id, (a00 * a22 - a02 * a20) * id, (a02 * a10 - a00 * a12) * id};
        const float r2[3] = {c02 * id, (a01 * a20 - a00 * a21) * id, (a00 * a11 - a01 * a10) * id};
        const float t[3] = {I.MVP[12], I.MVP[13], I.MVP[15]};
        const float e0 = -((r0[0] * t[0] + r0[1] * t[1]) + r0[2] * t[2]);
        const float e1 = -((r1[0] * t[0] + r1[1] * t[1]) + r1[2] * t[2]);
        const float e2 = -((r2[0] * t[0] + r2[1] * t[1]) + r2[2] * t[2]);
        const float dinf = fmaxf(fmaxf(fabsf(e0), fabsf(e1)), fabsf(e2)) - 0.5f;
        const float n0 = (fabsf(r0[0]) + fabsf(r0[1])) + fabsf(r0[2]);
        const float n1 = (fabsf(r1[0]) + fabsf(r1[1])) + fabsf(r1[2]);
        const float n2 = (fabsf(r2[0]) + fabsf(r2[1])) + fabsf(r2[2]);
        eps = 0.5f * dinf / fmaxf(fmaxf(n0, n1), n2);
        if (!(eps >= 1e-2f) || !(eps < INFINITY)) {
            whole = true;
        } else {
            cut = true;
            // rounding of a cut point's x and y, magnified by 1 / eps, in pixels (generous)
            const float sx = ((fabsf(a00) + fabsf(a01)) + fabsf(a02)) + fabsf(t[0]);
            const float sy = ((fabsf(a10) + fabsf(a11)) + fabsf(a12)) + fabsf(t[1]);
            slack += 16.0f * 1.1920929e-7f * fmaxf(sx / fp.sxn, sy / fp.syn) / eps;
        }
    }
    bool any = false;
    if (!whole) {
        auto add = [&](float X, float Y, float W) {
            const float fx = (X / W + 1.0f) / fp.sxn, fy = (Y / W + 1.0f) / fp.syn;
            if (!(fx == fx) || !(fy == fy)) { whole = true; return; }
            minx = fminf(minx, fx); maxx = fmaxf(maxx, fx);
            miny = fminf(miny, fy); maxy = fmaxf(maxy, fy);
            any = true;
        };
        for (int c = 0; c < 8; ++c)
            if (!cut || cW[c] >= eps) add(cX[c], cY[c], cW[c]);
        if (cut) {
            for (int c = 0; c < 8; ++c)
                for (int ax = 1; ax < 8; ax <<= 1) {
                    if (c & ax) continue;
                    const int d = c | ax;
                    if ((cW[c] >= eps) == (cW[d] >= eps)) continue;
                    const float sp = (eps - cW[c]) / (cW[d] - cW[c]);
                    add(cX[c] + sp * (cX[d] - cX[c]), cY[c] + sp * (cY[d] - cY[c]), eps);
                }
        }
    }
    if (whole) {
        I.bounds[0] = 0; I.bounds[1] = fp.width - 1; I.bounds[2] = 0; I.bounds[3] = fp.height - 1;
    } else if (!any) { // entirely behind the eye plane
        I.bounds[0] = 1; I.bounds[1] = 0; I.bounds[2] = 1; I.bounds[3] = 0;
    } else {
        const float lim = 1.0e9f;
        minx = fmaxf(minx - slack, -lim); maxx = fminf(maxx + slack, lim); miny = fmaxf(miny - slack, -lim); maxy = fminf(maxy + slack, lim);
        int x0 = __float2int_rd(minx), x1 = __float2int_rd(maxx);
        int y0 = __float2int_rd(miny), y1 = __float2int_rd(maxy);
        I.bounds[0] = x0 < 0 ? 0 : x0;
        I.bounds[1] = x1 > fp.width - 1 ? fp.width - 1 : x1;
        I.bounds[2] = y0 < 0 ? 0 : y0;
        I.bounds[3] = y1 > fp.height - 1 ? fp.height - 1 : y1;
    }
    out[i] = I;
}

cudaError_t launch_instance_setup(const float* instances, uint32_t n, const VolumeDesc* volumes, FrameParams fp,
                                  InstUniforms* out, uint32_t* flag, uint32_t flag_value, unsigned long long* zero_stats,
                                  cudaStream_t stream) {
    const int threads = 64;
    instance_setup_kernel<<<(n + threads - 1) / threads, threads, 0, stream>>>(instances, n, volumes, fp, out, flag, flag_value, zero_stats);
    return cudaGetLastError();
}

// -------------------------------------------------------------------------------------------
// bin_instances_kernel: one warp per 16x16-pixel bin.  Pass 1 counts the instances whose screen
// rectangle touches the bin, one atomic reserves the bin's segment of the list, pass 2 writes the
// instance indices in ascending order (= draw order, which the depth/blend rule depends on).
__global__ void bin_instances_kernel(const InstUniforms* __restrict__ inst, uint32_t n_inst, uint32_t bins_x, uint32_t bins_y,
                                     uint32_t* __restrict__ offset, uint32_t* __restrict__ count, uint32_t* __restrict__ list,
                                     uint32_t capacity, uint32_t* __restrict__ cursor) {
    const uint32_t bin = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    if (bin >= bins_x * bins_y) return;
    const int bx0 = (int)((bin % bins_x) << kBinShift), by0 = (int)((bin / bins_x) << kBinShift);
    const int bx1 = bx0 + (1 << kBinShift) - 1, by1 = by0 + (1 << kBinShift) - 1;
    uint32_t n = 0;
    for (uint32_t i0 = 0; i0 < n_inst; i0 += 32) {
        const uint32_t i = i0 + lane;
        bool t = false;
        if (i < n_inst) {
            const int4 b = *reinterpret_cast<const int4*>(inst[i].bounds);
            t = !(b.y < bx0 || b.x > bx1 || b.w < by0 || b.z > by1);
        }
        n += __popc(__ballot_sync(0xffffffffu, t));
    }
    uint32_t base = 0;
    if (lane == 0) base = n ? atomicAdd(cursor, n) : 0u;
    base = __shfl_sync(0xffffffffu, base, 0);
    if (lane == 0) { offset[bin] = base; count[bin] = (base + n <= capacity) ? n : 0xFFFFFFFFu; }
    if (base + n > capacity) return; // overflow: this bin's pixels visit every instance (bin_range); the host grows the list
    uint32_t w = base;
    for (uint32_t i0 = 0; i0 < n_inst; i0 += 32) {
        const uint32_t i = i0 + lane;
        bool t = false;
        if (i < n_inst) {
            const int4 b = *reinterpret_cast<const int4*>(inst[i].bounds);
            t = !(b.y < bx0 || b.x > bx1 || b.w < by0 || b.z > by1);
        }
        const uint32_t m = __ballot_sync(0xffffffffu, t);
        if (t) list[w + __popc(m & ((1u << lane) - 1u))] = i;
        w += __popc(m);
    }
}

cudaError_t launch_bin_instances(const InstUniforms* inst, uint32_t n_inst, uint32_t bins_x, uint32_t bins_y, uint32_t* offset,
                                 uint32_t* count, uint32_t* list, uint32_t capacity, uint32_t* cursor, cudaStream_t stream) {
    const uint32_t warps = bins_x * bins_y;
    const int threads = 256;
    bin_instances_kernel<<<(warps * 32 + threads - 1) / threads, threads, 0, stream>>>(inst, n_inst, bins_x, bins_y, offset, count, list,
                                                                                     capacity, cursor);
    return cudaGetLastError();
}

// Instances a pixel has to visit, in draw order: its 16x16 bin's segment of the binned list, or every instance when
// the scene is not binned — or when the bin's segment did not fit the list (count == 0xFFFFFFFF, see
// bin_instances_kernel): such a frame is still exact, only slower; the host grows the list for the next one.
__device__ __forceinline__ const uint32_t* bin_range(const BinTable& bins, int px, int py, uint32_t n_inst, uint32_t& k_begin,
                                                     uint32_t& k_end) {
    k_begin = 0;
    k_end = n_inst;
    if (!bins.enabled) return nullptr;
    const uint32_t bin = ((uint32_t)py >> kBinShift) * bins.bins_x + ((uint32_t)px >> kBinShift);
    const uint32_t cnt = __ldg(bins.count + bin);
    if (cnt == 0xFFFFFFFFu) return nullptr;
    k_begin = __ldg(bins.offset + bin);
    k_end = k_begin + cnt;
    return bins.list;
}

// -------------------------------------------------------------------------------------------
// rasteriser restatement: which point of the proxy cube's front faces covers the sample

// lo[k] = -0.5 - o[k], hi[k] = 0.5 - o[k] (precomputed per instance for camera rays)
__device__ __forceinline__ bool slab_unit_cube(const float o[3], const float lo3[3], const float hi3[3], const float d[3],
                                               float& tn_out, int& axis_out) {
    float tn = -INFINITY, tf = INFINITY;
    int axis = -1;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (d[k] == 0.0f) {
            if (o[k] < -0.5f || o[k] > 0.5f) return false;
            continue;
        }
        const float inv = 1.0f / d[k];
        const float t1 = lo3[k] * inv;
        const float t2 = hi3[k] * inv;
        const float lo = t1 < t2 ? t1 : t2;
        const float hi = t1 < t2 ? t2 : t1;
        if (lo > tn) { tn = lo; axis = k; }
        if (hi < tf) tf = hi;
    }
    if (axis < 0) return false;
    if (!(tn <= tf)) return false;
    if (!(tn > 0.0f)) return false; // inside / behind: only back faces -> culled (lib/pipeline.c:120-121)
    tn_out = tn;
    axis_out = axis;
    return true;
}

// the same with 1 / d[k] supplied (bit-identical when inv[k] == 1.0f / d[k])
__device__ __forceinline__ bool slab_unit_cube_inv(const float o[3], const float lo3[3], const float hi3[3], const float d[3],
                                                   const float inv3[3], float& tn_out, int& axis_out) {
    float tn = -INFINITY, tf = INFINITY;
    int axis = -1;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (d[k] == 0.0f) {
            if (o[k] < -0.5f || o[k] > 0.5f) return false;
            continue;
        }
        const float t1 = lo3[k] * inv3[k];
        const float t2 = hi3[k] * inv3[k];
        const float lo = t1 < t2 ? t1 : t2;
        const float hi = t1 < t2 ? t2 : t1;
        if (lo > tn) { tn = lo; axis = k; }
        if (hi < tf) tf = hi;
    }
    if (axis < 0) return false;
    if (!(tn <= tf)) return false;
    if (!(tn > 0.0f)) return false;
    tn_out = tn;
    axis_out = axis;
    return true;
}

__device__ __forceinline__ void entry_point(const float o[3], const float d[3], float tn, int axis, float mp[3]) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float p = o[k] + tn * d[k];
        p = p < -0.5f ? -0.5f : p;
        p = p > 0.5f ? 0.5f : p;
        mp[k] = (k == axis) ? (d[k] > 0.0f ? -0.5f : 0.5f) : p;
    }
}

// -------------------------------------------------------------------------------------------
// the DDA (trace.frag:63-89)

struct Dda {
    bool hit;
    int32_t v[3];       // model_ray_voxel at exit
    uint32_t steps;
    uint32_t last_mask; // axes advanced by the last executed iteration
    int32_t step[3];
    float side[3], delta[3], dir[3], pos[3], len;
};

// Volume view used by the march
struct Vol {
    uint32_t w, h, d, xb, yb;
    uint32_t mask_off;          // word offset of this volume's mask inside the arena
    const uint32_t* arena;      // global arena (used when the masks are not staged)
};

template <bool kSmem>
__device__ __forceinline__ uint32_t mask_word(const Vol& vol, uint32_t word) {
    if (kSmem) // indexing the __shared__ symbol directly keeps this an LDS, not a generic load
        return reinterpret_cast<const uint32_t*>(vt_smem + kSmemMaskOff)[vol.mask_off + word];
    else
        return __ldg(vol.arena + vol.mask_off + word);
}

enum DdaMode { kDdaDone = 0, kDdaFast = 1, kDdaSlow = 2 };

// trace.frag:63-71: ray state.  Returns how the march has to run and the start voxel's bit index.
__device__ __forceinline__ DdaMode dda_init(const Vol& vol, const float pos[3], const float dir[3], bool has_start,
                                            const int32_t sv[3], Dda& r, uint32_t& idx) {
    const int32_t isz[3] = {(int32_t)vol.w, (int32_t)vol.h, (int32_t)vol.d};
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
    idx = 0;
    // the start voxel must lie inside the padded mask (true for every caller: primary rays start
    // inside the volume, bounce rays start at most one voxel outside it)
    const bool in_pad = r.v[0] >= -1 && r.v[0] <= isz[0] && r.v[1] >= -1 && r.v[1] <= isz[1] && r.v[2] >= -1 && r.v[2] <= isz[2];
    if (!in_pad) return kDdaDone;
    idx = (uint32_t)(r.v[0] + 1) | ((uint32_t)(r.v[1] + 1) << vol.xb) | ((uint32_t)(r.v[2] + 1) << (vol.xb + vol.yb));
    const bool finite = isfinite(r.delta[0]) && isfinite(r.delta[1]) && isfinite(r.delta[2]) && isfinite(r.side[0]) &&
                        isfinite(r.side[1]) && isfinite(r.side[2]);
    return finite ? kDdaFast : kDdaSlow;
}

// One fast-path iteration (trace.frag:76-86) on scalar state; returns false when the walk stops.
// Fast path = no NaN/inf anywhere, so `side <= min(other two)` is two ordered compares,
// `vec3(mask) * delta` is a predicated add, every iteration advances >= 1 voxel (so the
// steps < W+H+D bound of :74-75 can never bind), and leaving the volume lands on a set border
// bit of the stop mask — no coordinate compares inside the loop.
template <bool kSmem>
__device__ __forceinline__ bool dda_step(const Vol& vol, float& sx, float& sy, float& sz, float dx, float dy, float dz,
                                         uint32_t& idx, uint32_t& prev, uint32_t& steps, uint32_t ix, uint32_t iy, uint32_t iz) {
    const uint32_t wv = mask_word<kSmem>(vol, idx >> 5);
    // Written in PTX so the instruction selection stays put: the single-bit mask is built while the
    // load is in flight (one LOP3 with predicate output after it), and the per-axis updates are three
    // predicated add pairs instead of select + 3-input add chains.
    uint32_t stop;
    asm("{\n"
        ".reg .u32 b;\n"
        ".reg .pred p;\n"
        "shf.l.wrap.b32 b, 0, 1, %2;\n"
        "and.b32 b, b, %1;\n"
        "setp.ne.u32 p, b, 0;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(stop)
        : "r"(wv), "r"(idx));
    if (stop) return false;
    prev = idx;
    // :83 mask = side <= min(other two); without NaNs that is side == min(all three).  :84-85 updates.
    asm("{\n"
        ".reg .pred px, py, pz;\n"
        ".reg .f32 m;\n"
        "min.f32 m, %0, %1;\n"
        "min.f32 m, m, %2;\n"
        "setp.eq.f32 px, %0, m;\n"
        "setp.eq.f32 py, %1, m;\n"
        "setp.eq.f32 pz, %2, m;\n"
        "@px add.rn.f32 %0, %0, %4;\n"
        "@py add.rn.f32 %1, %1, %5;\n"
        "@pz add.rn.f32 %2, %2, %6;\n"
        "@px add.s32 %3, %3, %7;\n"
        "@py add.s32 %3, %3, %8;\n"
        "@pz add.s32 %3, %3, %9;\n"
        "}\n"
        : "+f"(sx), "+f"(sy), "+f"(sz), "+r"(idx)
        : "f"(dx), "f"(dy), "f"(dz), "r"(ix), "r"(iy), "r"(iz));
    ++steps; // :86
    return true;
}

// After the fast walk stopped at bit index `idx`: exit voxel, hit flag, axes of the last iteration.
__device__ __forceinline__ void dda_finish_fast(const Vol& vol, Dda& r, uint32_t idx, uint32_t prev, uint32_t steps) {
    const uint32_t xb = vol.xb, zb = vol.xb + vol.yb;
    r.steps = steps;
    r.v[0] = (int32_t)(idx & ((1u << xb) - 1u)) - 1;
    r.v[1] = (int32_t)((idx >> xb) & ((1u << vol.yb) - 1u)) - 1;
    r.v[2] = (int32_t)(idx >> zb) - 1;
    r.hit = r.v[0] >= 0 && r.v[0] < (int32_t)vol.w && r.v[1] >= 0 && r.v[1] < (int32_t)vol.h && r.v[2] >= 0 && r.v[2] < (int32_t)vol.d;
    r.last_mask = 0;
    if (steps) {
        // axes advanced by the last iteration, recovered from the index delta
        const int32_t diff = (int32_t)(idx - prev);
        const int32_t qz = (diff + (1 << (zb - 1))) >> zb;
        const int32_t rem = diff - (qz << zb);
        const int32_t qy = (rem + (1 << (xb - 1))) >> xb;
        const int32_t qx = rem - (qy << xb);
        r.last_mask = (qx != 0 ? 1u : 0u) | (qy != 0 ? 2u : 0u) | (qz != 0 ? 4u : 0u);
    }
}

// Slow path: a direction component is exactly 0 (delta = inf, and 0 * inf = NaN from the first
// non-advancing iteration on) or something is NaN.  Literal transcription, including the
// steps < max_steps bound which CAN bind here.
template <bool kSmem>
__device__ __noinline__ void dda_slow_impl(const Vol& vol, Dda& r) {
    const int32_t isz[3] = {(int32_t)vol.w, (int32_t)vol.h, (int32_t)vol.d};
    const uint32_t xb = vol.xb, zb = vol.xb + vol.yb;
    const uint32_t max_steps = vol.w + vol.h + vol.d; // :74
    while (r.steps < max_steps && r.v[0] >= 0 && r.v[1] >= 0 && r.v[2] >= 0 && r.v[0] < isz[0] && r.v[1] < isz[1] &&
           r.v[2] < isz[2]) { // :75
        const uint32_t idx = (uint32_t)(r.v[0] + 1) | ((uint32_t)(r.v[1] + 1) << xb) | ((uint32_t)(r.v[2] + 1) << zb);
        const uint32_t wv = mask_word<kSmem>(vol, idx >> 5);
        if ((wv >> (idx & 31u)) & 1u) { r.hit = true; return; } // :78-80
        const bool m0 = r.side[0] <= vt_fmin(r.side[1], r.side[2]); // :83
        const bool m1 = r.side[1] <= vt_fmin(r.side[2], r.side[0]);
        const bool m2 = r.side[2] <= vt_fmin(r.side[0], r.side[1]);
        r.side[0] += (m0 ? 1.0f : 0.0f) * r.delta[0]; // :84
        r.side[1] += (m1 ? 1.0f : 0.0f) * r.delta[1];
        r.side[2] += (m2 ? 1.0f : 0.0f) * r.delta[2];
        r.v[0] += m0 ? r.step[0] : 0; // :85
        r.v[1] += m1 ? r.step[1] : 0;
        r.v[2] += m2 ? r.step[2] : 0;
        r.last_mask = (m0 ? 1u : 0u) | (m1 ? 2u : 0u) | (m2 ? 4u : 0u);
        ++r.steps; // :86
    }
}

// Out-of-line so the rare slow path costs the callers no registers; it works on stack copies so the
// callers' own ray state never has its address taken (and stays in registers).
template <bool kSmem>
__device__ __forceinline__ void dda_slow(const Vol& vol, Dda& r) {
    Dda tmp = r;
    Vol v = vol;
    dda_slow_impl<kSmem>(v, tmp);
    r = tmp;
}

template <bool kSmem>
__device__ __forceinline__ void dda_march(const Vol& vol, const float pos[3], const float dir[3], bool has_start,
                                          const int32_t sv[3], Dda& r) {
    uint32_t idx;
    const DdaMode mode = dda_init(vol, pos, dir, has_start, sv, r, idx);
    if (mode == kDdaFast) {
        float sx = r.side[0], sy = r.side[1], sz = r.side[2];
        const uint32_t ix = (uint32_t)r.step[0], iy = (uint32_t)r.step[1] << vol.xb, iz = (uint32_t)r.step[2] << (vol.xb + vol.yb);
        uint32_t prev = idx, steps = 0;
        while (dda_step<kSmem>(vol, sx, sy, sz, r.delta[0], r.delta[1], r.delta[2], idx, prev, steps, ix, iy, iz)) {}
        r.side[0] = sx; r.side[1] = sy; r.side[2] = sz;
        dda_finish_fast(vol, r, idx, prev, steps);
    } else if (mode == kDdaSlow) {
        dda_slow<kSmem>(vol, r);
    }
}

#include "bricks.cuh"

// texel colour at the hit voxel (the only volume-texel read of a ray)
__device__ __forceinline__ uchar4 fetch_texel(const uint8_t* rgba, uint32_t W, uint32_t H, uint32_t D, bool identity,
                                              const int32_t v[3]) {
    int32_t tx = v[0], ty = v[1], tz = v[2];
    if (!identity) { // sizes where floor(fl(v/s)*s) != v for some v: apply the reference's coordinate round trip
        tx = texel_of(v[0], (float)(int32_t)W, (int32_t)W);
        ty = texel_of(v[1], (float)(int32_t)H, (int32_t)H);
        tz = texel_of(v[2], (float)(int32_t)D, (int32_t)D);
    }
    const size_t t = (size_t)tx + (size_t)W * ((size_t)ty + (size_t)H * (size_t)tz);
    return __ldg(reinterpret_cast<const uchar4*>(rgba) + t);
}

// -------------------------------------------------------------------------------------------
// one fragment = rasteriser restatement + trace.frag prologue + DDA

struct Fragment {
    bool covered;
    float depth;
    int entry_axis;
    Dda dda;
};

// marches (pos, dir[, start voxel]) through the instance's volume, whatever its kind
template <bool kSmem, bool kBricks>
__device__ __forceinline__ void march_instance(const InstUniforms* __restrict__ Ip, const uint32_t* mask_base, const float pos[3],
                                               const float dir[3], bool has_start, const int32_t sv[3], Dda& r) {
    if (kBricks && Ip->bricks) {
        const BrickVolume bv = *Ip->bricks;
        dda_march_bricks(bv, Ip->w, Ip->h, Ip->d, pos, dir, has_start, sv, r);
    } else {
        Vol vol{Ip->w, Ip->h, Ip->d, Ip->xb, Ip->yb, Ip->mask_off, mask_base};
        dda_march<kSmem>(vol, pos, dir, has_start, sv, r);
    }
}

// colour of the instance's voxel v (dense: the texel; brick volumes: procedural)
template <bool kBricks>
__device__ __forceinline__ uchar4 instance_texel(const InstUniforms* __restrict__ Ip, const int32_t v[3]) {
    if (kBricks && Ip->bricks)
        return proc_color(Ip->bricks, Ip->h, (uint32_t)v[0], (uint32_t)v[1], (uint32_t)v[2]);
    return fetch_texel(Ip->rgba, Ip->w, Ip->h, Ip->d, Ip->remap_identity != 0, v);
}

// Where a secondary ray (bounce or shadow) leaves a hit: axis = first axis advanced by the last DDA
// iteration (or the box-entry axis when steps == 0), origin = hit point clamped to the hit voxel with
// the normal component on the face plane, start voxel = the neighbour across that face.
__device__ __forceinline__ void leave_hit(const Dda& r, int entry_axis, const float size[3], int& a_out, int& nsign_out, float p0[3],
                                          int32_t sv[3]) {
    const uint32_t lm = r.steps ? r.last_mask : (1u << entry_axis);
    const int a = (lm & 1u) ? 0 : ((lm & 2u) ? 1 : 2);
    const float t = r.steps ? ((a == 0 ? r.side[0] : (a == 1 ? r.side[1] : r.side[2])) -
                               (a == 0 ? r.delta[0] : (a == 1 ? r.delta[1] : r.delta[2])))
                            : 0.0f;
    const float tl = t / r.len;
    int nsign = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float p = r.pos[k] + r.dir[k] * tl;
        const float lo = (float)r.v[k], hi = (float)(r.v[k] + 1);
        p = p < lo ? lo : p;
        p = p > hi ? hi : p;
        p0[k] = p;
        sv[k] = r.v[k];
        if (k == a) {
            nsign = r.step[k] != 0 ? -r.step[k] : (r.pos[k] <= 0.5f * size[k] ? -1 : 1);
            p0[k] = (float)(r.v[k] + (nsign > 0 ? 1 : 0));
            sv[k] += nsign;
        }
    }
    a_out = a;
    nsign_out = nsign;
}

template <bool kSmem, bool kBricks = false>
__device__ __forceinline__ void run_fragment(const FrameParams& fp, const InstUniforms* __restrict__ Ip, const uint32_t* mask_base,
                                             int px, int py, float fx, float fy, Fragment& f) {
    f.covered = false;
    f.dda.hit = false;
    f.dda.steps = 0;
    // outside the instance's conservative screen rectangle (also empty for invalid instances)
    if (px < Ip->bounds[0] || px > Ip->bounds[1] || py < Ip->bounds[2] || py > Ip->bounds[3]) return;
    // SURVEY.md §A.2 step 1a
    const float x_ndc = fx * fp.sxn - 1.0f;
    const float y_ndc = fy * fp.syn - 1.0f;
    float d[3], o[3], lo3[3], hi3[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        d[k] = (Ip->dirm[0 * 3 + k] * x_ndc + Ip->dirm[1 * 3 + k] * y_ndc) + Ip->dirm[3 * 3 + k];
        o[k] = Ip->eye_m[k];
        lo3[k] = Ip->slab_lo[k];
        hi3[k] = Ip->slab_hi[k];
    }
    float tn;
    int axis;
    if (!slab_unit_cube(o, lo3, hi3, d, tn, axis)) return;
    float mp[3];
    entry_point(o, d, tn, axis, mp);
    // trace.vert:43-45 at the covered point
    float sp[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
        sp[i] = ((Ip->MVP[0 * 4 + i] * mp[0] + Ip->MVP[1 * 4 + i] * mp[1]) + Ip->MVP[2 * 4 + i] * mp[2]) + Ip->MVP[3 * 4 + i];
    if (!(sp[3] > 0.0f && sp[2] >= 0.0f && sp[2] <= sp[3])) return; // Vulkan clip volume
    f.covered = true;
    f.entry_axis = axis;
    f.depth = sp[2] / sp[3]; // trace.frag:46
    // trace.frag:59 ray_dir = normalize((RD * sp).xyz)
    float rr[3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
        rr[k] = ((fp.RD[0 * 4 + k] * sp[0] + fp.RD[1 * 4 + k] * sp[1]) + fp.RD[2 * 4 + k] * sp[2]) + fp.RD[3 * 4 + k] * sp[3];
    const float len = sqrtf((rr[0] * rr[0] + rr[1] * rr[1]) + rr[2] * rr[2]);
    const float rd[3] = {rr[0] / len, rr[1] / len, rr[2] / len};
    // trace.frag:65 model_ray_dir = (inverse(M) * vec4(ray_dir, 0)).xyz ; :66 model_ray_pos
    float dir[3], pos[3];
    const float size[3] = {(float)(int32_t)Ip->w, (float)(int32_t)Ip->h, (float)(int32_t)Ip->d};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        dir[k] = ((Ip->Mi[0 * 3 + k] * rd[0] + Ip->Mi[1 * 3 + k] * rd[1]) + Ip->Mi[2 * 3 + k] * rd[2]) + Ip->Mi[3 * 3 + k] * 0.0f;
        pos[k] = (mp[k] + 0.5f) * size[k];
    }
    const int32_t none[3] = {0, 0, 0};
    march_instance<kSmem, kBricks>(Ip, mask_base, pos, dir, false, none, f.dda);
}

__device__ __forceinline__ uint32_t face_bits(const Dda& r, int entry_axis) {
    const uint32_t mask = r.steps ? r.last_mask : (1u << entry_axis);
    const uint32_t neg = (r.step[0] < 0 ? 1u : 0u) | (r.step[1] < 0 ? 2u : 0u) | (r.step[2] < 0 ? 4u : 0u);
    return mask | (neg << 3);
}

__device__ __forceinline__ uint32_t srgb_encode(const float* __restrict__ thr, float x) {
    uint32_t k = 0;
#pragma unroll
    for (uint32_t bit = 128; bit; bit >>= 1)
        if (x >= __ldg(thr + (k | bit))) k |= bit;
    return k;
}

// CTA prologue shared by both trace kernels: stage the decode LUT and (optionally) the whole
// mask arena into shared memory.  Returns the base pointer masks are addressed from.
template <bool kSmem>
__device__ __forceinline__ void stage_tables(const uint32_t* mask_arena, uint32_t arena_words, const float* __restrict__ decode) {
    uint64_t* bar = reinterpret_cast<uint64_t*>(vt_smem);
    float* lut = reinterpret_cast<float*>(vt_smem + kSmemLutOff);
    uint32_t* smask = reinterpret_cast<uint32_t*>(vt_smem + kSmemMaskOff);
    if (kSmem) {
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t bytes = arena_words * 4u;
            mbar_expect_tx(bar, bytes);
            // one TMA bulk copy per <= 64 KiB chunk
            for (uint32_t off = 0; off < bytes; off += 65536u) {
                const uint32_t n = bytes - off < 65536u ? bytes - off : 65536u;
                bulk_g2s(reinterpret_cast<unsigned char*>(smask) + off, reinterpret_cast<const unsigned char*>(mask_arena) + off, n, bar);
            }
        }
    }
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = __ldg(decode + i);
    if (kSmem) mbar_wait(bar, 0);
    __syncthreads();
}

// Persistent warps pull work from a global counter (reset by the host before every launch):
// tile cost varies by orders of magnitude (sky vs. volume), static striding left SMs idle.
__device__ __forceinline__ int claim_tiles(unsigned long long* counter, int lane, int chunk) {
    unsigned long long c = 0;
    if (lane == 0) c = atomicAdd(counter, 1ull);
    c = __shfl_sync(0xffffffffu, c, 0);
    return (int)c * chunk;
}

// -------------------------------------------------------------------------------------------
// trace_primary_kernel: persistent warps over 8x4 pixel tiles; one thread per pixel.

template <bool kSmem, bool kBricks, bool kShadow>
__global__ void __launch_bounds__(kBlockThreads) trace_primary_kernel(const __grid_constant__ FrameParams fp,
                                                                      const InstUniforms* __restrict__ inst, const BinTable bins,
                                                                      const uint32_t* __restrict__ mask_arena,
                                                                      uint32_t arena_words, SrgbTables lut, FrameBuffers fb) {
    stage_tables<kSmem>(mask_arena, arena_words, lut.decode);
    const uint32_t* mask_base = mask_arena;
    const float* dec = reinterpret_cast<const float*>(vt_smem + kSmemLutOff);

    const int tiles_x = (fp.width + kTileW - 1) / kTileW;
    const int tiles_y = (fp.height + kTileH - 1) / kTileH;
    const int n_tiles = tiles_x * tiles_y;
    const int lane = threadIdx.x & 31;
    const int lx = lane & 7, ly = lane >> 3;

    // clear values, lib/command.c:56-61, as stored by the sRGB target (encoded once on the host)
    const uint32_t clear_r = fp.clear_rgba & 255u, clear_g = (fp.clear_rgba >> 8) & 255u, clear_b = (fp.clear_rgba >> 16) & 255u;
    const uint32_t tiles_x_magic = 0xFFFFFFFFu / (uint32_t)tiles_x + 1u; // tile / tiles_x == umulhi(tile, magic) for these sizes

    unsigned long long iter_sum = 0, shadow_rays = 0;
    // Primary rays are cheap (tens of microseconds for the whole frame), so tiles are dealt
    // round-robin over all resident warps instead of through a scheduler atomic: neighbouring
    // tiles (similar cost) land on different SMs, which balances as well and costs nothing.
    const int n_warps = gridDim.x * (kBlockThreads / 32);
    {
        for (int tile = blockIdx.x * (kBlockThreads / 32) + (threadIdx.x >> 5); tile < n_tiles; tile += n_warps) {
            const int tile_y = (int)__umulhi((uint32_t)tile, tiles_x_magic), tile_x = tile - tile_y * tiles_x;
            const int px = tile_x * kTileW + lx;
            const int py = tile_y * kTileH + ly;
            if (px >= fp.width || py >= fp.height) continue;
            const float fx = (float)px + 0.5f, fy = (float)py + 0.5f;
            uint32_t dst[4] = {clear_r, clear_g, clear_b, 255u};
            float zbuf = 1.0f; // lib/command.c:60
            HitRecord rec{VT_MISS, 0u, VT_MISS, 0u};
            // draw order = instance order (lib/command.c:102); with bins, only the instances whose screen
            // rectangle touches this pixel's 16x16 bin are visited (same order, same results)
            uint32_t k_begin, k_end;
            const uint32_t* bin_list = bin_range(bins, px, py, fp.n_inst, k_begin, k_end);
            for (uint32_t k = k_begin; k < k_end; ++k) {
                const uint32_t i = bin_list ? __ldg(bin_list + k) : k;
                Fragment f;
                run_fragment<kSmem, kBricks>(fp, inst + i, mask_base, px, py, fx, fy, f);
                if (!f.covered) continue;
                rec.iters += f.dda.steps;
                if (!f.dda.hit) continue;        // discard, trace.frag:89
                if (!(f.depth < zbuf)) continue; // VK_COMPARE_OP_LESS, lib/pipeline.c:148-150
                zbuf = f.depth;
                const InstUniforms* Ip = inst + i;
                const uchar4 s = instance_texel<kBricks>(Ip, f.dda.v);
                // extension: one shadow ray towards the sun, inside the fragment's own volume
                float shade = 1.0f;
                uint32_t shadow_bits = 0;
                if (kShadow) { // compiled in only for VT_FLAG_SHADOW_RAYS frames: the plain pass keeps its 64 registers
                    const float size[3] = {(float)(int32_t)Ip->w, (float)(int32_t)Ip->h, (float)(int32_t)Ip->d};
                    int ax, nsign;
                    float p0[3];
                    int32_t sv[3];
                    leave_hit(f.dda, f.entry_axis, size, ax, nsign, p0, sv);
                    const float sun_m[3] = {Ip->sun_m[0], Ip->sun_m[1], Ip->sun_m[2]};
                    bool lit = false;
                    if ((float)nsign * (ax == 0 ? sun_m[0] : (ax == 1 ? sun_m[1] : sun_m[2])) > 0.0f) { // the face looks at the sun
                        Dda sh;
                        march_instance<kSmem, kBricks>(Ip, mask_base, p0, sun_m, true, sv, sh);
                        rec.iters += sh.steps;
                        lit = !sh.hit;
                        shadow_bits = 1u | (lit ? 2u : 0u);
                        shadow_rays += 1;
                    }
                    shade = lit ? 1.0f : 0.35f;
                }
                if (s.w == 255 && shade == 1.0f) {
                    // a == 1: src*1 + dst*0 == src exactly and encode(decode(c)) == c by construction
                    dst[0] = s.x; dst[1] = s.y; dst[2] = s.z; dst[3] = 255u;
                } else {
                    // blend, lib/pipeline.c:129-137
                    const float a = (float)s.w / 255.0f;
                    const uint32_t sc[3] = {s.x, s.y, s.z};
#pragma unroll
                    for (int c = 0; c < 3; ++c) dst[c] = srgb_encode(lut.threshold, (dec[sc[c]] * shade) * a + dec[dst[c]] * (1.0f - a));
                    dst[3] = (uint32_t)__float2int_rz(floorf(a * 255.0f + 0.5f));
                }
                rec.hit_voxel = (uint32_t)f.dda.v[0] + Ip->w * ((uint32_t)f.dda.v[1] + Ip->h * (uint32_t)f.dda.v[2]);
                rec.packed = (f.dda.steps & 0xFFFFu) | (face_bits(f.dda, f.entry_axis) << 16) | (shadow_bits << 22);
                rec.instance = i;
            }
            const size_t p = (size_t)py * (size_t)fp.width + (size_t)px;
            if (fb.records) *reinterpret_cast<uint4*>(fb.records + p) = make_uint4(rec.hit_voxel, rec.packed, rec.instance, rec.iters);
            fb.color[p] = make_uchar4((unsigned char)dst[0], (unsigned char)dst[1], (unsigned char)dst[2], (unsigned char)dst[3]);
            if (fb.depth) fb.depth[p] = zbuf;
            iter_sum += rec.iters;
        }
    }
    // one atomic per warp
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        iter_sum += __shfl_xor_sync(0xffffffffu, iter_sum, o);
        shadow_rays += __shfl_xor_sync(0xffffffffu, shadow_rays, o);
    }
    if (lane == 0 && iter_sum) atomicAdd(fb.stats + 1, iter_sum);
    if (lane == 0 && shadow_rays) atomicAdd(fb.stats + 0, shadow_rays);
}


// -------------------------------------------------------------------------------------------
// path-tracing extension

__device__ __forceinline__ uint32_t vt_mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
struct Rng { uint32_t key, ctr; };
// counter-based stream keyed by (seed, pixel, sample): draw k is mix(key + k * phi)
__device__ __forceinline__ void rng_init(Rng& r, uint32_t seed, uint32_t pixel, uint32_t sample) {
    const uint32_t k = vt_mix(seed + pixel * 0x9E3779B9u);
    r.key = vt_mix(k ^ (sample * 0x85EBCA6Bu + 0xC2B2AE35u));
    r.ctr = 0;
}
// uniform in [0,1) with 23 random mantissa bits: as_float(0x3f800000 | bits) - 1 (no int->float conversion)
__device__ __forceinline__ float rng_u01(Rng& r) {
    const uint32_t x = vt_mix(r.key + (r.ctr++) * 0x9E3779B9u);
    return __uint_as_float(0x3f800000u | (x >> 9)) - 1.0f;
}
__device__ __forceinline__ void rng_sphere(Rng& r, float s[3]) {
    float a = 0.0f, b = 0.0f, q = 0.0f;
    bool ok = false;
    for (int attempt = 0; attempt < 16 && !ok; ++attempt) {
        a = rng_u01(r) * 2.0f - 1.0f;
        b = rng_u01(r) * 2.0f - 1.0f;
        q = a * a + b * b;
        ok = q < 1.0f;
    }
    if (!ok) { a = 0.0f; b = 0.0f; q = 0.0f; }
    const float w = sqrtf(1.0f - q);
    s[0] = (2.0f * a) * w;
    s[1] = (2.0f * b) * w;
    s[2] = 1.0f - 2.0f * q;
}

// -------------------------------------------------------------------------------------------
// trace_rays_kernel (extension): incoherent rays through instance 0's volume,
// in its voxel space.  Ray i: origin uniform in the volume, direction uniform on the sphere, from the
// RNG stream keyed (seed, i).  One thread per ray, grid-stride; masks of dense volumes stay in global
// memory here (the mode exists for volumes far larger than shared memory).
// Record: hit_voxel = x | y << 16, instance = z, packed = steps | face bits << 16, iters = steps.
#ifndef VT_RAYS_MIN_BLOCKS
#define VT_RAYS_MIN_BLOCKS 5 // 48 registers (a few spills) for 5 CTAs per SM: the walk waits on directory / pool loads (configs[4]: 36.3 -> 34.0 ms)
#endif
__global__ void __launch_bounds__(kBlockThreads, VT_RAYS_MIN_BLOCKS) trace_rays_kernel(const __grid_constant__ FrameParams fp,
                                                                   const InstUniforms* __restrict__ inst,
                                                                   const uint32_t* __restrict__ mask_arena, unsigned long long n,
                                                                   unsigned long long first, FrameBuffers fb) {
    const InstUniforms* Ip = inst;
    const float size[3] = {(float)(int32_t)Ip->w, (float)(int32_t)Ip->h, (float)(int32_t)Ip->d};
    const int lane = threadIdx.x & 31;
    unsigned long long iter_sum = 0;

    auto make_ray = [&](unsigned long long k, float pos[3], float dir[3]) {
        const unsigned long long i = first + k;
        Rng rng;
        rng_init(rng, fp.seed, (uint32_t)i, (uint32_t)(i >> 32));
#pragma unroll
        for (int c = 0; c < 3; ++c) pos[c] = rng_u01(rng) * size[c];
        rng_sphere(rng, dir);
    };
    auto write_ray = [&](unsigned long long k, const Dda& r) {
        uint4 rec = make_uint4(VT_MISS, 0u, VT_MISS, r.steps);
        uchar4 px = make_uchar4(0, 0, 0, 0);
        if (r.hit) {
            const uint32_t neg = (r.step[0] < 0 ? 1u : 0u) | (r.step[1] < 0 ? 2u : 0u) | (r.step[2] < 0 ? 4u : 0u);
            rec.x = (uint32_t)r.v[0] | ((uint32_t)r.v[1] << 16);
            rec.y = (r.steps & 0xFFFFu) | ((r.last_mask | (neg << 3)) << 16);
            rec.z = (uint32_t)r.v[2];
            px = instance_texel<true>(Ip, r.v);
        }
        if (fb.records) *reinterpret_cast<uint4*>(fb.records + k) = rec;
        fb.color[k] = px;
    };

    if (!Ip->valid || !Ip->bricks) {
        // dense volumes (or no volume): one ray per thread, grid-stride
        const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
        for (unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
            float pos[3], dir[3];
            make_ray(k, pos, dir);
            Dda r;
            const int32_t none[3] = {0, 0, 0};
            if (Ip->valid) march_instance<false, false>(Ip, mask_arena, pos, dir, false, none, r);
            else { r.hit = false; r.steps = 0; r.last_mask = 0; r.step[0] = r.step[1] = r.step[2] = 0; r.v[0] = r.v[1] = r.v[2] = 0; }
            iter_sum += r.steps;
            write_ray(k, r);
        }
    } else {
        // Brick volumes: rays differ in length by three orders of magnitude, so lanes are persistent
        // workers — a warp claims 2048 consecutive ray ids at a time; a lane whose ray ended writes its
        // record and starts the next id as soon as a quarter of the warp is idle.
        const BrickVolume bv = *Ip->bricks;
        const uint32_t W = Ip->w, H = Ip->h, D = Ip->d;
        constexpr unsigned long long kChunk = 2048;
        constexpr int kRayBurst = VT_RAY_BURST; // DDA iterations between brick lookups (see brick_walk_burst)
        const unsigned long long n_chunks = (n + kChunk - 1) / kChunk;
        unsigned long long chunk_next = 0, chunk_end = 0; // the warp's current range of ray ids
        bool more = true;
        bool active = false;
        int status = 0; // of the lane's walk: 0 = walking, 1 = hit, 2 = left the volume
        unsigned long long my_ray = 0;
        Dda r;
        BrickWalk k;
        r.hit = false; r.steps = 0; r.last_mask = 0; r.len = 1.0f;
#pragma unroll
        for (int c = 0; c < 3; ++c) { r.v[c] = 0; r.step[c] = 0; r.side[c] = r.delta[c] = r.dir[c] = r.pos[c] = 0.0f; }
        k.loc = k.ploc = kLocInside; k.cell = 0; k.ix = k.iy = k.iz = 0; k.sx = k.sy = k.sz = 0.0f; k.steps = k.last = 0; k.stop = kLocInside; k.slot = kSlotEmpty;
        for (;;) {
            // ---- refill idle lanes ----
            uint32_t idle = __ballot_sync(0xffffffffu, !active);
            while (idle && more) {
                if (chunk_next >= chunk_end) {
                    unsigned long long c = 0;
                    if (lane == 0) c = atomicAdd(fb.stats + 2, 1ull);
                    c = __shfl_sync(0xffffffffu, c, 0);
                    if (c >= n_chunks) { more = false; break; }
                    chunk_next = c * kChunk;
                    chunk_end = chunk_next + kChunk < n ? chunk_next + kChunk : n;
                }
                const unsigned long long avail = chunk_end - chunk_next;
                const uint32_t rank = __popc(idle & ((1u << lane) - 1u));
                if (!active && rank < avail) {
                    my_ray = chunk_next + rank;
                    float pos[3], dir[3];
                    make_ray(my_ray, pos, dir);
                    const int32_t none[3] = {0, 0, 0};
                    status = brick_walk_begin(bv, W, H, D, pos, dir, false, none, r, k);
                    active = true;
                }
                const unsigned long long want = __popc(idle);
                chunk_next += want < avail ? want : avail;
                idle = __ballot_sync(0xffffffffu, !active);
            }
            const uint32_t act = __ballot_sync(0xffffffffu, active);
            if (!act) break;
            // ---- walk until a quarter of the warp has ended (or everything, when no rays are left) ----
            const int stop_at = more ? 8 : 32;
            for (;;) {
#pragma unroll 1
                for (int u = 0; u < 2; ++u) {
                    if (status == 0 && active) status = brick_walk_burst<kRayBurst>(bv, r, k);
                    __syncwarp();
                }
                const int n_idle = __popc(__ballot_sync(0xffffffffu, !active || status != 0));
                if (n_idle >= stop_at || n_idle == 32) break;
            }
            // ---- ended rays write their record ----
            if (active && status != 0) {
                brick_walk_finish(bv, k, status == 1, r);
                iter_sum += r.steps;
                write_ray(my_ray, r);
                active = false;
            }
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) iter_sum += __shfl_xor_sync(0xffffffffu, iter_sum, o);
    if (lane == 0 && iter_sum) atomicAdd(fb.stats + 1, iter_sum);
}

cudaError_t launch_trace_rays(const FrameParams& fp, const InstUniforms* inst, const uint32_t* mask_arena, unsigned long long n,
                              unsigned long long first, FrameBuffers fb, int sm_count, cudaStream_t stream) {
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, trace_rays_kernel, kBlockThreads, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    unsigned long long blocks = (n + kBlockThreads - 1) / kBlockThreads;
    const unsigned long long cap = (unsigned long long)per_sm * sm_count; // persistent: one resident wave
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    trace_rays_kernel<<<(unsigned)blocks, kBlockThreads, 0, stream>>>(fp, inst, mask_arena, n, first, fb);
    return cudaGetLastError();
}

#include "world_grid.cuh"

struct PathHit {
    bool hit;
    uint32_t instance;
    int entry_axis;
    Dda dda;
};

// Nearest instance along a ray by box-entry parameter (ties: lower index); the first instance in
// that order whose DDA hits wins.  cam != nullptr: camera ray (o = eye in model space,
// d = dirm * (x_ndc, y_ndc, 1)); else world ray (o = Mi*(ow,1), d = Mi*(dw,0)).
template <bool kSmem, bool kBricks>
__device__ void trace_world(const FrameParams& fp, const InstUniforms* __restrict__ inst, const BinTable& bins, const WorldGridTable& wg,
                            const uint32_t* mask_base, uint32_t skip, const float* cam, int px, int py, const float ow[3],
                            const float dw[3], PathHit& out, unsigned long long& iters) {
    out.hit = false;
    float last_t = -INFINITY;
    uint32_t last_j = 0;
    bool have_last = false;
    // camera rays only need the instances binned to their pixel; world rays the ones registered in the grid
    // cells they pass (world_grid.cuh); without either, every instance is visited
    const bool binned = cam && bins.enabled;
    const bool gridded = !cam && wg.enabled && __ldg(&wg.hdr->overflow) == 0u;
    WorldWalk walk;
    if (gridded) {
        const WorldGrid g = *wg.hdr;
        world_walk_begin(g, ow, dw, walk);
    }
    // The ray's direction as an instance with inverse(M) = identity would see it, and its reciprocal: instances whose
    // inverse(M) is a diagonal of powers of two (InstUniforms::lin) scale both exactly, so they cost no division.
    float d0[3], inv0[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        d0[k] = cam ? (fp.RD[0 * 4 + k] * cam[0] + fp.RD[1 * 4 + k] * cam[1]) + fp.RD[3 * 4 + k] : dw[k];
        inv0[k] = 1.0f / d0[k];
    }
    // the ray in the space of instance j, and the box test of trace.frag's proxy cube
    auto candidate = [&](uint32_t j, float o[3], float d[3], float& tn, int& axis) -> bool {
        const InstUniforms* J = inst + j;
        float inv3[3];
        const bool lin = J->lin[0] != 0.0f;
        if (cam) {
#pragma unroll
            for (int k = 0; k < 3; ++k) o[k] = J->eye_m[k];
            if (!lin) {
#pragma unroll
                for (int k = 0; k < 3; ++k) d[k] = (J->dirm[0 * 3 + k] * cam[0] + J->dirm[1 * 3 + k] * cam[1]) + J->dirm[3 * 3 + k];
            }
        } else {
            if (lin) {
                // inverse(M)'s off-diagonal entries are +-0: the general sum below adds +-0 to lin[k] * ow[k], which changes
                // nothing but, possibly, the sign of an exact zero — and o[k] = +-0 gives the same slab bounds, the same
                // comparisons and the same entry point (its sign can only survive into mp[k] + 0.5)
#pragma unroll
                for (int k = 0; k < 3; ++k) o[k] = J->lin[k] * ow[k] + J->Mi[3 * 3 + k];
            } else {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    o[k] = ((J->Mi[0 * 3 + k] * ow[0] + J->Mi[1 * 3 + k] * ow[1]) + J->Mi[2 * 3 + k] * ow[2]) + J->Mi[3 * 3 + k];
                    d[k] = (J->Mi[0 * 3 + k] * dw[0] + J->Mi[1 * 3 + k] * dw[1]) + J->Mi[2 * 3 + k] * dw[2];
                }
            }
        }
        const float lo3[3] = {-0.5f - o[0], -0.5f - o[1], -0.5f - o[2]};
        const float hi3[3] = {0.5f - o[0], 0.5f - o[1], 0.5f - o[2]};
        if (lin) {
#pragma unroll
            for (int k = 0; k < 3; ++k) { d[k] = J->lin[k] * d0[k]; inv3[k] = J->ilin[k] * inv0[k]; }
            return slab_unit_cube_inv(o, lo3, hi3, d, inv3, tn, axis);
        }
        return slab_unit_cube(o, lo3, hi3, d, tn, axis);
    };
    for (;;) {
        bool found = false;
        float best_t = 0.0f;
        uint32_t best_j = 0;
        int best_axis = 0;
        float bo[3] = {0, 0, 0}, bd[3] = {0, 0, 0};
        uint32_t k_begin = 0, k_end = fp.n_inst;
        float t_lim = INFINITY; // grid: candidates entered beyond the current cell wait for their own cell
        bool last_cell = true;
        const uint32_t* list = nullptr;
        if (binned) {
            list = bin_range(bins, px, py, fp.n_inst, k_begin, k_end);
        } else if (gridded) {
            const uint32_t cell = world_walk_cell(walk, t_lim, last_cell);
            k_begin = __ldg(wg.offset + cell);
            k_end = k_begin + __ldg(wg.count + cell);
            list = wg.list;
        }
        for (uint32_t k = k_begin; k < k_end; ++k) {
            const uint32_t j = list ? __ldg(list + k) : k;
            const InstUniforms* J = inst + j;
            if (j == skip || !J->valid) continue;
            // camera rays: the conservative screen rectangle rejects most instances without arithmetic
            if (cam && (px < J->bounds[0] || px > J->bounds[1] || py < J->bounds[2] || py > J->bounds[3])) continue;
            float o[3], d[3], tn;
            int axis;
            if (!candidate(j, o, d, tn, axis)) continue;
            if (have_last && !(tn > last_t || (tn == last_t && j > last_j))) continue;
            if (tn > t_lim) continue;
            if (!found || tn < best_t || (tn == best_t && j < best_j)) { // (grid lists are unordered)
                found = true; best_t = tn; best_j = j; best_axis = axis;
#pragma unroll
                for (int k = 0; k < 3; ++k) { bo[k] = o[k]; bd[k] = d[k]; }
            }
        }
        if (!found) {
            if (last_cell) return;
            world_walk_next(walk); // nothing (more) entered inside this cell: on to the next one
            continue;
        }
        float mp[3], pos[3];
        entry_point(bo, bd, best_t, best_axis, mp);
        const InstUniforms* J = inst + best_j;
        const float size[3] = {(float)(int32_t)J->w, (float)(int32_t)J->h, (float)(int32_t)J->d};
#pragma unroll
        for (int k = 0; k < 3; ++k) pos[k] = (mp[k] + 0.5f) * size[k];
        const int32_t none[3] = {0, 0, 0};
        march_instance<kSmem, kBricks>(J, mask_base, pos, bd, false, none, out.dda);
        iters += out.dda.steps;
        if (out.dda.hit) {
            out.hit = true; out.instance = best_j; out.entry_axis = best_axis;
            return;
        }
        last_t = best_t; last_j = best_j; have_last = true;
    }
}

template <bool kSmem, bool kBricks>
__device__ void trace_path(const FrameParams& fp, const InstUniforms* __restrict__ inst, const BinTable& bins, const WorldGridTable& wg,
                           const uint32_t* mask_base, const float* __restrict__ dec, int px, int py, uint32_t sample, float L[3],
                           unsigned long long& rays, unsigned long long& iters) {
    Rng rng;
    rng_init(rng, fp.seed, (uint32_t)py * (uint32_t)fp.width + (uint32_t)px, sample);
    const float jx = rng_u01(rng), jy = rng_u01(rng);
    const float fx = (float)px + jx, fy = (float)py + jy;
    // camera segment: same visibility rule as every later segment
    PathHit cur;
    const float cam[2] = {fx * fp.sxn - 1.0f, fy * fp.syn - 1.0f};
    const float zero3[3] = {0.0f, 0.0f, 0.0f};
    trace_world<kSmem, kBricks>(fp, inst, bins, wg, mask_base, 0xFFFFFFFFu, cam, px, py, zero3, zero3, cur, iters);
    rays += 1;
    const float sky[3] = {53.0f / 100.0f, 81.0f / 100.0f, 92.0f / 100.0f}; // lib/command.c:57-59
    float thr[3] = {1.0f, 1.0f, 1.0f};
    L[0] = L[1] = L[2] = 0.0f;
    for (uint32_t b = 0;; ++b) {
        if (!cur.hit) {
#pragma unroll
            for (int c = 0; c < 3; ++c) L[c] = thr[c] * sky[c];
            return;
        }
        const InstUniforms* J = inst + cur.instance;
        const Dda& r = cur.dda;
        const uchar4 s = instance_texel<kBricks>(J, r.v);
        thr[0] = thr[0] * dec[s.x];
        thr[1] = thr[1] * dec[s.y];
        thr[2] = thr[2] * dec[s.z];
        if (b == fp.bounces) return;
        const float size[3] = {(float)(int32_t)J->w, (float)(int32_t)J->h, (float)(int32_t)J->d};
        const uint32_t lm = r.steps ? r.last_mask : (1u << cur.entry_axis);
        const int a = (lm & 1u) ? 0 : ((lm & 2u) ? 1 : 2);
        float p0[3], dn[3];
        int32_t sv[3];
        rng_sphere(rng, dn);
        int nsign = 0;
        // the per-axis quantities of the hit face are selected without dynamic indexing
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
            p0[k] = p;
            sv[k] = r.v[k];
            if (k == a) {
                nsign = r.step[k] != 0 ? -r.step[k] : (r.pos[k] <= 0.5f * size[k] ? -1 : 1);
                p0[k] = (float)(r.v[k] + (nsign > 0 ? 1 : 0));
                sv[k] += nsign;
                dn[k] += (float)nsign;
            }
        }
        const float l2 = (dn[0] * dn[0] + dn[1] * dn[1]) + dn[2] * dn[2];
        if (l2 < 1e-6f) {
#pragma unroll
            for (int k = 0; k < 3; ++k) dn[k] = (k == a) ? (float)nsign : 0.0f;
        } else {
            const float rl = 1.0f / sqrtf(l2);
            dn[0] *= rl; dn[1] *= rl; dn[2] *= rl;
        }
        rays += 1;
        PathHit next;
        next.hit = false;
        next.instance = cur.instance;
        next.entry_axis = a;
        march_instance<kSmem, kBricks>(J, mask_base, p0, dn, true, sv, next.dda);
        iters += next.dda.steps;
        next.hit = next.dda.hit;
        if (!next.hit && fp.n_inst > 1) {
            float pm[3], dm[3], ow[3], dw[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) { pm[k] = p0[k] / size[k] - 0.5f; dm[k] = dn[k] / size[k]; }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                ow[k] = ((J->M[0 * 3 + k] * pm[0] + J->M[1 * 3 + k] * pm[1]) + J->M[2 * 3 + k] * pm[2]) + J->M[3 * 3 + k];
                dw[k] = (J->M[0 * 3 + k] * dm[0] + J->M[1 * 3 + k] * dm[1]) + J->M[2 * 3 + k] * dm[2];
            }
            trace_world<kSmem, kBricks>(fp, inst, bins, wg, mask_base, cur.instance, nullptr, 0, 0, ow, dw, next, iters);
        }
        cur = next;
    }
}

// CTAs per SM the general path kernel is compiled for: brick scenes gain from a third CTA (heightmap 1024^3, 8 spp:
// 8.39 -> 7.19 ms) although it costs spills; dense scenes lose (entity grid 4.76 -> 5.47 ms)
template <bool kSmem, bool kBricks>
__global__ void __launch_bounds__(kBlockThreads, kBricks ? 3 : 2) trace_paths_kernel(const __grid_constant__ FrameParams fp,
                                                                    const InstUniforms* __restrict__ inst, const BinTable bins,
                                                                    const WorldGridTable wg, const uint32_t* __restrict__ mask_arena,
                                                                    uint32_t arena_words, SrgbTables lut, FrameBuffers fb) {
    stage_tables<kSmem>(mask_arena, arena_words, lut.decode);
    const uint32_t* mask_base = mask_arena;
    const float* dec = reinterpret_cast<const float*>(vt_smem + kSmemLutOff);

    const int tiles_x = (fp.width + kTileW - 1) / kTileW;
    const int tiles_y = (fp.height + kTileH - 1) / kTileH;
    const int n_tiles = tiles_x * tiles_y;
    const int lane = threadIdx.x & 31;
    const int lx = lane & 7, ly = lane >> 3;

    // radiance of a path that sees only sky: thr (1,1,1) * clear colour, in 2^-24 fixed point
    unsigned long long sky_q[3];
    {
        const float sky[3] = {53.0f / 100.0f, 81.0f / 100.0f, 92.0f / 100.0f};
#pragma unroll
        for (int c = 0; c < 3; ++c) sky_q[c] = __float2ull_rz((1.0f * sky[c]) * 16777216.0f);
    }

    unsigned long long rays = 0, iters = 0, analytic = 0;
    for (;;) {
        const int tile = claim_tiles(fb.stats + 2, lane, 1);
        if (tile >= n_tiles) break;
        const int px = (tile % tiles_x) * kTileW + lx;
        const int py = (tile / tiles_x) * kTileH + ly;
        if (px >= fp.width || py >= fp.height) continue;
        // Pixels outside every instance's screen rectangle: all spp paths leave through the sky after
        // the primary segment; their sum is known without tracing them (they still count as rays).
        bool may_hit = false;
        {
            uint32_t k_begin, k_end;
            const uint32_t* bin_list = bin_range(bins, px, py, fp.n_inst, k_begin, k_end);
            for (uint32_t k = k_begin; k < k_end; ++k) {
                const InstUniforms* Ip = inst + (bin_list ? __ldg(bin_list + k) : k);
                may_hit = may_hit || !(px < Ip->bounds[0] || px > Ip->bounds[1] || py < Ip->bounds[2] || py > Ip->bounds[3]);
            }
        }
        unsigned long long acc[3] = {0, 0, 0};
        if (!may_hit) {
#pragma unroll
            for (int c = 0; c < 3; ++c) acc[c] = sky_q[c] * fp.spp;
            rays += fp.spp;
            analytic += fp.spp;
        } else {
            for (uint32_t k = 0; k < fp.spp; ++k) {
                float L[3];
                trace_path<kSmem, kBricks>(fp, inst, bins, wg, mask_base, dec, px, py, fp.sample_first + k * fp.sample_stride, L, rays, iters);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float q = L[c] * 16777216.0f;
                    acc[c] += (q == q && q > 0.0f) ? __float2ull_rz(q) : 0ull;
                }
            }
        }
        const size_t p = (size_t)py * (size_t)fp.width + (size_t)px;
#pragma unroll
        for (int c = 0; c < 3; ++c) fb.accum[3 * p + c] += acc[c];
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        rays += __shfl_xor_sync(0xffffffffu, rays, o);
        iters += __shfl_xor_sync(0xffffffffu, iters, o);
        analytic += __shfl_xor_sync(0xffffffffu, analytic, o);
    }
    if (lane == 0) {
        if (rays) atomicAdd(fb.stats + 0, rays);
        if (iters) atomicAdd(fb.stats + 1, iters);
        if (analytic) atomicAdd(fb.stats + 3, analytic);
    }
}

#include "paths_wave.cuh"

__global__ void resolve_kernel(const unsigned long long* __restrict__ accum, uint32_t n_pixels, uint32_t total_spp, SrgbTables lut,
                               uchar4* __restrict__ color) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pixels) return;
    const float scale = 1.0f / ((float)total_spp * 16777216.0f);
    uint32_t c[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) c[k] = srgb_encode(lut.threshold, __ull2float_rn(accum[3 * (size_t)p + k]) * scale);
    color[p] = make_uchar4((unsigned char)c[0], (unsigned char)c[1], (unsigned char)c[2], 255);
}

// -------------------------------------------------------------------------------------------
// Fused multi-GPU reduction (single-instance scenes).  Every rank accumulates its own samples locally;
// only the tiles that touch the instance's screen rectangle can differ from "spp x sky".  push_partial
// streams exactly those pixels — 32-byte vector stores, coalesced — into this rank's slot of a buffer
// that lives in the ROOT GPU's memory (CUDA IPC mapping, NVLink), clearing the local sums behind it.
// After a stream barrier the root's resolve_partials adds the slots up (integers: order-free, bit-exact)
// and encodes the frame; pixels outside the rectangle are resolved analytically.
// pixels for which paths are traced at all: inside the instance's conservative screen rectangle
__device__ __forceinline__ bool covered_region(const InstUniforms* __restrict__ Ip, int px, int py) {
    return !(px < Ip->bounds[0] || px > Ip->bounds[1] || py < Ip->bounds[2] || py > Ip->bounds[3]);
}

// -------------------------------------------------------------------------------------------
// Single-instance frames through render_tick: only the instance's screen rectangle is ever accumulated.
// The pixels outside it see nothing but sky for every sample, so the frame neither clears, nor adds to,
// nor reads their accumulators: clear_rect zeroes the rectangle before the trace, resolve_rect encodes
// it afterwards and writes the (constant) sky colour everywhere else, and fill_sky materialises
// spp x sky in the outside accumulators only when somebody asks for them (vt_read_accum).
__global__ void __launch_bounds__(256) clear_rect_kernel(const InstUniforms* __restrict__ inst, unsigned long long* __restrict__ accum,
                                                         uint32_t width, uint32_t height) {
    const int x0 = max(inst->bounds[0], 0), x1 = min(inst->bounds[1], (int)width - 1);
    const int y0 = max(inst->bounds[2], 0), y1 = min(inst->bounds[3], (int)height - 1);
    for (int py = y0 + (int)blockIdx.x; py <= y1; py += (int)gridDim.x)
        for (int px = x0 + (int)threadIdx.x; px <= x1; px += (int)blockDim.x) {
            const size_t p = (size_t)py * width + (size_t)px;
            accum[3 * p + 0] = 0ull; accum[3 * p + 1] = 0ull; accum[3 * p + 2] = 0ull;
        }
}

// sky_only != 0: touch nothing but the accumulators outside the rectangle (fill_sky)
__global__ void __launch_bounds__(256) resolve_rect_kernel(const InstUniforms* __restrict__ inst, unsigned long long* __restrict__ accum,
                                                           uint32_t width, uint32_t height, uint32_t spp, uint32_t total_spp, SrgbTables lut,
                                                           uchar4* __restrict__ color, uint32_t sky_only,
                                                           const unsigned long long* __restrict__ stats, unsigned long long* host_stats) {
    if (host_stats && blockIdx.x == 0 && threadIdx.x == 0) { // the frame's counters go to (mapped, page-locked) host memory from here:
        host_stats[0] = stats[0]; host_stats[1] = stats[1];  // no copy-engine operation on the frame's critical path
        host_stats[2] = stats[2]; host_stats[3] = stats[3];
        __threadfence_system();
    }
    const float scale = 1.0f / ((float)total_spp * 16777216.0f);
    const float sky[3] = {53.0f / 100.0f, 81.0f / 100.0f, 92.0f / 100.0f};
    unsigned long long sky_sum[3];
    uint32_t sky_c[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        sky_sum[c] = __float2ull_rz((1.0f * sky[c]) * 16777216.0f) * spp;
        sky_c[c] = srgb_encode(lut.threshold, __ull2float_rn(sky_sum[c]) * scale);
    }
    const uchar4 sky_px = make_uchar4((unsigned char)sky_c[0], (unsigned char)sky_c[1], (unsigned char)sky_c[2], 255);
    for (uint32_t py = blockIdx.x; py < height; py += gridDim.x)
        for (uint32_t px = threadIdx.x; px < width; px += blockDim.x) {
            const size_t p = (size_t)py * width + px;
            if (!covered_region(inst, (int)px, (int)py)) {
                if (sky_only) { accum[3 * p + 0] = sky_sum[0]; accum[3 * p + 1] = sky_sum[1]; accum[3 * p + 2] = sky_sum[2]; }
                else color[p] = sky_px;
                continue;
            }
            if (sky_only) continue;
            uint32_t c[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) c[k] = srgb_encode(lut.threshold, __ull2float_rn(accum[3 * p + k]) * scale);
            color[p] = make_uchar4((unsigned char)c[0], (unsigned char)c[1], (unsigned char)c[2], 255);
        }
}

cudaError_t launch_clear_rect(const InstUniforms* inst, unsigned long long* accum, uint32_t width, uint32_t height, int sm_count,
                              cudaStream_t stream) {
    const int grid = sm_count * 4 < (int)height ? sm_count * 4 : (int)height;
    clear_rect_kernel<<<grid, 256, 0, stream>>>(inst, accum, width, height);
    return cudaGetLastError();
}

cudaError_t launch_resolve_rect(const InstUniforms* inst, unsigned long long* accum, uint32_t width, uint32_t height, uint32_t spp,
                                uint32_t total_spp, SrgbTables lut, uchar4* color, bool sky_only, const unsigned long long* stats,
                                unsigned long long* host_stats, int sm_count, cudaStream_t stream) {
    const int grid = sm_count * 4 < (int)height ? sm_count * 4 : (int)height;
    resolve_rect_kernel<<<grid, 256, 0, stream>>>(inst, accum, width, height, spp, total_spp, lut, color, sky_only ? 1u : 0u, stats, host_stats);
    return cudaGetLastError();
}


// Cross-GPU ordering of the fused accumulation without a collective: sequence-number flags in the root's
// memory.  A rank raises "my partial sums of frame s are in place" from the last block of its push kernel
// (every block fences its stores at system scope before it counts itself done); the root's summation kernel
// waits for every rank's flag before it reads the slots, and the root raises "frame s consumed" when it
// starts its next frame, so that the ranks may reuse that half of the double buffer.  Waits are bounded
// (~10 s): a rank that never arrives raises *err instead of hanging the GPU.
static constexpr unsigned long long kFlagWaitLimitNs = 10ull * 1000 * 1000 * 1000;

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// spins until *flag has reached `target` (sequence numbers compared modulo 2^32)
// (acquire / release at system scope: the flags live in another GPU's memory, the data they guard travels over NVLink)
__device__ __forceinline__ uint32_t flag_load_acquire(const uint32_t* flag) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    return v;
}
__device__ __forceinline__ void flag_store_release(uint32_t* flag, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(v) : "memory");
}
__device__ __forceinline__ void flag_wait(const uint32_t* flag, uint32_t target, uint32_t* err) {
    if ((int32_t)(flag_load_acquire(flag) - target) >= 0) return;
    const unsigned long long t0 = global_timer_ns();
    while ((int32_t)(flag_load_acquire(flag) - target) < 0) {
        __nanosleep(256);
        if (global_timer_ns() - t0 > kFlagWaitLimitNs) { atomicExch(err, 1u); break; }
    }
}

__device__ __forceinline__ void fused_sync_begin(const FusedSync& fs) {
    if (fs.host_stats && blockIdx.x == 0 && threadIdx.x == 0) {
        fs.host_stats[0] = fs.stats[0]; fs.host_stats[1] = fs.stats[1]; fs.host_stats[2] = fs.stats[2]; fs.host_stats[3] = fs.stats[3];
    }
    if (fs.wait_flags) {
        if (threadIdx.x < fs.wait_count) {
            flag_wait(fs.wait_flags + threadIdx.x, fs.wait_target, fs.err);
            __threadfence_system(); // the waiting threads' acquire; the barrier extends it to the block (fences are cumulative)
        }
        __syncthreads();
    }
}
__device__ __forceinline__ void fused_sync_end(const FusedSync& fs) {
    if (fs.signal_flag) {
        // The block's stores to the root's memory must be visible there before the flag can be: the barrier orders them before
        // thread 0's system-scope fence, which is cumulative (one fence per block instead of one per thread: a MEMBAR.SYS
        // waits for every store the SM has in flight).
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            if (atomicAdd(fs.done_counter, 1u) == gridDim.x - 1) { // last block
                *fs.done_counter = 0u;
                __threadfence_system();
                flag_store_release(fs.signal_flag, fs.signal_value);
            }
        }
    }
}

// kCompact: a rank's per-channel sums fit 32 bits (at most 255 samples of at most 2^24 each), so a pixel
// travels as ONE 16-byte store instead of two.  Blocks stride over the rows of the instance's screen
// rectangle only (the other pixels are sky for every rank and never travel).
template <bool kCompact>
__global__ void __launch_bounds__(256) push_partial_kernel(const InstUniforms* __restrict__ inst, unsigned long long* __restrict__ local_accum,
                                                           uint4* __restrict__ slot, uint32_t width, uint32_t height,
                                                           RowShare rows, FusedSync fs) {
    fused_sync_begin(fs);
    const int x0 = max(inst->bounds[0], 0), x1 = min(inst->bounds[1], (int)width - 1);
    const int y0 = max(inst->bounds[2], 0), y1 = min(inst->bounds[3], (int)height - 1);
    // the pixels of the rectangle's lines this rank owns, numbered consecutively: one grid-stride loop over all of them
    const int rect_w = x1 - x0 + 1;
    const int t0 = y0 / kTileH, t1 = y1 / kTileH;
    const int own_tile_rows = (y0 <= y1 && rect_w > 0) ? (int)rows.rows_in((uint32_t)t0, (uint32_t)t1) : 0;
    const long long total = (long long)own_tile_rows * kTileH * rect_w;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int line = (int)(i / rect_w), px = x0 + (int)(i - (long long)line * rect_w);
        const int py = (int)rows.row((uint32_t)t0, (uint32_t)(line / kTileH)) * kTileH + line % kTileH;
        if (py < y0 || py > y1) continue; // (enumerated rows may lie outside the rectangle; the first and last may be cut by it)
        const size_t p = (size_t)py * width + (size_t)px;
        const unsigned long long r = local_accum[3 * p + 0], g = local_accum[3 * p + 1], b = local_accum[3 * p + 2];
        local_accum[3 * p + 0] = 0ull; local_accum[3 * p + 1] = 0ull; local_accum[3 * p + 2] = 0ull;
        if (kCompact) {
            slot[p] = make_uint4((uint32_t)r, (uint32_t)g, (uint32_t)b, 0u);
        } else {
            slot[2 * p + 0] = make_uint4((uint32_t)r, (uint32_t)(r >> 32), (uint32_t)g, (uint32_t)(g >> 32));
            slot[2 * p + 1] = make_uint4((uint32_t)b, (uint32_t)(b >> 32), 0u, 0u);
        }
    }
    fused_sync_end(fs);
}

template <bool kCompact>
__global__ void __launch_bounds__(256) resolve_partials_kernel(const InstUniforms* __restrict__ inst, const uint4* __restrict__ partials,
                                                               uint32_t world, uint32_t width, uint32_t height, uint32_t total_spp,
                                                               SrgbTables lut, uchar4* __restrict__ color,
                                                               unsigned long long* __restrict__ accum_out, RowShare rows,
                                                               unsigned long long* __restrict__ root_local, uint4* __restrict__ root_slot,
                                                               FusedSync fs) {
    fused_sync_begin(fs);
    const size_t n_pix = (size_t)width * height;
    const float scale = 1.0f / ((float)total_spp * 16777216.0f);
    const float sky[3] = {53.0f / 100.0f, 81.0f / 100.0f, 92.0f / 100.0f};
    unsigned long long sky_sum[3];
    uint32_t sky_c[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { // a pixel outside the rectangle sees only sky, for every sample of every rank
        sky_sum[c] = __float2ull_rz((1.0f * sky[c]) * 16777216.0f) * total_spp;
        sky_c[c] = srgb_encode(lut.threshold, __ull2float_rn(sky_sum[c]) * scale);
    }
    const uchar4 sky_px = make_uchar4((unsigned char)sky_c[0], (unsigned char)sky_c[1], (unsigned char)sky_c[2], 255);
    for (uint32_t py = blockIdx.x; py < height; py += gridDim.x)
        for (uint32_t px = threadIdx.x; px < width; px += blockDim.x) {
            const size_t p = (size_t)py * width + px;
            if (!covered_region(inst, (int)px, (int)py)) {
                color[p] = sky_px;
                if (accum_out) { accum_out[3 * p + 0] = sky_sum[0]; accum_out[3 * p + 1] = sky_sum[1]; accum_out[3 * p + 2] = sky_sum[2]; }
                continue;
            }
            unsigned long long sum[3] = {0ull, 0ull, 0ull};
            // shared out by tile rows: the owner's slot holds the pixel's complete sums
            const uint32_t r_begin = rows.world > 1u ? rows.owner(py / (uint32_t)kTileH) : 0u;
            const uint32_t r_end = rows.world > 1u ? r_begin + 1u : world;
            for (uint32_t r = r_begin; r < r_end; ++r) { // (a slot is n_pix * 32 bytes whatever the layout; L1 is bypassed: peers wrote these lines)
                if (r == 0u && root_local) { // the root's own sums never travelled: take them from its accumulators
                    const unsigned long long lr = root_local[3 * p + 0], lg = root_local[3 * p + 1], lb = root_local[3 * p + 2];
                    root_local[3 * p + 0] = 0ull; root_local[3 * p + 1] = 0ull; root_local[3 * p + 2] = 0ull;
                    if (kCompact) {
                        root_slot[p] = make_uint4((uint32_t)lr, (uint32_t)lg, (uint32_t)lb, 0u);
                    } else {
                        root_slot[2 * p + 0] = make_uint4((uint32_t)lr, (uint32_t)(lr >> 32), (uint32_t)lg, (uint32_t)(lg >> 32));
                        root_slot[2 * p + 1] = make_uint4((uint32_t)lb, (uint32_t)(lb >> 32), 0u, 0u);
                    }
                    sum[0] += lr; sum[1] += lg; sum[2] += lb;
                    continue;
                }
                if (kCompact) {
                    const uint4 a = __ldcg(partials + r * n_pix * 2 + p);
                    sum[0] += a.x; sum[1] += a.y; sum[2] += a.z;
                } else {
                    const uint4 a = __ldcg(partials + (r * n_pix + p) * 2 + 0), b = __ldcg(partials + (r * n_pix + p) * 2 + 1);
                    sum[0] += (unsigned long long)a.x | ((unsigned long long)a.y << 32);
                    sum[1] += (unsigned long long)a.z | ((unsigned long long)a.w << 32);
                    sum[2] += (unsigned long long)b.x | ((unsigned long long)b.y << 32);
                }
            }
            uint32_t c[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) c[k] = srgb_encode(lut.threshold, __ull2float_rn(sum[k]) * scale);
            color[p] = make_uchar4((unsigned char)c[0], (unsigned char)c[1], (unsigned char)c[2], 255);
            if (accum_out) { accum_out[3 * p + 0] = sum[0]; accum_out[3 * p + 1] = sum[1]; accum_out[3 * p + 2] = sum[2]; }
        }
    fused_sync_end(fs);
}

cudaError_t launch_push_partial(const InstUniforms* inst, unsigned long long* local_accum, uint4* slot, uint32_t width, uint32_t height,
                                bool compact, RowShare rows, FusedSync fs, int sm_count, cudaStream_t stream) {
    const int grid = sm_count * 4; // grid-stride over the owned pixels
    if (compact) push_partial_kernel<true><<<grid, 256, 0, stream>>>(inst, local_accum, slot, width, height, rows, fs);
    else push_partial_kernel<false><<<grid, 256, 0, stream>>>(inst, local_accum, slot, width, height, rows, fs);
    return cudaGetLastError();
}

cudaError_t launch_resolve_partials(const InstUniforms* inst, const uint4* partials, uint32_t world, uint32_t width, uint32_t height,
                                    uint32_t total_spp, SrgbTables lut, uchar4* color, unsigned long long* accum_out, bool compact,
                                    RowShare rows, unsigned long long* root_local, uint4* root_slot, FusedSync fs, int sm_count,
                                    cudaStream_t stream) {
    const int grid = sm_count * 4 < (int)height ? sm_count * 4 : (int)height;
    if (compact) resolve_partials_kernel<true><<<grid, 256, 0, stream>>>(inst, partials, world, width, height, total_spp, lut, color, accum_out, rows, root_local, root_slot, fs);
    else resolve_partials_kernel<false><<<grid, 256, 0, stream>>>(inst, partials, world, width, height, total_spp, lut, color, accum_out, rows, root_local, root_slot, fs);
    return cudaGetLastError();
}

// -------------------------------------------------------------------------------------------
// launch wrappers


cudaError_t configure_kernels(int max_smem_optin) {
    cudaError_t e;
    e = cudaFuncSetAttribute(trace_primary_kernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(trace_primary_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(trace_primary_kernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(trace_primary_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(trace_paths_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(trace_paths_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(trace_paths_wave_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(trace_paths_wave_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    return e;
}

template <class K>
static int persistent_grid(K kernel, size_t smem, int sm_count, int n_tiles, int block_threads = kBlockThreads) {
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block_threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    long grid = (long)per_sm * sm_count; // whole number of CTAs per SM: one resident wave
    if (grid > n_tiles) grid = n_tiles;
    return (int)(grid < 1 ? 1 : grid);
}

cudaError_t launch_trace_primary(const FrameParams& fp, const InstUniforms* inst, BinTable bins, const uint32_t* mask_arena,
                                 uint32_t arena_words, bool masks_in_smem, SrgbTables lut, FrameBuffers fb, int sm_count,
                                 cudaStream_t stream) {
    const int n_tiles = ((fp.width + kTileW - 1) / kTileW) * ((fp.height + kTileH - 1) / kTileH);
    const size_t smem = trace_smem_bytes(arena_words, masks_in_smem);
    const int warps_needed = (n_tiles + 7) / 8;
    // variants: masks in shared memory or not x scenes with procedural brick volumes x shadow rays
    auto launch = [&](auto kernel) {
        const int grid = persistent_grid(kernel, smem, sm_count, warps_needed);
        kernel<<<grid, kBlockThreads, smem, stream>>>(fp, inst, bins, mask_arena, arena_words, lut, fb);
    };
    const bool shadow = (fp.flags & VT_FLAG_SHADOW_RAYS) != 0, bricks = fp.any_bricks != 0;
    if (masks_in_smem) {
        if (bricks) { if (shadow) launch(trace_primary_kernel<true, true, true>); else launch(trace_primary_kernel<true, true, false>); }
        else        { if (shadow) launch(trace_primary_kernel<true, false, true>); else launch(trace_primary_kernel<true, false, false>); }
    } else {
        if (bricks) { if (shadow) launch(trace_primary_kernel<false, true, true>); else launch(trace_primary_kernel<false, true, false>); }
        else        { if (shadow) launch(trace_primary_kernel<false, false, true>); else launch(trace_primary_kernel<false, false, false>); }
    }
    return cudaGetLastError();
}

// true when launch_trace_paths runs the single-instance wavefront kernel for this frame
bool paths_use_wave_kernel(const FrameParams& fp) {
    return !fp.any_bricks && fp.n_inst == 1 && !(fp.flags & VT_FLAG_PER_PIXEL_PATHS) && fp.max_idx_bits <= kWaveIdxBits;
}

cudaError_t launch_trace_paths(const FrameParams& fp, const InstUniforms* inst, BinTable bins, WorldGridTable wg, const uint32_t* mask_arena,
                               uint32_t arena_words, bool masks_in_smem, SrgbTables lut, FrameBuffers fb, int sm_count,
                               cudaStream_t stream) {
    const int n_tiles = ((fp.width + kTileW - 1) / kTileW) * ((fp.height + kTileH - 1) / kTileH);
    const size_t smem = trace_smem_bytes(arena_words, masks_in_smem);
    if (fp.any_bricks) { // scenes with brick volumes: the general kernel, marching through march_instance()
        if (masks_in_smem) {
            const int grid = persistent_grid(trace_paths_kernel<true, true>, smem, sm_count, n_tiles);
            trace_paths_kernel<true, true><<<grid, kBlockThreads, smem, stream>>>(fp, inst, bins, wg, mask_arena, arena_words, lut, fb);
        } else {
            const int grid = persistent_grid(trace_paths_kernel<false, true>, smem, sm_count, n_tiles);
            trace_paths_kernel<false, true><<<grid, kBlockThreads, smem, stream>>>(fp, inst, bins, wg, mask_arena, arena_words, lut, fb);
        }
        return cudaGetLastError();
    }
    if (paths_use_wave_kernel(fp)) {
        // single-instance scenes: warp-local wavefront engine (paths_wave.cuh)
        const int max_warps = 1 << 30; // persistent: one resident wave, work is claimed dynamically
        // the masks share the SM's shared memory with the warps' pools: a large arena is read through L1 instead
        int fits = 0;
        if (masks_in_smem &&
            (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fits, trace_paths_wave_kernel<true>, kWaveThreads,
                                                           wave_smem_bytes(arena_words, true)) != cudaSuccess || fits < 1)) {
            (void)cudaGetLastError();
            masks_in_smem = false;
        }
        const size_t wsmem = wave_smem_bytes(arena_words, masks_in_smem);
        if (masks_in_smem) {
            int grid = persistent_grid(trace_paths_wave_kernel<true>, wsmem, sm_count, max_warps, kWaveThreads);
            trace_paths_wave_kernel<true><<<grid, kWaveThreads, wsmem, stream>>>(fp, inst, mask_arena, arena_words, lut, fb);
        } else {
            int grid = persistent_grid(trace_paths_wave_kernel<false>, wsmem, sm_count, max_warps, kWaveThreads);
            trace_paths_wave_kernel<false><<<grid, kWaveThreads, wsmem, stream>>>(fp, inst, mask_arena, arena_words, lut, fb);
        }
        return cudaGetLastError();
    }
    if (masks_in_smem) {
        const int grid = persistent_grid(trace_paths_kernel<true, false>, smem, sm_count, n_tiles);
        trace_paths_kernel<true, false><<<grid, kBlockThreads, smem, stream>>>(fp, inst, bins, wg, mask_arena, arena_words, lut, fb);
    } else {
        const int grid = persistent_grid(trace_paths_kernel<false, false>, smem, sm_count, n_tiles);
        trace_paths_kernel<false, false><<<grid, kBlockThreads, smem, stream>>>(fp, inst, bins, wg, mask_arena, arena_words, lut, fb);
    }
    return cudaGetLastError();
}

#ifdef VT_WAVE_STATS
cudaError_t read_wave_stats(unsigned long long* out16) {
    cudaError_t e = cudaMemcpyFromSymbol(out16, vt_wave_stats, 32 * sizeof(unsigned long long));
    if (e != cudaSuccess) return e;
    unsigned long long zero[32] = {};
    return cudaMemcpyToSymbol(vt_wave_stats, zero, sizeof zero);
}
cudaError_t read_wave_times(unsigned int* out192) {
    cudaError_t e = cudaMemcpyFromSymbol(out192, vt_wave_times, 192 * sizeof(unsigned int));
    if (e != cudaSuccess) return e;
    unsigned int zero[192] = {};
    return cudaMemcpyToSymbol(vt_wave_times, zero, sizeof zero);
}
#endif

cudaError_t launch_resolve(const unsigned long long* accum, uint32_t n_pixels, uint32_t total_spp, SrgbTables lut, uchar4* color,
                           cudaStream_t stream) {
    const int threads = 256;
    resolve_kernel<<<(n_pixels + threads - 1) / threads, threads, 0, stream>>>(accum, n_pixels, total_spp, lut, color);
    return cudaGetLastError();
}

} // namespace vt

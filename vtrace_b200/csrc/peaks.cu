// peaks.cu — measured denominators for the cache-resident configurations:
// the .vox scenes are 256-500 KB and live in shared memory / L2, so "fraction of the HBM roofline" says
// little about them.  vt_measure_peak() times two plain streaming kernels on the device librender runs on:
//   kind 0: L2 read bandwidth  — 128-bit ld.global.cg (L1 bypassed) over a 32 MiB buffer that stays L2-resident,
//   kind 1: shared-memory read bandwidth — conflict-free LDS.128 from a 32 KiB tile per CTA,
//   kind 2: HBM read bandwidth — the same loads over a 2 GiB buffer (cross-check of MEASURED_PEAKS.json).
// bench.py reports the trace kernels' algorithmic GB/s against them next to the HBM figure.
#include "../../include/vtrace_abi.h"

#include <cstdint>
#include <cuda_runtime.h>

namespace {

__global__ void __launch_bounds__(256) stream_read_kernel(const uint4* __restrict__ src, size_t n_vec, int passes, uint32_t* sink) {
    uint32_t acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int p = 0; p < passes; ++p)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
            const uint4 v = __ldcg(src + i);
            acc ^= v.x ^ v.y ^ v.z ^ v.w;
        }
    if (acc == 0x12345679u) *sink = acc; // (never true for the zero-filled buffer; keeps the loads alive)
}

__global__ void __launch_bounds__(1024) lds_read_kernel(int iters, uint32_t* sink) {
    __shared__ uint4 tile[2048]; // 32 KiB
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) tile[i] = make_uint4(i, 0u, 0u, 0u);
    __syncthreads();
    uint32_t acc = 0, at = threadIdx.x;
#pragma unroll 8
    for (int i = 0; i < iters; ++i) {
        const uint4 v = tile[at & 2047u]; // consecutive lanes -> consecutive 16-byte words: conflict-free LDS.128
        acc ^= v.x ^ v.w;
        at += 1024u + (v.y & 1u); // (v.y is 0: the address stays data-dependent without changing)
    }
    if (acc == 0x12345679u) *sink = acc;
}

} // namespace

extern "C" int32_t vt_measure_peak(uint32_t kind, double* gb_per_s) {
    if (!gb_per_s || kind > 2) return -1;
    vt_config cfg;
    if (vt_get_config(&cfg) != 0) return -1; // entry() must have run: the measurement uses its device
    if (cudaSetDevice(cfg.device) != cudaSuccess) return -1;
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, cfg.device) != cudaSuccess) return -1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    uint32_t* sink = nullptr;
    cudaMalloc(&sink, 4);
    double best = 0.0;
    int32_t rc = 0;
    if (kind == 0 || kind == 2) {
        const size_t bytes = kind == 0 ? (size_t)32 << 20 : (size_t)2 << 30;
        const int passes = kind == 0 ? 64 : 1;
        uint4* buf = nullptr;
        if (cudaMalloc(&buf, bytes) != cudaSuccess) { (void)cudaGetLastError(); rc = -1; }
        else {
            cudaMemset(buf, 0, bytes);
            for (int rep = 0; rep < 5; ++rep) { // rep 0 warms the cache
                cudaEventRecord(e0);
                stream_read_kernel<<<prop.multiProcessorCount * 8, 256>>>(buf, bytes / 16, passes, sink);
                cudaEventRecord(e1);
                if (cudaEventSynchronize(e1) != cudaSuccess) { rc = -1; break; }
                float ms = 0.0f;
                cudaEventElapsedTime(&ms, e0, e1);
                const double gbs = (double)bytes * passes / (ms * 1e-3) / 1e9;
                if (rep > 0 && gbs > best) best = gbs;
            }
            cudaFree(buf);
        }
    } else {
        const int iters = 1 << 16;
        const int grid = prop.multiProcessorCount * 2;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            lds_read_kernel<<<grid, 1024>>>(iters, sink);
            cudaEventRecord(e1);
            if (cudaEventSynchronize(e1) != cudaSuccess) { rc = -1; break; }
            float ms = 0.0f;
            cudaEventElapsedTime(&ms, e0, e1);
            const double gbs = (double)grid * 1024 * 16 * iters / (ms * 1e-3) / 1e9;
            if (rep > 0 && gbs > best) best = gbs;
        }
    }
    if (cudaGetLastError() != cudaSuccess) rc = -1;
    cudaFree(sink);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *gb_per_s = best;
    return rc;
}

// FP64 FMA peak probe: the roofline denominator for the compute-bound batched kernels.  MEASURED_PEAKS.json
// records HBM and bf16 tensor peaks only, so the FP64 vector peak is measured here, on the same device, by a
// kernel that does nothing but independent DFMA chains (8 accumulators per thread, fully unrolled).
#include <cuda_runtime.h>
#include "../../include/gschur_cuda.h"

namespace {
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1.0, x2 = x0 + 2.0, x3 = x0 + 3.0;
    double x4 = x0 + 4.0, x5 = x0 + 5.0, x6 = x0 + 6.0, x7 = x0 + 7.0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            x0 = fma(x0, a, b);
            x1 = fma(x1, a, b);
            x2 = fma(x2, a, b);
            x3 = fma(x3, a, b);
            x4 = fma(x4, a, b);
            x5 = fma(x5, a, b);
            x6 = fma(x6, a, b);
            x7 = fma(x7, a, b);
        }
    }
    double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456) out[0] = s;   // keeps the chains live without a store in the common case
}

// L2 streaming probe: every thread reads, touches and writes back its own 16-byte words of a buffer that fits in L2,
// `passes` times, with the cache-global (L1-bypassing) accesses stage B's Z-warp uses.  Four independent words in
// flight per thread; n4 is a multiple of 4 * gridDim.x * blockDim.x.
__global__ void __launch_bounds__(256) l2_stream_kernel(uint4* buf, size_t n4, int passes) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int p = 0; p < passes; ++p) {
        for (size_t i = i0; i < n4; i += 4 * stride) {
            uint4 v0 = __ldcg(buf + i), v1 = __ldcg(buf + i + stride);
            uint4 v2 = __ldcg(buf + i + 2 * stride), v3 = __ldcg(buf + i + 3 * stride);
            v0.x += 1u;
            v1.x += 1u;
            v2.x += 1u;
            v3.x += 1u;
            __stcg(buf + i, v0);
            __stcg(buf + i + stride, v1);
            __stcg(buf + i + 2 * stride, v2);
            __stcg(buf + i + 3 * stride, v3);
        }
    }
}
// FP64 tensor (DMMA) probe: independent mma.sync.m8n8k4.f64 chains, operands in registers — the BLAS3 denominator of
// the large-matrix path (dgemm.cuh).  8 accumulator pairs per warp, 512 flop per instruction.
__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        c[u][0] = threadIdx.x * 1e-3 + u;
        c[u][1] = c[u][0] + 0.5;
    }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int u = 0; u < 8; ++u)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                             : "+d"(c[u][0]), "+d"(c[u][1])
                             : "d"(a), "d"(b));
        }
    }
    double s = 0.0;
#pragma unroll
    for (int u = 0; u < 8; ++u) s += c[u][0] + c[u][1];
    if (s == 123.456) out[0] = s;
}
}  // namespace

extern "C" int gschur_cuda_measure_dmma_peak(double* tflops, double* ms_out) {
    int dev = 0;
    if (!tflops) return GSCHUR_ERR_ARG;
    if (cudaGetDevice(&dev) != cudaSuccess) return GSCHUR_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return GSCHUR_ERR_CUDA;
    double* d = nullptr;
    if (cudaMalloc(&d, 8) != cudaSuccess) return GSCHUR_ERR_CUDA;
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 2048;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        dmma_peak_kernel<<<blocks, threads>>>(d, iters, 1e-9, 1e-9);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) {
            cudaFree(d);
            return GSCHUR_ERR_CUDA;
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    // per warp and instruction: m8n8k4 = 256 FMA = 512 flop; 32 instructions per iteration
    const double flops = 512.0 * 32.0 * (double)iters * (double)(threads / 32) * (double)blocks;
    *tflops = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return 0;
}

extern "C" int gschur_cuda_measure_fp64_peak(double* tflops, double* ms_out) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return GSCHUR_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return GSCHUR_ERR_CUDA;
    double* d = nullptr;
    if (cudaMalloc(&d, 8) != cudaSuccess) return GSCHUR_ERR_CUDA;
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        dfma_peak_kernel<<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) {
            cudaFree(d);
            return GSCHUR_ERR_CUDA;
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    double flops = 2.0 * 8.0 * 16.0 * (double)iters * (double)threads * (double)blocks;
    *tflops = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return 0;
}

extern "C" int gschur_cuda_measure_l2_bandwidth(double* gbs, double* ms_out) {
    int dev = 0;
    if (!gbs) return GSCHUR_ERR_ARG;
    if (cudaGetDevice(&dev) != cudaSuccess) return GSCHUR_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return GSCHUR_ERR_CUDA;
    const int blocks = prop.multiProcessorCount * 8, threads = 256, passes = 64;
    const size_t n4 = (size_t)blocks * threads * 8;   // 148 SMs: 2.4 M words = 38.8 MB, well inside the 126 MB L2
    uint4* d = nullptr;
    if (cudaMalloc(&d, n4 * sizeof(uint4)) != cudaSuccess) return GSCHUR_ERR_CUDA;
    cudaMemset(d, 0, n4 * sizeof(uint4));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {   // the first repetition also pulls the buffer into L2
        cudaEventRecord(e0);
        l2_stream_kernel<<<blocks, threads>>>(d, n4, passes);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) {
            cudaFree(d);
            return GSCHUR_ERR_CUDA;
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    *gbs = 2.0 * (double)(n4 * sizeof(uint4)) * passes / (best * 1e-3) / 1e9;   // bytes read + bytes written
    if (ms_out) *ms_out = best;
    return 0;
}

// Latency / issue micro-benchmarks for the FP64 chain design (development aid; not part of the product).
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
__global__ void k_dfma_dep(double* out, long long* cyc, double a, double b) {
    double x = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; ++i) x = fma(x, a, b);
    long long t1 = clock64();
    out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_dadd_dep(double* out, long long* cyc, double a) {
    double x = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; ++i) x = x + a;
    long long t1 = clock64();
    out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int ILP> __global__ void k_dfma_ilp(double* out, long long* cyc, double a, double b) {
    double x[ILP];
    for (int j = 0; j < ILP; ++j) x[j] = out[threadIdx.x] + j;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) x[j] = fma(x[j], a, b);
    }
    long long t1 = clock64();
    double s = 0; for (int j = 0; j < ILP; ++j) s += x[j];
    out[threadIdx.x] = s; if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_rcp_dep(double* out, long long* cyc) {
    double x = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; ++i) { double y; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); x = y; }
    long long t1 = clock64();
    out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_rsq_dep(double* out, long long* cyc) {
    double x = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; ++i) { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); x = y; }
    long long t1 = clock64();
    out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_shfl_dep(double* out, long long* cyc) {
    int x = (int)out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; ++i) x = __shfl_sync(0xffffffffu, x, (x + 1) & 31);
    long long t1 = clock64();
    out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_lds_dep(double* out, long long* cyc) {
    __shared__ int buf[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = (i * 17 + 5) & 1023;
    __syncthreads();
    int x = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; ++i) x = buf[x];
    long long t1 = clock64();
    out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_sts_lds(double* out, long long* cyc) {   // store by lane l, syncwarp, load by lane l^1 (128-bit)
    __shared__ double2 buf[64];
    double2 v = make_double2(out[threadIdx.x], 1.0);
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        buf[threadIdx.x] = v;
        __syncwarp();
        v = buf[threadIdx.x ^ 1];
        __syncwarp();
    }
    long long t1 = clock64();
    out[threadIdx.x] = v.x + v.y; if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// FP64 issue cost vs active lanes: ILP-8 DFMA with only `act` lanes active
__global__ void k_dfma_lanes(double* out, long long* cyc, double a, double b, int act) {
    double x[8];
    for (int j = 0; j < 8; ++j) x[j] = out[threadIdx.x] + j;
    long long t0 = clock64();
    if ((threadIdx.x & 31) < act) {
#pragma unroll 16
        for (int i = 0; i < N; ++i) {
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = fma(x[j], a, b);
        }
    }
    __syncwarp();
    long long t1 = clock64();
    double s = 0; for (int j = 0; j < 8; ++j) s += x[j];
    out[threadIdx.x] = s; if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// mixed: dependent DFMA chain + independent integer ALU work: does the int stream fill the FP64 latency?
int main() {
    double* out; long long* cyc; long long h[8];
    cudaMalloc(&out, 1024 * 8); cudaMemset(out, 0, 1024 * 8); cudaMalloc(&cyc, 64 * 8);
#define RUN(name, call, per) call; cudaMemcpy(h, cyc, 8 * 8, cudaMemcpyDeviceToHost); printf("%-28s %.2f cycles/op\n", name, (double)h[0] / (per));
    for (int rep = 0; rep < 2; ++rep) {
    RUN("dfma dependent (1 warp)", (k_dfma_dep<<<1, 32>>>(out, cyc, 1.0000001, 1e-9)), N)
    RUN("dadd dependent (1 warp)", (k_dadd_dep<<<1, 32>>>(out, cyc, 1e-9)), N)
    RUN("dfma ilp2 per-instr", (k_dfma_ilp<2><<<1, 32>>>(out, cyc, 1.0000001, 1e-9)), N * 2)
    RUN("dfma ilp4 per-instr", (k_dfma_ilp<4><<<1, 32>>>(out, cyc, 1.0000001, 1e-9)), N * 4)
    RUN("dfma ilp8 per-instr", (k_dfma_ilp<8><<<1, 32>>>(out, cyc, 1.0000001, 1e-9)), N * 8)
    RUN("dfma ilp8, 4 warps/SMSP(16w)", (k_dfma_ilp<8><<<1, 512>>>(out, cyc, 1.0000001, 1e-9)), N * 8)
    RUN("dfma ilp1, 16 warps/block", (k_dfma_ilp<1><<<1, 512>>>(out, cyc, 1.0000001, 1e-9)), N)
    RUN("rcp.approx.f64 dependent", (k_rcp_dep<<<1, 32>>>(out, cyc)), N)
    RUN("rsqrt.approx.f64 dependent", (k_rsq_dep<<<1, 32>>>(out, cyc)), N)
    RUN("shfl dependent", (k_shfl_dep<<<1, 32>>>(out, cyc)), N)
    RUN("lds dependent", (k_lds_dep<<<1, 32>>>(out, cyc)), N)
    RUN("sts+syncwarp+lds128+syncwarp", (k_sts_lds<<<1, 32>>>(out, cyc)), N)
    RUN("dfma ilp8 lanes=32", (k_dfma_lanes<<<1, 32>>>(out, cyc, 1.0000001, 1e-9, 32)), N * 8)
    RUN("dfma ilp8 lanes=16", (k_dfma_lanes<<<1, 32>>>(out, cyc, 1.0000001, 1e-9, 16)), N * 8)
    RUN("dfma ilp8 lanes=8", (k_dfma_lanes<<<1, 32>>>(out, cyc, 1.0000001, 1e-9, 8)), N * 8)
    RUN("dfma ilp8 lanes=1", (k_dfma_lanes<<<1, 32>>>(out, cyc, 1.0000001, 1e-9, 1)), N * 8)
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}

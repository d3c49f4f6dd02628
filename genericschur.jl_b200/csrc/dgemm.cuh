// FP64 tensor-core GEMM for regime (2) (one large matrix): C = alpha * op(A) * op(B) + beta * C, column-major.
// tcgen05 has no f64 kind; the Blackwell FP64 tensor path is the legacy warp-level `mma.sync.m8n8k4.f64`
// (SASS DMMA.8x8x4).  Tiling: 64x64 output tile per 128-thread CTA (4 warps in 2x2, 32x32 per warp = 4x4 DMMA
// tiles, 32 accumulator doubles per lane), K staged 16 at a time through shared memory with an 8-double row pad
// (fragment loads hit 32 distinct banks).  Global loads are coalesced along the contiguous dimension of each
// operand and double-buffered in registers.  The shapes of this code path are "thin": K = panel / window width
// (32..128) with M or N in the thousands, so the kernel favours simplicity over the last 20 % of DMMA peak.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gs {

constexpr int GEMM_BM = 64, GEMM_BN = 64, GEMM_BK = 16, GEMM_PAD = 8;

__device__ __forceinline__ void dmma8x8x4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// element (m, k) of op(A): TA = false: A[m + k*lda]; TA = true: A[k + m*lda].   Same for op(B) (k, n).
template <bool TA, bool TB>
__global__ void __launch_bounds__(128) dgemm_kernel(int M, int N, int K, double alpha, const double* __restrict__ A, int lda,
                                                    const double* __restrict__ B, int ldb, double beta, double* C, int ldc) {
    __shared__ double As[2][GEMM_BK][GEMM_BM + GEMM_PAD];
    __shared__ double Bs[2][GEMM_BK][GEMM_BN + GEMM_PAD];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 1, wn = warp >> 1;
    const int gid = lane >> 2, tig = lane & 3;
    const int m0 = blockIdx.x * GEMM_BM, n0 = blockIdx.y * GEMM_BN;

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // each thread stages 8 elements of the A tile and 8 of the B tile per K-step (64*16 / 128)
    double ra[8], rb[8];
    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int idx = tid + 128 * e;
            int m, k;
            if (!TA) { m = idx & 63; k = idx >> 6; }      // contiguous in m
            else     { k = idx & 15; m = idx >> 4; }      // contiguous in k
            const int gm = m0 + m, gk = k0 + k;
            double v = 0.0;
            if (gm < M && gk < K) v = TA ? A[(size_t)gk + (size_t)gm * lda] : A[(size_t)gm + (size_t)gk * lda];
            ra[e] = v;
            int n, kk;
            if (!TB) { kk = idx & 15; n = idx >> 4; }     // B[k + n*ldb]: contiguous in k
            else     { n = idx & 63; kk = idx >> 6; }     // B[n + k*ldb]: contiguous in n
            const int gn = n0 + n, gk2 = k0 + kk;
            double w = 0.0;
            if (gn < N && gk2 < K) w = TB ? B[(size_t)gn + (size_t)gk2 * ldb] : B[(size_t)gk2 + (size_t)gn * ldb];
            rb[e] = w;
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int idx = tid + 128 * e;
            int m, k;
            if (!TA) { m = idx & 63; k = idx >> 6; }
            else     { k = idx & 15; m = idx >> 4; }
            As[buf][k][m] = ra[e];
            int n, kk;
            if (!TB) { kk = idx & 15; n = idx >> 4; }
            else     { n = idx & 63; kk = idx >> 6; }
            Bs[buf][kk][n] = rb[e];
        }
    };

    const int nk = (K + GEMM_BK - 1) / GEMM_BK;
    if (nk > 0) {
        load_tiles(0);
        store_tiles(0);
    }
    __syncthreads();
    for (int t = 0; t < nk; ++t) {
        const int buf = t & 1;
        if (t + 1 < nk) load_tiles((t + 1) * GEMM_BK);
#pragma unroll
        for (int kk = 0; kk < GEMM_BK; kk += 4) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[buf][kk + tig][wm * 32 + i * 8 + gid];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[buf][kk + tig][wn * 32 + j * 8 + gid];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma8x8x4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        if (t + 1 < nk) store_tiles(buf ^ 1);
        __syncthreads();
    }
    // epilogue: lane holds C[row = gid][col = 2*tig + {0,1}] of each 8x8 tile
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int gm = m0 + wm * 32 + i * 8 + gid;
                const int gn = n0 + wn * 32 + j * 8 + 2 * tig + c;
                if (gm < M && gn < N) {
                    double* p = C + (size_t)gm + (size_t)gn * ldc;
                    const double v = alpha * acc[i][j][c];
                    *p = (beta == 0.0) ? v : v + beta * (*p);
                }
            }
}

void note_launch();

// C (M x N) = alpha * op(A) (M x K) * op(B) (K x N) + beta * C.  Must not alias C with A or B.
inline cudaError_t dgemm(cudaStream_t s, bool ta, bool tb, int M, int N, int K, double alpha, const double* A, int lda,
                         const double* B, int ldb, double beta, double* C, int ldc) {
    if (M <= 0 || N <= 0) return cudaSuccess;
    dim3 grid((M + GEMM_BM - 1) / GEMM_BM, (N + GEMM_BN - 1) / GEMM_BN);
    if (!ta && !tb) dgemm_kernel<false, false><<<grid, 128, 0, s>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
    else if (ta && !tb) dgemm_kernel<true, false><<<grid, 128, 0, s>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
    else if (!ta && tb) dgemm_kernel<false, true><<<grid, 128, 0, s>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
    else dgemm_kernel<true, true><<<grid, 128, 0, s>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
    note_launch();
    return cudaGetLastError();
}

}  // namespace gs

// Regime (2): ONE large Float64 matrix on one GPU (BASELINE config 4: n = 4096).
//   gschur_cuda_hessenberg_large : blocked WY Hessenberg reduction + Q (large_gehrd.cuh)
//   gschur_cuda_large            : + windowed multi-bulge QR iteration with DMMA GEMM updates (large_qr.cuh)
#include <cstring>
#include <string>
#include <vector>

#include "../../include/gschur_cuda.h"
#include "large_gehrd.cuh"

namespace gs {
void note_launch();
}
using namespace gs;

namespace {
thread_local std::string l_err;
}
extern "C" const char* gschur_cuda_large_last_error(void) { return l_err.c_str(); }

#define L_TRY(expr)                                                              \
    do {                                                                         \
        cudaError_t e__ = (expr);                                                \
        if (e__ != cudaSuccess) {                                                \
            l_err = std::string(#expr) + ": " + cudaGetErrorString(e__);         \
            return GSCHUR_ERR_CUDA;                                              \
        }                                                                        \
    } while (0)

extern "C" int gschur_cuda_hessenberg_large(int n, double* A, int lda, double* tau, double* Q, int ldq, uint32_t flags) {
    l_err.clear();
    if (n < 0 || lda < n || (Q && ldq < n) || !A) {
        l_err = "DimensionMismatch: bad n / lda / ldq";
        return GSCHUR_ERR_ARG;
    }
    if (n == 0) return 0;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        l_err = "no CUDA device available (there is no CPU fallback)";
        return GSCHUR_ERR_CUDA;
    }
    cudaStream_t s = 0;
    const bool dev = (flags & GSCHUR_FLAG_DEVICE_PTRS) != 0;
    const size_t nn = (size_t)n * n;
    double *dA = nullptr, *dQ = nullptr;
    if (dev && lda == n && (!Q || ldq == n)) {
        dA = A;
        dQ = Q;
    } else {
        L_TRY(cudaMallocAsync((void**)&dA, nn * sizeof(double), s));
        L_TRY(cudaMemcpy2DAsync(dA, (size_t)n * 8, A, (size_t)lda * 8, (size_t)n * 8, n, cudaMemcpyDefault, s));
        if (Q) L_TRY(cudaMallocAsync((void**)&dQ, nn * sizeof(double), s));
    }
    LargeWork w{};
    std::string err;
    int rc = lg_alloc(w, n, s, &err);
    w.A = dA;
    if (rc == 0) rc = lg_gehrd(w, dQ, s, &err);
    if (rc) {
        l_err = err;
        return rc;
    }
    if (tau && n > 1) L_TRY(cudaMemcpyAsync(tau, w.tau, (size_t)(n - 1) * 8, cudaMemcpyDefault, s));
    if (dA != A) {
        L_TRY(cudaMemcpy2DAsync(A, (size_t)lda * 8, dA, (size_t)n * 8, (size_t)n * 8, n, cudaMemcpyDefault, s));
        if (Q) L_TRY(cudaMemcpy2DAsync(Q, (size_t)ldq * 8, dQ, (size_t)n * 8, (size_t)n * 8, n, cudaMemcpyDefault, s));
    }
    lg_free(w, s);
    L_TRY(cudaStreamSynchronize(s));
    if (dA != A) {
        cudaFree(dA);
        if (dQ) cudaFree(dQ);
    }
    return 0;
}

// C = alpha op(A) op(B) + beta C with the library's DMMA kernel (device pointers); exposed for the test-suite.
extern "C" int gschur_cuda_dgemm(int ta, int tb, int M, int N, int K, double alpha, const double* A, int lda,
                                 const double* B, int ldb, double beta, double* C, int ldc) {
    l_err.clear();
    cudaError_t e = dgemm(0, ta != 0, tb != 0, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
    if (e == cudaSuccess) e = cudaStreamSynchronize(0);
    if (e != cudaSuccess) {
        l_err = cudaGetErrorString(e);
        return GSCHUR_ERR_CUDA;
    }
    return 0;
}

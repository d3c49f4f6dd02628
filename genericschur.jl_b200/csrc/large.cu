// Regime (2): ONE large Float64 matrix on one GPU (BASELINE config 4: n = 4096).
//   gschur_cuda_hessenberg_large : blocked WY Hessenberg reduction + Q (large_gehrd.cuh)
//   gschur_cuda_large            : + windowed multi-bulge QR iteration with DMMA GEMM updates (large_qr.cuh)
#include <cmath>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include "../../include/gschur_cuda.h"
#include "large_qr.cuh"

namespace gs {
void note_launch();
}
using namespace gs;

namespace {
thread_local std::string l_err;
}
// The stream-ordered scratch of the large-matrix path (and of the batched kernels it calls for small blocks and shifts:
// hundreds of calls per matrix) stays in the device's default pool across synchronisations instead of going back to the
// driver at each of them.
static void keep_default_pool() {
    int d = 0;
    cudaMemPool_t mp = nullptr;
    if (cudaGetDevice(&d) == cudaSuccess && cudaDeviceGetDefaultMemPool(&mp, d) == cudaSuccess && mp) {
        unsigned long long thr = ~0ULL;
        cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &thr);
    }
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

// frees what an entry point allocated on every way out (the error returns of L_TRY included)
namespace {
struct OnExit {
    std::function<void()> f;
    ~OnExit() {
        if (f) f();
    }
};
}  // namespace

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
    keep_default_pool();
    const bool dev = (flags & GSCHUR_FLAG_DEVICE_PTRS) != 0;
    const size_t nn = (size_t)n * n;
    double *dA = nullptr, *dQ = nullptr;
    const bool inplace = dev && lda == n && (!Q || ldq == n);
    LargeWork w{};
    OnExit cleanup{[&] {
        lg_free(w, s);
        if (!inplace) {
            cudaStreamSynchronize(s);
            if (dA) cudaFree(dA);
            if (dQ) cudaFree(dQ);
        }
    }};
    if (inplace) {
        dA = A;
        dQ = Q;
    } else {
        L_TRY(cudaMallocAsync((void**)&dA, nn * sizeof(double), s));
        L_TRY(cudaMemcpy2DAsync(dA, (size_t)n * 8, A, (size_t)lda * 8, (size_t)n * 8, n, cudaMemcpyDefault, s));
        if (Q) L_TRY(cudaMallocAsync((void**)&dQ, nn * sizeof(double), s));
    }
    std::string err;
    int rc = lg_alloc(w, n, s, &err);
    w.A = dA;
    if (rc == 0) rc = lg_gehrd(w, dQ, s, &err);
    if (rc) {
        l_err = err;
        return rc;
    }
    if (tau && n > 1) L_TRY(cudaMemcpyAsync(tau, w.tau, (size_t)(n - 1) * 8, cudaMemcpyDefault, s));
    if (!inplace) {
        L_TRY(cudaMemcpy2DAsync(A, (size_t)lda * 8, dA, (size_t)n * 8, (size_t)n * 8, n, cudaMemcpyDefault, s));
        if (Q) L_TRY(cudaMemcpy2DAsync(Q, (size_t)ldq * 8, dQ, (size_t)n * 8, (size_t)n * 8, n, cudaMemcpyDefault, s));
    }
    L_TRY(cudaStreamSynchronize(s));
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

namespace {
__global__ void lg_absmax_kernel(const double* A, size_t count, unsigned long long* out) {
    double m = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
        m = fmax(m, fabs(A[i]));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(m));   // m >= 0: bit order == value order
}
__global__ void lg_scale_kernel(double* A, size_t count, double mul) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) A[i] *= mul;
}
}  // namespace

// gschur!(A::Matrix{Float64}; wantZ = (Z != NULL), scale) for ONE large matrix (src/GenericSchur.jl:805-835).
//   A  in: matrix; out: quasi-triangular T        Z  out: Schur vectors (NULL ok)       w  out: n complex eigenvalues
//   info (NULL ok): 0, or k > 0 = iteration limit reached with the active block ending at row k.
//   stats3 (NULL ok): multishift sweeps, chase windows, blocks finished by the batched kernel.
extern "C" int gschur_cuda_large(int n, double* A, int lda, double* Z, int ldz, double* w, int scale, int* info,
                                 long long* stats3, uint32_t flags) {
    l_err.clear();
    if (n < 0 || lda < n || (Z && ldz < n) || (n > 0 && (!A || !w))) {
        l_err = "DimensionMismatch: bad n / lda / ldz / NULL pointer";
        return GSCHUR_ERR_ARG;
    }
    if (info) *info = 0;
    if (n == 0) return 0;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        l_err = "no CUDA device available (there is no CPU fallback)";
        return GSCHUR_ERR_CUDA;
    }
    cudaStream_t s = 0;
    keep_default_pool();
    const bool dev = (flags & GSCHUR_FLAG_DEVICE_PTRS) != 0;
    const size_t nn = (size_t)n * n;
    double *dA = nullptr, *dZ = nullptr, *dw = nullptr;
    const bool inplace = dev && lda == n && (!Z || ldz == n);
    OnExit cleanup{[&] {
        if (!inplace) {
            cudaStreamSynchronize(s);
            if (dA) cudaFree(dA);
            if (dZ) cudaFree(dZ);
            if (dw) cudaFree(dw);
        }
    }};
    if (inplace) {
        dA = A;
        dZ = Z;
        dw = w;
    } else {
        L_TRY(cudaMallocAsync((void**)&dA, nn * sizeof(double), s));
        L_TRY(cudaMemcpy2DAsync(dA, (size_t)n * 8, A, (size_t)lda * 8, (size_t)n * 8, n, cudaMemcpyDefault, s));
        if (Z) L_TRY(cudaMallocAsync((void**)&dZ, nn * sizeof(double), s));
        L_TRY(cudaMallocAsync((void**)&dw, (size_t)2 * n * sizeof(double), s));
    }
    // _scale! (src/util.jl:14-29)
    bool scaled = false;
    double cscale = 1.0, anrm = 1.0;
    if (scale) {
        unsigned long long* dmax = nullptr;
        L_TRY(cudaMallocAsync((void**)&dmax, 8, s));
        L_TRY(cudaMemsetAsync(dmax, 0, 8, s));
        lg_absmax_kernel<<<1024, 256, 0, s>>>(dA, nn, dmax);
        note_launch();
        unsigned long long bits = 0;
        L_TRY(cudaMemcpyAsync(&bits, dmax, 8, cudaMemcpyDeviceToHost, s));
        L_TRY(cudaStreamSynchronize(s));
        cudaFreeAsync(dmax, s);
        std::memcpy(&anrm, &bits, 8);
        const double smlnum = std::sqrt(2.2250738585072014e-308) / 2.220446049250313e-16, bignum = 1.0 / smlnum;
        if (anrm > 0.0 && anrm < smlnum) { scaled = true; cscale = smlnum; }
        else if (anrm > bignum) { scaled = true; cscale = bignum; }
        if (scaled) {
            lg_scale_kernel<<<1024, 256, 0, s>>>(dA, nn, cscale / anrm);
            note_launch();
        }
    }
    LargeWork wk{};
    std::string err;
    int rc = lg_alloc(wk, n, s, &err);
    wk.A = dA;
    if (rc == 0) rc = lg_gehrd(wk, dZ, s, &err);
    if (rc == 0) {
        lg_clear_tails_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, s>>>(dA, n);
        note_launch();
    }
    lg_free(wk, s);
    LargeQrStats st;
    int qinfo = 0;
    if (rc == 0) {
        qinfo = lg_qr(dA, dZ, n, dw, s, &err, &st);
        if (qinfo < 0) rc = qinfo;
    }
    if (rc) {
        l_err = err;
        return rc;
    }
    if (scaled) {
        lg_scale_kernel<<<1024, 256, 0, s>>>(dA, nn, anrm / cscale);
        lg_scale_kernel<<<64, 256, 0, s>>>(dw, (size_t)2 * n, anrm / cscale);
        note_launch();
        note_launch();
    }
    if (!inplace) {
        L_TRY(cudaMemcpy2DAsync(A, (size_t)lda * 8, dA, (size_t)n * 8, (size_t)n * 8, n, cudaMemcpyDefault, s));
        if (Z) L_TRY(cudaMemcpy2DAsync(Z, (size_t)ldz * 8, dZ, (size_t)n * 8, (size_t)n * 8, n, cudaMemcpyDefault, s));
        L_TRY(cudaMemcpyAsync(w, dw, (size_t)2 * n * 8, cudaMemcpyDefault, s));
    }
    L_TRY(cudaStreamSynchronize(s));
    if (info) *info = qinfo;
    if (stats3) {
        stats3[0] = st.sweeps;
        stats3[1] = st.windows;
        stats3[2] = st.small_blocks;
    }
    return qinfo > 0 ? 1 : 0;
}

// Balancing pre-step and triangularize post-step on the device — SURVEY.md section 8(f) rank 3.
//   gschur_cuda_balance_batched       balance!(A; scale, permute)                 src/balance.jl:33-199
//   gschur_cuda_balance_apply_batched lmul!(B::Balancer, V) / ldiv!(B, V)         src/balance.jl:203-260
//   gschur_cuda_triangularize_batched triangularize(S::Schur{<:Real})             src/triang.jl:9-43
// One warp per matrix (the matrix stays in global memory: every entry is touched a handful of times).  The reference's
// searches (first row / column whose off-diagonal part inside the active block is zero) keep their order: a ballot over
// 32 candidates at a time picks the one the serial loop would have stopped at.  The scaling loop is the reference's,
// row by row; the two-norms are formed from per-lane partial sums (index stride 32) combined by an xor butterfly, with
// explicitly unfused multiplies and adds — an order a CPU checker can reproduce exactly, so every power-of-two decision
// and therefore D, ilo, ihi, the permutation and the balanced matrix agree bit for bit.
#include <cuda_runtime.h>
#include <cstdint>
#include <string>
#include "../../include/gschur_cuda.h"
#include "launch.h"
#include "scalar.cuh"

namespace gs {

typedef cx<double> Cz;

GS_DEV bool bal_nz(double x) { return x != 0.0; }
GS_DEV bool bal_nz(const Cz& x) { return x.re != 0.0 || x.im != 0.0; }
GS_DEV double bal_abs2(double x) { return __dmul_rn(x, x); }
GS_DEV double bal_abs2(const Cz& x) { return __dadd_rn(__dmul_rn(x.re, x.re), __dmul_rn(x.im, x.im)); }
GS_DEV double bal_sabs2(double x, double s) {
    const double t = __dmul_rn(x, s);
    return __dmul_rn(t, t);
}
GS_DEV double bal_sabs2(const Cz& x, double s) {
    const double a = __dmul_rn(x.re, s), b = __dmul_rn(x.im, s);
    return __dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b));
}
GS_DEV double bal_mul(double x, double f) { return __dmul_rn(x, f); }
GS_DEV Cz bal_mul(const Cz& x, double f) { return mk_cx<double>(__dmul_rn(x.re, f), __dmul_rn(x.im, f)); }
GS_DEV double bal_wsum(double v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, m));
    return v;
}
GS_DEV double bal_wmax(double v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, m));
    return v;
}

struct BalParams {
    void* A;
    double* D;
    int* ilo_ihi;   // 3 ints per matrix: ilo, ihi, trivial
    int* sp;        // n ints per matrix (1-based targets; 0 = unused)
    int* info;      // may be null
    long long strideA, batch;
    int lda, n, scale, permute;
};

// (amax, norm2) of the strided vector x[0], x[st], ..., x[(len-1) st]: amax = sqrt(max |x|^2) (src/norm1est.jl:76-100),
// norm2 = amax * sqrt(sum (|x| / amax)^2)
template <class E> GS_DEV void bal_norms(const E* x, long long st, int len, int lane, double& amax, double& nrm) {
    double m2 = 0.0;
    for (int t = lane; t < len; t += 32) m2 = fmax(m2, bal_abs2(x[(long long)t * st]));
    m2 = bal_wmax(m2);
    amax = sqrt(m2);
    if (!(amax > 0.0) || !(amax < 1.7976931348623157e308)) {
        // zero vector, or overflow / NaN in the squares: fall back to the scaled maximum of the moduli
        double mx = 0.0;
        for (int t = lane; t < len; t += 32) {
            const E v = x[(long long)t * st];
            double a;
            if constexpr (sizeof(E) == 16) a = hypot(((const double*)&v)[0], ((const double*)&v)[1]);
            else a = fabs(((const double*)&v)[0]);
            mx = fmax(mx, a);
            if (a != a) mx = a;
        }
        mx = bal_wmax(mx);
        amax = mx;
    }
    if (!(amax > 0.0)) {
        nrm = amax;      // 0 (or NaN)
        return;
    }
    const double s = 1.0 / amax;
    double q = 0.0;
    for (int t = lane; t < len; t += 32) q = __dadd_rn(q, bal_sabs2(x[(long long)t * st], s));
    q = bal_wsum(q);
    nrm = __dmul_rn(amax, sqrt(q));
}

template <class E> __global__ void __launch_bounds__(128) gschur_balance_kernel(BalParams p) {
    const int lane = threadIdx.x & 31;
    const long long b = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= p.batch) return;
    const int n = p.n, lda = p.lda;
    E* A = reinterpret_cast<E*>(p.A) + b * p.strideA;
    double* D = p.D + b * n;
    int* sp = p.sp + b * n;
#define AA(i, j) A[(size_t)((i)-1) + (size_t)((j)-1) * lda]
    for (int i = lane; i < n; i += 32) {
        D[i] = 1.0;
        sp[i] = 0;
    }
    __syncwarp();
    int ilo = 1, ihi = n;
    bool trivial = true;
    int info = 0;
    auto swap_rc = [&](int js, int ms) {
        // columns js <-> ms in rows 1..ihi, rows js <-> ms in columns ilo..n (src/balance.jl:79-86, 118-125)
        for (int i = 1 + lane; i <= ihi; i += 32) {
            const E t = AA(i, js);
            AA(i, js) = AA(i, ms);
            AA(i, ms) = t;
        }
        __syncwarp();
        for (int i = ilo + lane; i <= n; i += 32) {
            const E t = AA(js, i);
            AA(js, i) = AA(ms, i);
            AA(ms, i) = t;
        }
        __syncwarp();
    };
    if (p.permute) {
        // ---- rows whose off-diagonal part in columns 1..ihi is zero go to the bottom (src/balance.jl:56-91) ----
        ihi = n + 1;
        while (ihi > 1) {
            ihi -= 1;
            int js = 0;
            for (int base = ihi; base >= 1 && !js; base -= 32) {
                const int j = base - lane;
                bool ok = j >= 1;
                if (ok)
                    for (int i = 1; i <= ihi; ++i)
                        if (i != j && bal_nz(AA(j, i))) {
                            ok = false;
                            break;
                        }
                const unsigned m = __ballot_sync(0xffffffffu, ok);
                if (m) js = base - (__ffs(m) - 1);
            }
            if (!js) break;
            const int ms = ihi;
            if (lane == 0) sp[ms - 1] = js;
            if (js != ms) {
                trivial = false;
                swap_rc(js, ms);
            }
        }
        // ---- columns whose off-diagonal part in rows ilo..ihi is zero go to the left (src/balance.jl:93-131) ----
        if (ihi > 1) {
            ilo = 0;
            while (ilo < n) {
                ilo += 1;
                int js = 0;
                for (int base = ilo; base <= ihi && !js; base += 32) {
                    const int j = base + lane;
                    bool ok = j <= ihi;
                    if (ok)
                        for (int i = ilo; i <= ihi; ++i)
                            if (i != j && bal_nz(AA(i, j))) {
                                ok = false;
                                break;
                            }
                    const unsigned m = __ballot_sync(0xffffffffu, ok);
                    if (m) js = base + (__ffs(m) - 1);
                }
                if (!js) break;
                const int ms = ilo;
                if (lane == 0) sp[ms - 1] = js;
                if (ms != js) {
                    trivial = false;
                    swap_rc(js, ms);
                }
            }
        }
    }
    if (p.scale) {
        // ---- diagonal similarity with powers of two (src/balance.jl:137-194, algo = :pr, p = 1) ----
        const double beta = 2.0, factor = 0.95;
        const double sfmin1 = 2.2250738585072014e-308 / 2.220446049250313e-16;
        const double sfmin2 = sfmin1 * beta, sfmax2 = 1.0 / sfmin2;
        bool converged = false;
        int guard = 0;
        while (!converged && info == 0 && guard++ < 10000) {
            converged = true;
            for (int i = ilo; i <= ihi; ++i) {
                double c, r, ca, ra;
                const int len = ihi - ilo + 1;
                bal_norms<E>(&AA(ilo, i), 1, len, lane, ca, c);
                bal_norms<E>(&AA(i, ilo), lda, len, lane, ra, r);
                if (c == 0.0 || r == 0.0) continue;
                double g = r / beta;
                const double s = c + r;
                double f = 1.0;
                while (c < r / beta) {
                    if (c >= g || (fmax(f, fmax(c, ca)) >= sfmax2) || (fmin(r, fmin(g, ra)) <= sfmin2)) break;
                    const double chk = c + f + ca + r + g + ra;
                    if (chk != chk) {
                        info = -5;      // error("NaN encountered while balancing")
                        break;
                    }
                    f *= beta;
                    c *= beta;
                    ca *= beta;
                    r /= beta;
                    g /= beta;
                    ra /= beta;
                }
                if (info) break;
                g = c / beta;
                while (r <= c / beta) {
                    if ((g < r) || (fmax(r, ra) >= sfmax2) || (fmin(fmin(f, c), fmin(g, ca)) <= sfmin2)) break;
                    f /= beta;
                    c /= beta;
                    g /= beta;
                    ca /= beta;
                    r *= beta;
                    ra *= beta;
                }
                if (f != 1.0) trivial = false;
                if (c + r >= factor * s) continue;
                converged = false;
                if (lane == 0) D[i - 1] *= f;
                const double rf = 1.0 / f;
                for (int j = ilo + lane; j <= n; j += 32) AA(i, j) = bal_mul(AA(i, j), rf);
                __syncwarp();
                for (int j = 1 + lane; j <= ihi; j += 32) AA(j, i) = bal_mul(AA(j, i), f);
                __syncwarp();
            }
        }
    }
    if (lane == 0) {
        p.ilo_ihi[3 * b + 0] = ilo;
        p.ilo_ihi[3 * b + 1] = ihi;
        p.ilo_ihi[3 * b + 2] = trivial ? 1 : 0;
        if (p.info) p.info[b] = info;
    }
#undef AA
}

// lmul!(B, V) (inv = 0: right eigenvectors) / ldiv!(B, V) (inv = 1: left eigenvectors): rows scaled by D (or 1 / D), then
// the row exchanges undone (src/balance.jl:203-260).  One warp per matrix; V is n x n complex or real.
template <class E> __global__ void __launch_bounds__(128) gschur_balance_apply_kernel(E* V, long long strideV, int ldv, int n,
                                                                                     long long batch, const double* D,
                                                                                     const int* ilo_ihi, const int* sp, int inv) {
    const int lane = threadIdx.x & 31;
    const long long b = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= batch) return;
    E* W = V + b * strideV;
    const int ilo = ilo_ihi[3 * b], ihi = ilo_ihi[3 * b + 1], trivial = ilo_ihi[3 * b + 2];
    if (trivial) return;
    const double* d = D + b * n;
    const int* s = sp + b * n;
    if (ilo != ihi) {
        for (int j = 0; j < n; ++j)
            for (int i = lane; i < n; i += 32) W[(size_t)i + (size_t)j * ldv] = bal_mul(W[(size_t)i + (size_t)j * ldv], inv ? 1.0 / d[i] : d[i]);
        __syncwarp();
    }
    auto swap_rows = [&](int j, int m) {
        for (int i = lane; i < n; i += 32) {
            const E t = W[(size_t)(j - 1) + (size_t)i * ldv];
            W[(size_t)(j - 1) + (size_t)i * ldv] = W[(size_t)(m - 1) + (size_t)i * ldv];
            W[(size_t)(m - 1) + (size_t)i * ldv] = t;
        }
        __syncwarp();
    };
    for (int j = ilo - 1; j >= 1; --j) {
        const int m = s[j - 1];
        if (m != j && m >= 1) swap_rows(j, m);
    }
    for (int j = ihi + 1; j <= n; ++j) {
        const int m = s[j - 1];
        if (m != j && m >= 1) swap_rows(j, m);
    }
}

// triangularize (src/triang.jl:9-43): the 2x2 blocks of a standardised real quasi-triangular T are rotated to complex
// upper triangular form; Tc, Zc are the complex copies of T, Z, updated in place.  One warp per matrix.
__global__ void __launch_bounds__(128) gschur_triangularize_kernel(const double* Tr, long long strideT, int ldt, const double* Zr,
                                                                    long long strideZ, int ldz, Cz* Tc, Cz* Zc, Cz* w, int n,
                                                                    long long batch) {
    const int lane = threadIdx.x & 31;
    const long long b = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= batch) return;
    const double* T0 = Tr + b * strideT;
    Cz* T = Tc + b * (long long)n * n;
    Cz* Z = Zc ? Zc + b * (long long)n * n : nullptr;
    for (int j = 0; j < n; ++j)
        for (int i = lane; i < n; i += 32) {
            T[(size_t)i + (size_t)j * n] = mk_cx<double>(T0[(size_t)i + (size_t)j * ldt], 0.0);
            if (Z) Z[(size_t)i + (size_t)j * n] = mk_cx<double>(Zr[b * strideZ + (size_t)i + (size_t)j * ldz], 0.0);
        }
    __syncwarp();
#define TT(i, j) T[(size_t)((i)-1) + (size_t)((j)-1) * n]
    for (int j = n; j >= 2; --j) {
        const double sub = T0[(size_t)(j - 1) + (size_t)(j - 2) * ldt];
        if (sub != 0.0) {
            const double s0 = sqrt(fabs(sub));
            const double c0 = sqrt(fabs(T0[(size_t)(j - 2) + (size_t)(j - 1) * ldt]));
            const double r = hypot(s0, c0);
            const Cz c = mk_cx<double>(c0 / r, 0.0);         // G = Givens(j-1, j, c, s), s = -i s0 / r
            const Cz s = mk_cx<double>(0.0, -(s0 / r));
            // lmul!(G, T): rows j-1, j:  [c s; -conj(s) c]
            for (int k = 1 + lane; k <= n; k += 32) {
                const Cz a1 = TT(j - 1, k), a2 = TT(j, k);
                TT(j - 1, k) = c * a1 + s * a2;
                TT(j, k) = -(cconj(s) * a1) + c * a2;
            }
            __syncwarp();
            // rmul!(T, G'), rmul!(Z, G'): columns j-1, j
            for (int k = 1 + lane; k <= n; k += 32) {
                const Cz a1 = TT(k, j - 1), a2 = TT(k, j);
                TT(k, j - 1) = a1 * c + a2 * cconj(s);
                TT(k, j) = -(a1 * s) + a2 * c;
                if (Z) {
                    const Cz z1 = Z[(size_t)(k - 1) + (size_t)(j - 2) * n], z2 = Z[(size_t)(k - 1) + (size_t)(j - 1) * n];
                    Z[(size_t)(k - 1) + (size_t)(j - 2) * n] = z1 * c + z2 * cconj(s);
                    Z[(size_t)(k - 1) + (size_t)(j - 1) * n] = -(z1 * s) + z2 * c;
                }
            }
            __syncwarp();
        }
    }
    // triu!(T); values = diag(T)
    for (int j = 0; j < n; ++j)
        for (int i = lane; i < n; i += 32)
            if (i > j) T[(size_t)i + (size_t)j * n] = mk_cx<double>(0.0, 0.0);
    __syncwarp();
    for (int i = lane; i < n; i += 32) w[b * n + i] = T[(size_t)i + (size_t)i * n];
#undef TT
}

}  // namespace gs

static thread_local std::string bal_err;
extern "C" const char* gschur_cuda_balance_last_error(void) { return bal_err.c_str(); }

namespace {
struct DevBuf {
    void* d = nullptr;
    void* h = nullptr;
    size_t bytes = 0;
    bool staged = false;
    cudaError_t in(const void* host, size_t nbytes, bool devp, bool copy_in) {
        bytes = nbytes;
        if (devp || !host) {
            d = const_cast<void*>(host);
            return cudaSuccess;
        }
        h = const_cast<void*>(host);
        staged = true;
        cudaError_t e = cudaMalloc(&d, nbytes ? nbytes : 1);
        if (e == cudaSuccess && copy_in) e = cudaMemcpy(d, host, nbytes, cudaMemcpyHostToDevice);
        return e;
    }
    cudaError_t out() { return staged && d ? cudaMemcpy(h, d, bytes, cudaMemcpyDeviceToHost) : cudaSuccess; }
    ~DevBuf() {
        if (staged && d) cudaFree(d);
    }
};
int bal_fail(const char* what, cudaError_t e) {
    bal_err = std::string(what) + ": " + cudaGetErrorString(e);
    return GSCHUR_ERR_CUDA;
}
bool have_device() {
    int nd = 0;
    return cudaGetDeviceCount(&nd) == cudaSuccess && nd >= 1;
}
}  // namespace

extern "C" int gschur_cuda_balance_batched(int kind, int n, int64_t batch, void* A, int lda, int64_t strideA, double* D,
                                           int32_t* ilo_ihi_trivial, int32_t* perm, int32_t* info, int scale, int permute,
                                           uint32_t flags) {
    using namespace gs;
    bal_err.clear();
    if (kind != GSCHUR_F64 && kind != GSCHUR_C64) {
        bal_err = "balance! is implemented for Float64 and ComplexF64";
        return GSCHUR_ERR_ARG;
    }
    if (n < 0 || batch < 0 || lda < n || (n > 0 && batch > 0 && (!A || !D || !ilo_ihi_trivial || !perm))) {
        bal_err = "DimensionMismatch: bad n / lda / NULL pointer";
        return GSCHUR_ERR_ARG;
    }
    if (n == 0 || batch == 0) return 0;
    if (!have_device()) {
        bal_err = "no CUDA device available (there is no CPU fallback)";
        return GSCHUR_ERR_CUDA;
    }
    const bool devp = (flags & GSCHUR_FLAG_DEVICE_PTRS) != 0;
    const size_t es = kind == GSCHUR_C64 ? 16 : 8;
    const size_t spanA = ((size_t)(batch - 1) * strideA + (size_t)(n - 1) * lda + n) * es;
    DevBuf bA, bD, bI, bP, bF;
    cudaError_t e;
    if ((e = bA.in(A, spanA, devp, true)) != cudaSuccess) return bal_fail("A", e);
    if ((e = bD.in(D, (size_t)batch * n * 8, devp, false)) != cudaSuccess) return bal_fail("D", e);
    if ((e = bI.in(ilo_ihi_trivial, (size_t)batch * 12, devp, false)) != cudaSuccess) return bal_fail("ilo/ihi", e);
    if ((e = bP.in(perm, (size_t)batch * n * 4, devp, false)) != cudaSuccess) return bal_fail("perm", e);
    if (info && (e = bF.in(info, (size_t)batch * 4, devp, false)) != cudaSuccess) return bal_fail("info", e);
    BalParams p;
    p.A = bA.d;
    p.D = (double*)bD.d;
    p.ilo_ihi = (int*)bI.d;
    p.sp = (int*)bP.d;
    p.info = info ? (int*)bF.d : nullptr;
    p.strideA = strideA;
    p.batch = batch;
    p.lda = lda;
    p.n = n;
    p.scale = scale;
    p.permute = permute;
    const unsigned grid = (unsigned)((batch + 3) / 4);
    if (kind == GSCHUR_C64) gschur_balance_kernel<Cz><<<grid, 128>>>(p);
    else gschur_balance_kernel<double><<<grid, 128>>>(p);
    note_launch();
    if ((e = cudaDeviceSynchronize()) != cudaSuccess) return bal_fail("balance kernel", e);
    if ((e = bA.out()) != cudaSuccess || (e = bD.out()) != cudaSuccess || (e = bI.out()) != cudaSuccess ||
        (e = bP.out()) != cudaSuccess || (info && (e = bF.out()) != cudaSuccess))
        return bal_fail("copy out", e);
    return 0;
}

extern "C" int gschur_cuda_balance_apply_batched(int kind, int n, int64_t batch, void* V, int ldv, int64_t strideV,
                                                 const double* D, const int32_t* ilo_ihi_trivial, const int32_t* perm,
                                                 int inverse, uint32_t flags) {
    using namespace gs;
    bal_err.clear();
    if (kind != GSCHUR_F64 && kind != GSCHUR_C64) {
        bal_err = "implemented for Float64 and ComplexF64";
        return GSCHUR_ERR_ARG;
    }
    if (n < 0 || batch < 0 || ldv < n || (n > 0 && batch > 0 && (!V || !D || !ilo_ihi_trivial || !perm))) {
        bal_err = "DimensionMismatch: bad n / ldv / NULL pointer";
        return GSCHUR_ERR_ARG;
    }
    if (n == 0 || batch == 0) return 0;
    if (!have_device()) {
        bal_err = "no CUDA device available (there is no CPU fallback)";
        return GSCHUR_ERR_CUDA;
    }
    const bool devp = (flags & GSCHUR_FLAG_DEVICE_PTRS) != 0;
    const size_t es = kind == GSCHUR_C64 ? 16 : 8;
    DevBuf bV, bD, bI, bP;
    cudaError_t e;
    if ((e = bV.in(V, ((size_t)(batch - 1) * strideV + (size_t)(n - 1) * ldv + n) * es, devp, true)) != cudaSuccess) return bal_fail("V", e);
    if ((e = bD.in(D, (size_t)batch * n * 8, devp, true)) != cudaSuccess) return bal_fail("D", e);
    if ((e = bI.in(ilo_ihi_trivial, (size_t)batch * 12, devp, true)) != cudaSuccess) return bal_fail("ilo/ihi", e);
    if ((e = bP.in(perm, (size_t)batch * n * 4, devp, true)) != cudaSuccess) return bal_fail("perm", e);
    const unsigned grid = (unsigned)((batch + 3) / 4);
    if (kind == GSCHUR_C64)
        gschur_balance_apply_kernel<Cz><<<grid, 128>>>((Cz*)bV.d, strideV, ldv, n, batch, (const double*)bD.d, (const int*)bI.d,
                                                       (const int*)bP.d, inverse);
    else
        gschur_balance_apply_kernel<double><<<grid, 128>>>((double*)bV.d, strideV, ldv, n, batch, (const double*)bD.d,
                                                           (const int*)bI.d, (const int*)bP.d, inverse);
    note_launch();
    if ((e = cudaDeviceSynchronize()) != cudaSuccess) return bal_fail("balance apply kernel", e);
    if ((e = bV.out()) != cudaSuccess) return bal_fail("copy out", e);
    return 0;
}

extern "C" int gschur_cuda_triangularize_batched(int n, int64_t batch, const double* T, int ldt, int64_t strideT, const double* Z,
                                                 int ldz, int64_t strideZ, void* Tc, void* Zc, void* w, uint32_t flags) {
    using namespace gs;
    bal_err.clear();
    if (n < 0 || batch < 0 || ldt < n || (Z && ldz < n) || (n > 0 && batch > 0 && (!T || !Tc || !w)) || (Z && !Zc)) {
        bal_err = "DimensionMismatch: bad n / leading dimension / NULL pointer";
        return GSCHUR_ERR_ARG;
    }
    if (n == 0 || batch == 0) return 0;
    if (!have_device()) {
        bal_err = "no CUDA device available (there is no CPU fallback)";
        return GSCHUR_ERR_CUDA;
    }
    const bool devp = (flags & GSCHUR_FLAG_DEVICE_PTRS) != 0;
    DevBuf bT, bZ, bTc, bZc, bw;
    cudaError_t e;
    if ((e = bT.in(T, ((size_t)(batch - 1) * strideT + (size_t)(n - 1) * ldt + n) * 8, devp, true)) != cudaSuccess) return bal_fail("T", e);
    if (Z && (e = bZ.in(Z, ((size_t)(batch - 1) * strideZ + (size_t)(n - 1) * ldz + n) * 8, devp, true)) != cudaSuccess) return bal_fail("Z", e);
    if ((e = bTc.in(Tc, (size_t)batch * n * n * 16, devp, false)) != cudaSuccess) return bal_fail("Tc", e);
    if (Z && (e = bZc.in(Zc, (size_t)batch * n * n * 16, devp, false)) != cudaSuccess) return bal_fail("Zc", e);
    if ((e = bw.in(w, (size_t)batch * n * 16, devp, false)) != cudaSuccess) return bal_fail("w", e);
    const unsigned grid = (unsigned)((batch + 3) / 4);
    gschur_triangularize_kernel<<<grid, 128>>>((const double*)bT.d, strideT, ldt, Z ? (const double*)bZ.d : nullptr, strideZ, ldz,
                                               (Cz*)bTc.d, Z ? (Cz*)bZc.d : nullptr, (Cz*)bw.d, n, batch);
    note_launch();
    if ((e = cudaDeviceSynchronize()) != cudaSuccess) return bal_fail("triangularize kernel", e);
    if ((e = bTc.out()) != cudaSuccess || (Z && (e = bZc.out()) != cudaSuccess) || (e = bw.out()) != cudaSuccess)
        return bal_fail("copy out", e);
    return 0;
}

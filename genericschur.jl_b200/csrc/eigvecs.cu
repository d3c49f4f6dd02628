// Eigenvectors from the Schur form on the device — SURVEY.md section 8(f) rank 2: what geigvecs / eigen! compute after
// gschur! (src/vectors.jl:45-131 right, :372-460 left; _usolve! / _cusolve! src/util.jl:128-461; _enormalize!
// :572-592), for ComplexF64 (the triangular case; the real quasi-triangular case goes through triangularize in the
// reference and is not built here).
//
// T and Z are already on the GPU after gschur!.  One CTA per matrix, ONE THREAD PER EIGENVECTOR: the back substitution
// for eigenvalue k_i is a strictly sequential recurrence on a private vector, and the n recurrences of a matrix are
// independent — exactly the thread-level parallelism of a batched problem.  T is read through the read-only path (the
// reference shifts its diagonal in place and restores it; here the shifted, floored pivot is formed on the fly), the
// private vectors live in a transposed scratch array (thread k_i owns W[k_i + i ld]: coalesced), and the multiplication
// by Z reads Z[j, i] at a warp-uniform address.  The overflow-guarded solves follow the reference statement by
// statement (growth bound, fast path, scaled path); the column 1-norms are shared by the threads and never modified
// (the reference scales them by tscale and back: the product is formed on the fly instead).
#include <cuda_runtime.h>
#include <cstdint>
#include <string>
#include "../../include/gschur_cuda.h"
#include "launch.h"
#include "scalar.cuh"

namespace gs {

#define TA(i, j) T[(size_t)((i)-1) + (size_t)((j)-1) * ldt]

typedef cx<double> Cz;

struct EvParams {
    const Cz* T;
    const Cz* Z;    // may be null: eigenvectors of T itself
    Cz* V;
    Cz* W;          // scratch, n x n per matrix
    long long strideT, strideZ, strideV, batch;
    int ldt, ldz, ldv, n;
    int normalize;  // apply _enormalize! (unit 2-norm, largest component real)
};

GS_DEV double ev_abs1(const Cz& z) { return fabs(z.re) + fabs(z.im); }
GS_DEV double ev_cabs1half(const Cz& z) { return fabs(z.re / 2) + fabs(z.im / 2); }
GS_DEV Cz ev_scale(const Cz& z, double s) { return mk_cx<double>(z.re * s, z.im * s); }

// private vector x[1..] of thread ki in the transposed scratch
struct EvVec {
    Cz* w;
    int ld;
    GS_DEV Cz get(int i) const { return w[(size_t)(i - 1) * ld]; }
    GS_DEV void set(int i, const Cz& v) const { w[(size_t)(i - 1) * ld] = v; }
};

// shifted diagonal of the system solved for eigenvalue lam (src/vectors.jl:87-90, 412-415): T[k,k] - lam, floored at smin
GS_DEV Cz ev_pivot(const Cz* T, int ldt, int k, const Cz& lam, double smin) {
    Cz d = T[(size_t)(k - 1) + (size_t)(k - 1) * ldt] - lam;
    if (ev_abs1(d) < smin) d = mk_cx<double>(smin, 0.0);
    return d;
}

// _usolve!(A, m, x, cnorm) (src/util.jl:128-298) for A = (T - lam I)[1:m, 1:m] with floored pivots; returns xscale
GS_DEV double ev_usolve(const Cz* T, int ldt, int m, const EvVec& x, const double* cnorm, const Cz& lam, double smin) {
    const double eps = 2.220446049250313e-16;
    const double half = 0.5;
    const double smallnum = 2.2250738585072014e-308 / eps;
    const double bignum = 1.0 / smallnum;
    double xscale = 1.0;
    double tmax = fabs(cnorm[0]);
    for (int j = 2; j <= m; ++j) tmax = fmax(tmax, fabs(cnorm[j - 1]));
    double tscale = 1.0;
    if (!(tmax <= bignum * half)) tscale = half / (smallnum * tmax);
    double xmax = ev_cabs1half(x.get(1));
    for (int j = 2; j <= m; ++j) xmax = fmax(xmax, ev_cabs1half(x.get(j)));
    double xbound = xmax, grow;
    if (tscale != 1.0) {
        grow = 0.0;
    } else {
        grow = half / fmax(xbound, smallnum);
        xbound = grow;
        bool toosmall = false;
        for (int j = m; j >= 1; --j) {
            toosmall = (grow <= smallnum);
            if (toosmall) break;
            const double tjj = ev_abs1(ev_pivot(T, ldt, j, lam, smin));
            if (tjj >= smallnum) xbound = fmin(xbound, fmin(1.0, tjj) * grow);
            else xbound = 0.0;
            if (tjj + cnorm[j - 1] >= smallnum) grow *= (tjj / (tjj + cnorm[j - 1]));
            else grow = 0.0;
        }
        if (!toosmall) grow = xbound;
    }
    if (grow * tscale > smallnum) {
        // the bound is fine: plain back substitution
        for (int j = m; j >= 1; --j) {
            const Cz xj = x.get(j) / ev_pivot(T, ldt, j, lam, smin);
            x.set(j, xj);
            for (int i = j - 1; i >= 1; --i) x.set(i, x.get(i) - TA(i, j) * xj);
        }
    } else {
        if (xmax > bignum * half) {
            xscale = (bignum * half) / xmax;
            for (int i = 1; i <= m; ++i) x.set(i, ev_scale(x.get(i), xscale));
            xmax = bignum;
        } else {
            xmax *= 2;
        }
        for (int j = m; j >= 1; --j) {
            double xj = ev_abs1(x.get(j));
            const Cz tjjs = ev_scale(ev_pivot(T, ldt, j, lam, smin), tscale);
            const double tjj = ev_abs1(tjjs);
            const double cnj = cnorm[j - 1] * tscale;
            if (tjj > smallnum) {
                if (tjj < 1.0) {
                    if (xj > tjj * bignum) {
                        const double rec = 1.0 / xj;
                        for (int i = 1; i <= m; ++i) x.set(i, ev_scale(x.get(i), rec));
                        xscale *= rec;
                        xmax *= rec;
                    }
                }
                x.set(j, x.get(j) / tjjs);
                xj = ev_abs1(x.get(j));
            } else if (tjj > 0.0) {
                if (xj > tjj * bignum) {
                    double rec = (tjj * bignum) / xj;
                    if (cnj > 1.0) rec /= cnj;
                    for (int i = 1; i <= m; ++i) x.set(i, ev_scale(x.get(i), rec));
                    xscale *= rec;
                    xmax *= rec;
                }
                x.set(j, x.get(j) / tjjs);
                xj = ev_abs1(x.get(j));
            } else {
                for (int i = 1; i <= m; ++i) x.set(i, mk_cx<double>(0.0, 0.0));
                x.set(j, mk_cx<double>(1.0, 0.0));
                xscale = 0.0;
                xmax = 0.0;
                xj = 1.0;
            }
            if (xj > 1.0) {
                double rec = 1.0 / xj;
                if (cnj > (bignum - xmax) * rec) {
                    rec *= half;
                    for (int i = 1; i <= m; ++i) x.set(i, ev_scale(x.get(i), rec));
                    xscale *= rec;
                }
            } else if (xj * cnj > bignum - xmax) {
                for (int i = 1; i <= m; ++i) x.set(i, ev_scale(x.get(i), half));
                xscale *= half;
            }
            if (j > 1) {
                const Cz xjt = ev_scale(x.get(j), tscale);
                for (int i = 1; i <= j - 1; ++i) x.set(i, x.get(i) - xjt * TA(i, j));
                xmax = ev_abs1(x.get(1));
                for (int i = 2; i <= j - 1; ++i) xmax = fmax(xmax, ev_abs1(x.get(i)));
            }
        }
        xscale /= tscale;
    }
    return xscale;
}

// _cusolve!(A, m, x, cnorm) (src/util.jl:300-461) for A = (T - lam I)[o+1 : o+m, o+1 : o+m] (conjugate-transposed solve);
// x and cnorm are indexed from 1 within the block
GS_DEV double ev_cusolve(const Cz* T, int ldt, int o, int m, const EvVec& x, const double* cnorm, const Cz& lam, double smin) {
    const double eps = 2.220446049250313e-16;
    const double half = 0.5;
    const double smallnum = 2.2250738585072014e-308 / eps;
    const double bignum = 1.0 / smallnum;
    double xscale = 1.0;
#define TB(i, j) T[(size_t)(o + (i)-1) + (size_t)(o + (j)-1) * ldt]
    double tmax = fabs(cnorm[0]);
    for (int j = 2; j <= m; ++j) tmax = fmax(tmax, fabs(cnorm[j - 1]));
    double tscale = 1.0;
    if (!(tmax <= bignum * half)) tscale = half / (smallnum * tmax);
    double xmax = ev_cabs1half(x.get(1));
    for (int j = 2; j <= m; ++j) xmax = fmax(xmax, ev_cabs1half(x.get(j)));
    double xbound = xmax, grow;
    if (tscale != 1.0) {
        grow = 0.0;
    } else {
        grow = half / fmax(xbound, smallnum);
        xbound = grow;
        bool toosmall = false;
        for (int j = 1; j <= m; ++j) {
            toosmall = (grow <= smallnum);
            if (toosmall) break;
            const double xj = 1.0 + cnorm[j - 1];
            grow = fmin(grow, xbound / xj);
            const double tjj = ev_abs1(ev_pivot(T, ldt, o + j, lam, smin));
            if (tjj >= smallnum) {
                if (xj > tjj) xbound *= (tjj / xj);
            } else {
                xbound = 0.0;
            }
        }
        if (!toosmall) grow = fmin(grow, xbound);
    }
    if (grow * tscale > smallnum) {
        for (int j = 1; j <= m; ++j) {
            Cz z = x.get(j);
            for (int i = 1; i <= j - 1; ++i) z = z - cconj(TB(i, j)) * x.get(i);
            x.set(j, z / cconj(ev_pivot(T, ldt, o + j, lam, smin)));
        }
    } else {
        if (xmax > bignum * half) {
            xscale = (bignum * half) / xmax;
            for (int i = 1; i <= m; ++i) x.set(i, ev_scale(x.get(i), xscale));
            xmax = bignum;
        } else {
            xmax *= 2;
        }
        for (int j = 1; j <= m; ++j) {
            double xj = ev_abs1(x.get(j));
            Cz uscale = mk_cx<double>(tscale, 0.0);
            bool used_diag = false;
            double rec = 1.0 / fmax(xmax, 1.0);
            const double cnj = cnorm[j - 1] * tscale;
            Cz tjjs = ev_scale(cconj(ev_pivot(T, ldt, o + j, lam, smin)), tscale);
            if (cnj > (bignum - xj) * rec) {
                rec *= half;
                const double tjj = ev_abs1(tjjs);
                if (tjj > 1.0) {
                    rec = fmin(1.0, rec * tjj);
                    uscale = uscale / tjjs;
                    used_diag = true;
                }
                if (rec < 1.0) {
                    for (int i = 1; i <= m; ++i) x.set(i, ev_scale(x.get(i), rec));
                    xscale *= rec;
                    xmax *= rec;
                }
            }
            Cz csumj = mk_cx<double>(0.0, 0.0);
            for (int i = 1; i <= j - 1; ++i) csumj = csumj + (cconj(TB(i, j)) * uscale) * x.get(i);
            if (!used_diag) {
                x.set(j, x.get(j) - csumj);
                xj = ev_abs1(x.get(j));
                const double tjj = ev_abs1(tjjs);
                if (tjj > smallnum) {
                    if (tjj < 1.0) {
                        if (xj > tjj * bignum) {
                            const double r2 = 1.0 / xj;
                            for (int i = 1; i <= m; ++i) x.set(i, ev_scale(x.get(i), r2));
                            xscale *= r2;
                            xmax *= r2;
                        }
                    }
                    x.set(j, x.get(j) / tjjs);
                } else if (tjj > 0.0) {
                    if (xj > tjj * bignum) {
                        const double r2 = (tjj * bignum) / xj;
                        for (int i = 1; i <= m; ++i) x.set(i, ev_scale(x.get(i), r2));
                        xscale *= r2;
                        xmax *= r2;
                    }
                    x.set(j, x.get(j) / tjjs);
                } else {
                    for (int i = 1; i <= m; ++i) x.set(i, mk_cx<double>(0.0, 0.0));
                    x.set(j, mk_cx<double>(1.0, 0.0));
                    xscale = 0.0;
                    xmax = 0.0;
                }
            } else {
                x.set(j, x.get(j) / tjjs - csumj);
            }
            xmax = fmax(xmax, ev_abs1(x.get(j)));
        }
        xscale /= tscale;
    }
    return xscale;
#undef TB
}

template <bool LEFT> __global__ void __launch_bounds__(128) gschur_eigvecs_kernel(EvParams p) {
    extern __shared__ double cnorm[];        // n column norms of the strictly upper part
    const int n = p.n, ki = threadIdx.x + 1;
    const long long b = blockIdx.x;
    const Cz* T = p.T + b * p.strideT;
    const Cz* Z = p.Z ? p.Z + b * p.strideZ : nullptr;
    Cz* V = p.V + b * p.strideV;
    const int ldt = p.ldt, ldv = p.ldv;
    // right: sum of moduli (src/vectors.jl:72-77); left: sum of abs1 (:394-399)
    for (int j = threadIdx.x + 1; j <= n; j += blockDim.x) {
        double s = 0.0;
        for (int i = 1; i <= j - 1; ++i) {
            const Cz t = TA(i, j);
            s += LEFT ? ev_abs1(t) : c_abs(t);
        }
        cnorm[j - 1] = s;
    }
    __syncthreads();
    if (ki > n) return;
    const double ulp = 2.220446049250313e-16;
    const double smallnum = 2.2250738585072014e-308 * (double)n;      // src/vectors.jl:62-63, 388-389
    const Cz lam = TA(ki, ki);
    const double smin = fmax(ulp * ev_abs1(lam), smallnum);
    EvVec x;
    x.w = p.W + b * (long long)n * n + (ki - 1);
    x.ld = n;
    double vscale = 1.0;
    if (!LEFT) {
        // (T[1:k,1:k] - lam I) x = -T[1:k, ki], k = ki - 1
        x.set(1, mk_cx<double>(1.0, 0.0));
        for (int k = 1; k <= ki - 1; ++k) x.set(k, -TA(k, ki));
        if (ki > 1) {
            vscale = ev_usolve(T, ldt, ki - 1, x, cnorm, lam, smin);
            x.set(ki, mk_cx<double>(vscale, 0.0));
        }
        // vectors[:, ki] = vscale Z[:, ki] + sum_{i < ki} Z[:, i] x[i]   (or x itself when there is no Z)
        if (Z) {
            for (int j = 1; j <= n; ++j) {
                Cz acc = ev_scale(Z[(size_t)(j - 1) + (size_t)(ki - 1) * p.ldz], vscale);
                for (int i = 1; i <= ki - 1; ++i) acc = acc + Z[(size_t)(j - 1) + (size_t)(i - 1) * p.ldz] * x.get(i);
                V[(size_t)(j - 1) + (size_t)(ki - 1) * ldv] = acc;
            }
        } else {
            for (int j = 1; j <= n; ++j)
                V[(size_t)(j - 1) + (size_t)(ki - 1) * ldv] = (j < ki) ? x.get(j) : (j == ki ? mk_cx<double>(ki > 1 ? vscale : 1.0, 0.0) : mk_cx<double>(0.0, 0.0));
        }
    } else {
        // (T[ki+1:n, ki+1:n] - lam I)^H x = -conj(T[ki, ki+1:n])
        EvVec y = x;                                    // block-local indexing: y[1] = x[ki+1]
        y.w = x.w + (size_t)ki * x.ld;
        for (int k = ki + 1; k <= n; ++k) x.set(k, -cconj(TA(ki, k)));
        if (ki < n) vscale = ev_cusolve(T, ldt, ki, n - ki, y, cnorm, lam, smin);
        // NOTE: the reference passes the FIRST n - ki entries of tnorms to _cusolve! (src/vectors.jl:417-420): restated as written
        x.set(ki, mk_cx<double>(vscale, 0.0));
        if (Z) {
            for (int j = 1; j <= n; ++j) {
                Cz acc = ev_scale(Z[(size_t)(j - 1) + (size_t)(ki - 1) * p.ldz], vscale);
                for (int i = ki + 1; i <= n; ++i) acc = acc + Z[(size_t)(j - 1) + (size_t)(i - 1) * p.ldz] * x.get(i);
                V[(size_t)(j - 1) + (size_t)(ki - 1) * ldv] = acc;
            }
        } else {
            for (int j = 1; j <= n; ++j)
                V[(size_t)(j - 1) + (size_t)(ki - 1) * ldv] = (j > ki) ? x.get(j) : (j == ki ? mk_cx<double>(ki < n ? vscale : 1.0, 0.0) : mk_cx<double>(0.0, 0.0));
        }
    }
    // normalise: largest abs1 component -> 1 (src/vectors.jl:113-121)
    Cz* vc = V + (size_t)(ki - 1) * ldv;
    double t0 = ev_abs1(vc[0]);
    for (int i = 2; i <= n; ++i) t0 = fmax(t0, ev_abs1(vc[i - 1]));
    const double remax = 1.0 / t0;
    for (int i = 1; i <= n; ++i) vc[i - 1] = ev_scale(vc[i - 1], remax);
    if (p.normalize) {
        // _enormalize! (src/util.jl:572-592): unit 2-norm, the component of largest modulus real
        double ssq = 0.0;
        for (int i = 1; i <= n; ++i) ssq += vc[i - 1].re * vc[i - 1].re + vc[i - 1].im * vc[i - 1].im;
        const double s = 1.0 / sqrt(ssq);
        double t = vc[0].re * vc[0].re + vc[0].im * vc[0].im;
        int i0 = 1;
        for (int i = 2; i <= n; ++i) {
            const double u = vc[i - 1].re * vc[i - 1].re + vc[i - 1].im * vc[i - 1].im;
            if (u > t) {
                i0 = i;
                t = u;
            }
        }
        const double st = s / sqrt(t);
        const Cz f = ev_scale(cconj(vc[i0 - 1]), st);
        for (int i = 1; i <= n; ++i) vc[i - 1] = vc[i - 1] * f;
        vc[i0 - 1].im = 0.0;
    }
}
#undef TA

}  // namespace gs

static thread_local std::string ev_err;

extern "C" const char* gschur_cuda_eigvecs_last_error(void) { return ev_err.c_str(); }

extern "C" int gschur_cuda_eigvecs_batched(int kind, int n, int64_t batch, const void* T, int ldt, int64_t strideT,
                                           const void* Z, int ldz, int64_t strideZ, void* V, int ldv, int64_t strideV,
                                           int left, uint32_t flags) {
    using namespace gs;
    ev_err.clear();
    if (kind != GSCHUR_C64) {
        ev_err = "eigenvectors from the Schur form are implemented for ComplexF64 (kind 1) only";
        return GSCHUR_ERR_ARG;
    }
    if (n < 0 || batch < 0 || (n > 0 && batch > 0 && (!T || !V)) || ldt < n || ldv < n || (Z && ldz < n)) {
        ev_err = "DimensionMismatch: bad n / leading dimension / NULL pointer";
        return GSCHUR_ERR_ARG;
    }
    if (n == 0 || batch == 0) return 0;
    if (n > 128) {
        ev_err = "n exceeds the batched-kernel limit 128";
        return GSCHUR_ERR_SIZE;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        ev_err = "no CUDA device available (there is no CPU fallback)";
        return GSCHUR_ERR_CUDA;
    }
    const bool devp = (flags & GSCHUR_FLAG_DEVICE_PTRS) != 0;
    const size_t es = sizeof(Cz);
    const Cz *dT = (const Cz*)T, *dZ = (const Cz*)Z;
    Cz *dV = (Cz*)V, *dW = nullptr, *hT = nullptr, *hZ = nullptr, *hV = nullptr;
    cudaError_t e = cudaSuccess;
    int rc = 0;
    auto fail = [&](const char* what) {
        ev_err = std::string(what) + ": " + cudaGetErrorString(e);
        rc = GSCHUR_ERR_CUDA;
    };
    const size_t spanT = (size_t)(batch - 1) * strideT + (size_t)(n - 1) * ldt + n;
    const size_t spanZ = Z ? (size_t)(batch - 1) * strideZ + (size_t)(n - 1) * ldz + n : 0;
    const size_t spanV = (size_t)(batch - 1) * strideV + (size_t)(n - 1) * ldv + n;
    do {
        if (!devp) {
            if ((e = cudaMalloc((void**)&hT, spanT * es)) != cudaSuccess) { fail("cudaMalloc(T)"); break; }
            if ((e = cudaMemcpy(hT, T, spanT * es, cudaMemcpyHostToDevice)) != cudaSuccess) { fail("cudaMemcpy(T)"); break; }
            dT = hT;
            if (Z) {
                if ((e = cudaMalloc((void**)&hZ, spanZ * es)) != cudaSuccess) { fail("cudaMalloc(Z)"); break; }
                if ((e = cudaMemcpy(hZ, Z, spanZ * es, cudaMemcpyHostToDevice)) != cudaSuccess) { fail("cudaMemcpy(Z)"); break; }
                dZ = hZ;
            }
            if ((e = cudaMalloc((void**)&hV, spanV * es)) != cudaSuccess) { fail("cudaMalloc(V)"); break; }
            dV = hV;
        }
        if ((e = cudaMalloc((void**)&dW, (size_t)batch * n * n * es)) != cudaSuccess) { fail("cudaMalloc(scratch)"); break; }
        EvParams p;
        p.T = dT;
        p.Z = dZ;
        p.V = dV;
        p.W = dW;
        p.strideT = strideT;
        p.strideZ = strideZ;
        p.strideV = strideV;
        p.batch = batch;
        p.ldt = ldt;
        p.ldz = ldz;
        p.ldv = ldv;
        p.n = n;
        p.normalize = (flags & 0x10u) ? 0 : 1;
        const int threads = (n + 31) & ~31;
        const size_t smem = (size_t)n * sizeof(double);
        for (int64_t b0 = 0; b0 < batch && rc == 0; b0 += 0x40000000LL) {
            const int64_t cnt = (batch - b0 < 0x40000000LL) ? batch - b0 : 0x40000000LL;
            EvParams q = p;
            q.T = p.T + b0 * strideT;
            if (q.Z) q.Z = p.Z + b0 * strideZ;
            q.V = p.V + b0 * strideV;
            q.W = p.W + b0 * (int64_t)n * n;
            if (left) gschur_eigvecs_kernel<true><<<(unsigned)cnt, threads, smem>>>(q);
            else gschur_eigvecs_kernel<false><<<(unsigned)cnt, threads, smem>>>(q);
            note_launch();
        }
        if ((e = cudaDeviceSynchronize()) != cudaSuccess) { fail("eigvecs kernel"); break; }
        if (!devp && (e = cudaMemcpy(V, hV, spanV * es, cudaMemcpyDeviceToHost)) != cudaSuccess) { fail("cudaMemcpy(V)"); break; }
    } while (0);
    if (dW) cudaFree(dW);
    if (hT) cudaFree(hT);
    if (hZ) cudaFree(hZ);
    if (hV) cudaFree(hV);
    return rc;
}

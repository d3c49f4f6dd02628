// Regime (2), part 1: blocked (compact-WY) Householder reduction of ONE large Float64 matrix to Hessenberg form and
// explicit formation of Q — what _hessenberg! (src/hessenberg.jl:3-17) and _materializeQ (:150-166) compute, organised
// as LAPACK xGEHRD/xLAHR2/xORGHR organise it so that all but the panel work is GEMM:
//   per panel column: one small single-CTA kernel (apply the panel's previous reflectors to the column, generate the
//   reflector, T column) and ONE memory-bound kernel over the whole GPU — y = A(:, trailing) * v — which streams the
//   trailing matrix once (coalesced 128-bit loads, split over row blocks x column chunks).  Its algorithmic traffic
//   is 8 (n-k)(n-k-i) bytes per column, (8/3) n^3 in total: the HBM roofline of SURVEY.md §8d.
//   per panel: the right and left block-reflector updates and Y's top rows as FP64 DMMA GEMMs (dgemm.cuh).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <string>
#include "dgemm.cuh"

namespace gs {

constexpr int LG_NB = 32;          // panel width
constexpr int LG_PT = 512;         // threads of the single-CTA panel kernel (128 registers each: 32 running sums)
constexpr int LG_GEMV_ROWS = 512;  // rows per gemv CTA (two rows per thread, 128-bit loads: 4 KB contiguous per column and CTA)
constexpr int LG_MAXCHUNKS = 64;   // column chunks of the panel gemv (partial sums are reduced by the next panel kernel)

struct LargeWork {
    int n, nb;
    double* A;       // n x n (lda = n): in -> Hessenberg H (+ reflector tails below the sub-diagonal)
    double* V;       // n x n: explicit unit-lower-trapezoidal reflector blocks (panel at its own columns)
    double* T;       // nb x nb per panel (upper triangular, zero-filled)
    double* Y;       // n x nb
    double* W1;      // nb x n
    double* W2;      // n x nb  (also nb x n)
    double* ypart;   // chunks x n partial gemv sums
    double* vcur;    // n: current reflector as a full-length vector (zeros above its leading 1)
    double* tsave;   // nb: V2' v of the current column; [nb] holds tau
    double* tau;     // n
    int chunks;
};

// ---- warp / block reductions ---------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, m));
    return v;
}

// Single CTA.  Panel starts at column p (0-based), local column i (1-based, 1..ib): global column c = p + i - 1.
// Rows of the panel's V / Y blocks: r0 = p + 1 .. n-1 (index rr = row - r0).
__global__ void __launch_bounds__(LG_PT) lg_panel_col_kernel(LargeWork w, int p, int i) {
    extern __shared__ double sm[];
    const int n = w.n, r0 = p + 1, m = n - r0;       // m rows in the block
    const int c = p + i - 1;
    double* b = sm;                                   // m doubles: the column being reduced
    double* red = sm + ((m + 31) & ~31);              // 32 x 32 reduction scratch
    double* wv = red + 32 * 32;                       // nb: small vectors
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nbv = i - 1;                            // number of previous reflectors in this panel
    double* Tp = w.T + (size_t)(p / w.nb) * w.nb * w.nb;
    double* Vp = w.V + (size_t)r0 + (size_t)p * n;    // V block: Vp[rr + j*n]
    double* Yp = w.Y + r0;                            // Y block rows r0..: Yp[rr + j*n]

    // (0) finish Y(:, i-1) and T(:, i-1) from the previous column's gemv partial sums
    if (nbv >= 1) {
        const int j = nbv - 1;                        // 0-based index of the previous reflector
        const double tauj = w.tsave[w.nb];
        for (int rr = tid; rr < m; rr += LG_PT) {
            double y = w.ypart[r0 + rr];          // accumulated by the gemv CTAs with atomicAdd
            w.ypart[r0 + rr] = 0.0;               // ready for the next gemv
            for (int q = 0; q < j; ++q) y -= Yp[rr + (size_t)q * n] * w.tsave[q];
            Yp[rr + (size_t)j * n] = tauj * y;
        }
        if (warp == 0) {
            // T(0:j-1, j) = -tau * T(0:j-1, 0:j-1) * tsave ;  T(j, j) = tau
            double acc = 0.0;
            if (lane < j)
                for (int q = lane; q < j; ++q) acc += Tp[lane + q * w.nb] * w.tsave[q];
            if (lane < j) Tp[lane + j * w.nb] = -tauj * acc;
            if (lane == j) Tp[j + j * w.nb] = tauj;
            if (lane > j && lane < w.nb) Tp[lane + j * w.nb] = 0.0;
        }
    }
    __syncthreads();
    // (1) load the column, apply the right-update of the previous reflectors: b -= Y(:, 0:nbv-1) * V(i-2, 0:nbv-1)'
    for (int rr = tid; rr < m; rr += LG_PT) {
        double v = w.A[(size_t)(r0 + rr) + (size_t)c * n];
        for (int q = 0; q < nbv; ++q) v -= Yp[rr + (size_t)q * n] * Vp[(nbv - 1) + (size_t)q * n];
        b[rr] = v;
    }
    __syncthreads();
    // (2) left-apply (I - V T' V') to b
    if (nbv >= 1) {
        // wv = V' b   (nbv reductions over m rows)
        double part[LG_NB];
#pragma unroll
        for (int q = 0; q < LG_NB; ++q) part[q] = 0.0;
        for (int rr = tid; rr < m; rr += LG_PT) {
            const double bv = b[rr];
#pragma unroll
            for (int q = 0; q < LG_NB; ++q)
                if (q < nbv) part[q] += Vp[rr + (size_t)q * n] * bv;
        }
#pragma unroll
        for (int q = 0; q < LG_NB; ++q) {
            if (q < nbv) {
                double s = warp_sum(part[q]);
                if (lane == 0) red[warp * 32 + q] = s;
            }
        }
        __syncthreads();
        if (warp == 0) {
            double s = 0.0;
            if (lane < nbv)
                for (int ww = 0; ww < LG_PT / 32; ++ww) s += red[ww * 32 + lane];
            // wv = T' * (V'b): T upper triangular -> (T' x)_q = sum_{r <= q} T[r][q] x_r
            double x = s;
            double tq = 0.0;
            for (int r = 0; r < nbv; ++r) {
                double xr = __shfl_sync(0xffffffffu, x, r);
                if (lane < nbv && r <= lane) tq += Tp[r + lane * w.nb] * xr;
            }
            if (lane < nbv) wv[lane] = tq;
        }
        __syncthreads();
        for (int rr = tid; rr < m; rr += LG_PT) {
            double v = b[rr];
            for (int q = 0; q < nbv; ++q) v -= Vp[rr + (size_t)q * n] * wv[q];
            b[rr] = v;
        }
        __syncthreads();
    }
    // (3) reflector on b[i-1 ..] (rows p+i ..): xLARFG as in src/householder.jl:12-54
    const int h = i - 1;                 // index of alpha within b
    const int tl = m - h - 1;            // tail length
    double amax = 0.0;
    for (int rr = h + 1 + tid; rr < m; rr += LG_PT) amax = fmax(amax, fabs(b[rr]));
    amax = warp_max(amax);
    if (lane == 0) red[warp] = amax;
    __syncthreads();
    amax = red[0];
    for (int ww = 1; ww < LG_PT / 32; ++ww) amax = fmax(amax, red[ww]);
    __syncthreads();
    double xnorm = 0.0;
    if (amax > 0.0 && tl > 0) {
        const double rs = 1.0 / amax;
        double ssq = 0.0;
        for (int rr = h + 1 + tid; rr < m; rr += LG_PT) {
            const double t = b[rr] * rs;
            ssq += t * t;
        }
        ssq = warp_sum(ssq);
        if (lane == 0) red[warp] = ssq;
        __syncthreads();
        ssq = 0.0;
        for (int ww = 0; ww < LG_PT / 32; ++ww) ssq += red[ww];
        __syncthreads();
        xnorm = amax * sqrt(ssq);
    }
    const double alpha = b[h];
    double tau = 0.0, beta = alpha, scal = 0.0;
    if (xnorm != 0.0) {
        beta = -copysign(hypot(alpha, xnorm), alpha);
        // (the sub-sfmin rescaling loop of xLARFG is not needed: _scale! has brought max|a_ij| into [smlnum, bignum])
        tau = (beta - alpha) / beta;
        scal = 1.0 / (alpha - beta);
    }
    // write back: H entries of column c (rows r0 .. p+i), reflector tail into A and into the dense V block
    for (int rr = tid; rr < m; rr += LG_PT) {
        double a_out, v_out;
        if (rr < h) { a_out = b[rr]; v_out = 0.0; }
        else if (rr == h) { a_out = beta; v_out = 1.0; }
        else { v_out = b[rr] * scal; a_out = v_out; }
        w.A[(size_t)(r0 + rr) + (size_t)c * n] = a_out;
        Vp[rr + (size_t)h * n] = v_out;
        w.vcur[r0 + rr] = v_out;
        b[rr] = v_out;
    }
    if (tid == 0) {
        w.tau[c] = tau;
        w.tsave[w.nb] = tau;
    }
    __syncthreads();
    // (4) tsave = V(:, 0:h-1)' v
    if (h >= 1) {
        double part[LG_NB];
#pragma unroll
        for (int q = 0; q < LG_NB; ++q) part[q] = 0.0;
        for (int rr = h + tid; rr < m; rr += LG_PT) {
            const double bv = b[rr];
#pragma unroll
            for (int q = 0; q < LG_NB; ++q)
                if (q < h) part[q] += Vp[rr + (size_t)q * n] * bv;
        }
#pragma unroll
        for (int q = 0; q < LG_NB; ++q) {
            if (q < h) {
                double s = warp_sum(part[q]);
                if (lane == 0) red[warp * 32 + q] = s;
            }
        }
        __syncthreads();
        if (warp == 0 && lane < h) {
            double s = 0.0;
            for (int ww = 0; ww < LG_PT / 32; ++ww) s += red[ww * 32 + lane];
            w.tsave[lane] = s;
        }
    }
}

// After the last column of the panel: finish Y(:, ib-1) and T(:, ib-1)
__global__ void __launch_bounds__(LG_PT) lg_panel_finish_kernel(LargeWork w, int p, int ib) {
    const int n = w.n, r0 = p + 1, m = n - r0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* Tp = w.T + (size_t)(p / w.nb) * w.nb * w.nb;
    double* Yp = w.Y + r0;
    const int j = ib - 1;
    const double tauj = w.tsave[w.nb];
    for (int rr = tid; rr < m; rr += LG_PT) {
        double y = w.ypart[r0 + rr];
        w.ypart[r0 + rr] = 0.0;
        for (int q = 0; q < j; ++q) y -= Yp[rr + (size_t)q * n] * w.tsave[q];
        Yp[rr + (size_t)j * n] = tauj * y;
    }
    if (warp == 0) {
        double acc = 0.0;
        if (lane < j)
            for (int q = lane; q < j; ++q) acc += Tp[lane + q * w.nb] * w.tsave[q];
        if (lane < j) Tp[lane + j * w.nb] = -tauj * acc;
        if (lane == j) Tp[j + j * w.nb] = tauj;
        if (lane > j && lane < w.nb) Tp[lane + j * w.nb] = 0.0;
    }
}

// The memory-bound kernel: y[r] += sum_{col in chunk} A(r, col) * v[col] for rows r0..n-1, columns c0..n-1.
// One row per thread (coalesced along the column-major rows), LG_GEMV_ROWS rows per CTA, `chunks` column chunks whose
// partial sums are combined with FP64 atomicAdd in L2 (y is cleared by the panel kernel that consumes it).
__global__ void __launch_bounds__(LG_GEMV_ROWS / 2) lg_gemv_kernel(LargeWork w, int r0, int c0) {
    // two consecutive rows per thread through one 128-bit load (rows start at an even index; n is even for that path)
    const int n = w.n;
    const bool vec2 = (n % 2 == 0);
    const int rbase = vec2 ? (r0 & ~1) : r0;
    const int r = rbase + blockIdx.x * LG_GEMV_ROWS + 2 * threadIdx.x;
    const int ncols = n - c0;
    const int per = (ncols + w.chunks - 1) / w.chunks;
    const int cb = c0 + blockIdx.y * per;
    const int ce = min(cb + per, n);
    __shared__ double vs[512];
    double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;      // a*: row r, b*: row r+1 (two accumulators each)
    for (int cc = cb; cc < ce; cc += 512) {
        const int cnt = min(512, ce - cc);
        __syncthreads();
        for (int t = threadIdx.x; t < cnt; t += LG_GEMV_ROWS / 2) vs[t] = w.vcur[cc + t];
        __syncthreads();
        if (r + 1 < n && vec2) {
            const double2* ap = reinterpret_cast<const double2*>(w.A + (size_t)r + (size_t)cc * n);
            const size_t st = (size_t)n / 2;
            int t = 0;
            for (; t + 8 <= cnt; t += 8) {       // 8 independent 128-bit loads in flight per thread
                double2 x[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) x[u] = __ldg(ap + (size_t)(t + u) * st);
#pragma unroll
                for (int u = 0; u < 8; u += 2) {
                    a0 = fma(x[u].x, vs[t + u], a0);
                    b0 = fma(x[u].y, vs[t + u], b0);
                    a1 = fma(x[u + 1].x, vs[t + u + 1], a1);
                    b1 = fma(x[u + 1].y, vs[t + u + 1], b1);
                }
            }
            for (; t < cnt; ++t) {
                const double2 x = __ldg(ap + (size_t)t * st);
                a0 = fma(x.x, vs[t], a0);
                b0 = fma(x.y, vs[t], b0);
            }
        } else if (r < n) {
            for (int t = 0; t < cnt; ++t) {
                a0 = fma(w.A[(size_t)r + (size_t)(cc + t) * n], vs[t], a0);
                if (r + 1 < n) b0 = fma(w.A[(size_t)r + 1 + (size_t)(cc + t) * n], vs[t], b0);
            }
        }
    }
    if (cb < ce) {
        if (r >= r0 && r < n) atomicAdd(&w.ypart[r], a0 + a1);
        if (r + 1 >= r0 && r + 1 < n) atomicAdd(&w.ypart[r + 1], b0 + b1);
    }
}

__global__ void lg_set_identity_kernel(double* Q, int n) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < (size_t)n * n) Q[idx] = ((idx % n) == (idx / n)) ? 1.0 : 0.0;
}
__global__ void lg_zero_kernel(double* p, size_t count) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < count) p[idx] = 0.0;
}
// zero everything below the first sub-diagonal (the reflector tails) once Q has been formed
__global__ void lg_clear_tails_kernel(double* A, int n) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < (size_t)n * n) {
        const int i = idx % n, j = idx / n;
        if (i > j + 1) A[idx] = 0.0;
    }
}

#define LG_TRY(expr)                                                            \
    do {                                                                        \
        cudaError_t e__ = (expr);                                               \
        if (e__ != cudaSuccess) {                                               \
            *err = std::string(#expr) + ": " + cudaGetErrorString(e__);         \
            return -2;                                                          \
        }                                                                       \
    } while (0)

// A (n x n, lda = n, device) -> Hessenberg (reflector tails below the sub-diagonal), tau; Q (n x n) if non-null.
// W must have been allocated by lg_alloc.  All work is enqueued on `s`.
inline int lg_gehrd(LargeWork& w, double* Q, cudaStream_t s, std::string* err) {
    const int n = w.n, nb = w.nb;
    const size_t nn = (size_t)n * n;
    lg_zero_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, s>>>(w.V, nn);
    note_launch();
    const size_t smem_panel = ((size_t)((n + 31) & ~31) + 32 * 32 + nb + 8) * sizeof(double);
    LG_TRY(cudaFuncSetAttribute(lg_panel_col_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_panel));
    int npanels = 0;
    for (int p = 0; p < n - 1; p += nb, ++npanels) {
        const int ib = (n - 1 - p < nb) ? n - 1 - p : nb;
        const int r0 = p + 1, m = n - r0;
        lg_zero_kernel<<<(nb * nb + 255) / 256, 256, 0, s>>>(w.T + (size_t)npanels * nb * nb, (size_t)nb * nb);
        note_launch();
        for (int i = 1; i <= ib; ++i) {
            lg_panel_col_kernel<<<1, LG_PT, smem_panel, s>>>(w, p, i);
            note_launch();
            // y = A(r0.., p+i ..) * v    (v has its leading 1 at row p+i)
            const int rowblocks = (m + 1 + LG_GEMV_ROWS - 1) / LG_GEMV_ROWS;
            int ch = (1184 + rowblocks - 1) / rowblocks;          // ~8 CTAs per SM
            ch = ch < 8 ? 8 : (ch > LG_MAXCHUNKS ? LG_MAXCHUNKS : ch);
            if (ch > (n - (p + i) + 63) / 64) ch = (n - (p + i) + 63) / 64 > 0 ? (n - (p + i) + 63) / 64 : 1;
            w.chunks = ch;
            dim3 grid(rowblocks, ch);
            lg_gemv_kernel<<<grid, LG_GEMV_ROWS / 2, 0, s>>>(w, r0, p + i);
            note_launch();
        }
        lg_panel_finish_kernel<<<1, LG_PT, 0, s>>>(w, p, ib);
        note_launch();
        const double* Vd = w.V + (size_t)r0 + (size_t)p * n;                 // m x ib, ld n
        const double* Tp = w.T + (size_t)npanels * nb * nb;                   // ib x ib, ld nb
        // (a) Y(0:r0-1, :) = A(0:r0-1, r0:n-1) * V * T
        LG_TRY(dgemm(s, false, false, r0, ib, m, 1.0, w.A + (size_t)r0 * n, n, Vd, n, 0.0, w.W2, n));
        LG_TRY(dgemm(s, false, false, r0, ib, ib, 1.0, w.W2, n, Tp, nb, 0.0, w.Y, n));
        // (b) right update: top rows, all columns r0..n-1;  lower rows, columns p+ib..n-1
        LG_TRY(dgemm(s, false, true, r0, m, ib, -1.0, w.Y, n, Vd, n, 1.0, w.A + (size_t)r0 * n, n));
        const int c1 = p + ib, mt = n - c1;                                   // trailing columns
        if (mt > 0) {
            LG_TRY(dgemm(s, false, true, m, mt, ib, -1.0, w.Y + r0, n, Vd + (c1 - r0), n, 1.0,
                         w.A + (size_t)r0 + (size_t)c1 * n, n));
            // (c) left update of A(r0:n-1, c1:n-1) with (I - V T' V')
            LG_TRY(dgemm(s, true, false, ib, mt, m, 1.0, Vd, n, w.A + (size_t)r0 + (size_t)c1 * n, n, 0.0, w.W1, nb));
            LG_TRY(dgemm(s, true, false, ib, mt, ib, 1.0, Tp, nb, w.W1, nb, 0.0, w.W2, nb));
            LG_TRY(dgemm(s, false, false, m, mt, ib, -1.0, Vd, n, w.W2, nb, 1.0, w.A + (size_t)r0 + (size_t)c1 * n, n));
        }
    }
    if (Q) {
        lg_set_identity_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, s>>>(Q, n);
        note_launch();
        for (int pi = npanels - 1; pi >= 0; --pi) {
            const int p = pi * nb;
            const int ib = (n - 1 - p < nb) ? n - 1 - p : nb;
            const int r0 = p + 1, m = n - r0;
            const double* Vd = w.V + (size_t)r0 + (size_t)p * n;
            const double* Tp = w.T + (size_t)pi * nb * nb;
            double* Qs = Q + (size_t)r0 + (size_t)r0 * n;
            // Q(r0:, r0:) <- (I - V T V') Q(r0:, r0:)
            LG_TRY(dgemm(s, true, false, ib, m, m, 1.0, Vd, n, Qs, n, 0.0, w.W1, nb));
            LG_TRY(dgemm(s, false, false, ib, m, ib, 1.0, Tp, nb, w.W1, nb, 0.0, w.W2, nb));
            LG_TRY(dgemm(s, false, false, m, m, ib, -1.0, Vd, n, w.W2, nb, 1.0, Qs, n));
        }
    }
    LG_TRY(cudaGetLastError());
    return 0;
}

inline int lg_alloc(LargeWork& w, int n, cudaStream_t s, std::string* err) {
    w.n = n;
    w.nb = LG_NB;
    w.chunks = LG_MAXCHUNKS;
    const size_t nn = (size_t)n * n;
    const int npanels = (n + w.nb - 1) / w.nb + 1;
    double* base = nullptr;
    const size_t total = nn + (size_t)npanels * w.nb * w.nb + 3 * (size_t)n * w.nb + (size_t)w.chunks * n + 2 * (size_t)n + w.nb + 64;
    LG_TRY(cudaMallocAsync((void**)&base, total * sizeof(double), s));
    w.V = base;
    w.T = w.V + nn;
    w.Y = w.T + (size_t)npanels * w.nb * w.nb;
    w.W1 = w.Y + (size_t)n * w.nb;
    w.W2 = w.W1 + (size_t)n * w.nb;
    w.ypart = w.W2 + (size_t)n * w.nb;
    w.vcur = w.ypart + (size_t)w.chunks * n;
    w.tau = w.vcur + n;
    w.tsave = w.tau + n;
    LG_TRY(cudaMemsetAsync(w.ypart, 0, (size_t)n * sizeof(double), s));
    return 0;
}
inline void lg_free(LargeWork& w, cudaStream_t s) {
    if (w.V) cudaFreeAsync(w.V, s);
    w.V = nullptr;
}

}  // namespace gs

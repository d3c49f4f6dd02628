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
#include <cooperative_groups.h>
#include <math.h>
#include <string>
#include "dgemm.cuh"

namespace gs {

constexpr int LG_NB = 32;          // panel width
constexpr int LG_PT = 512;         // threads of the single-CTA panel finish kernel
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

// The panel column kernel.  ONE thread-block cluster of LG_NC CTAs (distributed shared memory): the m rows of the
// panel's V / Y blocks are split over the CTAs of the cluster, every CTA keeps its rows of the column in its own shared
// memory, and the full-column reductions (V'b, the norm, V'v) are exchanged through DSMEM — every CTA publishes its
// partial sums in its own shared memory, cluster.sync(), every CTA adds up the partials of all ranks.  (Round 1 ran this
// on a single CTA: 82 us per column against 23 us for the gemv that streams the trailing matrix.)
// Panel starts at column p (0-based), local column i (1-based, 1..ib): global column c = p + i - 1.
// Rows of the panel's V / Y blocks: r0 = p + 1 .. n-1 (index rr = row - r0).
constexpr int LG_NC = 16;          // CTAs per cluster (non-portable size, allowed on sm_100)
constexpr int LG_CT = 256;         // threads per CTA of the panel kernel

__device__ __forceinline__ void lg_block_sum32(double (&part)[LG_NB], int nv, double* red, double* out, int tid) {
    // sums part[q] (q < nv) over the CTA; result in out[q] (shared), valid after the trailing __syncthreads
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int q = 0; q < LG_NB; ++q) {
        if (q < nv) {
            const double s = warp_sum(part[q]);
            if (lane == 0) red[warp * 32 + q] = s;
        }
    }
    __syncthreads();
    if (warp == 0) {
        double s = 0.0;
        if (lane < nv)
            for (int ww = 0; ww < LG_CT / 32; ++ww) s += red[ww * 32 + lane];
        if (lane < LG_NB) out[lane] = (lane < nv) ? s : 0.0;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(LG_CT) lg_panel_col_kernel(LargeWork w, int p, int i) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    extern __shared__ double sm[];
    const int n = w.n, r0 = p + 1, m = n - r0;       // m rows in the block
    const int c = p + i - 1;
    const int rpc = (((m + LG_NC - 1) / LG_NC) + 31) & ~31;     // rows per CTA
    const int lo = rank * rpc, hi = min(m, lo + rpc);            // this CTA's rows [lo, hi)
    double* b = sm;                                   // rpc doubles: this CTA's rows of the column being reduced
    double* red = sm + rpc;                           // (LG_CT / 32) x 32 reduction scratch
    double* xch = red + (LG_CT / 32) * 32;            // 40 doubles published to the cluster: [0..31] sums, [32] amax / ssq, [33] alpha
    double* wv = xch + 40;                            // nb: small vectors
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nbv = i - 1;                            // number of previous reflectors in this panel
    double* Tp = w.T + (size_t)(p / w.nb) * w.nb * w.nb;
    double* Vp = w.V + (size_t)r0 + (size_t)p * n;    // V block: Vp[rr + j*n]
    double* Yp = w.Y + r0;                            // Y block rows r0..: Yp[rr + j*n]
    double* bc = wv + 3 * LG_NB;                      // 4 doubles: cluster-wide scalars broadcast to the CTA
    // Cluster-wide sums of the published values (call after cluster.sync()).  The remote reads are spread over the lanes
    // (one DSMEM round trip instead of LG_NC dependent ones).
    auto gather_vec = [&]() {                         // returns, in warp 0 lane q, the sum over the ranks of xch[q]
        static_assert(LG_NC == 2 * (LG_CT / 32), "two ranks per warp");
        red[warp * 32 + lane] = cluster.map_shared_rank(xch, warp)[lane] + cluster.map_shared_rank(xch, warp + LG_CT / 32)[lane];
        __syncthreads();
        double sres = 0.0;
        if (warp == 0)
            for (int ww = 0; ww < LG_CT / 32; ++ww) sres += red[ww * 32 + lane];
        __syncthreads();
        return sres;
    };
    auto gather_scalars = [&](bool want_max) {        // bc[0] = max or sum over the ranks of xch[32], bc[1] = sum of xch[33]
        if (warp == 0) {
            double a = 0.0, s1 = 0.0;
            if (lane < LG_NC) {
                const double* x = cluster.map_shared_rank(xch, lane);
                a = x[32];
                s1 = x[33];
            }
            a = want_max ? warp_max(a) : warp_sum(a);
            s1 = warp_sum(s1);
            if (lane == 0) {
                bc[0] = a;
                bc[1] = s1;
            }
        }
        __syncthreads();
    };

    // (0) + (1) in one pass over the rows: finish Y(:, i-1) (and, rank 0, T(:, i-1)) from the previous column's gemv
    //     sums; load the column and apply the right-update of the previous reflectors,
    //     b = A(:, c) - Y(:, 0:nbv-1) * V(i-2, 0:nbv-1)'.  All loads of a row are issued together (one L2 round trip).
    double* ts = wv + LG_NB;                          // tsave (V2' v of the previous column) and the row V(c, 0:nbv-1)
    double* vr = ts + LG_NB;
    if (tid < LG_NB) {
        ts[tid] = (tid < nbv) ? w.tsave[tid] : 0.0;
        vr[tid] = (tid < nbv) ? Vp[(nbv - 1) + (size_t)tid * n] : 0.0;
    }
    __syncthreads();
    {
        const int j = nbv - 1;                        // 0-based index of the previous reflector (if any)
        const double tauj = (nbv >= 1) ? w.tsave[w.nb] : 0.0;
        for (int rr = lo + tid; rr < hi; rr += LG_CT) {
            double yrow[LG_NB];
#pragma unroll
            for (int q = 0; q < LG_NB; ++q) yrow[q] = (q < j) ? Yp[rr + (size_t)q * n] : 0.0;
            double v = w.A[(size_t)(r0 + rr) + (size_t)c * n];
            if (nbv >= 1) {
                double y = w.ypart[r0 + rr];          // accumulated by the gemv CTAs with atomicAdd
                w.ypart[r0 + rr] = 0.0;               // ready for the next gemv
#pragma unroll
                for (int q = 0; q < LG_NB; ++q) y -= yrow[q] * ts[q];      // ts[q] = 0 for q >= j: yrow is 0 there anyway
                y *= tauj;
                Yp[rr + (size_t)j * n] = y;
                v -= y * vr[j];
            }
#pragma unroll
            for (int q = 0; q < LG_NB; ++q) v -= yrow[q] * vr[q];
            b[rr - lo] = v;
        }
        if (nbv >= 1 && rank == 0 && warp == 0) {
            // T(0:j-1, j) = -tau * T(0:j-1, 0:j-1) * tsave ;  T(j, j) = tau
            double acc = 0.0;
            if (lane < j)
                for (int q = lane; q < j; ++q) acc += Tp[lane + q * w.nb] * ts[q];
            if (lane < j) Tp[lane + j * w.nb] = -tauj * acc;
            if (lane == j) Tp[j + j * w.nb] = tauj;
            if (lane > j && lane < w.nb) Tp[lane + j * w.nb] = 0.0;
            __threadfence();
        }
    }
    __syncthreads();
    // (2) left-apply (I - V T' V') to b
    if (nbv >= 1) {
        double part[LG_NB];
#pragma unroll
        for (int q = 0; q < LG_NB; ++q) part[q] = 0.0;
        for (int rr = lo + tid; rr < hi; rr += LG_CT) {
            const double bv = b[rr - lo];
#pragma unroll
            for (int q = 0; q < LG_NB; ++q)
                if (q < nbv) part[q] += Vp[rr + (size_t)q * n] * bv;
        }
        lg_block_sum32(part, nbv, red, xch, tid);
        cluster.sync();                               // partial V'b of every rank (and rank 0's T column) visible
        const double s = gather_vec();
        if (warp == 0) {
            // wv = T' * (V'b): T upper triangular -> (T' x)_q = sum_{r <= q} T[r][q] x_r
            double tq = 0.0;
            for (int r = 0; r < nbv; ++r) {
                const double xr = __shfl_sync(0xffffffffu, s, r);
                if (lane < nbv && r <= lane) tq += __ldcg(&Tp[r + lane * w.nb]) * xr;
            }
            if (lane < nbv) wv[lane] = tq;
        }
        cluster.sync();                               // everybody has read xch: it may be overwritten
        if (tid >= nbv && tid < LG_NB) wv[tid] = 0.0;
        __syncthreads();
        for (int rr = lo + tid; rr < hi; rr += LG_CT) {
            double v = b[rr - lo];
            double vrow[LG_NB];
#pragma unroll
            for (int q = 0; q < LG_NB; ++q) vrow[q] = (q < nbv) ? Vp[rr + (size_t)q * n] : 0.0;
#pragma unroll
            for (int q = 0; q < LG_NB; ++q) v -= vrow[q] * wv[q];
            b[rr - lo] = v;
        }
        __syncthreads();
    }
    // (3) reflector on b[i-1 ..] (rows p+i ..): xLARFG as in src/householder.jl:12-54
    const int h = i - 1;                 // index of alpha within b
    const int tl = m - h - 1;            // tail length
    {
        double amax = 0.0;
        for (int rr = max(lo, h + 1) + tid; rr < hi; rr += LG_CT) amax = fmax(amax, fabs(b[rr - lo]));
        amax = warp_max(amax);
        if (lane == 0) red[warp] = amax;
        __syncthreads();
        if (tid == 0) {
            double a = red[0];
            for (int ww = 1; ww < LG_CT / 32; ++ww) a = fmax(a, red[ww]);
            xch[32] = a;
            xch[33] = (h >= lo && h < hi) ? b[h - lo] : 0.0;      // alpha lives in exactly one CTA
        }
    }
    cluster.sync();
    gather_scalars(true);
    const double amax = bc[0], alpha = bc[1];
    cluster.sync();
    double xnorm = 0.0;
    if (amax > 0.0 && tl > 0) {          // uniform over the cluster
        const double rs = 1.0 / amax;
        double ssq = 0.0;
        for (int rr = max(lo, h + 1) + tid; rr < hi; rr += LG_CT) {
            const double t = b[rr - lo] * rs;
            ssq += t * t;
        }
        ssq = warp_sum(ssq);
        if (lane == 0) red[warp] = ssq;
        __syncthreads();
        if (tid == 0) {
            double q = 0.0;
            for (int ww = 0; ww < LG_CT / 32; ++ww) q += red[ww];
            xch[32] = q;
        }
        if (tid == 0) xch[33] = 0.0;
        cluster.sync();
        gather_scalars(false);
        xnorm = amax * sqrt(bc[0]);
        cluster.sync();
    }
    double tau = 0.0, beta = alpha, scal = 0.0;
    if (xnorm != 0.0) {
        beta = -copysign(hypot(alpha, xnorm), alpha);
        // (the sub-sfmin rescaling loop of xLARFG is not needed: _scale! has brought max|a_ij| into [smlnum, bignum])
        tau = (beta - alpha) / beta;
        scal = 1.0 / (alpha - beta);
    }
    // write back: H entries of column c (rows r0 .. p+i), reflector tail into A and into the dense V block
    for (int rr = lo + tid; rr < hi; rr += LG_CT) {
        double a_out, v_out;
        const double bv = b[rr - lo];
        if (rr < h) { a_out = bv; v_out = 0.0; }
        else if (rr == h) { a_out = beta; v_out = 1.0; }
        else { v_out = bv * scal; a_out = v_out; }
        w.A[(size_t)(r0 + rr) + (size_t)c * n] = a_out;
        Vp[rr + (size_t)h * n] = v_out;
        w.vcur[r0 + rr] = v_out;
        b[rr - lo] = v_out;
    }
    if (rank == 0 && tid == 0) {
        w.tau[c] = tau;
        w.tsave[w.nb] = tau;
    }
    __syncthreads();
    // (4) tsave = V(:, 0:h-1)' v
    if (h >= 1) {
        double part[LG_NB];
#pragma unroll
        for (int q = 0; q < LG_NB; ++q) part[q] = 0.0;
        for (int rr = max(lo, h) + tid; rr < hi; rr += LG_CT) {
            const double bv = b[rr - lo];
#pragma unroll
            for (int q = 0; q < LG_NB; ++q)
                if (q < h) part[q] += Vp[rr + (size_t)q * n] * bv;
        }
        lg_block_sum32(part, h, red, xch, tid);
        cluster.sync();
        const double tsum = gather_vec();
        if (rank == 0 && warp == 0 && lane < h) w.tsave[lane] = tsum;
        cluster.sync();                               // remote shared memory stays valid until every read is done
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
    const size_t smem_panel = ((size_t)((((n + LG_NC - 1) / LG_NC) + 31) & ~31) + (LG_CT / 32) * 32 + 40 + 3 * LG_NB + 8) * sizeof(double);   // b | red | xch | wv, ts, vr | bc
    LG_TRY(cudaFuncSetAttribute(lg_panel_col_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_panel));
    LG_TRY(cudaFuncSetAttribute(lg_panel_col_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t pcfg = {};
    cudaLaunchAttribute pattr[1];
    pattr[0].id = cudaLaunchAttributeClusterDimension;
    pattr[0].val.clusterDim.x = LG_NC;
    pattr[0].val.clusterDim.y = 1;
    pattr[0].val.clusterDim.z = 1;
    pcfg.gridDim = dim3(LG_NC, 1, 1);
    pcfg.blockDim = dim3(LG_CT, 1, 1);
    pcfg.dynamicSmemBytes = smem_panel;
    pcfg.stream = s;
    pcfg.attrs = pattr;
    pcfg.numAttrs = 1;
    int npanels = 0;
    for (int p = 0; p < n - 1; p += nb, ++npanels) {
        const int ib = (n - 1 - p < nb) ? n - 1 - p : nb;
        const int r0 = p + 1, m = n - r0;
        lg_zero_kernel<<<(nb * nb + 255) / 256, 256, 0, s>>>(w.T + (size_t)npanels * nb * nb, (size_t)nb * nb);
        note_launch();
        for (int i = 1; i <= ib; ++i) {
            LG_TRY(cudaLaunchKernelEx(&pcfg, lg_panel_col_kernel, w, p, i));
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

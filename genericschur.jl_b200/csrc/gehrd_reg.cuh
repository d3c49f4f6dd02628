// Stage A for Float64, n <= 64, third generation: scale -> Householder Hessenberg reduction -> explicit Q with every
// COLUMN OF THE MATRIX IN THE REGISTERS OF ONE THREAD.
//
// The thread-per-column kernel of gehrd.cuh keeps the tile in shared memory and walks it with dependent scalar loops:
// ncu (profiles/r02_stageA_f64.summary.txt) shows 192 k warp instructions per matrix for 19 k warp-FMAs of useful work,
// 2.8 short-scoreboard stall cycles per issue and an FP64 pipe at 12 %.  Here a column lives in N statically indexed
// registers, so the two rank-1 updates of a Householder step are straight-line FMA code:
//     left   w_j = v' a_j (v broadcast from shared memory, 128-bit loads), a_j -= tau w_j v          — thread-local
//     right  every thread publishes a_j v_j (its share of A v) to a conflict-free shared N x (N+1) array, thread r adds
//            up row r, w is broadcast back, a_j -= tau v_j w                                          — two barriers
// Rows above the reflector are handled by zeros in v (no predicates); columns left of it by tau = 0.  The reflector
// itself is formed by the thread that owns column i from its registers (src/householder.jl:12-54, with the
// sub-sfmin rescaling loop).  Q is accumulated backwards in registers with the reflector tails parked in the same
// shared array (src/hessenberg.jl:150-166 computes the same product applied to the identity).
// Same outputs as gehrd_q_kernel: A <- H (zeros below the sub-diagonal), Z <- Q, scratch <- (scaled?, cscale, anrm).
#pragma once
#include "gehrd.cuh"

namespace gs {

template <int N> struct gehrd_reg_layout {
    static constexpr int LD = N + 1;
    __host__ __device__ static constexpr size_t off_c() { return 0; }                                   // N x LD doubles
    __host__ __device__ static constexpr size_t off_v() { return (size_t)N * LD * sizeof(double); }     // N doubles
    __host__ __device__ static constexpr size_t off_w() { return off_v() + N * sizeof(double); }        // N doubles
    __host__ __device__ static constexpr size_t off_tau() { return off_w() + N * sizeof(double); }      // N doubles
    __host__ __device__ static constexpr size_t off_red() { return off_tau() + N * sizeof(double); }    // 8 doubles
    __host__ __device__ static constexpr size_t bytes() { return off_red() + 64; }
};

template <int N> __global__ void __launch_bounds__(N, (N == 64 ? 4 : 12)) gehrd_reg_kernel(BatchedParams p) {
    typedef gehrd_reg_layout<N> GL;
    constexpr int LD = GL::LD;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* C = reinterpret_cast<double*>(smem_raw + GL::off_c());
    double* vs = reinterpret_cast<double*>(smem_raw + GL::off_v());
    double* ws = reinterpret_cast<double*>(smem_raw + GL::off_w());
    double* taus = reinterpret_cast<double*>(smem_raw + GL::off_tau());
    double* red = reinterpret_cast<double*>(smem_raw + GL::off_red());
    __shared__ long long s_next;
    const int n = p.n, tid = threadIdx.x;
    const bool wantZ = p.Z != nullptr;
    const bool mine = tid < n;

    for (;;) {
        if (tid == 0) s_next = (long long)atomicAdd(p.counter, 1ULL);
        __syncthreads();
        const long long b = s_next;
        __syncthreads();
        if (b >= p.batch) break;
        double* gA = reinterpret_cast<double*>(p.A) + b * p.strideA;
        double a[N];
#pragma unroll
        for (int r = 0; r < N; ++r) a[r] = (mine && r < n) ? gA[r + (size_t)tid * p.lda] : 0.0;

        // ---- _scale! (src/util.jl:14-29) ----
        bool scaled = false;
        double cscale = 1.0, anrm = 1.0;
        if (p.scale) {
            double m = 0.0;
#pragma unroll
            for (int r = 0; r < N; ++r) m = fmax(m, fabs(a[r]));
            m = block_max<double, N>(m, red);
            anrm = m;
            const double smlnum = r_sqrt(r_safemin<double>()) / rtraits<double>::eps();
            const double bignum = 1.0 / smlnum;
            if (anrm > 0.0 && anrm < smlnum) {
                scaled = true;
                cscale = smlnum;
            } else if (anrm > bignum) {
                scaled = true;
                cscale = bignum;
            }
            if (scaled) {
                safescale_apply<double, double, N>(anrm, cscale, [&](double mul) {
#pragma unroll
                    for (int r = 0; r < N; ++r) a[r] *= mul;
                });
            }
        }
        // ---- _hessenberg! (src/hessenberg.jl:3-17): reflector i (0-based column i-1 ... here 0-based step c) ----
        for (int c = 0; c < n - 1; ++c) {
            __syncthreads();                              // the previous step's reads of vs / ws are done
            // ---- reflector from column c, rows c+1 .. n-1 (head at row h = c+1), src/householder.jl:12-54.  The owner
            //      publishes its column; the norm is a block reduction with one entry per thread; every thread forms
            //      beta, tau and the scaling factor; thread r writes v[r] ----
            const int h = c + 1;
            if (tid == c) {
#pragma unroll
                for (int r = 0; r < N; ++r) vs[r] = a[r];
            }
            __syncthreads();
            double x = (tid > h) ? vs[tid] : 0.0;          // tail entry of this thread (0 outside the tail)
            const double alpha = vs[h];
            double tau = 0.0, beta = alpha, scal = 0.0;
            const bool trivial = (n - 1 - c) <= 1;      // a real length-1 reflector is the identity
            double amax = block_max<double, N>(fabs(x), red);
            double xnorm = 0.0;
            if (!trivial && amax > 0.0) {
                const double t = x * (1.0 / amax);
                xnorm = amax * r_sqrt(block_sum<double, N>(t * t, red));
            }
            if (!trivial && xnorm != 0.0) {
                beta = -copysign(r_hypot4(alpha, 0.0, xnorm, 0.0), alpha);
                const double sfmin = 2.0 * rtraits<double>::floatmin() / rtraits<double>::eps();
                int kount = 0;
                double al = alpha;
                if (fabs(beta) < sfmin) {            // uniform over the block
                    const double rsfmin = 1.0 / sfmin;
                    bool smallb = true;
                    while (smallb) {
                        kount += 1;
                        x *= rsfmin;
                        beta *= rsfmin;
                        al *= rsfmin;
                        smallb = (fabs(beta) < sfmin) && (kount < 20);
                    }
                    amax = block_max<double, N>(fabs(x), red);
                    xnorm = 0.0;
                    if (amax > 0.0) {
                        const double t = x * (1.0 / amax);
                        xnorm = amax * r_sqrt(block_sum<double, N>(t * t, red));
                    }
                    beta = -copysign(r_hypot4(al, 0.0, xnorm, 0.0), al);
                }
                tau = (beta - al) / beta;
                scal = 1.0 / (al - beta);
                for (int q = 0; q < kount; ++q) beta *= sfmin;
                x *= scal;
            }
            __syncthreads();                              // everybody has read the published column
            vs[tid] = (tid < h) ? 0.0 : (tid == h ? 1.0 : x);
            if (tid == 0) taus[c] = tau;
            __syncthreads();
            if (tid == c && !trivial && xnorm != 0.0) {
                // the reflector tail stays in the owner's column below the sub-diagonal, beta on it
#pragma unroll
                for (int r = 0; r < N; r += 2) {
                    const double2 v2 = *reinterpret_cast<const double2*>(vs + r);
                    a[r] = (r > h) ? v2.x : (r == h ? beta : a[r]);
                    a[r + 1] = (r + 1 > h) ? v2.y : (r + 1 == h ? beta : a[r + 1]);
                }
            }
            const double teff = (tid > c) ? tau : 0.0;          // columns c+1 .. n-1 only
            const double vj = vs[tid < N ? tid : 0];
            // ---- lmul!(H', A[c+1:, c+1:]): w = v' a_j, a_j -= tau w v ----
            {
                // rows above the reflector carry v = 0: blocks of 8 rows that lie entirely above it are skipped
                double w0 = 0.0, w1 = 0.0, w2 = 0.0, w3 = 0.0;
#pragma unroll
                for (int blk = 0; blk < N / 8; ++blk) {
                    if (8 * blk + 7 >= h) {
#pragma unroll
                        for (int r = 8 * blk; r < 8 * blk + 8; r += 4) {
                            const double2 va = *reinterpret_cast<const double2*>(vs + r);
                            const double2 vb = *reinterpret_cast<const double2*>(vs + r + 2);
                            w0 = fma(va.x, a[r], w0);
                            w1 = fma(va.y, a[r + 1], w1);
                            w2 = fma(vb.x, a[r + 2], w2);
                            w3 = fma(vb.y, a[r + 3], w3);
                        }
                    }
                }
                const double tw = teff * ((w0 + w1) + (w2 + w3));
#pragma unroll
                for (int blk = 0; blk < N / 8; ++blk) {
                    if (8 * blk + 7 >= h) {
#pragma unroll
                        for (int r = 8 * blk; r < 8 * blk + 8; r += 2) {
                            const double2 v2 = *reinterpret_cast<const double2*>(vs + r);
                            a[r] = fma(-tw, v2.x, a[r]);
                            a[r + 1] = fma(-tw, v2.y, a[r + 1]);
                        }
                    }
                }
            }
            // ---- rmul!(A[:, c+1:], H): x = A v, A -= tau x v' ----
#pragma unroll
            for (int r = 0; r < N; ++r) C[r * LD + tid] = a[r] * vj;      // v_j = 0 for j <= c
            __syncthreads();
            {
                double x0 = 0.0, x1 = 0.0, x2 = 0.0, x3 = 0.0;
                const double* row = C + tid * LD;
#pragma unroll
                for (int blk = 0; blk < N / 8; ++blk) {
                    if (8 * blk + 7 >= h) {       // columns left of the reflector contributed zeros
#pragma unroll
                        for (int j = 8 * blk; j < 8 * blk + 8; j += 4) {
                            x0 += row[j];
                            x1 += row[j + 1];
                            x2 += row[j + 2];
                            x3 += row[j + 3];
                        }
                    }
                }
                ws[tid] = tau * ((x0 + x1) + (x2 + x3));
            }
            __syncthreads();
#pragma unroll
            for (int r = 0; r < N; r += 2) {
                const double2 x2 = *reinterpret_cast<const double2*>(ws + r);
                a[r] = fma(-x2.x, vj, a[r]);
                a[r + 1] = fma(-x2.y, vj, a[r + 1]);
            }
            // the next step's first barrier orders the reads of vs / ws above against its writes
        }
        __syncthreads();
        // ---- H out: upper Hessenberg part, zeros below ----
        if (mine) {
#pragma unroll
            for (int r = 0; r < N; ++r)
                if (r < n) gA[r + (size_t)tid * p.lda] = (r <= tid + 1) ? a[r] : 0.0;
        }
        if (tid == 0 && p.scratch) {
            double* sc = p.scratch + 8 * b;
            sc[0] = scaled ? 1.0 : 0.0;
            sc[1] = cscale;
            sc[2] = 0.0;
            sc[3] = anrm;
            sc[4] = 0.0;
        }
        if (wantZ) {
            // ---- _materializeQ (src/hessenberg.jl:150-166): Q = H_1 H_2 ... H_{n-1}, accumulated backwards on the identity.
            //      Reflector tails (with the leading 1, zeros above) are parked column-wise in shared memory. ----
            double* V = C;      // V[c * N + r]: reflector c (from column c), entry r
#pragma unroll
            for (int r = 0; r < N; ++r) {
                if (tid < N) V[tid * N + r] = (r <= tid) ? 0.0 : (r == tid + 1 ? 1.0 : a[r]);
            }
            __syncthreads();
            double q[N];
#pragma unroll
            for (int r = 0; r < N; ++r) q[r] = (r == tid) ? 1.0 : 0.0;
            for (int c = n - 2; c >= 0; --c) {
                const double tau = taus[c];
                const double* v = V + c * N;
                const int h = c + 1;
                double w0 = 0.0, w1 = 0.0, w2 = 0.0, w3 = 0.0;
#pragma unroll
                for (int blk = 0; blk < N / 8; ++blk) {
                    if (8 * blk + 7 >= h) {
#pragma unroll
                        for (int r = 8 * blk; r < 8 * blk + 8; r += 4) {
                            const double2 va = *reinterpret_cast<const double2*>(v + r);
                            const double2 vb = *reinterpret_cast<const double2*>(v + r + 2);
                            w0 = fma(va.x, q[r], w0);
                            w1 = fma(va.y, q[r + 1], w1);
                            w2 = fma(vb.x, q[r + 2], w2);
                            w3 = fma(vb.y, q[r + 3], w3);
                        }
                    }
                }
                const double tw = tau * ((w0 + w1) + (w2 + w3));
#pragma unroll
                for (int blk = 0; blk < N / 8; ++blk) {
                    if (8 * blk + 7 >= h) {
#pragma unroll
                        for (int r = 8 * blk; r < 8 * blk + 8; r += 2) {
                            const double2 v2 = *reinterpret_cast<const double2*>(v + r);
                            q[r] = fma(-tw, v2.x, q[r]);
                            q[r + 1] = fma(-tw, v2.y, q[r + 1]);
                        }
                    }
                }
            }
            if (mine) {
                double* gZ = reinterpret_cast<double*>(p.Z) + b * p.strideZ;
#pragma unroll
                for (int r = 0; r < N; ++r)
                    if (r < n) gZ[r + (size_t)tid * p.ldz] = q[r];
            }
        }
        __syncthreads();
    }
}

template <int N> int launch_gehrd_reg(const BatchedParams& p, int dev_sms, cudaStream_t stream, std::string* err) {
    auto kern = gehrd_reg_kernel<N>;
    const size_t smem = gehrd_reg_layout<N>::bytes();
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int per_sm = 0;
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, N, smem);
    if (e != cudaSuccess) {
        *err = std::string("gehrd (register) kernel setup: ") + cudaGetErrorString(e);
        return -2;
    }
    if (per_sm < 1) {
        *err = "gehrd (register) kernel does not fit on an SM";
        return -3;
    }
    long long grid = (long long)per_sm * dev_sms;
    if (grid > p.batch) grid = p.batch;
    kern<<<(unsigned)grid, N, smem, stream>>>(p);
    note_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) {
        *err = std::string("gehrd (register) kernel launch: ") + cudaGetErrorString(e);
        return -2;
    }
    return 0;
}

}  // namespace gs

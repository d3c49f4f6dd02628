// Stage A for Float64, n <= 64, third generation: scale -> Householder Hessenberg reduction -> explicit Q with THE
// MATRIX IN THE REGISTERS OF ONE WARP, tiled in two dimensions.
//
// The thread-per-column kernel of gehrd.cuh keeps the tile in shared memory and walks it with dependent scalar loops:
// ncu (profiles/r02_stageA_f64.summary.txt) shows 192 k warp instructions per matrix for 19 k warp-FMAs of useful work,
// 2.8 short-scoreboard stall cycles per issue and an FP64 pipe at 12 %.  A first register version (one COLUMN per
// thread, v broadcast from shared memory) was no faster: 128 broadcast loads and a 64 x 64 transposition through shared
// memory per step made it shared-memory-bandwidth bound.  Here thread t owns the (N/8) x 8 block (rows (N/8) (t % 8) ..,
// columns 8 (t / 8) ..) of the N x N matrix — 8 x 8 entries and two warps for N = 64, 4 x 8 and one warp for N = 32 —
//     left update   w_j = v' a_j   = N/8 FMAs per owned column + an xor-butterfly over the 8 lanes of a block column,
//     right update  x_r = a_r' v   = 8 FMAs per owned row + a butterfly over the 4 block columns of a warp (and, for
//                                    N = 64, one exchange of the two warps' partial sums through shared memory),
// both rank-1 updates are straight-line FMAs on statically indexed registers, and a lane reads only ITS 8 + 16 entries of
// v from shared memory.  Rows above the reflector are handled by zeros in v, columns left of it by tau = 0 (no
// predicates).  The reflector (src/householder.jl:12-54, with the sub-sfmin rescaling loop) is formed by the whole warp:
// the 8 lanes that own the column extract it, the norm is a warp reduction.  Q is accumulated backwards in registers
// with the reflector tails parked in shared memory (src/hessenberg.jl:150-166 computes the same product applied to
// the identity).  Same outputs as gehrd_q_kernel: A <- H (zeros below the sub-diagonal), Z <- Q,
// scratch <- (scaled?, cscale, anrm).
#pragma once
#include "gehrd.cuh"

namespace gs {

template <int N> struct gehrd_reg_layout {
    static constexpr int NT = (N == 64) ? 64 : 32;                                                  // threads per matrix
    __host__ __device__ static constexpr size_t off_v() { return (size_t)N * N * sizeof(double); }   // after the parked tails
    __host__ __device__ static constexpr size_t off_tau() { return off_v() + N * sizeof(double); }
    __host__ __device__ static constexpr size_t off_xs() { return off_tau() + N * sizeof(double); }  // 2 parities x 2 warps x N
    __host__ __device__ static constexpr size_t bytes() { return off_xs() + 4 * N * sizeof(double) + 16; }
};

GS_DEV double gr_wsum(double v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}
GS_DEV double gr_wmax(double v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, m));
    return v;
}

template <int N> __global__ void __launch_bounds__((N == 64 ? 64 : 32), (N == 64 ? 4 : 12)) gehrd_reg_kernel(BatchedParams p) {
    typedef gehrd_reg_layout<N> GL;
    constexpr int NT = GL::NT;
    constexpr int RB = N / 8, CB = 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* V = reinterpret_cast<double*>(smem_raw);                       // parked reflectors: V[c * N + r]
    double* vs = reinterpret_cast<double*>(smem_raw + GL::off_v());
    double* taus = reinterpret_cast<double*>(smem_raw + GL::off_tau());
    double* xs = reinterpret_cast<double*>(smem_raw + GL::off_xs());
    __shared__ long long s_next;
    const int n = p.n, lane = threadIdx.x, wid = threadIdx.x >> 5;
    const int bi = lane & 7, bj = lane >> 3;
    const int r0 = RB * bi, c0 = CB * bj;
    const bool wantZ = p.Z != nullptr;

    for (;;) {
        if (lane == 0) s_next = (long long)atomicAdd(p.counter, 1ULL);
        __syncthreads();
        const long long b = s_next;
        __syncthreads();
        if (b >= p.batch) break;
        double* gA = reinterpret_cast<double*>(p.A) + b * p.strideA;
        double a[RB][CB];
#pragma unroll
        for (int lc = 0; lc < CB; ++lc)
#pragma unroll
            for (int lr = 0; lr < RB; ++lr)
                a[lr][lc] = (r0 + lr < n && c0 + lc < n) ? gA[(r0 + lr) + (size_t)(c0 + lc) * p.lda] : 0.0;

        // ---- _scale! (src/util.jl:14-29) ----
        bool scaled = false;
        double cscale = 1.0, anrm = 1.0;
        if (p.scale) {
            double m = 0.0;
#pragma unroll
            for (int lc = 0; lc < CB; ++lc)
#pragma unroll
                for (int lr = 0; lr < RB; ++lr) m = fmax(m, fabs(a[lr][lc]));
            anrm = gr_wmax(m);
            if constexpr (NT == 64) {
                if ((lane & 31) == 0) xs[wid] = anrm;
                __syncthreads();
                anrm = fmax(xs[0], xs[1]);
                __syncthreads();
            }
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
                safescale_apply<double, double, NT>(anrm, cscale, [&](double mul) {
#pragma unroll
                    for (int lc = 0; lc < CB; ++lc)
#pragma unroll
                        for (int lr = 0; lr < RB; ++lr) a[lr][lc] *= mul;
                });
            }
        }
        // ---- _hessenberg! (src/hessenberg.jl:3-17): step c = reflector from column c, rows c+1 .. n-1 ----
        for (int c = 0; c < n - 1; ++c) {
            const int h = c + 1;
            const int jb = c / CB, lc0 = c % CB;
            const bool owner = bj == jb;
            const bool owner_warp = (jb >> 2) == wid;        // four block columns per warp
            double tau = 0.0;
            if (owner_warp) {
            // the owning lanes extract their rows of column c
            double x[RB];
#pragma unroll
            for (int lr = 0; lr < RB; ++lr) {
                double t = 0.0;
#pragma unroll
                for (int lc = 0; lc < CB; ++lc) t = (lc == lc0) ? a[lr][lc] : t;
                x[lr] = owner ? t : 0.0;
            }
            // alpha = A[h, c], the largest modulus and the (unscaled) sum of squares of the tail: ONE butterfly over the 8
            // lanes of the owning block column for the three values, then a broadcast to the warp
            double alpha = 0.0, amax = 0.0, ssq = 0.0;
#pragma unroll
            for (int lr = 0; lr < RB; ++lr) {
                alpha = (r0 + lr == h) ? x[lr] : alpha;
                x[lr] = (r0 + lr > h) ? x[lr] : 0.0;      // the tail only
                amax = fmax(amax, fabs(x[lr]));
                ssq = fma(x[lr], x[lr], ssq);
            }
#pragma unroll
            for (int m = 1; m <= 4; m <<= 1) {
                alpha += __shfl_xor_sync(0xffffffffu, alpha, m);
                amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, m));
                ssq += __shfl_xor_sync(0xffffffffu, ssq, m);
            }
            {
                const int src = 8 * (jb & 3);
                alpha = __shfl_sync(0xffffffffu, alpha, src);
                amax = __shfl_sync(0xffffffffu, amax, src);
                ssq = __shfl_sync(0xffffffffu, ssq, src);
            }
            double beta = alpha, scal = 0.0;
            const bool trivial = (n - 1 - c) <= 1;       // a real length-1 reflector is the identity
            double xnorm = 0.0;
            if (!trivial && amax > 0.0) {
                if (q_exp_in(ssq, 1023u - 900u, 1023u + 900u)) {
                    xnorm = fast_sqrt(ssq);               // the squares neither overflowed nor lost bits to underflow
                } else {
                    // badly scaled column: the scaled two-pass form (src/util.jl:506-557)
                    const double rs = 1.0 / amax;
                    double q2 = 0.0;
#pragma unroll
                    for (int lr = 0; lr < RB; ++lr) {
                        const double t = x[lr] * rs;
                        q2 = fma(t, t, q2);
                    }
                    xnorm = amax * r_sqrt(gr_wsum(q2));
                }
            }
            const bool have = !trivial && xnorm != 0.0;
            if (have) {
                beta = -copysign(r_hypot4(alpha, 0.0, xnorm, 0.0), alpha);
                const double sfmin = 2.0 * rtraits<double>::floatmin() / rtraits<double>::eps();
                int kount = 0;
                double al = alpha;
                if (fabs(beta) < sfmin) {
                    const double rsfmin = 1.0 / sfmin;
                    bool smallb = true;
                    while (smallb) {
                        kount += 1;
#pragma unroll
                        for (int lr = 0; lr < RB; ++lr) x[lr] *= rsfmin;
                        beta *= rsfmin;
                        al *= rsfmin;
                        smallb = (fabs(beta) < sfmin) && (kount < 20);
                    }
                    double am2 = 0.0;
#pragma unroll
                    for (int lr = 0; lr < RB; ++lr) am2 = fmax(am2, fabs(x[lr]));
                    am2 = gr_wmax(am2);
                    xnorm = 0.0;
                    if (am2 > 0.0) {
                        const double rs2 = 1.0 / am2;
                        double ssq = 0.0;
#pragma unroll
                        for (int lr = 0; lr < RB; ++lr) {
                            const double t = x[lr] * rs2;
                            ssq = fma(t, t, ssq);
                        }
                        xnorm = am2 * r_sqrt(gr_wsum(ssq));
                    }
                    beta = -copysign(r_hypot4(al, 0.0, xnorm, 0.0), al);
                }
                tau = (beta - al) * q_rcp(beta);
                scal = q_rcp(al - beta);
                for (int q = 0; q < kount; ++q) beta *= sfmin;
#pragma unroll
                for (int lr = 0; lr < RB; ++lr) x[lr] *= scal;
            }
            // publish v (zeros above the head, 1 at the head) and park it for the Q phase; the owners keep the tail
            // and beta in their column
            if (owner) {
#pragma unroll
                for (int lr = 0; lr < RB; ++lr) {
                    const int r = r0 + lr;
                    const double vr = (r < h) ? 0.0 : (r == h ? 1.0 : x[lr]);
                    vs[r] = vr;
                    V[c * N + r] = vr;
                }
                if (have) {
#pragma unroll
                    for (int lr = 0; lr < RB; ++lr) {
                        const int r = r0 + lr;
#pragma unroll
                        for (int lc = 0; lc < CB; ++lc)
                            if (lc == lc0) a[lr][lc] = (r > h) ? x[lr] : (r == h ? beta : a[lr][lc]);
                    }
                }
            }
            if ((lane & 31) == 0) taus[c] = tau;
            }   // owner warp
            if constexpr (NT == 64) __syncthreads();
            else __syncwarp();
            tau = taus[c];
            double vr[RB], vc[CB];
#pragma unroll
            for (int lr = 0; lr < RB; lr += 2) {
                const double2 t = *reinterpret_cast<const double2*>(vs + r0 + lr);
                vr[lr] = t.x;
                vr[lr + 1] = t.y;
            }
#pragma unroll
            for (int lc = 0; lc < CB; lc += 2) {
                const double2 t = *reinterpret_cast<const double2*>(vs + c0 + lc);
                vc[lc] = t.x;
                vc[lc + 1] = t.y;
            }
            // (vs is rewritten by the next step's owner warp only after the exchange barrier below / the warp barrier)
            if constexpr (NT == 32) __syncwarp();
            // ---- lmul!(H', A[c+1:, c+1:]): w_j = v' a_j, a_j -= tau w_j v (columns j > c) ----
#pragma unroll
            for (int lc = 0; lc < CB; ++lc) {
                double w = 0.0;
#pragma unroll
                for (int lr = 0; lr < RB; ++lr) w = fma(vr[lr], a[lr][lc], w);
                w += __shfl_xor_sync(0xffffffffu, w, 1);
                w += __shfl_xor_sync(0xffffffffu, w, 2);
                w += __shfl_xor_sync(0xffffffffu, w, 4);
                const double tw = (c0 + lc > c) ? tau * w : 0.0;
#pragma unroll
                for (int lr = 0; lr < RB; ++lr) a[lr][lc] = fma(-tw, vr[lr], a[lr][lc]);
            }
            // ---- rmul!(A[:, c+1:], H): x_r = a_r' v, a_r -= tau x_r v' (v_j = 0 for j <= c) ----
            double xr[RB];
#pragma unroll
            for (int lr = 0; lr < RB; ++lr) {
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int lc = 0; lc < CB; lc += 2) {
                    s0 = fma(a[lr][lc], vc[lc], s0);
                    s1 = fma(a[lr][lc + 1], vc[lc + 1], s1);
                }
                double s = s0 + s1;
                s += __shfl_xor_sync(0xffffffffu, s, 8);
                s += __shfl_xor_sync(0xffffffffu, s, 16);
                xr[lr] = s;
            }
            if constexpr (NT == 64) {
                // the two warps hold the partial sums of four block columns each: exchange through shared memory
                double* mine = xs + ((c & 1) * 2 + wid) * N;
                const double* other = xs + ((c & 1) * 2 + (1 - wid)) * N;
                if ((lane & 24) == 0) {
#pragma unroll
                    for (int lr = 0; lr < RB; lr += 2) *reinterpret_cast<double2*>(mine + r0 + lr) = make_double2(xr[lr], xr[lr + 1]);
                }
                __syncthreads();
#pragma unroll
                for (int lr = 0; lr < RB; lr += 2) {
                    const double2 t = *reinterpret_cast<const double2*>(other + r0 + lr);
                    xr[lr] += t.x;
                    xr[lr + 1] += t.y;
                }
            }
#pragma unroll
            for (int lr = 0; lr < RB; ++lr) {
                const double ts = tau * xr[lr];
#pragma unroll
                for (int lc = 0; lc < CB; ++lc) a[lr][lc] = fma(-ts, vc[lc], a[lr][lc]);
            }
        }
        // ---- H out: upper Hessenberg part, zeros below ----
#pragma unroll
        for (int lc = 0; lc < CB; ++lc)
#pragma unroll
            for (int lr = 0; lr < RB; ++lr) {
                const int r = r0 + lr, j = c0 + lc;
                if (r < n && j < n) gA[r + (size_t)j * p.lda] = (r <= j + 1) ? a[lr][lc] : 0.0;
            }
        if (threadIdx.x == 0 && p.scratch) {
            double* sc = p.scratch + 8 * b;
            sc[0] = scaled ? 1.0 : 0.0;
            sc[1] = cscale;
            sc[2] = 0.0;
            sc[3] = anrm;
            sc[4] = 0.0;
        }
        if (wantZ) {
            // ---- _materializeQ (src/hessenberg.jl:150-166): Q = H_1 H_2 ... H_{n-1}, accumulated backwards on the identity ----
            __syncthreads();
#pragma unroll
            for (int lc = 0; lc < CB; ++lc)
#pragma unroll
                for (int lr = 0; lr < RB; ++lr) a[lr][lc] = (r0 + lr == c0 + lc) ? 1.0 : 0.0;
            for (int c = n - 2; c >= 0; --c) {
                const double tau = taus[c];
                double vr[RB];
#pragma unroll
                for (int lr = 0; lr < RB; lr += 2) {
                    const double2 t = *reinterpret_cast<const double2*>(V + c * N + r0 + lr);
                    vr[lr] = t.x;
                    vr[lr + 1] = t.y;
                }
#pragma unroll
                for (int lc = 0; lc < CB; ++lc) {
                    double w = 0.0;
#pragma unroll
                    for (int lr = 0; lr < RB; ++lr) w = fma(vr[lr], a[lr][lc], w);
                    w += __shfl_xor_sync(0xffffffffu, w, 1);
                    w += __shfl_xor_sync(0xffffffffu, w, 2);
                    w += __shfl_xor_sync(0xffffffffu, w, 4);
                    const double tw = tau * w;
#pragma unroll
                    for (int lr = 0; lr < RB; ++lr) a[lr][lc] = fma(-tw, vr[lr], a[lr][lc]);
                }
            }
            double* gZ = reinterpret_cast<double*>(p.Z) + b * p.strideZ;
#pragma unroll
            for (int lc = 0; lc < CB; ++lc)
#pragma unroll
                for (int lr = 0; lr < RB; ++lr) {
                    const int r = r0 + lr, j = c0 + lc;
                    if (r < n && j < n) gZ[r + (size_t)j * p.ldz] = a[lr][lc];
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
    constexpr int NT = gehrd_reg_layout<N>::NT;
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem);
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
    kern<<<(unsigned)grid, NT, smem, stream>>>(p);
    note_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) {
        *err = std::string("gehrd (register) kernel launch: ") + cudaGetErrorString(e);
        return -2;
    }
    return 0;
}

}  // namespace gs

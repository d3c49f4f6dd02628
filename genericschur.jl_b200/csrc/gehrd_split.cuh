// Stage A for Float64 / ComplexF64 matrices with n <= 64, second generation: the same computation as
// gehrd_q_kernel (scale -> Householder Hessenberg reduction -> explicit Q, src/hessenberg.jl:3-17, 150-166,
// src/householder.jl:12-102, 140-172) with the work of one reflector application split over G lanes per
// column / row instead of one thread per column.
//
// Why: with one thread per column every dot product v^H a_j is a serial chain of up to n dependent complex FMAs
// and a 64x64 matrix keeps 2 warps busy; three matrices per SM (shared-memory bound) are 6 warps per SM and
// the kernel ran at 10 % FP64-pipe utilisation, all stalls short-scoreboard / fixed-latency.  Here a CTA has
// 4 n threads; G = 8 (Float64) or 16 (ComplexF64) lanes share a column: each lane loads n/G entries ONCE into
// registers, the partial dot products meet in a log2(G)-step shuffle butterfly, and the rank-1 update is applied
// to the register copy and stored — 1 load + 1 store per entry and reflector, chains G times shorter, 4x the
// warps.  Each lane also keeps its share of the reflector tail in registers for all column passes of one application
// (the first version re-read it from shared memory per entry and was shared-memory-bandwidth bound: 80 % of the
// LSU wavefront peak in ncu, half of it the tail).  G grows as the trailing matrix shrinks so the lanes stay busy.  Leading dimensions are chosen so that
// both access patterns (lanes along a column, lanes across a row) are bank-conflict free: n | 1 for 16-byte
// elements, 4 mod 16 for 8-byte elements (which also makes Float64 columns 16-byte aligned: TMA for both kinds).
//
// The reflector itself (scaled 2-norm, beta, tau, 1/(alpha - beta)) is computed REDUNDANTLY by every warp with
// shuffle reductions — no block-wide reduction, no barrier — and the scaled tail is published through a small
// vector in shared memory: three barriers per column instead of eleven.
// Q is accumulated backwards in place without the xORGHR column shift: reflector i lives in column i, the Q block
// it acts on in columns i+1.., so nothing has to move.
#pragma once
#include <cstdlib>
#include <type_traits>
#include "gehrd.cuh"
#include "gehrd_reg.cuh"

namespace gs {

GS_DEV double shfl_xor_e(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
GS_DEV cx<double> shfl_xor_e(const cx<double>& v, int m) {
    return mk_cx<double>(__shfl_xor_sync(0xffffffffu, v.re, m), __shfl_xor_sync(0xffffffffu, v.im, m));
}

template <class T> struct split_g0;
template <> struct split_g0<double> { static constexpr int value = 8; };
template <> struct split_g0<cx<double>> { static constexpr int value = 16; };

template <class T, int NMAX> struct GehrdSplit {
    typedef typename etraits<T>::real R;
    typedef smem_layout<T> L;
    static constexpr bool CPLX = etraits<T>::is_complex;
    static constexpr int NT = 4 * NMAX;
    static constexpr int G0 = split_g0<T>::value;
    static constexpr int SMAX = NMAX / 32;   // reflector tail entries per lane
    __host__ __device__ static int ld(int n) { return CPLX ? (n | 1) : ((n + 11) / 16) * 16 + 4; }
    // [tile n*ld T][v: NMAX T][tau: NMAX T][red: 32 R][mbar]
    __host__ __device__ static size_t off_v(int n) { return L::up16((size_t)n * ld(n) * sizeof(T)); }
    __host__ __device__ static size_t off_tau(int n) { return off_v(n) + (size_t)NMAX * sizeof(T); }
    __host__ __device__ static size_t off_red(int n) { return off_tau(n) + (size_t)NMAX * sizeof(T); }
    __host__ __device__ static size_t off_mbar(int n) { return off_red(n) + L::up16(32 * sizeof(R)); }
    __host__ __device__ static size_t bytes(int n) { return off_mbar(n) + 16; }

    // (I - tl v v^H) from the left on columns j0..j1 (1-based); v = (1, vt[0..nv)) acts on rows r0, r0+1..r0+nv.
    // HEADZERO: row r0 is logically zero on entry (Q accumulation writes it for the first time).
    template <int G, bool HEADZERO>
    GS_DEV static void left_apply(T* H, int ld, const T* vt, const T tl, int r0, int nv, int j0, int j1, int tid) {
        constexpr int NC = NT / G, K = NMAX / G;
        const int q = tid % G, c = tid / G;
        T v[K];   // this lane's share of the reflector tail: loaded once, used for every column pass
#pragma unroll
        for (int t = 0; t < K; ++t) {
            const int e = q + G * t;
            v[t] = e_zero<T>();
            if (e < nv) v[t] = vt[e];
        }
        for (int jb = j0; jb <= j1; jb += NC) {
            const int j = jb + c;
            const bool on = j <= j1;
            // Lanes past the last column read column j0 and store nothing: while another group rewrites that column in the same
            // pass the values they get are undefined and unused (compute-sanitizer's racecheck reports these reads; loading
            // under the `on` predicate instead was measured 6.6 % slower for the whole kernel).
            T* col = H + (size_t)((on ? j : j0) - 1) * ld + r0;   // col[e] = row r0+1+e, col[-1] = row r0
            T a[K];
            T acc = e_zero<T>();
            T head = e_zero<T>();
            if (!HEADZERO) head = col[-1];
#pragma unroll
            for (int t = 0; t < K; ++t) {
                const int e = q + G * t;
                a[t] = e_zero<T>();
                if (e < nv) a[t] = col[e];
                acc = e_fma_cja(v[t], a[t], acc);
            }
#pragma unroll
            for (int m = G / 2; m >= 1; m >>= 1) acc = acc + shfl_xor_e(acc, m);
            const T va = tl * (head + acc);
            if (on) {
                if (q == 0) col[-1] = head - va;
#pragma unroll
                for (int t = 0; t < K; ++t) {
                    const int e = q + G * t;
                    if (e < nv) col[e] = e_fnma(va, v[t], a[t]);
                }
            }
        }
    }
    // (I - tau v v^H) from the right on rows 1..n: columns c0 (head), c0+1..c0+nv
    template <int G>
    GS_DEV static void right_apply(T* H, int ld, const T* vt, const T tau, int c0, int nv, int n, int tid) {
        constexpr int NR = NT / G, K = NMAX / G;
        const int q = tid % G, c = tid / G;
        T v[K];
#pragma unroll
        for (int t = 0; t < K; ++t) {
            const int e = q + G * t;
            v[t] = e_zero<T>();
            if (e < nv) v[t] = vt[e];
        }
        for (int rb = 1; rb <= n; rb += NR) {
            const int r = rb + c;
            const bool on = r <= n;
            T* row = H + ((on ? r : 1) - 1) + (size_t)c0 * ld;   // row[e ld] = H(r, c0+1+e), row[-ld] = H(r, c0)
            T a[K];
            T acc = e_zero<T>();
            const T head = row[-ld];
#pragma unroll
            for (int t = 0; t < K; ++t) {
                const int e = q + G * t;
                a[t] = e_zero<T>();
                if (e < nv) a[t] = row[(size_t)e * ld];
                acc = e_fma(a[t], v[t], acc);
            }
#pragma unroll
            for (int m = G / 2; m >= 1; m >>= 1) acc = acc + shfl_xor_e(acc, m);
            const T tx = tau * (head + acc);
            if (on) {
                if (q == 0) row[-ld] = head - tx;
#pragma unroll
                for (int t = 0; t < K; ++t) {
                    const int e = q + G * t;
                    if (e < nv) row[(size_t)e * ld] = e_fnma_cjb(tx, v[t], a[t]);
                }
            }
        }
    }
    template <bool HEADZERO>
    GS_DEV static void left_apply_auto(T* H, int ld, const T* vt, const T tl, int r0, int nv, int j0, int j1, int tid) {
        const int ncols = j1 - j0 + 1;
        if (ncols <= 0) return;
        constexpr int G1 = G0 * 2 <= 32 ? G0 * 2 : 32, G2 = G0 * 4 <= 32 ? G0 * 4 : 32;
        if (G2 > G1 && ncols * G2 <= NT) left_apply<G2, HEADZERO>(H, ld, vt, tl, r0, nv, j0, j1, tid);
        else if (G1 > G0 && ncols * G1 <= NT) left_apply<G1, HEADZERO>(H, ld, vt, tl, r0, nv, j0, j1, tid);
        else left_apply<G0, HEADZERO>(H, ld, vt, tl, r0, nv, j0, j1, tid);
    }

    GS_DEV static double abs_q(double a) { return fabs(a); }
    GS_DEV static double abs_q(const cx<double>& a) { return c_abs_q(a); }
    GS_DEV static R warp_max(R v) {
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) v = r_max(v, shfl_xor(v, m));
        return v;
    }
    GS_DEV static R warp_sum(R v) {
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) v = v + shfl_xor(v, m);
        return v;
    }
    GS_DEV static R tail_norm(const T (&x)[SMAX]) {
        const R zero = r_const<R>(0.0), one = r_const<R>(1.0);
        R amax = zero;
#pragma unroll
        for (int s = 0; s < SMAX; ++s) amax = r_max(amax, e_maxpart(x[s]));
        amax = warp_max(amax);
        if (!(amax > zero)) return zero;
        const R rs = q_rcp(amax);
        R ssq = zero;
#pragma unroll
        for (int s = 0; s < SMAX; ++s) ssq = ssq + e_sq_scaled(x[s], rs);
        ssq = warp_sum(ssq);
        return amax * q_sqrt(ssq);
    }
    // sqrt(a^2 + b^2 + c^2) scaled by the largest magnitude (src/util.jl:562-570) with the guarded fast reciprocal / sqrt
    GS_DEV static R hypot3_q(const R& a, const R& b, const R& c) {
        const R aa = r_abs(a), ab = r_abs(b), ac = r_abs(c);
        const R w = r_max(r_max(aa, ab), ac);
        if (w == r_const<R>(0.0) || r_isnan(w)) return w + aa + ab + ac;
        const R rw = q_rcp(w);
        const R x = aa * rw, y = ab * rw, z = ac * rw;
        return w * q_sqrt(x * x + y * y + z * z);
    }
    // _reflector!(view(A, i+1:n, i)), src/householder.jl:12-102, computed by ONE warp (shuffle reductions, no block
    // barrier); divisions and square roots go through the guarded MUFU + Newton primitives of scalar.cuh (IEEE
    // fallbacks outside the safe exponent window).  colp = &A(i+1, i); x[] receives the scaled tail.
    // Returns false when H = I (tau = 0, nothing to write).
    GS_DEV static bool reflector(const T* colp, int nv, int lane, T& tau, R& beta, T (&x)[SMAX]) {
        const R zero = r_const<R>(0.0), one = r_const<R>(1.0);
        tau = e_zero<T>();
        beta = zero;
        if (!CPLX && nv <= 0) return false;   // a real length-1 reflector is the identity
        const T alpha = colp[0];
#pragma unroll
        for (int s = 0; s < SMAX; ++s) {
            const int e = lane + 32 * s;
            x[s] = e_zero<T>();
            if (e < nv) x[s] = colp[1 + e];
        }
        R xnorm = tail_norm(x);
        R ar, ai;
        if constexpr (CPLX) {
            ar = alpha.re;
            ai = alpha.im;
        } else {
            ar = alpha;
            ai = zero;
        }
        if (CPLX ? (xnorm == zero && ai == zero) : (xnorm == zero)) return false;
        beta = -r_copysign(hypot3_q(ar, ai, xnorm), ar);
        const R sfmin = CPLX ? rtraits<R>::floatmin() / rtraits<R>::eps()
                             : r_const<R>(2.0) * rtraits<R>::floatmin() / rtraits<R>::eps();
        int kount = 0;
        if (r_abs(beta) < sfmin) {
            const R rsfmin = one / sfmin;
            bool smallb = true;
            while (smallb) {
                kount += 1;
#pragma unroll
                for (int s = 0; s < SMAX; ++s) x[s] = e_scale(x[s], rsfmin);
                beta = beta * rsfmin;
                ar = ar * rsfmin;
                ai = ai * rsfmin;
                smallb = (r_abs(beta) < sfmin) && (kount < 20);
            }
            xnorm = tail_norm(x);
            beta = -r_copysign(hypot3_q(ar, ai, xnorm), ar);
        }
        T tscal;
        const R rbeta = q_rcp(beta);
        if constexpr (CPLX) {
            tau = mk_cx<R>((beta - ar) * rbeta, -ai * rbeta);
            tscal = c_div_q(mk_cx<R>(one, zero), mk_cx<R>(ar - beta, ai));
        } else {
            tau = (beta - ar) * rbeta;
            tscal = q_rcp(ar - beta);
        }
#pragma unroll
        for (int s = 0; s < SMAX; ++s) x[s] = x[s] * tscal;
        for (int j = 0; j < kount; ++j) beta = beta * sfmin;
        return true;
    }
};

template <class T, int NMAX>
__global__ void __launch_bounds__(4 * NMAX, 768 / (4 * NMAX)) gehrd_q_split_kernel(BatchedParams p) {
    typedef typename etraits<T>::real R;
    typedef GehrdSplit<T, NMAX> GS_;
    constexpr bool CPLX = etraits<T>::is_complex;
    constexpr int NT = GS_::NT, SMAX = GS_::SMAX, G0 = GS_::G0;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = p.n;
    const int ld = GS_::ld(n);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    T* H = reinterpret_cast<T*>(smem_raw);
    T* sV = reinterpret_cast<T*>(smem_raw + GS_::off_v(n));
    T* sTau = reinterpret_cast<T*>(smem_raw + GS_::off_tau(n));
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw + GS_::off_mbar(n));
    BatchedSolver<T, NT> S;   // for scale_in only
    S.n = n;
    S.ld = ld;
    S.tid = tid;
    S.lane = lane;
    S.H = H;
    S.Z = nullptr;
    S.sTau = sTau;
    S.sW = nullptr;
    S.sRed = reinterpret_cast<R*>(smem_raw + GS_::off_red(n));
    __shared__ long long s_next;
    __shared__ int s_did;
    const bool wantZ = (p.Z != nullptr);
#define AA(i, j) H[((i)-1) + (size_t)((j)-1) * ld]

    if (tid == 0) mbar_init(mbar, 1);
    __syncthreads();
    uint32_t parity = 0;
    const uint32_t col_bytes = (uint32_t)(n * sizeof(T));

    for (;;) {
        if (tid == 0) s_next = (long long)atomicAdd(p.counter, 1ULL);
        __syncthreads();
        const long long b = s_next;
        __syncthreads();
        if (b >= p.batch) break;
        T* gA = reinterpret_cast<T*>(p.A) + b * p.strideA;
        T* gZ = wantZ ? reinterpret_cast<T*>(p.Z) + b * p.strideZ : nullptr;

        // ---- stage the tile: one TMA bulk copy per column when columns are 16-byte aligned multiples of 16 bytes ----
        const bool use_tma = (col_bytes % 16 == 0) && ((reinterpret_cast<uintptr_t>(gA) & 15) == 0) &&
                             (((size_t)p.lda * sizeof(T)) % 16 == 0) && (((size_t)ld * sizeof(T)) % 16 == 0);
        if (use_tma) {
            if (tid == 0) {
                fence_proxy_async();
                mbar_expect_tx(mbar, col_bytes * (uint32_t)n);
            }
            __syncthreads();
            if (tid < 32)
                for (int j = tid; j < n; j += 32) tma_bulk_g2s(H + (size_t)j * ld, gA + (size_t)j * p.lda, col_bytes, mbar);
            mbar_wait(mbar, parity);
            parity ^= 1;
        } else {
            for (int e = tid; e < n * n; e += NT) {
                int i = e % n, j = e / n;
                H[i + (size_t)j * ld] = gA[i + (size_t)j * p.lda];
            }
        }
        __syncthreads();

        bool scaled = false;
        R cscale = r_const<R>(1.0), anrm = r_const<R>(1.0);
        if (p.scale) {
            // _scale! (src/util.jl:14-29): max |a_ij| with the guarded fast modulus, then the rare rescaling
            const R zero = r_const<R>(0.0);
            R m = zero;
            for (int e = tid; e < n * n; e += NT) {
                int i = e % n, j = e / n;
                m = r_max(m, GS_::abs_q(H[i + (size_t)j * ld]));
            }
            anrm = block_max<R, NT>(m, S.sRed);
            const R smlnum = r_sqrt(r_safemin<R>()) / rtraits<R>::eps();
            const R bignum = r_const<R>(1.0) / smlnum;
            if (anrm > zero && anrm < smlnum) {
                scaled = true;
                cscale = smlnum;
            } else if (anrm > bignum) {
                scaled = true;
                cscale = bignum;
            }
            if (scaled) {
                safescale_apply<T, R, NT>(anrm, cscale, [&](R mul) {
                    for (int e = tid; e < n * n; e += NT) {
                        int i = e % n, j = e / n;
                        H[i + (size_t)j * ld] = e_scale(H[i + (size_t)j * ld], mul);
                    }
                });
            }
            __syncthreads();
        }

        // ---- _hessenberg! ----
        for (int i = 1; i <= n - 1; ++i) {
            const int nv = n - i - 1;
            // the reflector is formed by warp 0 alone; tau and the scaled tail reach the others through shared memory
            T x[SMAX];
            R beta = r_const<R>(0.0);
            if (warp == 0) {
                T tau0;
                const bool did0 = GS_::reflector(&AA(i + 1, i), nv, lane, tau0, beta, x);
                if (did0) {
#pragma unroll
                    for (int s = 0; s < SMAX; ++s) {
                        const int e = lane + 32 * s;
                        if (e < nv) sV[e] = x[s];
                    }
                }
                if (lane == 0) {
                    sTau[i - 1] = tau0;
                    s_did = did0 ? 1 : 0;
                }
            }
            __syncthreads();
            const T tau = sTau[i - 1];
            const bool did = s_did != 0;
            if (did) {
                if (warp == 0) {
#pragma unroll
                    for (int s = 0; s < SMAX; ++s) {
                        const int e = lane + 32 * s;
                        if (e < nv) AA(i + 2 + e, i) = x[s];
                    }
                    if (lane == 0) {
                        if constexpr (CPLX) AA(i + 1, i) = mk_cx<R>(beta, r_const<R>(0.0));
                        else AA(i + 1, i) = beta;
                    }
                }
                // lmul!(H', view(A, i+1:n, i+1:n)); rmul!(view(A, :, i+1:n), H)
                GS_::template left_apply_auto<false>(H, ld, sV, cconj(tau), i + 1, nv, i + 1, n, tid);
                __syncthreads();
                GS_::template right_apply<G0>(H, ld, sV, tau, i + 1, nv, n, tid);
            }
            __syncthreads();
        }
        // ---- H out (upper Hessenberg part, zeros below) ----
        for (int e = tid; e < n * n; e += NT) {
            int i = e % n, j = e / n;
            gA[i + (size_t)j * p.lda] = (i <= j + 1) ? H[i + (size_t)j * ld] : e_zero<T>();
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
            __syncthreads();
            // ---- _materializeQ: Q = H_1 ... H_{n-1} accumulated backwards, in place.  Reflector i has its tail in
            //      A(i+2:n, i); before it is applied the Q block occupies A(i+2:n, i+2:n), row and column i+1 are e_{i+1} ----
            for (int i = n - 1; i >= 1; --i) {
                const int nv = n - i - 1;
                const T taui = sTau[i - 1];
                const T* vt = &AA(i + 2, i);
                GS_::template left_apply_auto<true>(H, ld, vt, taui, i + 1, nv, i + 2, n, tid);
                for (int e = tid; e < nv; e += NT) AA(i + 2 + e, i + 1) = -(taui * vt[e]);
                if (tid == 0) AA(i + 1, i + 1) = e_one<T>() - taui;
                __syncthreads();
            }
            for (int e = tid; e < n; e += NT) {
                AA(1, e + 1) = (e == 0) ? e_one<T>() : e_zero<T>();
                AA(e + 1, 1) = (e == 0) ? e_one<T>() : e_zero<T>();
            }
            __syncthreads();
            for (int e = tid; e < n * n; e += NT) {
                int i = e % n, j = e / n;
                gZ[i + (size_t)j * p.ldz] = H[i + (size_t)j * ld];
            }
        }
        __syncthreads();
    }
#undef AA
}

template <class T, int NMAX> int launch_gehrd_split(const BatchedParams& p, int dev_sms, cudaStream_t stream, std::string* err) {
    auto kern = gehrd_q_split_kernel<T, NMAX>;
    constexpr int NT = 4 * NMAX;
    size_t smem = GehrdSplit<T, NMAX>::bytes(p.n);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int per_sm = 0;
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem);
    if (e != cudaSuccess) {
        *err = std::string("gehrd (split) kernel setup: ") + cudaGetErrorString(e);
        return -2;
    }
    if (per_sm < 1) {
        *err = "gehrd (split) kernel does not fit on an SM";
        return -3;
    }
    long long grid = (long long)per_sm * dev_sms;
    if (grid > p.batch) grid = p.batch;
    kern<<<(unsigned)grid, NT, smem, stream>>>(p);
    note_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) {
        *err = std::string("gehrd (split) kernel launch: ") + cudaGetErrorString(e);
        return -2;
    }
    return 0;
}

// stage A dispatch of the two-kernel path.  Measured on B200 (16384 matrices, device-resident): ComplexF64 64x64
// 20.4 ms (one thread per column) -> 15.2 ms (split); Float64 has a quarter of the flops per entry and the split
// kernel's shuffle / predicate overhead outweighs the shorter chains (64x64: 7 ms vs 10 ms), so Float64 keeps the
// thread-per-column kernel.  GSCHUR_GEHRD=v1|v2 forces one or the other (profiling knob).
template <class T> int launch_stage_a(const BatchedParams& p, int dev_sms, cudaStream_t stream, std::string* err) {
    if constexpr (std::is_same<T, double>::value) {
        // Float64, n <= 64: the register-tiled kernel (gehrd_reg.cuh); GSCHUR_GEHRD=v1 / v2 force the older kernels
        const char* force = std::getenv("GSCHUR_GEHRD");
        if (!(force && force[0] == 'v') && p.mode == MODE_SCHUR) {
            if (p.n <= 32) return launch_gehrd_reg<32>(p, dev_sms, stream, err);
            if (p.n <= 64) return launch_gehrd_reg<64>(p, dev_sms, stream, err);
        }
    }
    if constexpr (std::is_same<T, double>::value || std::is_same<T, cx<double>>::value) {
        const char* force = std::getenv("GSCHUR_GEHRD");
        bool split = std::is_same<T, cx<double>>::value;
        if (force && force[0] == 'v' && force[1] == '1') split = false;
        if (force && force[0] == 'v' && force[1] == '2') split = true;
        if (split) {
            if (p.n <= 32) return launch_gehrd_split<T, 32>(p, dev_sms, stream, err);
            if (p.n <= 64) return launch_gehrd_split<T, 64>(p, dev_sms, stream, err);
        }
    }
    return launch_gehrd<T, 64>(p, dev_sms, stream, err);
}

}  // namespace gs

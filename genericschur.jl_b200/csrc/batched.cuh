// Batched Schur kernel, regime (1): one matrix per CTA, A and Z resident in shared memory for the whole
// decomposition (scale -> Householder Hessenberg -> form Q -> Francis QR sweeps with deflation and Z
// accumulation -> unscale -> store T, Z, w).  One persistent grid pulls matrices from an atomic queue.
//
// The algorithm (shift strategy, deflation criterion, iteration caps, 2x2 standardisation, quirks) is that
// of the reference; each routine cites the reference file:line whose behaviour it reproduces.  The
// implementation is not a translation: scans are warp-parallel ballots, reflector applications are
// thread-per-row / thread-per-column over padded (odd leading dimension => conflict-free) shared memory,
// norms are shuffle reductions, the matrix is staged by a TMA bulk copy.
//
// Indices in this file are 1-based through the HH()/ZZ() accessors so the decision rules read like
// SURVEY.md Appendix A.
#pragma once
#include "launch.h"
#include "scalar.cuh"

namespace gs {

template <int NT> GS_DEV void block_sync() {
    if (NT == 32) __syncwarp();
    else __syncthreads();
}

// ---- TMA bulk copy (global -> shared), mbarrier-signalled --------------------------------------------------
GS_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
GS_DEV void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
GS_DEV void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
GS_DEV void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
GS_DEV void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
GS_DEV void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- block reductions (result replicated in every thread) --------------------------------------------------
template <class R, int NT> GS_DEV R block_max(R v, R* sred) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v = r_max(v, shfl_xor(v, m));
    if (NT > 32) {
        __syncthreads();
        if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = v;
        __syncthreads();
        v = sred[0];
#pragma unroll
        for (int i = 1; i < NT / 32; ++i) v = r_max(v, sred[i]);
    }
    return v;
}
template <class R, int NT> GS_DEV R block_sum(R v, R* sred) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v = v + shfl_xor(v, m);
    if (NT > 32) {
        __syncthreads();
        if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = v;
        __syncthreads();
        v = sred[0];
#pragma unroll
        for (int i = 1; i < NT / 32; ++i) v = v + sred[i];
    }
    return v;
}

// ---- small reflectors used inside the QR sweeps ------------------------------------------------------------
// Real xLARFG on a 2- or 3-vector held in registers (src/householder.jl:12-54).  Returns tau; v0 <- beta,
// v1, v2 <- scaled tail.  The 2-norm of the tail and beta are formed with one scaled hypot.
template <class R> GS_DEV R reflector_real_small_generic(R& v0, R& v1, R& v2, int nr) {
    const R zero = r_const<R>(0.0), one = r_const<R>(1.0);
    if (nr <= 1) return zero;
    if (nr == 2) v2 = zero;
    if (v1 == zero && v2 == zero) return zero;
    R alpha = v0;
    R beta = -r_copysign(r_hypot4(alpha, v1, v2, zero), alpha);
    const R sfmin = r_const<R>(2.0) * rtraits<R>::floatmin() / rtraits<R>::eps();
    int kount = 0;
    if (r_abs(beta) < sfmin) {
        const R rsfmin = one / sfmin;
        bool smallb = true;
        while (smallb) {
            kount += 1;
            v1 = v1 * rsfmin;
            v2 = v2 * rsfmin;
            beta = beta * rsfmin;
            alpha = alpha * rsfmin;
            smallb = (r_abs(beta) < sfmin) && (kount < 20);
        }
        beta = -r_copysign(r_hypot4(alpha, v1, v2, zero), alpha);
    }
    R tau = (beta - alpha) / beta;
    R t = one / (alpha - beta);
    v1 = v1 * t;
    v2 = v2 * t;
    for (int j = 0; j < kount; ++j) beta = beta * sfmin;
    v0 = beta;
    return tau;
}

// Double-double fast path: the same quantities as the generic routine (beta = -sign(alpha) ||x||, tau = (beta - alpha) /
// beta, tail / (alpha - beta)) from one reciprocal square root and one reciprocal (scalar.cuh: dd_rsqrt_fast, dd_rcp_fast)
// instead of a scaled hypot and two divisions; out-of-range input takes the generic routine.
GS_DEV dd_t reflector_real_small(dd_t& v0, dd_t& v1, dd_t& v2, int nr) {
    const dd_t zero = mk_dd(0.0);
    if (nr <= 1) return zero;
    if (nr == 2) v2 = zero;
    if (v1 == zero && v2 == zero) return zero;
    const double w = fmax(fabs(v0.hi), fmax(fabs(v1.hi), fabs(v2.hi)));
    if (!(w >= 1e-100 && w <= 1e100)) return reflector_real_small_generic<dd_t>(v0, v1, v2, nr);
    const dd_t alpha = v0;
    const dd_t q = alpha * alpha + (v1 * v1 + v2 * v2);
    const dd_t rs = dd_rsqrt_fast(q);
    const dd_t nrm = q * rs;
    const bool neg = signbit(alpha.hi);
    const dd_t beta = neg ? nrm : -nrm, rbeta = neg ? rs : -rs;
    const dd_t tau = (beta - alpha) * rbeta;
    const dd_t t = dd_rcp_fast(alpha - beta);
    v1 = v1 * t;
    v2 = v2 * t;
    v0 = beta;
    return tau;
}
// Float64 fast path: same quantities (beta = -sign(alpha) ||x||, tau = (beta - alpha)/beta, tail / (alpha - beta)),
// computed on a copy scaled by an exact power of two with one fast sqrt and two fast reciprocals.
GS_DEV double reflector_real_small(double& v0, double& v1, double& v2, int nr) {
    if (nr <= 1) return 0.0;
    if (nr == 2) v2 = 0.0;
    if (v1 == 0.0 && v2 == 0.0) return 0.0;
    const double w = fmax(fabs(v0), fmax(fabs(v1), fabs(v2)));
    const double sfmin = 2.0 * 2.2250738585072014e-308 / 2.220446049250313e-16;
    if (!(w >= sfmin && w <= 1e300)) return reflector_real_small_generic<double>(v0, v1, v2, nr);   // rare
    double dn, up;
    pow2_scales(w, dn, up);
    const double a = v0 * dn, c = v1 * dn, d = v2 * dn;
    const double q = fma(a, a, fma(c, c, d * d));
    const double beta = -copysign(fast_sqrt(q), a);
    const double tau = (beta - a) * fast_rcp(beta);
    const double t = fast_rcp(a - beta);
    v1 = c * t;
    v2 = d * t;
    v0 = beta * up;
    return tau;
}

// Complex xLARFG on a 2-vector (src/householder.jl:56-102): beta real, tau complex.
template <class R> GS_DEV cx<R> reflector_cplx2_generic(cx<R>& v0, cx<R>& v1) {
    const R zero = r_const<R>(0.0), one = r_const<R>(1.0);
    R ar = v0.re, ai = v0.im;
    if (v1.re == zero && v1.im == zero && ai == zero) return mk_cx<R>(zero, zero);
    R beta = -r_copysign(r_hypot4(ar, ai, v1.re, v1.im), ar);
    const R sfmin = rtraits<R>::floatmin() / rtraits<R>::eps();
    int kount = 0;
    if (r_abs(beta) < sfmin) {
        const R rsfmin = one / sfmin;
        bool smallb = true;
        while (smallb) {
            kount += 1;
            v1 = v1 * rsfmin;
            beta = beta * rsfmin;
            ar = ar * rsfmin;
            ai = ai * rsfmin;
            smallb = (r_abs(beta) < sfmin) && (kount < 20);
        }
        beta = -r_copysign(r_hypot4(ar, ai, v1.re, v1.im), ar);
    }
    cx<R> tau = mk_cx<R>((beta - ar) / beta, -ai / beta);
    cx<R> t = mk_cx<R>(one, zero) / mk_cx<R>(ar - beta, ai);
    v1 = v1 * t;
    for (int j = 0; j < kount; ++j) beta = beta * sfmin;
    v0 = mk_cx<R>(beta, zero);
    return tau;
}

// Complex double-double fast path (see reflector_real_small above): tau = (beta - alpha) / beta, v1 <- v1 / (alpha - beta)
// with 1 / (alpha - beta) = conj(d) / |d|^2.
GS_DEV cx<dd_t> reflector_cplx2(cx<dd_t>& v0, cx<dd_t>& v1) {
    const dd_t zero = mk_dd(0.0);
    const dd_t ar = v0.re, ai = v0.im;
    if (v1.re == zero && v1.im == zero && ai == zero) return mk_cx<dd_t>(zero, zero);
    const double w = fmax(fmax(fabs(ar.hi), fabs(ai.hi)), fmax(fabs(v1.re.hi), fabs(v1.im.hi)));
    if (!(w >= 1e-100 && w <= 1e100)) return reflector_cplx2_generic<dd_t>(v0, v1);
    const dd_t q = (ar * ar + ai * ai) + (v1.re * v1.re + v1.im * v1.im);
    const dd_t rs = dd_rsqrt_fast(q);
    const dd_t nrm = q * rs;
    const bool neg = signbit(ar.hi);
    const dd_t beta = neg ? nrm : -nrm, rbeta = neg ? rs : -rs;
    const cx<dd_t> tau = mk_cx<dd_t>((beta - ar) * rbeta, -(ai * rbeta));
    const dd_t dr = ar - beta;
    const dd_t rm = dd_rcp_fast(dr * dr + ai * ai);
    const cx<dd_t> t = mk_cx<dd_t>(dr * rm, -(ai * rm));
    v1 = v1 * t;
    v0 = mk_cx<dd_t>(beta, zero);
    return tau;
}
GS_DEV cx<double> reflector_cplx2(cx<double>& v0, cx<double>& v1) {
    const double ar = v0.re, ai = v0.im;
    if (v1.re == 0.0 && v1.im == 0.0 && ai == 0.0) return mk_cx<double>(0.0, 0.0);
    const double w = fmax(fmax(fabs(ar), fabs(ai)), fmax(fabs(v1.re), fabs(v1.im)));
    const double sfmin = 2.2250738585072014e-308 / 2.220446049250313e-16;
    if (!(w >= sfmin && w <= 1e300)) return reflector_cplx2_generic<double>(v0, v1);   // rare
    double dn, up;
    pow2_scales(w, dn, up);
    const double a = ar * dn, b = ai * dn, c = v1.re * dn, d = v1.im * dn;
    const double q = fma(a, a, fma(b, b, fma(c, c, d * d)));
    const double beta = -copysign(fast_sqrt(q), a);
    const double rb = fast_rcp(beta);
    cx<double> tau = mk_cx<double>((beta - a) * rb, -b * rb);
    // 1/(alpha - beta) = conj(alpha - beta) / |alpha - beta|^2; |a - beta| = |a| + ||x|| >= 1 on the scaled copy
    const double amb = a - beta;
    const double rm = fast_rcp(fma(amb, amb, b * b));
    const double tr = amb * rm, ti = -b * rm;
    v1 = mk_cx<double>(c * tr - d * ti, c * ti + d * tr);
    v0 = mk_cx<double>(beta * up, 0.0);
    return tau;
}

// dlanv2 (src/GenericSchur.jl:716-803): standardise a real 2x2 block.
template <class R> GS_DEV void gs2x2(R& a, R& b, R& c, R& d, R& cs, R& sn, cx<R>& w1, cx<R>& w2) {
    const R zero = r_const<R>(0.0), one = r_const<R>(1.0), half = r_const<R>(0.5);
#define GS_SGN(x) (((x) < zero) ? -one : one)
    const R small = r_const<R>(4.0) * rtraits<R>::eps();
    if (c == zero) {
        cs = one;
        sn = zero;
    } else if (b == zero) {
        cs = zero;
        sn = one;
        R a0 = a, c0 = c, d0 = d;
        a = d0;
        b = -c0;
        c = zero;
        d = a0;
    } else if ((a - d) == zero && (b * c < zero)) {
        cs = one;
        sn = zero;
    } else {
        R asubd = a - d;
        R p = half * asubd;
        R bcmax = r_max(r_abs(b), r_abs(c));
        R bcmis = r_min(r_abs(b), r_abs(c)) * GS_SGN(b) * GS_SGN(c);
        R scale = r_max(r_abs(p), bcmax);
        R z = (p / scale) * p + (bcmax / scale) * bcmis;
        if (z >= small) {
            z = p + r_sqrt(scale) * r_sqrt(z) * GS_SGN(p);
            a = d + z;
            d = d - (bcmax / z) * bcmis;
            R tau = r_hypot(c, z);
            cs = z / tau;
            sn = c / tau;
            b = b - c;
            c = zero;
        } else {
            R sigma = b + c;
            R tau = r_hypot(sigma, asubd);
            cs = r_sqrt(half * (one + r_abs(sigma) / tau));
            sn = -(p / (tau * cs)) * GS_SGN(sigma);
            R aa = a * cs + b * sn, bb = -a * sn + b * cs;
            R cc = c * cs + d * sn, dd = -c * sn + d * cs;
            a = aa * cs + cc * sn;
            b = bb * cs + dd * sn;
            c = -aa * sn + cc * cs;
            d = -bb * sn + dd * cs;
            R midad = half * (a + d);
            a = midad;
            d = a;
            if (c != zero) {
                if (b != zero) {
                    if (b * c >= zero) {
                        R sab = r_sqrt(r_abs(b)), sac = r_sqrt(r_abs(c));
                        p = sab * sac * GS_SGN(c);
                        tau = one / r_sqrt(r_abs(b + c));
                        a = midad + p;
                        d = midad - p;
                        b = b - c;
                        c = zero;
                        R cs1 = sab * tau, sn1 = sac * tau;
                        R csn = cs * cs1 - sn * sn1, snn = cs * sn1 + sn * cs1;
                        cs = csn;
                        sn = snn;
                    }
                } else {
                    b = -c;
                    c = zero;
                    R cs0 = cs;
                    cs = -sn;
                    sn = cs0;
                }
            }
        }
    }
    if (c == zero) {
        w1 = mk_cx<R>(a, zero);
        w2 = mk_cx<R>(d, zero);
    } else {
        R rti = r_sqrt(r_abs(b)) * r_sqrt(r_abs(c));
        w1 = mk_cx<R>(a, rti);
        w2 = mk_cx<R>(d, -rti);
    }
#undef GS_SGN
}

// xLASCL multiplier sequence (src/util.jl:41-80), applied to `count` elements by the whole block.
template <class T, class R, int NT, class F> GS_DEV void safescale_apply(R cfrom, R cto, F&& apply_mul) {
    const R smlnum = r_safemin<R>();
    const R bignum = r_const<R>(1.0) / smlnum;
    const R zero = r_const<R>(0.0);
    R cfromc = cfrom, ctoc = cto, mul = zero;
    bool done = false;
    int guard = 0;
    while (!done && guard < 64) {
        ++guard;
        R cfrom1 = cfromc * smlnum;
        if (cfrom1 == cfromc) {
            mul = ctoc / cfromc;
            done = true;
        } else {
            R cto1 = ctoc / bignum;
            if (cto1 == ctoc) {
                mul = cto;
                done = true;
                cfromc = r_const<R>(1.0);
            } else if (r_abs(cfrom1) > r_abs(ctoc) && ctoc != zero) {
                mul = smlnum;
                done = false;
                cfromc = r_const<R>(1.0);
            } else if (r_abs(cto1) > r_abs(cfromc)) {
                mul = bignum;
                done = false;
                cfromc = cfrom1;
            } else {
                mul = ctoc / cfromc;
                done = true;
            }
        }
        apply_mul(mul);
    }
}

// =================================================================================================
template <class T, int NT> struct BatchedSolver {
    typedef typename etraits<T>::real R;
    typedef cx<R> C;
    static constexpr bool CPLX = etraits<T>::is_complex;

    int n, ld, tid, lane;
    T* H;
    T* Z;      // shared; valid storage even when !wantZ
    T* sTau;
    C* sW;
    R* sRed;
    bool wantZ;
    unsigned* stp;   // per-matrix counters: sweeps, reflector applications, exceptional shifts, iterations

#define HH(i, j) H[((i)-1) + ((j)-1) * ld]
#define ZZ(i, j) Z[((i)-1) + ((j)-1) * ld]

    // ---- _scale! (src/util.jl:14-29) ------------------------------------------------------------
    GS_DEV bool scale_in(R& cscale, R& anrm) {
        const R zero = r_const<R>(0.0);
        R m = zero;
        for (int e = tid; e < n * n; e += NT) {
            int i = e % n, j = e / n;
            m = r_max(m, e_abs(H[i + j * ld]));
        }
        anrm = block_max<R, NT>(m, sRed);
        const R smlnum = r_sqrt(r_safemin<R>()) / rtraits<R>::eps();
        const R bignum = r_const<R>(1.0) / smlnum;
        bool scaled = false;
        cscale = r_const<R>(1.0);
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
                    H[i + j * ld] = e_scale(H[i + j * ld], mul);
                }
            });
        }
        block_sync<NT>();
        return scaled;
    }

    // ---- _hessenberg! (src/hessenberg.jl:3-17, src/householder.jl:12-102,140-172) ------------------
    GS_DEV void hessenberg() {
        const R zero = r_const<R>(0.0), one = r_const<R>(1.0);
        for (int i = 1; i <= n - 1; ++i) {
            const int len = n - i;        // reflector acts on rows i+1..n
            const int nv = len - 1;       // stored tail HH(i+2..n, i)
            // ---- _reflector!(view(A, i+1:n, i)) ----
            T alpha = HH(i + 1, i);
            T tau = e_zero<T>();
            bool trivial;
            if (CPLX) trivial = false;    // a complex length-1 "reflector" is a phase (src/householder.jl:59-60)
            else trivial = (len <= 1);
            if (!trivial) {
                // scaled 2-norm of the tail (src/util.jl:506-557 computes the same quantity serially)
                R amax = zero;
                for (int r = tid; r < nv; r += NT) amax = r_max(amax, e_maxpart(HH(i + 2 + r, i)));
                amax = block_max<R, NT>(amax, sRed);
                R xnorm = zero;
                if (amax > zero) {
                    R rs = one / amax, ssq = zero;
                    for (int r = tid; r < nv; r += NT) ssq = ssq + e_sq_scaled(HH(i + 2 + r, i), rs);
                    ssq = block_sum<R, NT>(ssq, sRed);
                    xnorm = amax * r_sqrt(ssq);
                }
                R ar, ai;
                if constexpr (CPLX) {
                    ar = alpha.re;
                    ai = alpha.im;
                } else {
                    ar = alpha;
                    ai = zero;
                }
                bool nothing = CPLX ? (xnorm == zero && ai == zero) : (xnorm == zero);
                if (!nothing) {
                    R beta = -r_copysign(r_hypot4(ar, ai, xnorm, zero), ar);
                    const R sfmin = CPLX ? rtraits<R>::floatmin() / rtraits<R>::eps()
                                         : r_const<R>(2.0) * rtraits<R>::floatmin() / rtraits<R>::eps();
                    int kount = 0;
                    if (r_abs(beta) < sfmin) {
                        const R rsfmin = one / sfmin;
                        bool smallb = true;
                        while (smallb) {
                            kount += 1;
                            for (int r = tid; r < nv; r += NT) HH(i + 2 + r, i) = e_scale(HH(i + 2 + r, i), rsfmin);
                            beta = beta * rsfmin;
                            ar = ar * rsfmin;
                            ai = ai * rsfmin;
                            smallb = (r_abs(beta) < sfmin) && (kount < 20);
                        }
                        block_sync<NT>();
                        R am2 = zero;
                        for (int r = tid; r < nv; r += NT) am2 = r_max(am2, e_maxpart(HH(i + 2 + r, i)));
                        am2 = block_max<R, NT>(am2, sRed);
                        xnorm = zero;
                        if (am2 > zero) {
                            R rs = one / am2, ssq = zero;
                            for (int r = tid; r < nv; r += NT) ssq = ssq + e_sq_scaled(HH(i + 2 + r, i), rs);
                            ssq = block_sum<R, NT>(ssq, sRed);
                            xnorm = am2 * r_sqrt(ssq);
                        }
                        beta = -r_copysign(r_hypot4(ar, ai, xnorm, zero), ar);
                    }
                    T tscal;
                    if constexpr (CPLX) {
                        tau = mk_cx<R>((beta - ar) / beta, -ai / beta);
                        tscal = mk_cx<R>(one, zero) / mk_cx<R>(ar - beta, ai);
                    } else {
                        tau = (beta - ar) / beta;
                        tscal = one / (ar - beta);
                    }
                    for (int r = tid; r < nv; r += NT) HH(i + 2 + r, i) = HH(i + 2 + r, i) * tscal;
                    for (int j = 0; j < kount; ++j) beta = beta * sfmin;
                    if (tid == 0) {
                        if constexpr (CPLX) HH(i + 1, i) = mk_cx<R>(beta, zero);
                        else HH(i + 1, i) = beta;
                    }
                }
            }
            if (tid == 0) sTau[i - 1] = tau;
            block_sync<NT>();
            // ---- lmul!(H', view(A, i+1:n, i+1:n)) : one thread per column ----
            const T tauc = cconj(tau);
            for (int j = i + 1 + tid; j <= n; j += NT) {
                T va = HH(i + 1, j);
                for (int r = 0; r < nv; ++r) va = va + cconj(HH(i + 2 + r, i)) * HH(i + 2 + r, j);
                va = tauc * va;
                HH(i + 1, j) = HH(i + 1, j) - va;
                for (int r = 0; r < nv; ++r) HH(i + 2 + r, j) = HH(i + 2 + r, j) - va * HH(i + 2 + r, i);
            }
            block_sync<NT>();
            // ---- rmul!(view(A, :, i+1:n), H) : one thread per row ----
            for (int r = 1 + tid; r <= n; r += NT) {
                T x = HH(r, i + 1);
                for (int c = 0; c < nv; ++c) x = x + HH(r, i + 2 + c) * HH(i + 2 + c, i);
                T tx = tau * x;
                HH(r, i + 1) = HH(r, i + 1) - tx;
                for (int c = 0; c < nv; ++c) {
                    // the reflector tail of column i lives in rows i+2.. of column i; row r of that column is
                    // only read here when r >= i+2 as an operand of another row's update, never written
                    HH(r, i + 2 + c) = HH(r, i + 2 + c) - tx * cconj(HH(i + 2 + c, i));
                }
            }
            block_sync<NT>();
        }
    }

    // ---- _materializeQ (src/hessenberg.jl:150-166): backward accumulation, one thread per column ----
    GS_DEV void form_q() {
        for (int e = tid; e < n * n; e += NT) {
            int i = e % n, j = e / n;
            Z[i + j * ld] = (i == j) ? e_one<T>() : e_zero<T>();
        }
        block_sync<NT>();
        for (int j = 2 + tid; j <= n; j += NT) {
            // reflector k touches column j only once k+1 <= j
            for (int k = (j - 1 < n - 1 ? j - 1 : n - 1); k >= 1; --k) {
                const int nv = n - k - 1;
                T vb = ZZ(k + 1, j);
                for (int r = 0; r < nv; ++r) vb = vb + cconj(HH(k + 2 + r, k)) * ZZ(k + 2 + r, j);
                vb = sTau[k - 1] * vb;
                ZZ(k + 1, j) = ZZ(k + 1, j) - vb;
                for (int r = 0; r < nv; ++r) ZZ(k + 2 + r, j) = ZZ(k + 2 + r, j) - HH(k + 2 + r, k) * vb;
            }
        }
        block_sync<NT>();
    }

    GS_DEV void zero_below_subdiag() {
        for (int e = tid; e < n * n; e += NT) {
            int i = e % n, j = e / n;
            if (i > j + 1) H[i + j * ld] = e_zero<T>();
        }
        block_sync<NT>();
    }

    // =============================================================================================
    // complex single-shift QR (src/GenericSchur.jl:194-335, 374-504)
    // =============================================================================================
    GS_DEV bool split_test_c(int c, const R& smallnum, const R& ulp) {
        const R zero = r_const<R>(0.0);
        C h10 = HH(c + 1, c);
        if (abs1(h10) <= smallnum) return true;
        C hcc = HH(c, c), hc1 = HH(c + 1, c + 1);
        R tst = abs1(hcc) + abs1(hc1);
        if (tst == zero) {
            if (c - 1 >= 1) tst = tst + r_abs(HH(c, c - 1).re);
            if (c + 2 <= n) tst = tst + r_abs(HH(c + 2, c + 1).re);
        }
        if (r_abs(h10.re) <= ulp * tst) {
            R a1 = abs1(h10), a2 = abs1(HH(c, c + 1));
            R ab = r_max(a1, a2), ba = r_min(a1, a2);
            R d1 = abs1(hc1), d2 = abs1(hcc - hc1);
            R aa = r_max(d1, d2), bb = r_min(d1, d2);
            R s = aa + ab;
            if (ba * (ab / s) <= r_max(smallnum, ulp * (bb * (aa / s)))) return true;
        }
        return false;
    }

    GS_DEV void sweep_complex(const C& shift, int istart, int iend) {
        const R zero = r_const<R>(0.0), one = r_const<R>(1.0);
        const R ulp = rtraits<R>::eps();
        // start row: largest mm in [istart+1, iend-1] passing the two-small-subdiagonals test
        int istart1 = 0;
        for (int base = iend - 1; base >= istart + 1 && !istart1; base -= 32) {
            int mm = base - lane;
            bool hit = false;
            if (mm >= istart + 1) {
                C h11 = HH(mm, mm), h22 = HH(mm + 1, mm + 1);
                C h11s = h11 - shift;
                R h21 = HH(mm + 1, mm).re;
                R s = abs1(h11s) + r_abs(h21);
                h11s = mk_cx<R>(h11s.re / s, h11s.im / s);
                h21 = h21 / s;
                R h10 = HH(mm, mm - 1).re;
                hit = r_abs(h10) * r_abs(h21) <= ulp * (abs1(h11s) * (abs1(h11) + abs1(h22)));
            }
            unsigned m = __ballot_sync(0xffffffffu, hit);
            if (m) istart1 = base - (__ffs(m) - 1);
        }
        if (!istart1) istart1 = istart;
        C v0, v1;
        {
            C h11s = HH(istart1, istart1) - shift;
            R h21 = HH(istart1 + 1, istart1).re;
            R s = abs1(h11s) + r_abs(h21);
            v0 = mk_cx<R>(h11s.re / s, h11s.im / s);
            v1 = mk_cx<R>(h21 / s, zero);
        }
        for (int k = istart1; k <= iend - 1; ++k) {
            if (k > istart1) {
                v0 = HH(k, k - 1);
                v1 = HH(k + 1, k - 1);
            }
            C tau1 = reflector_cplx2(v0, v1);
            stp[1] += 1;
            const C v2 = v1, v2c = cconj(v1), tau1c = cconj(tau1);
            const R tau2 = (tau1 * v2).re;
            // phase 1: left update of rows k, k+1 (one thread per column) and Z update (one thread per row)
            for (int j = k + tid; j <= n; j += NT) {
                C a = HH(k, j), b = HH(k + 1, j);
                C ss = tau1c * a + tau2 * b;
                HH(k, j) = a - ss;
                HH(k + 1, j) = b - ss * v2;
            }
            if (wantZ) {
                for (int r = 1 + tid; r <= n; r += NT) {
                    C a = ZZ(r, k), b = ZZ(r, k + 1);
                    C ss = tau1 * a + tau2 * b;
                    ZZ(r, k) = a - ss;
                    ZZ(r, k + 1) = b - ss * v2c;
                }
            }
            block_sync<NT>();
            // phase 2: right update of columns k, k+1 (one thread per row); deferred sub-diagonal writes
            const int jmax = (k + 2 < iend) ? k + 2 : iend;
            for (int r = 1 + tid; r <= jmax; r += NT) {
                C a = HH(r, k), b = HH(r, k + 1);
                C ss = tau1 * a + tau2 * b;
                HH(r, k) = a - ss;
                HH(r, k + 1) = b - ss * v2c;
            }
            if (tid == 0 && k > istart1) {
                HH(k, k - 1) = v0;
                HH(k + 1, k - 1) = mk_cx<R>(zero, zero);
            }
            block_sync<NT>();
            if (k == istart1 && istart1 > istart) {
                // late start: rescale so that HH[istart1, istart1-1] stays real (src/GenericSchur.jl:461-482)
                C t = mk_cx<R>(one, zero) - tau1;
                R at = c_abs(t);
                t = mk_cx<R>(t.re / at, t.im / at);
                C tc = cconj(t);
                if (tid == 0) {
                    HH(istart1 + 1, istart1) = HH(istart1 + 1, istart1) * tc;
                    if (istart1 + 2 <= iend) HH(istart1 + 2, istart1 + 1) = HH(istart1 + 2, istart1 + 1) * t;
                }
                block_sync<NT>();
                for (int j = istart1; j <= iend; ++j) {
                    if (j == istart1 + 1) continue;
                    for (int c = j + 1 + tid; c <= n; c += NT) HH(j, c) = HH(j, c) * t;
                    for (int r = 1 + tid; r <= j - 1; r += NT) HH(r, j) = HH(r, j) * tc;
                    if (wantZ)
                        for (int r = 1 + tid; r <= n; r += NT) ZZ(r, j) = ZZ(r, j) * tc;
                    block_sync<NT>();
                }
            }
        }
        // make the tail sub-diagonal real (src/GenericSchur.jl:486-500)
        C t = HH(iend, iend - 1);
        if (t.im != zero) {
            R rt = c_abs(t);
            t = mk_cx<R>(t.re / rt, t.im / rt);
            C tc = cconj(t);
            for (int c = iend + 1 + tid; c <= n; c += NT) HH(iend, c) = HH(iend, c) * tc;
            for (int r = 1 + tid; r <= iend - 1; r += NT) HH(r, iend) = HH(r, iend) * t;
            if (wantZ)
                for (int r = 1 + tid; r <= n; r += NT) ZZ(r, iend) = ZZ(r, iend) * t;
            block_sync<NT>();
            if (tid == 0) HH(iend, iend - 1) = mk_cx<R>(rt, zero);
        }
        block_sync<NT>();
    }

    // returns info (0 ok, else iend at failure); st[0..3] = sweeps, applications, exceptional, iterations
    GS_DEV int qr_complex(int maxiter, unsigned* st) {
        stp = st;
        const R zero = r_const<R>(0.0), half = r_const<R>(0.5), threeq = r_const<R>(0.75);
        const R ulp = rtraits<R>::eps();
        const R smallnum = r_safemin<R>() * (r_const<R>((double)n) / ulp);
        const int maxinner = 30 * n;
        int istart = 1, iend = n, it = 0;
        while (iend >= 1) {
            istart = 1;
            for (int its = 0; its <= maxinner; ++its) {
                it += 1;
                if (it > maxiter) {
                    st[3] = it;
                    return iend;
                }
                // lowest-positioned negligible sub-diagonal: largest c in [istart, iend-1] that splits
                int found = 0;
                for (int base = iend - 1; base >= istart && !found; base -= 32) {
                    int c = base - lane;
                    bool hit = (c >= istart) ? split_test_c(c, smallnum, ulp) : false;
                    unsigned m = __ballot_sync(0xffffffffu, hit);
                    if (m) found = base - (__ffs(m) - 1);
                }
                if (found) istart = found + 1;
                block_sync<NT>();
                if (istart > 1 && tid == 0) HH(istart, istart - 1) = mk_cx<R>(zero, zero);
                block_sync<NT>();
                if (istart >= iend) {
                    iend -= 1;
                    break;
                }
                C t;
                if (its % 30 == 10) {
                    R s = threeq * r_abs(HH(istart + 1, istart).re);
                    t = HH(istart, istart);
                    t.re = t.re + s;
                    st[2] += 1;
                } else if (its % 30 == 20) {
                    R s = threeq * r_abs(HH(iend, iend - 1).re);
                    t = HH(iend, iend);
                    t.re = t.re + s;
                    st[2] += 1;
                } else {
                    // Wilkinson shift (src/GenericSchur.jl:309-324)
                    t = HH(iend, iend);
                    C u = c_sqrt(HH(iend - 1, iend)) * c_sqrt(HH(iend, iend - 1));
                    R s = abs1(u);
                    if (s != zero) {
                        C x = half * (HH(iend - 1, iend - 1) - t);
                        R sx = abs1(x);
                        s = r_max(s, sx);
                        C xs = mk_cx<R>(x.re / s, x.im / s), us = mk_cx<R>(u.re / s, u.im / s);
                        C y = s * c_sqrt(xs * xs + us * us);
                        if (sx > zero) {
                            if ((x.re / sx) * y.re + (x.im / sx) * y.im < zero) y = -y;
                        }
                        t = t - u * (u / (x + y));
                    }
                }
                st[0] += 1;
                sweep_complex(t, istart, iend);
            }
        }
        st[3] = it;
        return 0;
    }

    // =============================================================================================
    // real double-shift QR (src/GenericSchur.jl:513-699, 837-952)
    // =============================================================================================
    GS_DEV bool split_test_r(int k, const R& smallnum, const R& eps) {
        const R zero = r_const<R>(0.0);
        R h = r_abs(HH(k, k - 1));
        if (h < smallnum) return true;
        R Hkk = HH(k, k), Hk1 = HH(k - 1, k - 1);
        R t = r_abs(Hk1) + r_abs(Hkk);
        if (t == zero) {
            if (k > 2) t = t + r_abs(HH(k - 1, k - 2));
            if (k + 1 <= n) t = t + r_abs(HH(k + 1, k));
        }
        if (h <= t * eps) {
            R o = r_abs(HH(k - 1, k));
            R ab = r_max(h, o), ba = r_min(h, o);
            R d1 = r_abs(Hkk), d2 = r_abs(Hk1 - Hkk);
            R aa = r_max(d1, d2), bb = r_min(d1, d2);
            R s = aa + bb;   // as the reference has it (src/GenericSchur.jl:586)
            if (ba * (ab / s) <= r_max(smallnum, eps * (bb * (aa / s)))) return true;
        }
        return false;
    }

    GS_DEV void first_column_r(int m, const R& r1r, const R& r1i, const R& r2r, const R& r2i, R& v0, R& v1, R& v2) {
        R hmm = HH(m, m);
        R H21s = HH(m + 1, m);
        R s = r_abs(hmm - r2r) + r_abs(r2i) + r_abs(H21s);
        H21s = H21s / s;
        v0 = H21s * HH(m, m + 1) + (hmm - r1r) * ((hmm - r2r) / s) - r1i * (r2i / s);
        v1 = H21s * (hmm + HH(m + 1, m + 1) - r1r - r2r);
        v2 = H21s * HH(m + 2, m + 1);
        s = r_abs(v0) + r_abs(v1) + r_abs(v2);
        v0 = v0 / s;
        v1 = v1 / s;
        v2 = v2 / s;
    }

    GS_DEV void sweep_real(const R& r1r, const R& r1i, const R& r2r, const R& r2i, int istart, int iend) {
        const R zero = r_const<R>(0.0), one = r_const<R>(1.0);
        const R eps = rtraits<R>::eps();
        int mx = 0;
        for (int base = iend - 2; base >= istart + 1 && !mx; base -= 32) {
            int m = base - lane;
            bool hit = false;
            if (m >= istart + 1) {
                R v0, v1, v2;
                first_column_r(m, r1r, r1i, r2r, r2i, v0, v1, v2);
                hit = r_abs(HH(m, m - 1)) * (r_abs(v1) + r_abs(v2)) <=
                      eps * r_abs(v0) * (r_abs(HH(m - 1, m - 1)) + r_abs(HH(m, m)) + r_abs(HH(m + 1, m + 1)));
            }
            unsigned msk = __ballot_sync(0xffffffffu, hit);
            if (msk) mx = base - (__ffs(msk) - 1);
        }
        if (!mx) mx = istart;
        R v0, v1, v2;
        first_column_r(mx, r1r, r1i, r2r, r2i, v0, v1, v2);
        for (int k = mx; k <= iend - 1; ++k) {
            const int nr = (iend - k + 1 < 3) ? iend - k + 1 : 3;
            if (k > mx) {
                v0 = HH(k, k - 1);
                v1 = HH(k + 1, k - 1);
                v2 = (nr == 3) ? HH(k + 2, k - 1) : zero;
            }
            const R tau1 = reflector_real_small(v0, v1, v2, nr);
            stp[1] += 1;
            const R tau2 = tau1 * v1;
            if (nr == 3) {
                const R tau3 = tau1 * v2;
                for (int j = k + tid; j <= n; j += NT) {
                    R a = HH(k, j), b = HH(k + 1, j), c = HH(k + 2, j);
                    R ss = a + v1 * b + v2 * c;
                    HH(k, j) = a - ss * tau1;
                    HH(k + 1, j) = b - ss * tau2;
                    HH(k + 2, j) = c - ss * tau3;
                }
                if (wantZ) {
                    for (int r = 1 + tid; r <= n; r += NT) {
                        R a = ZZ(r, k), b = ZZ(r, k + 1), c = ZZ(r, k + 2);
                        R ss = a + v1 * b + v2 * c;
                        ZZ(r, k) = a - ss * tau1;
                        ZZ(r, k + 1) = b - ss * tau2;
                        ZZ(r, k + 2) = c - ss * tau3;
                    }
                }
                block_sync<NT>();
                const int jmax = (k + 3 < iend) ? k + 3 : iend;
                for (int r = 1 + tid; r <= jmax; r += NT) {
                    R a = HH(r, k), b = HH(r, k + 1), c = HH(r, k + 2);
                    R ss = a + v1 * b + v2 * c;
                    HH(r, k) = a - ss * tau1;
                    HH(r, k + 1) = b - ss * tau2;
                    HH(r, k + 2) = c - ss * tau3;
                }
            } else {
                for (int j = k + tid; j <= n; j += NT) {
                    R a = HH(k, j), b = HH(k + 1, j);
                    R ss = a + v1 * b;
                    HH(k, j) = a - ss * tau1;
                    HH(k + 1, j) = b - ss * tau2;
                }
                if (wantZ) {
                    for (int r = 1 + tid; r <= n; r += NT) {
                        R a = ZZ(r, k), b = ZZ(r, k + 1);
                        R ss = a + v1 * b;
                        ZZ(r, k) = a - ss * tau1;
                        ZZ(r, k + 1) = b - ss * tau2;
                    }
                }
                block_sync<NT>();
                for (int r = 1 + tid; r <= iend; r += NT) {
                    R a = HH(r, k), b = HH(r, k + 1);
                    R ss = a + v1 * b;
                    HH(r, k) = a - ss * tau1;
                    HH(r, k + 1) = b - ss * tau2;
                }
            }
            if (tid == 0) {
                if (k > mx) {
                    HH(k, k - 1) = v0;
                    HH(k + 1, k - 1) = zero;
                    if (k < iend - 1) HH(k + 2, k - 1) = zero;
                } else if (mx > istart) {
                    HH(k, k - 1) = HH(k, k - 1) * (one - tau1);
                }
            }
            block_sync<NT>();
        }
    }

    GS_DEV int qr_real(int maxiter, unsigned* st) {
        stp = st;
        const R zero = r_const<R>(0.0);
        const R eps = rtraits<R>::eps();
        const R smallnum = rtraits<R>::floatmin() * (r_const<R>((double)n) / eps);
        const R threeq = r_const<R>(0.75), m7_16 = r_const<R>(-0.4375);
        int istart = 1, iend = n, iwcur = n, iter = 0;
        while (iend >= 1) {
            istart = 1;
            int iterqr = 0;
            while (true) {
                iter += 1;
                if (iter > maxiter) {
                    st[3] = iter;
                    return iend;
                }
                int found = 0;
                for (int base = iend; base >= istart + 1 && !found; base -= 32) {
                    int k = base - lane;
                    bool hit = (k >= istart + 1) ? split_test_r(k, smallnum, eps) : false;
                    unsigned m = __ballot_sync(0xffffffffu, hit);
                    if (m) found = base - (__ffs(m) - 1);
                }
                istart = found ? found : 1;
                block_sync<NT>();
                if (istart > 1 && tid == 0) HH(istart, istart - 1) = zero;
                block_sync<NT>();
                if (istart >= iend - 1) break;
                iterqr += 1;
                R H11, H12, H21, H22;
                if (iterqr == 10) {
                    R s = r_abs(HH(istart + 1, istart)) + r_abs(HH(istart + 2, istart + 1));
                    H11 = threeq * s + HH(istart, istart);
                    H12 = m7_16 * s;
                    H21 = s;
                    H22 = H11;
                    st[2] += 1;
                } else if (iterqr == 20) {
                    R s = r_abs(HH(iend, iend - 1)) + r_abs(HH(iend - 1, iend - 2));
                    H11 = threeq * s + HH(iend, iend);
                    H12 = m7_16 * s;
                    H21 = s;
                    H22 = H11;
                    st[2] += 1;
                } else {
                    H11 = HH(iend - 1, iend - 1);
                    H21 = HH(iend, iend - 1);
                    H12 = HH(iend - 1, iend);
                    H22 = HH(iend, iend);
                }
                R s = r_abs(H11) + r_abs(H12) + r_abs(H21) + r_abs(H22);
                R r1r = zero, r2r = zero, r1i = zero, r2i = zero;
                if (!(s == zero)) {
                    H11 = H11 / s;
                    H12 = H12 / s;
                    H21 = H21 / s;
                    H22 = H22 / s;
                    R tr = (H11 + H22) * r_const<R>(0.5);
                    R d = (H11 - tr) * (H22 - tr) - H12 * H21;
                    R rtd = r_sqrt(r_abs(d));
                    if (d >= zero) {
                        r1r = tr * s;
                        r2r = r1r;
                        r1i = rtd * s;
                        r2i = -r1i;
                    } else {
                        r1r = tr + rtd;
                        r2r = tr - rtd;
                        if (r_abs(r1r - H22) <= r_abs(r2r - H22)) {
                            r1r = r1r * s;
                            r2r = r1r;
                        } else {
                            r2r = r2r * s;
                            r1r = r2r;
                        }
                    }
                }
                st[0] += 1;
                sweep_real(r1r, r1i, r2r, r2i, istart, iend);
            }
            // deflation (src/GenericSchur.jl:668-688)
            if (istart >= iend) {
                if (tid == 0) sW[iwcur - 1] = mk_cx<R>(HH(iend, iend), zero);
                iwcur -= 1;
            } else if (istart + 1 == iend) {
                R a = HH(iend - 1, iend - 1), b = HH(iend - 1, iend), c = HH(iend, iend - 1), d = HH(iend, iend);
                R cs, sn;
                C w1, w2;
                gs2x2(a, b, c, d, cs, sn, w1, w2);
                if (tid == 0) {
                    sW[iwcur - 1] = w2;
                    sW[iwcur - 2] = w1;
                }
                iwcur -= 2;
                block_sync<NT>();
                // lmul!(G2, view(HH, :, istart:n))
                for (int j = istart + tid; j <= n; j += NT) {
                    R a1 = HH(iend - 1, j), a2 = HH(iend, j);
                    HH(iend - 1, j) = cs * a1 + sn * a2;
                    HH(iend, j) = -sn * a1 + cs * a2;
                }
                if (wantZ) {
                    for (int r = 1 + tid; r <= n; r += NT) {
                        R a1 = ZZ(r, iend - 1), a2 = ZZ(r, iend);
                        ZZ(r, iend - 1) = a1 * cs + a2 * sn;
                        ZZ(r, iend) = -a1 * sn + a2 * cs;
                    }
                }
                block_sync<NT>();
                // rmul!(view(HH, 1:iend, :), G2')
                for (int r = 1 + tid; r <= iend; r += NT) {
                    R a1 = HH(r, iend - 1), a2 = HH(r, iend);
                    HH(r, iend - 1) = a1 * cs + a2 * sn;
                    HH(r, iend) = -a1 * sn + a2 * cs;
                }
                block_sync<NT>();
                if (tid == 0) {
                    HH(iend - 1, iend - 1) = a;
                    HH(iend - 1, iend) = b;
                    HH(iend, iend - 1) = c;
                    HH(iend, iend) = d;
                    if (iend > 2) HH(iend - 1, iend - 2) = zero;
                }
                block_sync<NT>();
            }
            iend = istart - 1;
        }
        st[3] = iter;
        return 0;
    }
#undef HH
#undef ZZ
};

// =================================================================================================
// kernel
// =================================================================================================
template <class T> struct smem_layout {
    typedef typename etraits<T>::real R;
    // [H: n*ld T][Z: n*ld T][tau: n T][w: n cx<R>][red: 32 R][mbar: 8 B], every region 16-byte aligned
    __host__ __device__ static int ld(int n) { return n | 1; }
    __host__ __device__ static size_t up16(size_t x) { return (x + 15) & ~(size_t)15; }
    __host__ __device__ static size_t off_Z(int n) { return up16((size_t)n * ld(n) * sizeof(T)); }
    __host__ __device__ static size_t off_tau(int n) { return off_Z(n) + up16((size_t)n * ld(n) * sizeof(T)); }
    __host__ __device__ static size_t off_w(int n) { return off_tau(n) + up16((size_t)n * sizeof(T)); }
    __host__ __device__ static size_t off_red(int n) { return off_w(n) + up16((size_t)n * 2 * sizeof(R)); }
    __host__ __device__ static size_t off_mbar(int n) { return off_red(n) + up16(32 * sizeof(R)); }
    __host__ __device__ static size_t bytes(int n) { return off_mbar(n) + 16; }
};

template <class T, int NT>
__global__ void __launch_bounds__(NT) gschur_batched_kernel(BatchedParams p) {
    typedef typename etraits<T>::real R;
    typedef cx<R> C;
    constexpr bool CPLX = etraits<T>::is_complex;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = p.n, ld = smem_layout<T>::ld(n);
    const int tid = threadIdx.x;

    BatchedSolver<T, NT> S;
    S.n = n;
    S.ld = ld;
    S.tid = tid;
    S.lane = tid & 31;
    typedef smem_layout<T> L;
    S.H = reinterpret_cast<T*>(smem_raw);
    S.Z = reinterpret_cast<T*>(smem_raw + L::off_Z(n));
    S.sTau = reinterpret_cast<T*>(smem_raw + L::off_tau(n));
    S.sW = reinterpret_cast<C*>(smem_raw + L::off_w(n));
    S.sRed = reinterpret_cast<R*>(smem_raw + L::off_red(n));
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw + L::off_mbar(n));
    __shared__ long long s_next;
    S.wantZ = (p.Z != nullptr);

    if (tid == 0) mbar_init(mbar, 1);
    __syncthreads();
    uint32_t parity = 0;

    const size_t mat_bytes = (size_t)n * n * sizeof(T);
    const bool dense = (p.lda == n);

    for (;;) {
        if (tid == 0) s_next = (long long)atomicAdd(p.counter, 1ULL);
        __syncthreads();
        const long long b = s_next;
        __syncthreads();
        if (b >= p.batch) break;

        T* gA = reinterpret_cast<T*>(p.A) + b * p.strideA;
        T* gZ = S.wantZ ? reinterpret_cast<T*>(p.Z) + b * p.strideZ : nullptr;
        const R zero = r_const<R>(0.0);

        // ---- stage A_b: one TMA bulk copy into the (still unused) Z region, then repack with the padded ld ----
        const bool use_tma = dense && (mat_bytes % 16 == 0) && ((reinterpret_cast<uintptr_t>(gA) & 15) == 0);
        if (use_tma) {
            if (tid == 0) {
                fence_proxy_async();
                mbar_expect_tx(mbar, (uint32_t)mat_bytes);
                tma_bulk_g2s(S.Z, gA, (uint32_t)mat_bytes, mbar);
            }
            mbar_wait(mbar, parity);
            parity ^= 1;
            for (int e = tid; e < n * n; e += NT) {
                int i = e % n, j = e / n;
                S.H[i + j * ld] = S.Z[e];
            }
        } else {
            for (int e = tid; e < n * n; e += NT) {
                int i = e % n, j = e / n;
                S.H[i + j * ld] = gA[i + (size_t)j * p.lda];
            }
        }
        __syncthreads();

        unsigned st[4] = {0u, 0u, 0u, 0u};
        int info = 0;
        bool scaled = false;
        R cscale = r_const<R>(1.0), anrm = r_const<R>(1.0);

        if (p.mode == MODE_HESSENBERG) {
            S.hessenberg();
            if (S.wantZ) S.form_q();
            // store factors, tau, Q
            for (int e = tid; e < n * n; e += NT) {
                int i = e % n, j = e / n;
                gA[i + (size_t)j * p.lda] = S.H[i + j * ld];
                if (S.wantZ) gZ[i + (size_t)j * p.ldz] = S.Z[i + j * ld];
            }
            T* gtau = reinterpret_cast<T*>(p.tau) + b * (long long)(n > 1 ? n - 1 : 0);
            for (int e = tid; e < n - 1; e += NT) gtau[e] = S.sTau[e];
            __syncthreads();
            continue;
        }

        if (p.flags & F_HESS_INPUT) {
            // gschur!(H::Hessenberg, Z): Z is in/out (src/GenericSchur.jl:194-198, 513-518)
            if (S.wantZ)
                for (int e = tid; e < n * n; e += NT) {
                    int i = e % n, j = e / n;
                    S.Z[i + j * ld] = gZ[i + (size_t)j * p.ldz];
                }
            if (CPLX && (p.flags & F_CHECK_SUBDIAG)) {
                int bad = 0;
                if constexpr (CPLX) {
                    for (int j = 1 + tid; j <= n - 1; j += NT)
                        if (S.H[j + (j - 1) * ld].im != zero) bad = 1;
                }
                bad = __syncthreads_or(bad);
                if (bad) info = -4;
            }
            __syncthreads();
        } else {
            if (p.scale) scaled = S.scale_in(cscale, anrm);
            S.hessenberg();
            if (S.wantZ) S.form_q();
        }
        if (info == 0) {
            S.zero_below_subdiag();
            const int maxiter = p.maxiter > 0 ? p.maxiter : 100 * n;
            if constexpr (CPLX) info = S.qr_complex(maxiter, st);
            else info = S.qr_real(maxiter, st);
        }
        __syncthreads();

        // ---- unscale (src/GenericSchur.jl:367-370, 830-833) ----
        if (scaled) {
            safescale_apply<T, R, NT>(cscale, anrm, [&](R mul) {
                for (int e = tid; e < n * n; e += NT) {
                    int i = e % n, j = e / n;
                    S.H[i + j * ld] = e_scale(S.H[i + j * ld], mul);
                }
                if (!CPLX)
                    for (int e = tid; e < n; e += NT) S.sW[e] = mk_cx<R>(S.sW[e].re * mul, S.sW[e].im * mul);
            });
            __syncthreads();
        }

        // ---- store T (exact zeros below the (quasi-)triangle), Z, w, info, stats ----
        for (int e = tid; e < n * n; e += NT) {
            int i = e % n, j = e / n;
            T v = S.H[i + j * ld];
            bool keep = CPLX ? (i <= j) : (i <= j + 1);
            gA[i + (size_t)j * p.lda] = keep ? v : e_zero<T>();
            if (S.wantZ) gZ[i + (size_t)j * p.ldz] = S.Z[i + j * ld];
        }
        C* gw = reinterpret_cast<C*>(p.w) + b * (long long)n;
        for (int e = tid; e < n; e += NT) {
            if constexpr (CPLX) gw[e] = S.H[e + e * ld];
            else gw[e] = S.sW[e];
        }
        if (tid == 0) {
            if (p.info) p.info[b] = info;
            if (p.stats) {
                p.stats[4 * b + 0] = st[0];
                p.stats[4 * b + 1] = st[1];
                p.stats[4 * b + 2] = st[2];
                p.stats[4 * b + 3] = st[3];
            }
        }
        __syncthreads();
    }
}

// ---- host launcher shared by the per-kind translation units ------------------------------------------------
template <class T, int NT> int launch_nt(const BatchedParams& p, int dev_sms, cudaStream_t stream, std::string* err) {
    auto kern = gschur_batched_kernel<T, NT>;
    size_t smem = smem_layout<T>::bytes(p.n);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 0;
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem);
    if (e != cudaSuccess) {
        *err = std::string("kernel setup: ") + cudaGetErrorString(e);
        return -2;
    }
    if (per_sm < 1) {
        *err = "kernel does not fit on an SM";
        return -3;
    }
    long long grid = (long long)per_sm * dev_sms;
    if (grid > p.batch) grid = p.batch;
    kern<<<(unsigned)grid, NT, smem, stream>>>(p);
    note_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) {
        *err = std::string("kernel launch: ") + cudaGetErrorString(e);
        return -2;
    }
    return 0;
}
template <class T> int launch_t(const BatchedParams& p, int dev_sms, cudaStream_t stream, std::string* err) {
    const int n = p.n;
    if (n <= 32) return launch_nt<T, 32>(p, dev_sms, stream, err);
    if (n <= 64) return launch_nt<T, 64>(p, dev_sms, stream, err);
    return launch_nt<T, 128>(p, dev_sms, stream, err);
}

}  // namespace gs

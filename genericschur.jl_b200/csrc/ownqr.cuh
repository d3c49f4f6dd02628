// Stage B of the three-stage path, third generation: "owner computes" sweeps.
//
// The chain kernels (chainqr.cuh) keep the diagonal block of a bulge step in registers and update it redundantly in
// all 32 lanes of the warp: ~45 (real) / ~40 (complex) FP64 instructions per step that serve one lane's worth of data —
// and an FP64 instruction costs the SM sub-partition 2.3 issue cycles whatever the number of useful lanes.  Measured
// (ncu, r02a): 64x64 Float64 with 12 warps per SM is bound by exactly that issue time, 64x64 ComplexF64 with 6 warps
// per SM spends two thirds of a step's 630 cycles issuing its own 240 instructions.
// Here nothing is replicated except the reflector itself.  A step (reflector k, rows / columns k..k+2 real, k..k+1
// complex) is:
//     block-L   the lanes that own columns k..k+2 apply the reflector from the left to rows k..k+2 of their column
//               (shared memory -> registers -> shared memory),
//     __syncwarp
//     block-R   the lanes that own rows k..k+3 apply it from the right to columns k..k+2 of their row; this creates
//               the bulge in column k,
//     __syncwarp
//     all lanes read the bulge (three / two entries) and form reflector k+1 (the only replicated arithmetic),
// while the FAR items — columns right of the block, rows above it — run with register carries exactly as in the chain
// kernels; they depend on nothing but the reflector and fill the stall slots of the serial part.  Per step: 40 instead of
// ~90 FP64 instructions (real), 69 instead of 92 (complex), two warp-level barriers, no register-resident block: the
// start and the end of a sweep need no load / write-back of a block.
// Decision rules, shifts, deflation, fix-ups, the log: unchanged (drivers inherited from ChainR / ChainC).
// The arithmetic per entry is the reference's (src/GenericSchur.jl:877-946 real, :426-460 complex); every entry sees
// the same operations in the same order.
#pragma once
#include <type_traits>
#include "chainqr.cuh"

namespace gs {

// =====================================================================================================================
// Float64, double shift
// =====================================================================================================================
template <int CPL> struct OwnR : ChainR<32, CPL> {
    typedef ChainR<32, CPL> B;
    typedef double R;
    typedef cx<double> C;
    static constexpr int EX = 3;
    static constexpr uint32_t ES = 8;

    // One Francis double-shift sweep (src/GenericSchur.jl:837-952).  Window sizes >= 3 (the driver deflates smaller ones).
    GS_DEV void sweep(bool want, R r1r, R r1i, R r2r, R r2i) {
        const R eps = 2.220446049250313e-16;
        const int n = this->n, sub = this->sub, iend = this->iend, istart = this->istart;
        const uint32_t hb = this->hb;
        int mx = 1, len = 0;
        bool scale_sub = false;
        uint32_t ca[CPL], ib[CPL];
        int jl[CPL], jr[CPL];
        R c1[CPL], c2[CPL];
        R tau1 = 0, tau2 = 0, tau3 = 0, v1 = 0, v2 = 0, beta = 0;
        uint32_t ak = hb;
#pragma unroll
        for (int s = 0; s < CPL; ++s) {
            const int j = sub + 1 + 32 * s;
            const bool valid = j <= n;
            jl[s] = valid ? j : -(1 << 28);
            jr[s] = valid ? j : (1 << 28);
            ca[s] = hb + ES * (uint32_t)(B::colbase(valid ? j : 1) - 1);
            ib[s] = ES * (uint32_t)(valid ? j : 1);
            c1[s] = 0.0;
            c2[s] = 0.0;
        }
        if (want) {
            int m0 = 0;
            for (int base = iend - 2; base >= istart + 1 && !m0; base -= 32) {
                const int m = base - sub;
                bool hit = false;
                if (m >= istart + 1) {
                    R a0, a1, a2;
                    this->first_column(m, r1r, r1i, r2r, r2i, a0, a1, a2);
                    hit = fabs(this->ld(m, m - 1)) * (fabs(a1) + fabs(a2)) <=
                          eps * fabs(a0) * (fabs(this->ld(m - 1, m - 1)) + fabs(this->ld(m, m)) + fabs(this->ld(m + 1, m + 1)));
                }
                const unsigned msk = this->sballot(hit);
                if (msk) m0 = base - (__ffs(msk) - 1);
            }
            mx = m0 ? m0 : istart;
            R v0;
            this->first_column(mx, r1r, r1i, r2r, r2i, v0, v1, v2);
            scale_sub = mx > istart;
            len = iend - 1 - mx;          // three-row reflectors at k = mx .. iend-2; the two-row one follows the loop
            this->st[1] += (unsigned)(iend - mx);
            ak = hb + ES * (uint32_t)(B::colbase(mx) - 1);
            const uint32_t ak1 = ak + ES * (uint32_t)(mx + EX);
#pragma unroll
            for (int s = 0; s < CPL; ++s) {
                if (jl[s] >= mx + 3) {
                    c1[s] = lds_e<R>(ca[s] + ES * mx);
                    c2[s] = lds_e<R>(ca[s] + ES * (mx + 1));
                } else if (jr[s] <= mx - 1) {
                    c1[s] = lds_e<R>(ak + ib[s]);
                    c2[s] = lds_e<R>(ak1 + ib[s]);
                }
            }
            tau1 = reflector_real_small(v0, v1, v2, 3);
            beta = v0;
            tau2 = tau1 * v1;
            tau3 = tau1 * v2;
            this->lg.put_hdr(LOG_REFL3, mx, len, iend, 0.0, 0.0);
        }
        if (want && scale_sub) {
            // the sweep starts inside the block: the entry left of it is scaled by the first reflector (src/GenericSchur.jl:896-899)
            this->ssync();
            if (sub == 0) this->stc(mx, mx - 1, this->ld(mx, mx - 1) * (1.0 - tau1));
        }
        this->ssync();
#ifdef GS_QR_PROFILE
        const long long tl0 = clock64();
#endif
        int t = 0;
        R f10 = 0, f20 = 0, f30 = 0;
        auto& lg = this->lg;
        for (;;) {
            if (t >= len) break;
            int chunk = len - t;
            if (lg.on && !lg.ovf) {
                if (lg.left == 0) lg.new_page();
                if (!lg.ovf && lg.left < chunk) chunk = lg.left;
            }
            const bool logp = lg.on && !lg.ovf && sub == 0;
            unsigned char* lp = lg.cur;
            bool ok = true;
            int i = 0;
            // With two columns / rows per lane (n > 32) the slot of the high ones (j >= 33) is a plain column item of every lane
            // while k <= KLO, the slot of the low ones (j <= 32) a plain row item once k >= KHI: the step loop is instantiated
            // for the three regimes, and the constant roles drop out of two thirds of the steps (the kernel is issue bound).
            constexpr int KLO = 29, KHI = 34;
            int mode = 0;
            if constexpr (CPL == 2) {
                const int kc = mx + t;
                if (kc <= KLO) {
                    mode = 1;
                    if (chunk > KLO - kc + 1) chunk = KLO - kc + 1;
                } else if (kc >= KHI) {
                    mode = 2;
                } else if (chunk > KHI - kc) {
                    chunk = KHI - kc;
                }
            }
            auto steps = [&](auto MODE_) {
            constexpr int MODE = decltype(MODE_)::value;
            // slot s has a constant role in this regime: LO = column item (never in the block), HI = row item
            auto is_lo = [](int s) { return MODE == 1 && s == CPL - 1; };
            auto is_hi = [](int s) { return MODE == 2 && s == 0; };
#pragma unroll 1
            for (; i < chunk && ok; ++i) {
                const int k = mx + t + i;
                const uint32_t kb = ES * (uint32_t)k;
                const uint32_t ak1 = ak + ES * (uint32_t)(k + EX);        // column k+1
                const uint32_t ak2 = ak1 + ES * (uint32_t)(k + 1 + EX);   // column k+2
                // ---- block-L: column j in k..k+2 (at most one of a lane's columns), rows k..k+2 ----
                {
                    uint32_t ba = ca[0] + kb;
                    bool pL = (is_lo(0) || is_hi(0)) ? false : ((unsigned)(jl[0] - k) <= 2u);
#pragma unroll
                    for (int s = 1; s < CPL; ++s) {
                        const bool in = (is_lo(s) || is_hi(s)) ? false : ((unsigned)(jl[s] - k) <= 2u);
                        ba = (is_hi(0) && s == 1) ? ca[s] + kb : (in ? ca[s] + kb : ba);
                        pL = pL || in;
                    }
                    const R x0 = lds_e<R>(ba), x1 = lds_e<R>(ba + ES), x2 = lds_e<R>(ba + 2 * ES);
                    const R ss = fma(v2, x2, fma(v1, x1, x0));
                    sts_f64_if(ba, fma(-ss, tau1, x0), pL);
                    sts_f64_if(ba + ES, fma(-ss, tau2, x1), pL);
                    sts_f64_if(ba + 2 * ES, fma(-ss, tau3, x2), pL);
                }
                this->ssync();
                // ---- far items: loads (third entry; the carries of a row that left the block in the previous step) ----
                uint32_t sa[CPL];
                bool act[CPL];
                R y[CPL];
#pragma unroll
                for (int s = 0; s < CPL; ++s) {
                    const bool isR = is_lo(s) ? false : (is_hi(s) ? true : (jr[s] <= k - 1));
                    act[s] = is_lo(s) ? (jl[s] > 0) : (is_hi(s) ? true : (isR || (jl[s] >= k + 3)));
                    sa[s] = isR ? ak + ib[s] : ca[s] + kb;
                    const uint32_t ya = isR ? ak2 + ib[s] : sa[s] + 2 * ES;
                    if (!is_lo(s) && !is_hi(s)) {
                        lds_f64_if(c1[s], sa[s], jr[s] == k - 1);
                        lds_f64_if(c2[s], ak1 + ib[s], jr[s] == k - 1);
                    }
                    y[s] = lds_e<R>(ya);
                }
                // ---- block-R: row i in k..min(k+3, iend) (at most one of a lane's rows), columns k..k+2 ----
                {
                    uint32_t ro = ib[0];
                    bool pR = (is_lo(0) || is_hi(0)) ? false : ((unsigned)(jr[0] - k) <= 3u && jr[0] <= iend);
#pragma unroll
                    for (int s = 1; s < CPL; ++s) {
                        const bool in = (is_lo(s) || is_hi(s)) ? false : ((unsigned)(jr[s] - k) <= 3u && jr[s] <= iend);
                        ro = (is_hi(0) && s == 1) ? ib[s] : (in ? ib[s] : ro);
                        pR = pR || in;
                    }
                    const R y0 = lds_e<R>(ak + ro), y1 = lds_e<R>(ak1 + ro), y2 = lds_e<R>(ak2 + ro);
                    const R tt = fma(v2, y2, fma(v1, y1, y0));
                    sts_f64_if(ak + ro, fma(-tt, tau1, y0), pR);
                    sts_f64_if(ak1 + ro, fma(-tt, tau2, y1), pR);
                    sts_f64_if(ak2 + ro, fma(-tt, tau3, y2), pR);
                }
                {   // column k-1 below the diagonal: (beta, 0, 0) — its bulge entries were read at the end of the previous step
                    const uint32_t akm = ak - ES * (uint32_t)(k - 1 + EX);
                    const bool psub = sub == 0 && (t + i > 0);
                    sts_f64_if(akm + kb, beta, psub);
                    sts_f64_if(akm + kb + ES, 0.0, psub);
                    sts_f64_if(akm + kb + 2 * ES, 0.0, psub);
                    stg_2f64_if(lp, tau1, v1, logp);
                    stg_2f64_if(lp + 16, v2, 0.0, logp);
                    lp += 32;
                }
                this->ssync();
                // ---- the bulge: H[k+1..k+3, k] ----
                f10 = lds_e<R>(ak + kb + ES);
                f20 = lds_e<R>(ak + kb + 2 * ES);
                f30 = lds_e<R>(ak + kb + 3 * ES);
                // ---- far items: arithmetic and stores (independent of the block: they fill the stall slots below) ----
                const R otau1 = tau1, otau2 = tau2, otau3 = tau3, ov1 = v1, ov2 = v2;
                // ---- reflector k+1 from (f10, f20, f30) (src/householder.jl:12-54): beta = -sign(a) ||x||, tau = 1 - a / beta,
                //      tau v_i = -x_i / beta, v_i = (tau v_i) / tau.  (After the last step its result is not used.) ----
                {
                    const double q = fma(f10, f10, fma(f20, f20, f30 * f30));
                    const unsigned tz = ((unsigned)(__double2hiint(f20) | __double2hiint(f30)) << 1) |
                                        (unsigned)(__double2loint(f20) | __double2loint(f30));
                    ok = (q_exp_in(q, 1023u - 900u, 1023u + 900u) && (tz != 0u)) || (t + i + 1 >= len);
                    double yr0;
                    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(yr0) : "d"(q));
                    const double qy = q * yr0;
                    const double e = fma(-qy, yr0, 1.0);
                    const double cf = fma(e, 0.375, 0.5);
                    const double yr = fma(yr0 * e, cf, yr0);   // 1/sqrt(q), one cubic step
                    const double rb = -copysign(yr, f10);      // 1/beta
                    beta = -copysign(q * yr, f10);
                    tau1 = fma(-f10, rb, 1.0);                 // in [1, 2]
                    tau2 = -f20 * rb;
                    tau3 = -f30 * rb;
                    double y0;
                    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(tau1));
                    const double e2 = fma(-tau1, y0, 1.0);
                    const double rt = fma(y0, fma(e2, e2, e2), y0);
                    v1 = tau2 * rt;
                    v2 = tau3 * rt;
                }
#pragma unroll
                for (int s = 0; s < CPL; ++s) {
                    const R ss = fma(ov2, y[s], fma(ov1, c2[s], c1[s]));
                    const R sv = fma(-ss, otau1, c1[s]);
                    c1[s] = fma(-ss, otau2, c2[s]);
                    c2[s] = fma(-ss, otau3, y[s]);
                    sts_f64_if(sa[s], sv, act[s]);
                    if (!is_lo(s) && !is_hi(s)) {
                        const bool own3 = jl[s] == k + 3;      // column k+3 enters the block next step
                        sts_f64_if(sa[s] + ES, c1[s], own3);
                        sts_f64_if(sa[s] + 2 * ES, c2[s], own3);
                    }
                }
                ak = ak1;
            }
            };
            if (mode == 1) steps(std::integral_constant<int, 1>{});
            else if (mode == 2) steps(std::integral_constant<int, 2>{});
            else steps(std::integral_constant<int, 0>{});
            if (lg.on && !lg.ovf) {
                lg.cur += 32 * i;
                lg.left -= i;
                lg.nrec += i;
            }
            t += i;
            if (!ok) {   // out-of-range or degenerate input: the general routine forms the reflector of step t
                R w0 = f10, w1 = f20, w2 = f30;
                tau1 = reflector_real_small(w0, w1, w2, 3);
                beta = w0;
                v1 = w1;
                v2 = w2;
                tau2 = tau1 * v1;
                tau3 = tau1 * v2;
            }
        }
#ifdef GS_QR_PROFILE
        this->prof_loop += clock64() - tl0;
#endif
        if (want) {
            // ---- last step: the two-row reflector at k = iend-1 (src/GenericSchur.jl:927-946).  The far items are on
            //      registers (rows iend-1, iend of the lanes' columns / columns iend-1, iend of their rows); the block
            //      and row iend-2 are read from shared memory by every lane and written back by lane 0 ----
            const int k = iend - 1;
            this->ssync();
            const R L01 = this->ld(k - 1, k), L02 = this->ld(k - 1, k + 1);
            const R L11 = this->ld(k, k), L12 = this->ld(k, k + 1), L21 = this->ld(k + 1, k), L22 = this->ld(k + 1, k + 1);
            R w0 = f10, w1 = f20, w2 = 0.0;
            const R t1 = reflector_real_small(w0, w1, w2, 2);
            const R t2 = t1 * w1;
            lg.put_hdr(LOG_REFL2, k, 1, k, 0.0, 0.0);
            lg.put4(t1, w1, 0.0, 0.0);
            const R sa0 = L11 + w1 * L21, sa1 = L12 + w1 * L22;
            const R g11 = L11 - sa0 * t1, g21 = L21 - sa0 * t2, g12 = L12 - sa1 * t1, g22 = L22 - sa1 * t2;
            const R sr0 = L01 + w1 * L02, sr1 = g11 + w1 * g12, sr2 = g21 + w1 * g22;
            this->ssync();
#pragma unroll
            for (int s = 0; s < CPL; ++s) {
                const R ss = c1[s] + w1 * c2[s];
                const R n1 = c1[s] - ss * t1, n2 = c2[s] - ss * t2;
                if (jl[s] >= iend + 1) {
                    this->stc(iend - 1, jl[s], n1);
                    this->stc(iend, jl[s], n2);
                }
                if (jr[s] <= iend - 3) {
                    this->stc(jr[s], iend - 1, n1);
                    this->stc(jr[s], iend, n2);
                }
            }
            if (sub == 0) {
                this->stc(k, k - 1, w0);
                this->stc(k + 1, k - 1, 0.0);
                this->stc(k - 1, k, L01 - sr0 * t1);
                this->stc(k - 1, k + 1, L02 - sr0 * t2);
                this->stc(k, k, g11 - sr1 * t1);
                this->stc(k, k + 1, g12 - sr1 * t2);
                this->stc(k + 1, k, g21 - sr2 * t1);
                this->stc(k + 1, k + 1, g22 - sr2 * t2);
            }
            this->ssync();
        }
    }

    GS_DEV void run() {
#ifdef GS_QR_PROFILE
        this->prof_loop = 0;
        this->prof_ns = 0;
        const long long tr0 = clock64();
#endif
        for (;;) {
            R r1r = 0, r1i = 0, r2r = 0, r2i = 0;
            bool want = false;
#ifdef GS_QR_PROFILE
            const long long tn0 = clock64();
#endif
            if (this->alive) want = this->next_sweep(r1r, r1i, r2r, r2i);
            __syncwarp();
#ifdef GS_QR_PROFILE
            this->prof_ns += clock64() - tn0;
#endif
            if (!want) break;
            sweep(want, r1r, r1i, r2r, r2i);
            if (this->lg.ovf) {
                this->alive = false;
                this->info = LOG_OVERFLOW_RC;
            }
        }
        this->st[3] = (unsigned)this->it;
#ifdef GS_QR_PROFILE
        this->st[0] = (unsigned)((clock64() - tr0) >> 6);
        this->st[2] = (unsigned)(this->prof_ns >> 6);
        this->st[3] = (unsigned)(this->prof_loop >> 6);
#endif
    }
};

// =====================================================================================================================
// ComplexF64, single shift
// =====================================================================================================================
template <int CPL> struct OwnC : ChainC<32, CPL> {
    typedef ChainC<32, CPL> B;
    typedef double R;
    typedef cx<double> C;
    static constexpr int EX = 2;
    static constexpr uint32_t ES = 16;

    GS_DEV static double flip_if(double x, unsigned m) { return B::flip_if(x, m); }

    // One single-shift sweep (src/GenericSchur.jl:374-504)
    GS_DEV void sweep(bool want, const C& shift) {
        const R ulp = 2.220446049250313e-16;
        const int n = this->n, sub = this->sub, iend = this->iend, istart = this->istart;
        const uint32_t hb = this->hb;
        auto& lg = this->lg;
        int kf = 1, len = 0;
        bool store_sub = false;
        uint32_t ca[CPL], ib[CPL];
        int jl[CPL], jr[CPL];
        C c[CPL];
        C tau1 = mk_cx<R>(0.0, 0.0), v2 = tau1, nv0 = tau1;
        R tau2 = 0.0, beta = 0.0, nv1 = 0.0;
        uint32_t ak = hb;
#pragma unroll
        for (int s = 0; s < CPL; ++s) {
            const int j = sub + 1 + 32 * s;
            const bool valid = j <= n;
            jl[s] = valid ? j : -(1 << 28);
            jr[s] = valid ? j : (1 << 28);
            ca[s] = hb + ES * (uint32_t)(B::colbase(valid ? j : 1) - 1);
            ib[s] = ES * (uint32_t)(valid ? j : 1);
            c[s] = mk_cx<R>(0.0, 0.0);
        }
#ifdef GS_QR_PROFILE
        const long long ty0 = clock64();
        long long ty1 = ty0;
#endif
        if (want) {
            // ---- start row (src/GenericSchur.jl:390-420) ----
            int istart1 = 0;
            for (int base = iend - 1; base >= istart + 1 && !istart1; base -= 32) {
                const int mm = base - sub;
                bool hit = false;
                if (mm >= istart + 1) {
                    const C h11 = this->ld(mm, mm), h22 = this->ld(mm + 1, mm + 1);
                    const C h11s = h11 - shift;
                    const R h21 = this->ld(mm + 1, mm).re;
                    const R rs = q_rcp(abs1(h11s) + fabs(h21));
                    const R h10 = this->ld(mm, mm - 1).re;
                    hit = fabs(h10) * fabs(h21 * rs) <= ulp * ((fabs(h11s.re * rs) + fabs(h11s.im * rs)) * (abs1(h11) + abs1(h22)));
                }
                const unsigned m = this->sballot(hit);
                if (m) istart1 = base - (__ffs(m) - 1);
            }
            if (!istart1) istart1 = istart;
#ifdef GS_QR_PROFILE
            ty1 = clock64();
#endif
            const int k0 = istart1;
            C v0, v1;
            {
                const C h11s = this->ld(k0, k0) - shift;
                const R h21 = this->ld(k0 + 1, k0).re;
                const R rs = q_rcp(abs1(h11s) + fabs(h21));
                v0 = mk_cx<R>(h11s.re * rs, h11s.im * rs);
                v1 = mk_cx<R>(h21 * rs, 0.0);
            }
            kf = k0;
            unsigned napplied = 0;
            if (k0 > istart) {
                this->late_start_step(k0, v0, v1);
                napplied = 1;
                kf = k0 + 1;
                store_sub = true;
                if (kf <= iend - 1) {
                    v0 = this->ld(kf, kf - 1);
                    v1 = this->ld(kf + 1, kf - 1);
                    this->ssync();
                    if (sub == 0) this->stc(kf + 1, kf - 1, mk_cx<R>(0.0, 0.0));
                }
            }
            if (kf > iend - 1) {
                // the late-start step was the only one
                this->st[1] += napplied;
                this->tail_fix_smem();
            } else {
                len = iend - kf;
                this->st[1] += napplied + (unsigned)len;
                ak = hb + ES * (uint32_t)(B::colbase(kf) - 1);
#pragma unroll
                for (int s = 0; s < CPL; ++s) {
                    if (jl[s] >= kf + 2) c[s] = lds_e<C>(ca[s] + ES * kf);
                    else if (jr[s] <= kf - 1) c[s] = cconj(lds_e<C>(ak + ib[s]));
                }
                tau1 = reflector_cplx2(v0, v1);
                beta = v0.re;
                v2 = v1;
                tau2 = tau1.re * v2.re - tau1.im * v2.im;
                lg.put_hdr(LOG_REFL, kf, len, iend, 0.0, 0.0);
            }
        }
        this->ssync();
#ifdef GS_QR_PROFILE
        const long long tl0 = clock64();
        this->prof_x[2] += ty1 - ty0;
        this->prof_x[3] += tl0 - ty1;
#endif
        int t = 0;
        // The operands of block-L are fetched one step ahead (right after the barrier that publishes block-R of the previous
        // step, next to the bulge loads), so their latency sits under the reflector chain; the entry H[k+1, k+2] that the far
        // part of the previous step produces for the same lane comes from its carry register instead of shared memory.
        C px0 = mk_cx<R>(0.0, 0.0), px1 = px0;
        if (len > 0) {
            uint32_t ba = ca[0] + ES * (uint32_t)kf;
#pragma unroll
            for (int s = 1; s < CPL; ++s) ba = ((unsigned)(jl[s] - kf) <= 1u) ? ca[s] + ES * (uint32_t)kf : ba;
            px0 = lds_e<C>(ba);
            px1 = lds_e<C>(ba + ES);
        }
        for (;;) {
            if (t >= len) break;
            int chunk = len - t;
            if (lg.on && !lg.ovf) {
                if (lg.left == 0) lg.new_page();
                if (!lg.ovf && lg.left < chunk) chunk = lg.left;
            }
            const bool logp = lg.on && !lg.ovf && sub == 0;
            unsigned char* lp = lg.cur;
            bool ok = true;
            int i = 0;
            // three instances of the step loop by the regime of k (see OwnR::sweep): the slot of the high columns / rows is a
            // plain column item while k <= KLO, the slot of the low ones a plain row item once k >= KHI
            constexpr int KLO = 30, KHI = 34;
            int mode = 0;
            if constexpr (CPL == 2) {
                const int kc = kf + t;
                if (kc <= KLO) {
                    mode = 1;
                    if (chunk > KLO - kc + 1) chunk = KLO - kc + 1;
                } else if (kc >= KHI) {
                    mode = 2;
                } else if (chunk > KHI - kc) {
                    chunk = KHI - kc;
                }
            }
            auto steps = [&](auto MODE_) {
            constexpr int MODE = decltype(MODE_)::value;
            auto is_lo = [](int s) { return MODE == 1 && s == CPL - 1; };
            auto is_hi = [](int s) { return MODE == 2 && s == 0; };
#pragma unroll 1
            for (; i < chunk && ok; ++i) {
                const int k = kf + t + i;
                const uint32_t kb = ES * (uint32_t)k;
                const uint32_t ak1 = ak + ES * (uint32_t)(k + EX);   // column k+1
                // ---- block-L: column j in {k, k+1} (at most one of a lane's columns), rows k, k+1 ----
                {
                    uint32_t ba = ca[0] + kb;
                    bool pL = (is_lo(0) || is_hi(0)) ? false : ((unsigned)(jl[0] - k) <= 1u);
#pragma unroll
                    for (int s = 1; s < CPL; ++s) {
                        const bool in = (is_lo(s) || is_hi(s)) ? false : ((unsigned)(jl[s] - k) <= 1u);
                        ba = (is_hi(0) && s == 1) ? ca[s] + kb : (in ? ca[s] + kb : ba);
                        pL = pL || in;
                    }
                    const C x0 = px0, x1 = px1;
                    const C ss = mk_cx<R>(fma(tau1.re, x0.re, fma(tau1.im, x0.im, tau2 * x1.re)),
                                          fma(tau1.re, x0.im, fma(-tau1.im, x0.re, tau2 * x1.im)));
                    sts_c64_if(ba, x0 - ss, pL);
                    sts_c64_if(ba + ES, e_fnma(ss, v2, x1), pL);
                }
                this->ssync();
                // ---- far items: loads.  Index j is a LEFT item (column j, rows k, k+1) while j >= k+2 and a RIGHT item (row
                //      j, columns k, k+1; held conjugated so that one instruction stream serves both) once j <= k-1 ----
                uint32_t sa[CPL];
                unsigned sg[CPL];
                bool act[CPL];
                C y[CPL];
#pragma unroll
                for (int s = 0; s < CPL; ++s) {
                    const bool isR = is_lo(s) ? false : (is_hi(s) ? true : (jr[s] <= k - 1));
                    act[s] = is_lo(s) ? (jl[s] > 0) : (is_hi(s) ? true : (isR || (jl[s] >= k + 2)));
                    sa[s] = isR ? ak + ib[s] : ca[s] + kb;
                    const uint32_t ya = isR ? ak1 + ib[s] : sa[s] + ES;
                    sg[s] = isR ? 0x80000000u : 0u;
                    if (!is_lo(s) && !is_hi(s)) {
                        const bool ent = jr[s] == k - 1;          // row k-1 left the block in the previous step
                        lds_c64_if(c[s], sa[s], ent);
                        c[s].im = flip_if(c[s].im, ent ? 0x80000000u : 0u);
                    }
                    y[s] = lds_e<C>(ya);
                    if (!is_lo(s)) y[s].im = flip_if(y[s].im, sg[s]);
                }
                // ---- block-R: row i in k..min(k+2, iend) (at most one of a lane's rows), columns k, k+1 ----
                {
                    uint32_t ro = ib[0];
                    bool pR = (is_lo(0) || is_hi(0)) ? false : ((unsigned)(jr[0] - k) <= 2u && jr[0] <= iend);
#pragma unroll
                    for (int s = 1; s < CPL; ++s) {
                        const bool in = (is_lo(s) || is_hi(s)) ? false : ((unsigned)(jr[s] - k) <= 2u && jr[s] <= iend);
                        ro = (is_hi(0) && s == 1) ? ib[s] : (in ? ib[s] : ro);
                        pR = pR || in;
                    }
                    const C y0 = lds_e<C>(ak + ro), y1 = lds_e<C>(ak1 + ro);
                    const C sr = mk_cx<R>(fma(tau1.re, y0.re, fma(-tau1.im, y0.im, tau2 * y1.re)),
                                          fma(tau1.re, y0.im, fma(tau1.im, y0.re, tau2 * y1.im)));
                    sts_c64_if(ak + ro, y0 - sr, pR);
                    sts_c64_if(ak1 + ro, e_fnma_cjb(sr, v2, y1), pR);
                }
                {   // column k-1 below the diagonal: (beta, 0) — its bulge entry was read at the end of the previous step
                    const uint32_t akm = ak - ES * (uint32_t)(k - 1 + EX);
                    const bool psub = sub == 0 && (t + i > 0 || store_sub);
                    sts_c64_if(akm + kb, mk_cx<R>(beta, 0.0), psub);
                    sts_c64_if(akm + kb + ES, mk_cx<R>(0.0, 0.0), psub);
                    stg_2f64_if(lp, tau1.re, tau1.im, logp);
                    stg_2f64_if(lp + 16, v2.re, v2.im, logp);
                    lp += 32;
                }
                this->ssync();
                // ---- the bulge: H[k+1, k] (complex), H[k+2, k] (real) ----
                nv0 = lds_e<C>(ak + kb + ES);
                nv1 = lds_e<R>(ak + kb + 2 * ES);
                {   // block-L operands of step k+1: column j in {k+1, k+2}, rows k+1, k+2
                    uint32_t bn = ca[0] + kb + ES;
#pragma unroll
                    for (int s = 1; s < CPL; ++s)
                        bn = (is_hi(0) && s == 1) ? ca[s] + kb + ES
                                                  : ((!is_lo(s) && (unsigned)(jl[s] - k - 1) <= 1u) ? ca[s] + kb + ES : bn);
                    px0 = lds_e<C>(bn);
                    px1 = lds_e<C>(bn + ES);
                }
                const C otau1 = tau1, ov2 = v2;
                const R otau2 = tau2;
                // ---- reflector k+1 (src/householder.jl:56-102) from (nv0, nv1): beta = -sign(Re a) ||.||,
                //      tau = 1 - a / beta, tau2 = Re(tau v2) = -x2 / beta, v2 = tau2 conj(tau) / |tau|^2 (= x2 / (a - beta)) ----
                {
                    const double a = nv0.re, b = nv0.im, cc = nv1;
                    const double q = fma(a, a, fma(b, b, cc * cc));
                    const unsigned tz = ((unsigned)(__double2hiint(cc) | __double2hiint(b)) << 1) |
                                        (unsigned)(__double2loint(cc) | __double2loint(b));
                    ok = (q_exp_in(q, 1023u - 900u, 1023u + 900u) && (tz != 0u)) || (t + i + 1 >= len);
                    double yr0;
                    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(yr0) : "d"(q));
                    const double qy = q * yr0;
                    const double e = fma(-qy, yr0, 1.0);
                    const double cf = fma(e, 0.375, 0.5);
                    const double yr = fma(yr0 * e, cf, yr0);   // 1/sqrt(q), one cubic step
                    const double rb = -copysign(yr, a);        // 1/beta
                    beta = -copysign(q * yr, a);
                    tau1 = mk_cx<R>(fma(-a, rb, 1.0), -b * rb);
                    tau2 = -cc * rb;
                    const double m2 = fma(tau1.re, tau1.re, tau1.im * tau1.im);   // |tau|^2 in [1, 4]
                    double y0;
                    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(m2));
                    const double e2 = fma(-m2, y0, 1.0);
                    const double rm = fma(y0, fma(e2, e2, e2), y0);
                    const double sc = tau2 * rm;
                    v2 = mk_cx<R>(sc * tau1.re, -sc * tau1.im);
                }
                // ---- far items: arithmetic and stores ----
#pragma unroll
                for (int s = 0; s < CPL; ++s) {
                    const C x = c[s];
                    const C ss = mk_cx<R>(fma(otau1.re, x.re, fma(otau1.im, x.im, otau2 * y[s].re)),
                                          fma(otau1.re, x.im, fma(-otau1.im, x.re, otau2 * y[s].im)));
                    C sv = x - ss;
                    if (!is_lo(s)) sv.im = flip_if(sv.im, sg[s]);
                    c[s] = e_fnma(ss, ov2, y[s]);
                    sts_c64_if(sa[s], sv, act[s]);
                    // H[k+1, k+2] enters the block next step: handed to block-L in a register (its shared-memory copy is
                    // rewritten by block-L of step k+1 or, after the last step, by the sweep's write-back of the carries)
                    if (!is_lo(s) && !is_hi(s)) {
                        const bool own2 = jl[s] == k + 2;
                        px0.re = own2 ? c[s].re : px0.re;
                        px0.im = own2 ? c[s].im : px0.im;
                    }
                }
                ak = ak1;
            }
            };
            if (mode == 1) steps(std::integral_constant<int, 1>{});
            else if (mode == 2) steps(std::integral_constant<int, 2>{});
            else steps(std::integral_constant<int, 0>{});
            if (lg.on && !lg.ovf) {
                lg.cur += 32 * i;
                lg.left -= i;
                lg.nrec += i;
            }
            t += i;
            if (!ok) {   // out-of-range or degenerate input: the general routine forms the reflector of step t
                C w0 = nv0, w1 = mk_cx<R>(nv1, 0.0);
                tau1 = reflector_cplx2_generic<R>(w0, w1);
                beta = w0.re;
                v2 = w1;
                tau2 = tau1.re * v2.re - tau1.im * v2.im;
            }
        }
#ifdef GS_QR_PROFILE
        const long long tl1 = clock64();
        this->prof_loop += tl1 - tl0;
#endif
        if (want && len > 0) {
            // ---- the far carries are written back after the last step (k = iend-1); the unit-modulus factor that makes
            //      H[iend, iend-1] real (src/GenericSchur.jl:486-500) is applied on the way ----
            C tph = mk_cx<R>(1.0, 0.0);
            R fsr = nv0.re;
            const bool fix = nv0.im != 0.0;
            if (fix) {
                fsr = c_abs_q(nv0);
                const R ri = q_rcp(fsr);
                tph = mk_cx<R>(nv0.re * ri, nv0.im * ri);
                lg.put_hdr(LOG_SCALE, iend, 0, iend, tph.re, tph.im);
            }
            const C tphc = cconj(tph);
            this->ssync();
            const C hrow = this->ld(iend - 1, iend);     // row iend-1 was a block row of the last step
#pragma unroll
            for (int s = 0; s < CPL; ++s) {
                C v = c[s];
                if (fix) v = v * tphc;
                if (jl[s] >= iend + 1) this->stc(iend, jl[s], v);
                if (jr[s] <= iend - 2) this->stc(jr[s], iend, cconj(v));
            }
            if (sub == 0) {
                this->stc(iend, iend - 1, mk_cx<R>(fsr, 0.0));
                if (fix) this->stc(iend - 1, iend, hrow * tph);
            }
            this->ssync();
        }
#ifdef GS_QR_PROFILE
        this->prof_x[4] += clock64() - tl1;
#endif
    }

    GS_DEV void run() {
#ifdef GS_QR_PROFILE
        for (int q = 0; q < 6; ++q) this->prof_x[q] = 0;
        this->prof_loop = 0;
        this->prof_ns = 0;
        const long long tr0 = clock64();
#endif
        for (;;) {
            C shift = mk_cx<R>(0.0, 0.0);
            bool want = false;
#ifdef GS_QR_PROFILE
            const long long tn0 = clock64();
#endif
            if (this->alive) want = this->next_sweep(shift);
            __syncwarp();
#ifdef GS_QR_PROFILE
            this->prof_ns += clock64() - tn0;
#endif
            if (!want) break;
            sweep(want, shift);
            this->its += 1;
            if (this->lg.ovf) {
                this->alive = false;
                this->info = LOG_OVERFLOW_RC;
            }
        }
        this->st[3] = (unsigned)this->it;
#ifdef GS_QR_PROFILE
        this->st[0] = (unsigned)((clock64() - tr0) >> 6);
        this->st[2] = (unsigned)(this->prof_ns >> 6);
        this->st[3] = (unsigned)(this->prof_loop >> 6);
        if (this->sub == 0 && blockIdx.x == 5)
            printf("profile matrix (CTA 5): scan %lld shift %lld | start-row %lld first-reflector %lld epilogue %lld | loops %lld total %lld\n",
                   this->prof_x[0], this->prof_x[1], this->prof_x[2], this->prof_x[3], this->prof_x[4], this->prof_loop, clock64() - tr0);
#endif
    }
};

template <int CPL> struct chain_traits<cx<double>, 32, CPL, 1> {
    typedef OwnC<CPL> Solver;
};
template <int CPL> struct chain_traits<double, 32, CPL, 1> {
    typedef OwnR<CPL> Solver;
};

}  // namespace gs

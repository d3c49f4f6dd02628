// Stage B of the three-stage path, second generation: the "chain" kernels.
//
// What bounds stage B is not FP64 throughput but ISSUE SLOTS spent on the serial chain of the sweep
// (reflector k -> 2x2 block -> reflector k+1), which a warp computes redundantly in all of its lanes: per bulge step a
// 64x64 ComplexF64 matrix has ~25 warp-instructions of useful bulk FP64 work (the row / column pairs of H) against ~60
// FP64 instructions of chain.  Two consequences shape this file:
//   * SLOTS.  A warp is split into 32 / LPM slots of LPM lanes and each slot works on its own matrix.  The slots run the
//     same instruction stream, so one chain instruction serves 32 / LPM matrices; the bulk items of a matrix are spread over
//     its LPM lanes (CPL = ceil(n / LPM) indices per lane).  All cross-lane operations (ballots, shuffles, syncs) are
//     confined to the slot, control flow that depends on a matrix's state diverges per slot and re-converges by
//     itself; the step loop has a per-slot trip count, so the slots stay converged for the common iterations.
//   * A LEAN STEP.  Per bulge step: 39 FP64 operations for the block, 20 for the next reflector (tau from 1/beta,
//     v2 = tau2 conj(tau1) / |tau1|^2 — no second norm), 12 per bulk item; the row of the block that leaves it is handed
//     to its owner lane through shared memory (predicated 16-byte loads instead of register selects); inactive items
//     are not masked but left to compute garbage that is never stored.
// Decision rules, shifts, deflation, fix-ups: src/GenericSchur.jl:194-335, 374-504 (complex), 513-699, 837-952 (real) —
// unchanged.  Every right-hand transformation is logged for stage C (qrlog.cuh).
#pragma once
#include "fastqr.cuh"

namespace gs {

template <int LPM> struct SlotGeom {
    int sub, sbase;
    unsigned smask;
    GS_DEV void init(int lane) {
        sub = lane % LPM;
        sbase = lane - sub;
        smask = (LPM == 32) ? 0xffffffffu : (((1u << (LPM & 31)) - 1u) << sbase);
    }
};

// =====================================================================================================================
// ComplexF64, single shift
// =====================================================================================================================
template <int LPM, int CPL> struct ChainC {
    typedef double R;
    typedef cx<double> C;
    static constexpr int EX = 2;
    static constexpr uint32_t ES = 16;
    int n, sub, sbase;
    unsigned smask;
    uint32_t hb;   // shared byte address of the packed matrix: &H(i, j) = hb + ES * (colbase(j) + i - 1)
    LogWriter<double> lg;
    unsigned st[4];
    // driver state (uniform across the slot)
    int istart, iend, its, it, maxiter, maxinner, info;
    bool alive;
#ifdef GS_QR_PROFILE
    long long prof_loop, prof_ns, prof_x[6];
#endif

    __host__ __device__ static int colbase(int j) { return ((j - 1) * (j + 2 * EX)) / 2; }
    __host__ __device__ static int packed_elems(int n) { return (n * (n + 1)) / 2 + EX * n; }
    GS_DEV uint32_t ad(int i, int j) const { return hb + ES * (uint32_t)(colbase(j) + i - 1); }
    GS_DEV C ld(int i, int j) const { return lds_e<C>(ad(i, j)); }
    GS_DEV void stc(int i, int j, const C& v) const { sts_e<C>(ad(i, j), v); }
    GS_DEV void ssync() const {
        if constexpr (LPM == 32) __syncwarp();
        else __syncwarp(smask);
    }
    GS_DEV unsigned sballot(bool p) const { return __ballot_sync(smask, p) >> sbase; }

    // src/GenericSchur.jl:240-286 (Ahues & Tisseur)
    GS_DEV bool split_test(int c, R smallnum, R ulp) const {
        const C h10 = ld(c + 1, c);
        if (abs1(h10) <= smallnum) return true;
        const C hcc = ld(c, c), hc1 = ld(c + 1, c + 1);
        R tst = abs1(hcc) + abs1(hc1);
        if (tst == 0.0) {
            if (c - 1 >= 1) tst = tst + fabs(ld(c, c - 1).re);
            if (c + 2 <= n) tst = tst + fabs(ld(c + 2, c + 1).re);
        }
        if (fabs(h10.re) <= ulp * tst) {
            const R a1 = abs1(h10), a2 = abs1(ld(c, c + 1));
            const R ab = fmax(a1, a2), ba = fmin(a1, a2);
            const R d1 = abs1(hc1), d2 = abs1(hcc - hc1);
            const R aa = fmax(d1, d2), bb = fmin(d1, d2);
            const R rs = q_rcp(aa + ab);
            if (ba * (ab * rs) <= fmax(smallnum, ulp * (bb * (aa * rs)))) return true;
        }
        return false;
    }

    GS_DEV void begin(int n_, int maxiter_, bool have) {
        n = n_;
        maxiter = maxiter_;
        maxinner = 30 * n_;
        istart = 1;
        iend = n_;
        its = 0;
        it = 0;
        info = 0;
        alive = have;
        st[0] = st[1] = st[2] = st[3] = 0u;
    }

    // Advance the driver (src/GenericSchur.jl:226-325) until a sweep is due; returns false when the matrix is finished
    // (converged, or out of iterations: info = row where the active block ends).
    GS_DEV bool next_sweep(C& t) {
        const R ulp = 2.220446049250313e-16;
        const R smallnum = 2.2250738585072014e-308 * ((double)n / ulp);
        for (;;) {
            if (iend < 1) {
                alive = false;
                return false;
            }
            if (its > maxinner) {   // the inner loop ran out: the outer loop starts over on the same block
                istart = 1;
                its = 0;
            }
            it += 1;
            if (it > maxiter) {
                alive = false;
                info = iend;
                return false;
            }
#ifdef GS_QR_PROFILE
            const long long tx0 = clock64();
#endif
            int found = 0;
            for (int base = iend - 1; base >= istart && !found; base -= LPM) {
                const int c = base - sub;
                const bool hit = (c >= istart) ? split_test(c, smallnum, ulp) : false;
                const unsigned m = sballot(hit);
                if (m) found = base - (__ffs(m) - 1);
            }
            if (found) istart = found + 1;
            ssync();
            if (istart > 1 && sub == 0) stc(istart, istart - 1, mk_cx<R>(0.0, 0.0));
            ssync();
#ifdef GS_QR_PROFILE
            prof_x[0] += clock64() - tx0;
#endif
            if (istart >= iend) {
                iend -= 1;
                istart = 1;
                its = 0;
                continue;
            }
#ifdef GS_QR_PROFILE
            const long long tx1 = clock64();
#endif
            if (its % 30 == 10) {
                const R s = 0.75 * fabs(ld(istart + 1, istart).re);
                t = ld(istart, istart);
                t.re = t.re + s;
                st[2] += 1;
            } else if (its % 30 == 20) {
                const R s = 0.75 * fabs(ld(iend, iend - 1).re);
                t = ld(iend, iend);
                t.re = t.re + s;
                st[2] += 1;
            } else {
                t = ld(iend, iend);
                const C u = c_sqrt_q(ld(iend - 1, iend)) * c_sqrt_q(ld(iend, iend - 1));
                R s = abs1(u);
                if (s != 0.0) {
                    const C x = 0.5 * (ld(iend - 1, iend - 1) - t);
                    const R sx = abs1(x);
                    s = fmax(s, sx);
                    const R rs = q_rcp(s);
                    const C xs = mk_cx<R>(x.re * rs, x.im * rs), us = mk_cx<R>(u.re * rs, u.im * rs);
                    C y = s * c_sqrt_q(xs * xs + us * us);
                    if (sx > 0.0) {
                        const R rsx = q_rcp(sx);
                        if ((x.re * rsx) * y.re + (x.im * rsx) * y.im < 0.0) y = -y;
                    }
                    t = t - u * c_div_q(u, x + y);
                }
            }
#ifdef GS_QR_PROFILE
            prof_x[1] += clock64() - tx1 + (long long)(t.re == 1.2345e300);
#endif
            st[0] += 1;
            return true;
        }
    }

    // the first step of a sweep that starts inside the active block (src/GenericSchur.jl:461-482), in shared memory
    GS_DEV void late_start_step(int k0, C v0, C v1) {
        const C tau1 = reflector_cplx2(v0, v1);
        const C v2 = v1, v2c = cconj(v1), tau1c = cconj(tau1);
        const R tau2 = (tau1 * v2).re;
        lg.put_hdr(LOG_REFL, k0, 1, k0, 0.0, 0.0);
        lg.put4(tau1.re, tau1.im, v2.re, v2.im);
        for (int j = k0 + sub; j <= n; j += LPM) {
            const C a = ld(k0, j), b = ld(k0 + 1, j);
            const C ss = tau1c * a + tau2 * b;
            stc(k0, j, a - ss);
            stc(k0 + 1, j, b - ss * v2);
        }
        ssync();
        const int jmax = (k0 + 2 < iend) ? k0 + 2 : iend;
        for (int i = 1 + sub; i <= jmax; i += LPM) {
            const C d = ld(i, k0), e = ld(i, k0 + 1);
            const C ss = tau1 * d + tau2 * e;
            stc(i, k0, d - ss);
            stc(i, k0 + 1, e - ss * v2c);
        }
        ssync();
        C t = mk_cx<R>(1.0, 0.0) - tau1;
        const R at = c_abs(t);
        t = mk_cx<R>(t.re / at, t.im / at);
        const C tc = cconj(t);
        if (sub == 0) {
            stc(k0 + 1, k0, ld(k0 + 1, k0) * tc);
            if (k0 + 2 <= iend) stc(k0 + 2, k0 + 1, ld(k0 + 2, k0 + 1) * t);
        }
        ssync();
        for (int j = k0; j <= iend; ++j) {
            if (j == k0 + 1) continue;
            for (int c = j + 1 + sub; c <= n; c += LPM) stc(j, c, ld(j, c) * t);
            for (int r = 1 + sub; r <= j - 1; r += LPM) stc(r, j, ld(r, j) * tc);
            ssync();
        }
        lg.put_hdr(LOG_SCALE, k0, 0, k0, tc.re, tc.im);
        if (iend >= k0 + 2) lg.put_hdr(LOG_SCALE, k0 + 2, 0, iend, tc.re, tc.im);
    }

    // make the tail sub-diagonal real in shared memory (src/GenericSchur.jl:486-500); used when a sweep has no pipelined part
    GS_DEV void tail_fix_smem() {
        C t = ld(iend, iend - 1);
        if (t.im != 0.0) {
            const R rt = c_abs(t);
            t = mk_cx<R>(t.re / rt, t.im / rt);
            const C tc = cconj(t);
            for (int c = iend + 1 + sub; c <= n; c += LPM) stc(iend, c, ld(iend, c) * tc);
            for (int r = 1 + sub; r <= iend - 1; r += LPM) stc(r, iend, ld(r, iend) * t);
            lg.put_hdr(LOG_SCALE, iend, 0, iend, t.re, t.im);
            ssync();
            if (sub == 0) stc(iend, iend - 1, mk_cx<R>(rt, 0.0));
        }
        ssync();
    }

    GS_DEV static double flip_if(double x, unsigned m) {   // x with its sign bit xor-ed by m (0 or 0x80000000)
        return __hiloint2double(__double2hiint(x) ^ (int)m, __double2loint(x));
    }

    // One single-shift sweep (src/GenericSchur.jl:374-504) on the slot's matrix; slots with want == false skip it.
    GS_DEV void sweep(bool want, const C& shift) {
        const R ulp = 2.220446049250313e-16;
        int kf = 1, len = 0;
        bool store_sub = false;
        uint32_t ca[CPL], ib[CPL];
        int jl[CPL], jr[CPL];
        C c[CPL];
        C tau1 = mk_cx<R>(0.0, 0.0), v2 = tau1, d00 = tau1, d10 = tau1, nv0 = tau1;
        R tau2 = 0.0, beta = 0.0;
        uint32_t ak = hb;
#pragma unroll
        for (int s = 0; s < CPL; ++s) {
            const int j = sub + 1 + LPM * s;
            const bool valid = j <= n;
            jl[s] = valid ? j : -(1 << 28);
            jr[s] = valid ? j : (1 << 28);
            ca[s] = hb + ES * (uint32_t)(colbase(valid ? j : 1) - 1);
            ib[s] = ES * (uint32_t)(valid ? j : 1);
            c[s] = mk_cx<R>(0.0, 0.0);
        }
        if (want) {
            // ---- start row (src/GenericSchur.jl:390-420) ----
            int istart1 = 0;
            for (int base = iend - 1; base >= istart + 1 && !istart1; base -= LPM) {
                const int mm = base - sub;
                bool hit = false;
                if (mm >= istart + 1) {
                    const C h11 = ld(mm, mm), h22 = ld(mm + 1, mm + 1);
                    const C h11s = h11 - shift;
                    const R h21 = ld(mm + 1, mm).re;
                    const R rs = q_rcp(abs1(h11s) + fabs(h21));
                    const R h10 = ld(mm, mm - 1).re;
                    hit = fabs(h10) * fabs(h21 * rs) <= ulp * ((fabs(h11s.re * rs) + fabs(h11s.im * rs)) * (abs1(h11) + abs1(h22)));
                }
                const unsigned m = sballot(hit);
                if (m) istart1 = base - (__ffs(m) - 1);
            }
            if (!istart1) istart1 = istart;
            const int k0 = istart1;
            C v0, v1;
            {
                const C h11s = ld(k0, k0) - shift;
                const R h21 = ld(k0 + 1, k0).re;
                const R rs = q_rcp(abs1(h11s) + fabs(h21));
                v0 = mk_cx<R>(h11s.re * rs, h11s.im * rs);
                v1 = mk_cx<R>(h21 * rs, 0.0);
            }
            kf = k0;
            unsigned napplied = 0;
            if (k0 > istart) {
                late_start_step(k0, v0, v1);
                napplied = 1;
                kf = k0 + 1;
                store_sub = true;
                if (kf <= iend - 1) {
                    v0 = ld(kf, kf - 1);
                    v1 = ld(kf + 1, kf - 1);
                    ssync();
                    if (sub == 0) stc(kf + 1, kf - 1, mk_cx<R>(0.0, 0.0));   // the bulge lives in registers from here on
                }
            }
            if (kf > iend - 1) {
                // the late-start step was the only one
                st[1] += napplied;
                tail_fix_smem();
            } else {
                len = iend - kf;
                st[1] += napplied + (unsigned)len;
                ak = hb + ES * (uint32_t)(colbase(kf) - 1);
#pragma unroll
                for (int s = 0; s < CPL; ++s) {
                    if (jl[s] >= kf + 2) c[s] = lds_e<C>(ca[s] + ES * kf);
                    else if (jr[s] <= kf - 1) c[s] = cconj(lds_e<C>(ak + ib[s]));
                }
                d00 = ld(kf, kf);
                d10 = ld(kf + 1, kf);
                tau1 = reflector_cplx2(v0, v1);
                beta = v0.re;
                v2 = v1;
                tau2 = tau1.re * v2.re - tau1.im * v2.im;
                lg.put_hdr(LOG_REFL, kf, len, iend, 0.0, 0.0);
            }
        }
        // ---- the step loop.  Branch-free inside (taken branches and re-convergence points cost 20-50 cycles each on a
        //      warp that has an SM sub-partition to itself): the log space is reserved per chunk of steps outside, a
        //      reflector outside the fast path's range ends the chunk and is formed by the general routine between chunks.
        //      Slots keep the same chunk length, so they stay converged. ----
#ifdef GS_QR_PROFILE
        const long long tl0 = clock64();
#endif
        int t = 0;
        R nv1 = 0.0;
        for (;;) {
            int chunk = 0x7fffffff;
            if (t < len) {
                chunk = len - t;
                if (lg.on && !lg.ovf) {
                    if (lg.left == 0) lg.new_page();
                    if (!lg.ovf && lg.left < chunk) chunk = lg.left;
                }
            }
            if constexpr (LPM < 32) {
#pragma unroll
                for (int m = LPM; m < 32; m <<= 1) {
                    const int o = __shfl_xor_sync(0xffffffffu, chunk, m);
                    chunk = o < chunk ? o : chunk;
                }
            }
            if (chunk == 0x7fffffff) break;
            if (t < len) {
                const bool logp = lg.on && !lg.ovf && sub == 0;
                unsigned char* lp = lg.cur;
                bool ok = true;
                int i = 0;
#pragma unroll 1
                for (; i < chunk && ok; ++i) {
                    const int k = kf + t + i;
                    ssync();
                    const uint32_t kb = ES * (uint32_t)k;
                    const uint32_t ak1 = ak + ES * (uint32_t)(k + EX);   // column k+1
                    const C d01 = lds_e<C>(ak1 + kb), d11 = lds_e<C>(ak1 + kb + ES);
                    R e1 = 0.0;
                    lds_f64_if(e1, ak1 + kb + 2 * ES, k + 2 <= iend);
                    // second entry of every bulk item (the row that leaves the block re-reads its two entries below)
                    uint32_t sa[CPL], ya[CPL];
                    unsigned sg[CPL];
                    bool act[CPL];
                    C y[CPL];
#pragma unroll
                    for (int s = 0; s < CPL; ++s) {
                        const bool isR = jr[s] <= k;
                        act[s] = isR || (jl[s] >= k + 2);
                        sa[s] = isR ? ak + ib[s] : ca[s] + kb;
                        ya[s] = isR ? ak1 + ib[s] : sa[s] + ES;
                        sg[s] = isR ? 0x80000000u : 0u;
                        y[s] = lds_e<C>(ya[s]);
                    }
                    // ---- chain, part 1: rows k, k+1 of columns k, k+1 after the left update ----
                    const C ss0 = mk_cx<R>(fma(tau1.re, d00.re, fma(tau1.im, d00.im, tau2 * d10.re)),
                                           fma(tau1.re, d00.im, fma(-tau1.im, d00.re, tau2 * d10.im)));
                    const C a00 = d00 - ss0, a10 = e_fnma(ss0, v2, d10);
                    const C ss1 = mk_cx<R>(fma(tau1.re, d01.re, fma(tau1.im, d01.im, tau2 * d11.re)),
                                           fma(tau1.re, d01.im, fma(-tau1.im, d01.re, tau2 * d11.im)));
                    const C a01 = d01 - ss1, a11 = e_fnma(ss1, v2, d11);
                    // Row k leaves the block: its owner lane picks (conj H[k,k], H[k,k+1]) up from shared memory below.  Every
                    // lane of the slot stores the same two values (no predicate, no branch).
                    sts_e<C>(ak + kb, mk_cx<R>(a00.re, flip_if(a00.im, 0x80000000u)));
                    sts_e<C>(ak1 + kb, a01);
                    ssync();
                    // ---- bulk items: index j is a LEFT item (column j, rows k, k+1) while j >= k+2 and a RIGHT item (row j,
                    //      columns k, k+1; held conjugated so that one instruction stream serves both) once j <= k.
                    //      Source order = issue order of the memory operations: every load precedes the stores below. ----
                    C sv[CPL];
#pragma unroll
                    for (int s = 0; s < CPL; ++s) {
                        const bool isD = jr[s] == k;
                        lds_c64_if(c[s], sa[s], isD);                  // the row that just left the block (stored conjugated)
                        lds_c64_if(y[s], ya[s], isD);
                        y[s].im = flip_if(y[s].im, sg[s]);
                    }
#pragma unroll
                    for (int s = 0; s < CPL; ++s) {
                        const C x = c[s];
                        const C ss = mk_cx<R>(fma(tau1.re, x.re, fma(tau1.im, x.im, tau2 * y[s].re)),
                                              fma(tau1.re, x.im, fma(-tau1.im, x.re, tau2 * y[s].im)));
                        sv[s] = x - ss;
                        sv[s].im = flip_if(sv[s].im, sg[s]);
                        c[s] = e_fnma(ss, v2, y[s]);
                    }
                    // ---- chain, part 2: columns k, k+1 of rows k+1, k+2 after the right update ----
                    const C sr1 = tau1 * a10 + tau2 * a11;
                    nv0 = a10 - sr1;                                   // H[k+1, k]: head of the next reflector
                    const C n_d00 = e_fnma_cjb(sr1, v2, a11);          // H[k+1, k+1]
                    nv1 = -tau2 * e1;                                  // H[k+2, k]: the bulge (real)
                    const C n_d10 = mk_cx<R>(fma(nv1, v2.re, e1), -nv1 * v2.im);   // H[k+2, k+1]
#pragma unroll
                    for (int s = 0; s < CPL; ++s) {
                        sts_c64_if(sa[s], sv[s], act[s]);
                        sts_c64_if(sa[s] + ES, c[s], jl[s] == k + 2);  // H[k+1, k+2] enters the block next step
                    }
                    {
                        const uint32_t akm = ak - ES * (uint32_t)(k - 1 + EX);
                        sts_c64_if(akm + kb, mk_cx<R>(beta, 0.0), sub == 0 && (t + i > 0 || store_sub));
                        stg_2f64_if(lp, tau1.re, tau1.im, logp);
                        stg_2f64_if(lp + 16, v2.re, v2.im, logp);
                        lp += 32;
                    }
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
                    d00 = n_d00;
                    d10 = n_d10;
                    ak = ak1;
                }
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
        }
#ifdef GS_QR_PROFILE
        prof_loop += clock64() - tl0;
#endif
        if (want && len > 0) {
            // ---- write back what is still in registers after the last step (k = iend-1); the unit-modulus factor that
            //      makes H[iend, iend-1] real (src/GenericSchur.jl:486-500) is applied on the way ----
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
            ssync();
#pragma unroll
            for (int s = 0; s < CPL; ++s) {
                C v = c[s];
                if (fix) v = v * tphc;
                if (jl[s] >= iend + 1) stc(iend, jl[s], v);
                if (jr[s] <= iend - 1) stc(jr[s], iend, cconj(v));
            }
            if (sub == 0) {
                stc(iend, iend - 1, mk_cx<R>(fsr, 0.0));
                stc(iend, iend, d00);
            }
            ssync();
        }
    }

    // All 32 lanes of the warp call this together; slots without a matrix come in with alive == false.
    GS_DEV void run() {
#ifdef GS_QR_PROFILE
        prof_loop = 0;
        const long long tr0 = clock64();
#endif
        for (;;) {
            C shift = mk_cx<R>(0.0, 0.0);
            bool want = false;
            if (alive) want = next_sweep(shift);
            __syncwarp();
            if (!__any_sync(0xffffffffu, want)) break;
            sweep(want, shift);
            if (want) {
                its += 1;
                if (lg.ovf) {
                    alive = false;
                    info = LOG_OVERFLOW_RC;
                }
            }
        }
        st[3] = (unsigned)it;
#ifdef GS_QR_PROFILE
        st[0] = (unsigned)((clock64() - tr0) >> 6);
        st[2] = 0u;
        st[3] = (unsigned)(prof_loop >> 6);
#endif
    }
};

// =====================================================================================================================
// Float64, double shift
// =====================================================================================================================
template <int LPM, int CPL> struct ChainR {
    typedef double R;
    typedef cx<double> C;
    static constexpr int EX = 3;
    static constexpr uint32_t ES = 8;
    int n, sub, sbase;
    unsigned smask;
    uint32_t hb;   // &H(i, j) = hb + ES * (colbase(j) + i - 1)
    C* gw;         // eigenvalues of this matrix in global memory (written as blocks deflate)
    LogWriter<double> lg;
    unsigned st[4];
    int istart, iend, iterqr, it, iwcur, maxiter, info;
    bool alive;
#ifdef GS_QR_PROFILE
    long long prof_loop, prof_ns;
#endif

    __host__ __device__ static int colbase(int j) { return ((j - 1) * (j + 2 * EX)) / 2; }
    __host__ __device__ static int packed_elems(int n) { return (n * (n + 1)) / 2 + EX * n; }
    GS_DEV uint32_t ad(int i, int j) const { return hb + ES * (uint32_t)(colbase(j) + i - 1); }
    GS_DEV R ld(int i, int j) const { return lds_e<R>(ad(i, j)); }
    GS_DEV void stc(int i, int j, R v) const { sts_e<R>(ad(i, j), v); }
    GS_DEV void ssync() const {
        if constexpr (LPM == 32) __syncwarp();
        else __syncwarp(smask);
    }
    GS_DEV unsigned sballot(bool p) const { return __ballot_sync(smask, p) >> sbase; }

    // src/GenericSchur.jl:563-600 (note s = aa + bb, :586)
    GS_DEV bool split_test(int k, R smallnum, R eps) const {
        const R h = fabs(ld(k, k - 1));
        if (h < smallnum) return true;
        const R Hkk = ld(k, k), Hk1 = ld(k - 1, k - 1);
        R t = fabs(Hk1) + fabs(Hkk);
        if (t == 0.0) {
            if (k > 2) t = t + fabs(ld(k - 1, k - 2));
            if (k + 1 <= n) t = t + fabs(ld(k + 1, k));
        }
        if (h <= t * eps) {
            const R o = fabs(ld(k - 1, k));
            const R ab = fmax(h, o), ba = fmin(h, o);
            const R d1 = fabs(Hkk), d2 = fabs(Hk1 - Hkk);
            const R aa = fmax(d1, d2), bb = fmin(d1, d2);
            const R rs = q_rcp(aa + bb);
            if (ba * (ab * rs) <= fmax(smallnum, eps * (bb * (aa * rs)))) return true;
        }
        return false;
    }
    // src/GenericSchur.jl:851-872
    GS_DEV void first_column(int m, R r1r, R r1i, R r2r, R r2i, R& v0, R& v1, R& v2) const {
        const R hmm = ld(m, m);
        R H21s = ld(m + 1, m);
        R s = q_rcp(fabs(hmm - r2r) + fabs(r2i) + fabs(H21s));
        H21s = H21s * s;
        v0 = H21s * ld(m, m + 1) + (hmm - r1r) * ((hmm - r2r) * s) - r1i * (r2i * s);
        v1 = H21s * (hmm + ld(m + 1, m + 1) - r1r - r2r);
        v2 = H21s * ld(m + 2, m + 1);
        s = q_rcp(fabs(v0) + fabs(v1) + fabs(v2));
        v0 = v0 * s;
        v1 = v1 * s;
        v2 = v2 * s;
    }

    GS_DEV void begin(int n_, int maxiter_, bool have, C* gw_) {
        n = n_;
        maxiter = maxiter_;
        istart = 1;
        iend = n_;
        iwcur = n_;
        iterqr = 0;
        it = 0;
        info = 0;
        alive = have;
        gw = gw_;
        st[0] = st[1] = st[2] = st[3] = 0u;
    }

    // Advance the driver (src/GenericSchur.jl:540-690) until a sweep is due; returns false when the matrix is finished.
    GS_DEV bool next_sweep(R& r1r, R& r1i, R& r2r, R& r2i) {
        const R eps = 2.220446049250313e-16;
        const R smallnum = 2.2250738585072014e-308 * ((double)n / eps);
        for (;;) {
            if (iend < 1) {
                alive = false;
                return false;
            }
            it += 1;
            if (it > maxiter) {
                alive = false;
                info = iend;
                return false;
            }
            int found = 0;
            for (int base = iend; base >= istart + 1 && !found; base -= LPM) {
                const int k = base - sub;
                const bool hit = (k >= istart + 1) ? split_test(k, smallnum, eps) : false;
                const unsigned m = sballot(hit);
                if (m) found = base - (__ffs(m) - 1);
            }
            istart = found ? found : 1;
            ssync();
            if (istart > 1 && sub == 0) stc(istart, istart - 1, 0.0);
            ssync();
            if (istart >= iend - 1) {
                // ---- deflation of a 1x1 or 2x2 block (src/GenericSchur.jl:668-690) ----
                if (istart >= iend) {
                    if (sub == 0) gw[iwcur - 1] = mk_cx<R>(ld(iend, iend), 0.0);
                    iwcur -= 1;
                } else {
                    R a = ld(iend - 1, iend - 1), b = ld(iend - 1, iend), c = ld(iend, iend - 1), d = ld(iend, iend);
                    R cs, sn;
                    C w1, w2;
                    gs2x2(a, b, c, d, cs, sn, w1, w2);
                    if (sub == 0) {
                        gw[iwcur - 1] = w2;
                        gw[iwcur - 2] = w1;
                    }
                    iwcur -= 2;
                    lg.put_hdr(LOG_GIVENS, iend - 1, 0, iend - 1, cs, sn);
                    ssync();
                    for (int j = istart + sub; j <= n; j += LPM) {
                        const R a1 = ld(iend - 1, j), a2 = ld(iend, j);
                        stc(iend - 1, j, cs * a1 + sn * a2);
                        stc(iend, j, -sn * a1 + cs * a2);
                    }
                    ssync();
                    for (int r = 1 + sub; r <= iend; r += LPM) {
                        const R a1 = ld(r, iend - 1), a2 = ld(r, iend);
                        stc(r, iend - 1, a1 * cs + a2 * sn);
                        stc(r, iend, -a1 * sn + a2 * cs);
                    }
                    ssync();
                    if (sub == 0) {
                        stc(iend - 1, iend - 1, a);
                        stc(iend - 1, iend, b);
                        stc(iend, iend - 1, c);
                        stc(iend, iend, d);
                        if (iend > 2) stc(iend - 1, iend - 2, 0.0);
                    }
                    ssync();
                }
                iend = istart - 1;
                istart = 1;
                iterqr = 0;
                continue;
            }
            iterqr += 1;
            R H11, H12, H21, H22;
            if (iterqr == 10) {
                const R s = fabs(ld(istart + 1, istart)) + fabs(ld(istart + 2, istart + 1));
                H11 = 0.75 * s + ld(istart, istart);
                H12 = -0.4375 * s;
                H21 = s;
                H22 = H11;
                st[2] += 1;
            } else if (iterqr == 20) {
                const R s = fabs(ld(iend, iend - 1)) + fabs(ld(iend - 1, iend - 2));
                H11 = 0.75 * s + ld(iend, iend);
                H12 = -0.4375 * s;
                H21 = s;
                H22 = H11;
                st[2] += 1;
            } else {
                H11 = ld(iend - 1, iend - 1);
                H21 = ld(iend, iend - 1);
                H12 = ld(iend - 1, iend);
                H22 = ld(iend, iend);
            }
            const R s = fabs(H11) + fabs(H12) + fabs(H21) + fabs(H22);
            r1r = r2r = r1i = r2i = 0.0;
            if (!(s == 0.0)) {
                const R rs = q_rcp(s);
                H11 = H11 * rs;
                H12 = H12 * rs;
                H21 = H21 * rs;
                H22 = H22 * rs;
                const R tr = (H11 + H22) * 0.5;
                const R d = (H11 - tr) * (H22 - tr) - H12 * H21;
                const R rtd = q_sqrt(fabs(d));
                if (d >= 0.0) {
                    r1r = tr * s;
                    r2r = r1r;
                    r1i = rtd * s;
                    r2i = -r1i;
                } else {
                    r1r = tr + rtd;
                    r2r = tr - rtd;
                    if (fabs(r1r - H22) <= fabs(r2r - H22)) {
                        r1r = r1r * s;
                        r2r = r1r;
                    } else {
                        r2r = r2r * s;
                        r1r = r2r;
                    }
                }
            }
            st[0] += 1;
            return true;
        }
    }

    // One Francis double-shift sweep (src/GenericSchur.jl:837-952).  Window sizes >= 3 (the driver deflates smaller ones).
    GS_DEV void sweep(bool want, R r1r, R r1i, R r2r, R r2i) {
        const R eps = 2.220446049250313e-16;
        int mx = 1, len = 0;
        bool scale_sub = false;
        uint32_t ca[CPL], ib[CPL];
        int jl[CPL], jr[CPL];
        R c1[CPL], c2[CPL];
        R b00 = 0, b10 = 0, b20 = 0, b01 = 0, b11 = 0, b21 = 0;
        R tau1 = 0, tau2 = 0, tau3 = 0, v1 = 0, v2 = 0, beta = 0;
        R L10 = 0, L20 = 0, L11 = 0, L21 = 0, L12 = 0, L22 = 0, L01 = 0, L02 = 0;
        uint32_t ak = hb;
#pragma unroll
        for (int s = 0; s < CPL; ++s) {
            const int j = sub + 1 + LPM * s;
            const bool valid = j <= n;
            jl[s] = valid ? j : -(1 << 28);
            jr[s] = valid ? j : (1 << 28);
            ca[s] = hb + ES * (uint32_t)(colbase(valid ? j : 1) - 1);
            ib[s] = ES * (uint32_t)(valid ? j : 1);
            c1[s] = 0.0;
            c2[s] = 0.0;
        }
        if (want) {
            int m0 = 0;
            for (int base = iend - 2; base >= istart + 1 && !m0; base -= LPM) {
                const int m = base - sub;
                bool hit = false;
                if (m >= istart + 1) {
                    R a0, a1, a2;
                    first_column(m, r1r, r1i, r2r, r2i, a0, a1, a2);
                    hit = fabs(ld(m, m - 1)) * (fabs(a1) + fabs(a2)) <=
                          eps * fabs(a0) * (fabs(ld(m - 1, m - 1)) + fabs(ld(m, m)) + fabs(ld(m + 1, m + 1)));
                }
                const unsigned msk = sballot(hit);
                if (msk) m0 = base - (__ffs(msk) - 1);
            }
            mx = m0 ? m0 : istart;
            R v0;
            first_column(mx, r1r, r1i, r2r, r2i, v0, v1, v2);
            scale_sub = mx > istart;
            len = iend - 1 - mx;          // three-row reflectors at k = mx .. iend-2; the two-row one follows the loop
            st[1] += (unsigned)(iend - mx);
            ak = hb + ES * (uint32_t)(colbase(mx) - 1);
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
            b00 = ld(mx, mx);
            b10 = ld(mx + 1, mx);
            b20 = ld(mx + 2, mx);
            b01 = ld(mx, mx + 1);
            b11 = ld(mx + 1, mx + 1);
            b21 = ld(mx + 2, mx + 1);
            tau1 = reflector_real_small(v0, v1, v2, 3);
            beta = v0;
            tau2 = tau1 * v1;
            tau3 = tau1 * v2;
            lg.put_hdr(LOG_REFL3, mx, len, iend, 0.0, 0.0);
        }
        if (want && scale_sub) {
            // the sweep starts inside the block: the entry left of it is scaled by the first reflector (src/GenericSchur.jl:896-899)
            ssync();
            if (sub == 0) stc(mx, mx - 1, ld(mx, mx - 1) * (1.0 - tau1));
        }
        // ---- the step loop: branch-free chunks (see ChainC::sweep) ----
#ifdef GS_QR_PROFILE
        const long long tl0 = clock64();
#endif
        int t = 0;
        R f10 = 0, f20 = 0, f30 = 0, f01 = 0, f02 = 0;
        for (;;) {
            int chunk = 0x7fffffff;
            if (t < len) {
                chunk = len - t;
                if (lg.on && !lg.ovf) {
                    if (lg.left == 0) lg.new_page();
                    if (!lg.ovf && lg.left < chunk) chunk = lg.left;
                }
            }
            if constexpr (LPM < 32) {
#pragma unroll
                for (int m = LPM; m < 32; m <<= 1) {
                    const int o = __shfl_xor_sync(0xffffffffu, chunk, m);
                    chunk = o < chunk ? o : chunk;
                }
            }
            if (chunk == 0x7fffffff) break;
            if (t < len) {
                const bool logp = lg.on && !lg.ovf && sub == 0;
                unsigned char* lp = lg.cur;
                bool ok = true;
                int i = 0;
#pragma unroll 1
                for (; i < chunk && ok; ++i) {
                    const int k = mx + t + i;
                    ssync();
                    const uint32_t kb = ES * (uint32_t)k;
                    const uint32_t ak1 = ak + ES * (uint32_t)(k + EX);        // column k+1
                    const uint32_t ak2 = ak1 + ES * (uint32_t)(k + 1 + EX);   // column k+2
                    // ---- loads (all of the step's loads precede its stores): column k+2 of the block, third entry of every item ----
                    const R b02 = lds_e<R>(ak2 + kb), b12 = lds_e<R>(ak2 + kb + ES), b22 = lds_e<R>(ak2 + kb + 2 * ES);
                    R e3 = 0.0;
                    lds_f64_if(e3, ak2 + kb + 3 * ES, k + 3 <= iend);
                    uint32_t sa[CPL];
                    bool act[CPL];
                    R y[CPL];
#pragma unroll
                    for (int s = 0; s < CPL; ++s) {
                        const bool isR = jr[s] <= k - 1;
                        act[s] = isR || (jl[s] >= k + 3);
                        sa[s] = isR ? ak + ib[s] : ca[s] + kb;
                        const uint32_t ya = isR ? ak2 + ib[s] : sa[s] + 2 * ES;
                        lds_f64_if(c1[s], sa[s], jr[s] == k - 1);             // row k-1 left the block in the previous step
                        lds_f64_if(c2[s], ak1 + ib[s], jr[s] == k - 1);
                        y[s] = lds_e<R>(ya);
                    }
                    // ---- chain: rows k..k+2 of columns k..k+2 (left), then columns k..k+2 of rows k..k+3 (right) ----
                    const R s0 = fma(v2, b20, fma(v1, b10, b00));
                    const R a00 = fma(-s0, tau1, b00), a10 = fma(-s0, tau2, b10), a20 = fma(-s0, tau3, b20);
                    const R s1 = fma(v2, b21, fma(v1, b11, b01));
                    const R a01 = fma(-s1, tau1, b01), a11 = fma(-s1, tau2, b11), a21 = fma(-s1, tau3, b21);
                    const R s2 = fma(v2, b22, fma(v1, b12, b02));
                    const R a02 = fma(-s2, tau1, b02), a12 = fma(-s2, tau2, b12), a22 = fma(-s2, tau3, b22);
                    const R t1 = fma(v2, a12, fma(v1, a11, a10));
                    const R t2 = fma(v2, a22, fma(v1, a21, a20));
                    const R t3 = v2 * e3;
                    f10 = fma(-t1, tau1, a10);
                    f20 = fma(-t2, tau1, a20);
                    f30 = -t3 * tau1;
                    const R f11 = fma(-t1, tau2, a11), f12 = fma(-t1, tau3, a12);   // row k+1
                    const R f21 = fma(-t2, tau2, a21), f22 = fma(-t2, tau3, a22);   // row k+2
                    const R f31 = -t3 * tau2, f32 = fma(-t3, tau3, e3);             // row k+3
                    const R t0 = fma(v2, a02, fma(v1, a01, a00));
                    const R f00 = fma(-t0, tau1, a00);                              // row k: final
                    f01 = fma(-t0, tau2, a01);
                    f02 = fma(-t0, tau3, a02);
                    // ---- bulk items (the same formulas serve columns on the left and rows on the right) ----
                    R sv[CPL];
#pragma unroll
                    for (int s = 0; s < CPL; ++s) {
                        const R ss = fma(v2, y[s], fma(v1, c2[s], c1[s]));
                        sv[s] = fma(-ss, tau1, c1[s]);
                        c1[s] = fma(-ss, tau2, c2[s]);
                        c2[s] = fma(-ss, tau3, y[s]);
                    }
                    // ---- stores ----
                    {
                        const bool l0 = sub == 0;
                        const uint32_t akm = ak - ES * (uint32_t)(k - 1 + EX);
                        const bool psub = l0 && (t + i > 0);      // column k-1: (beta, 0, 0) below the diagonal
                        sts_f64_if(akm + kb, beta, psub);
                        sts_f64_if(akm + kb + ES, 0.0, psub);
                        sts_f64_if(akm + kb + 2 * ES, 0.0, psub);
                        sts_f64_if(ak + kb, f00, l0);
                        sts_f64_if(ak1 + kb, f01, l0);
                        sts_f64_if(ak2 + kb, f02, l0);
                        stg_2f64_if(lp, tau1, v1, logp);
                        stg_2f64_if(lp + 16, v2, 0.0, logp);
                        lp += 32;
                    }
#pragma unroll
                    for (int s = 0; s < CPL; ++s) {
                        sts_f64_if(sa[s], sv[s], act[s]);
                        const bool own3 = jl[s] == k + 3;          // column k+3 enters the block next step
                        sts_f64_if(sa[s] + ES, c1[s], own3);
                        sts_f64_if(sa[s] + 2 * ES, c2[s], own3);
                    }
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
                    b00 = f11;
                    b10 = f21;
                    b20 = f31;
                    b01 = f12;
                    b11 = f22;
                    b21 = f32;
                    ak = ak1;
                }
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
        }
        L10 = f10; L20 = f20; L01 = f01; L02 = f02;
        L11 = b00; L21 = b10; L12 = b01; L22 = b11;
#ifdef GS_QR_PROFILE
        prof_loop += clock64() - tl0;
#endif
        if (want) {
            // ---- last step: the two-row reflector at k = iend-1 (src/GenericSchur.jl:927-946), entirely on registers: the
            //      lanes hold rows iend-1, iend of their columns (left items) / columns iend-1, iend of their rows (right
            //      items), the block and row iend-2 are replicated; everything is written back afterwards ----
            const int k = iend - 1;
            R w0 = L10, w1 = L20, w2 = 0.0;
            const R t1 = reflector_real_small(w0, w1, w2, 2);
            const R t2 = t1 * w1;
            lg.put_hdr(LOG_REFL2, k, 1, k, 0.0, 0.0);
            lg.put4(t1, w1, 0.0, 0.0);
            const R sa0 = L11 + w1 * L21, sa1 = L12 + w1 * L22;
            const R g11 = L11 - sa0 * t1, g21 = L21 - sa0 * t2, g12 = L12 - sa1 * t1, g22 = L22 - sa1 * t2;
            const R sr0 = L01 + w1 * L02, sr1 = g11 + w1 * g12, sr2 = g21 + w1 * g22;
            ssync();
#pragma unroll
            for (int s = 0; s < CPL; ++s) {
                const R ss = c1[s] + w1 * c2[s];
                const R n1 = c1[s] - ss * t1, n2 = c2[s] - ss * t2;
                if (jl[s] >= iend + 1) {
                    stc(iend - 1, jl[s], n1);
                    stc(iend, jl[s], n2);
                }
                if (jr[s] <= iend - 3) {
                    stc(jr[s], iend - 1, n1);
                    stc(jr[s], iend, n2);
                }
            }
            if (sub == 0) {
                stc(k, k - 1, w0);
                stc(k + 1, k - 1, 0.0);
                stc(k - 1, k, L01 - sr0 * t1);
                stc(k - 1, k + 1, L02 - sr0 * t2);
                stc(k, k, g11 - sr1 * t1);
                stc(k, k + 1, g12 - sr1 * t2);
                stc(k + 1, k, g21 - sr2 * t1);
                stc(k + 1, k + 1, g22 - sr2 * t2);
            }
            ssync();
        }
    }

    GS_DEV void run() {
#ifdef GS_QR_PROFILE
        prof_loop = 0;
        const long long tr0 = clock64();
#endif
        for (;;) {
            R r1r = 0, r1i = 0, r2r = 0, r2i = 0;
            bool want = false;
            if (alive) want = next_sweep(r1r, r1i, r2r, r2i);
            __syncwarp();
            if (!__any_sync(0xffffffffu, want)) break;
            sweep(want, r1r, r1i, r2r, r2i);
            if (want && lg.ovf) {
                alive = false;
                info = LOG_OVERFLOW_RC;
            }
        }
        st[3] = (unsigned)it;
#ifdef GS_QR_PROFILE
        st[0] = (unsigned)((clock64() - tr0) >> 6);
        st[2] = 0u;
        st[3] = (unsigned)(prof_loop >> 6);
#endif
    }
};

// ---------------------------------------------------------------------------------------------------------------------
// kernel: one warp per CTA, 32 / LPM matrices per warp
// ---------------------------------------------------------------------------------------------------------------------
// VAR = 0: the chain solvers of this file; VAR = 1: the owner-computes solvers of ownqr.cuh (LPM = 32 only)
template <class T, int LPM, int CPL, int VAR = 0> struct chain_traits;
template <int LPM, int CPL> struct chain_traits<cx<double>, LPM, CPL, 0> {
    typedef ChainC<LPM, CPL> Solver;
};
template <int LPM, int CPL> struct chain_traits<double, LPM, CPL, 0> {
    typedef ChainR<LPM, CPL> Solver;
};

template <class T, int LPM, int CPL> struct chain_layout {
    typedef typename chain_traits<T, LPM, CPL>::Solver S;
    static constexpr int NS = 32 / LPM;
    __host__ __device__ static size_t slot_bytes(int n) { return ((size_t)S::packed_elems(n) * sizeof(T) + 15) & ~(size_t)15; }
    __host__ __device__ static size_t bytes(int n) { return NS * slot_bytes(n); }
};

template <class T, int LPM, int CPL, int MINB, int VAR = 0>
__global__ void __launch_bounds__(32, MINB) gschur_chain_kernel(BatchedParams p) {
    typedef typename etraits<T>::real R;
    typedef cx<R> C;
    constexpr bool CPLX = etraits<T>::is_complex;
    typedef typename chain_traits<T, LPM, CPL, VAR>::Solver S;
    typedef chain_layout<T, LPM, CPL> CL;
    constexpr int NS = CL::NS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = p.n;
    const int lane = threadIdx.x;
    SlotGeom<LPM> G;
    G.init(lane);
    const int slot = lane / LPM;
    const bool wantZ = (p.Z != nullptr);
    T* H = reinterpret_cast<T*>(smem_raw + (size_t)slot * CL::slot_bytes(n));
    S F;
    F.sub = G.sub;
    F.sbase = G.sbase;
    F.smask = G.smask;
    F.hb = smem_u32(H);
    const R zero = r_const<R>(0.0);

    for (;;) {
        long long b0 = 0;
        if (lane == 0) b0 = (long long)atomicAdd(p.counter, (unsigned long long)NS);
        b0 = __shfl_sync(0xffffffffu, b0, 0);
        if (b0 >= p.batch) break;
        const long long b = b0 + slot;
        const bool have = b < p.batch;
        T* gA = reinterpret_cast<T*>(p.A) + (have ? b : b0) * p.strideA;

        // ---- load the Hessenberg part into packed storage (four columns in flight per lane); bulge slots start at zero ----
        int bad = 0;
        if (have) {
            for (int j0 = 1; j0 <= n; j0 += 4) {
                T v[4][CPL + 1];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int j = j0 + q;
#pragma unroll
                    for (int s = 0; s <= CPL; ++s) {
                        const int i = 1 + G.sub + LPM * s;
                        v[q][s] = e_zero<T>();
                        if (j <= n && i <= j + 1 && i <= n) v[q][s] = gA[(i - 1) + (size_t)(j - 1) * p.lda];
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int j = j0 + q;
                    if (j > n) continue;
                    const int cb = S::colbase(j);
#pragma unroll
                    for (int s = 0; s <= CPL; ++s) {
                        const int i = 1 + G.sub + LPM * s;
                        if (i <= j + S::EX) H[cb + i - 1] = v[q][s];
                        if constexpr (CPLX) {
                            if (i == j + 1 && i <= n && (p.flags & F_CHECK_SUBDIAG) && v[q][s].im != zero) bad = 1;
                        }
                    }
                }
            }
            bad = __ballot_sync(G.smask, bad) != 0u;
        }
        __syncwarp();
        const int maxiter = p.maxiter > 0 ? p.maxiter : 100 * n;
        C* gw = reinterpret_cast<C*>(p.w) + (have ? b : 0) * (long long)n;
        if constexpr (CPLX) F.begin(n, maxiter, have && !bad);
        else F.begin(n, maxiter, have && !bad, gw);
        F.lg.init(p, have ? b : 0, lane, wantZ && have && !bad, G.smask, G.sbase);
        F.run();
        __syncwarp();
        if (have) {
            int info = bad ? -4 : F.info;
            F.lg.finish();
            if (info == LOG_OVERFLOW_RC) {
                // leave H (and Q) untouched in global memory: the fused kernel redoes this matrix after stage C
                if (G.sub == 0) {
                    const unsigned idx = atomicAdd(p.redo_count, 1u);
                    p.redo_list[idx] = b;
                }
            } else {
                // ---- unscale (src/GenericSchur.jl:367-370, 830-833) with the factors stage A recorded ----
                bool scaled = false;
                R cscale = r_const<R>(1.0), anrm = r_const<R>(1.0);
                if (p.scratch) {
                    const double* sc = p.scratch + 8 * b;
                    scaled = sc[0] != 0.0;
                    cscale = sc[1];
                    anrm = sc[3];
                }
                if (scaled) {
                    const int total = S::packed_elems(n);
                    safescale_apply<T, R, 32>(cscale, anrm, [&](R mul) {
                        for (int e = G.sub; e < total; e += LPM) H[e] = e_scale(H[e], mul);
                        if constexpr (!CPLX) {   // the real path rescales the eigenvalues separately (src/GenericSchur.jl:832)
                            for (int e = G.sub; e < n; e += LPM) gw[e] = mk_cx<R>(gw[e].re * mul, gw[e].im * mul);
                        }
                    });
                    __syncwarp(G.smask);
                }
                // ---- store T (exact zeros below the (quasi-)triangle), w, info, stats ----
                for (int j = 1; j <= n; ++j) {
                    const int cb = S::colbase(j);
                    for (int i = 1 + G.sub; i <= n; i += LPM) {
                        const bool keep = CPLX ? (i <= j) : (i <= j + 1);
                        gA[(i - 1) + (size_t)(j - 1) * p.lda] = keep ? H[cb + i - 1] : e_zero<T>();
                    }
                }
                for (int e = G.sub; e < n; e += LPM) {
                    if constexpr (CPLX) gw[e] = H[S::colbase(e + 1) + e];
                }
                if (G.sub == 0) {
                    if (p.info) p.info[b] = info;
                    if (p.stats) {
                        p.stats[4 * b + 0] = F.st[0];
                        p.stats[4 * b + 1] = F.st[1];
                        p.stats[4 * b + 2] = F.st[2];
                        p.stats[4 * b + 3] = F.st[3];
                    }
                }
            }
        }
        __syncwarp();
    }
}

}  // namespace gs

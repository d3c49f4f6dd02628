// Warp-specialised QR-iteration stage for Float64 / ComplexF64 matrices with n <= 64.
//
// The Francis sweep is a serial chain: reflector k+1 cannot be formed before reflector k has been applied to
// the entries next to the diagonal.  Everything else — the far columns of H and the whole of Z — only needs the
// reflectors, not the other way round.  This stage therefore splits the CTA in two roles:
//
//   * the H-warp (warp 0) owns the Hessenberg matrix.  Lane l owns columns and rows l+1, l+33 (CPL = 2).  It keeps
//     the running entries of its columns (left update) and rows (right update) in registers ("carries"), so a
//     bulge step costs one shared-memory load and one store per owned column/row instead of 2-3 each, and the
//     only synchronisation on the critical path is __syncwarp.  Lanes on the diagonal publish / refresh their
//     carries through shared memory, which is also where the next reflector's vector is picked up.
//   * the Z-warp (warp 1) applies the reflectors to the Schur vectors.  It receives them 16 at a time through a ring of
//     record buffers in shared memory (named barriers full[b] / empty[b]; three buffers for Float64 / ComplexF64, two for
//     the double-double kinds) and streams them through carry registers as well (one load + one store per row and step).
//     It never blocks the H-warp unless it falls two buffers (32 bulge steps) behind.
//
// Storage.  The Hessenberg matrix is kept in PACKED form in shared memory — column j holds rows 1..j+EX, EX = 2
// (complex: one bulge entry below the sub-diagonal) or 3 (real double shift: two) — which halves its footprint
// (64x64 ComplexF64: 35 KB instead of 65 KB).  Z is NOT kept on chip: nothing on the critical path ever reads it, so
// the Z-warp streams it through L2 (one coalesced column load and store per reflector, software-prefetched) in
// place in the caller's Z buffer.  Together that is 37.6 KB of shared memory per matrix instead of 142 KB, i.e. six
// CTAs per SM instead of one — occupancy is what this latency-bound algorithm needs.
//
// The arithmetic per reflector application and every decision rule are those of batched.cuh (and of the
// reference, src/GenericSchur.jl:194-335, 374-504, 513-699, 837-952); only the schedule and the layout differ.
#pragma once
#include <cstdlib>
#include "gehrd_split.cuh"
#include "qrlog.cuh"

#ifndef GS_QR_MINB_F64_1
#define GS_QR_MINB_F64_1 9   // 32x32 Float64: CTAs per SM the register budget is sized for (9 x 7 named barriers fit an SM)
#endif
#ifndef GS_QR_MINB_F64_2
#define GS_QR_MINB_F64_2 8   // 64x64 Float64
#endif
#ifndef GS_QR_MINB_C64_1
#define GS_QR_MINB_C64_1 8   // n <= 32 ComplexF64: 128 registers, 8 CTAs per SM (measured: 6 -> 8 CTAs = -12 % stage-B time, 10 = spills)
#endif
#ifndef GS_ZRUN
#define GS_ZRUN 8
#endif

// smallest CPL (columns per lane) at which the complex double-double kernel uses several H- and Z-warps per matrix
// (measured at 64x64, CPL = 2: 6890 matrices/s with two of each against 7176 with one — several CTAs share an SM there)
#ifndef GS_CDD_NH_MINCPL
#define GS_CDD_NH_MINCPL 3
#endif

namespace gs {

enum { ZOP_REFL = 1, ZOP_SCALE = 2, ZOP_REFL3 = 3, ZOP_REFL2 = 4, ZOP_GIVENS = 5 };
enum { LOG_OVERFLOW_RC = -100 };   // FastSolver<..., LOG>::qr_*: the matrix's reflector log is full

template <bool CPLX, class R> struct zop_t;
template <class R> struct __align__(16) zop_t<true, R> {    // 16 B + 4 reals
    int op, k, k2, pad;
    R a[4];   // REFL: tau1.re, tau1.im, v2.re, v2.im ; SCALE: t.re, t.im (columns k..k2)
};
template <class R> struct __align__(16) zop_t<false, R> {   // Float64: 32 B
    int op, k;
    R a[3];   // REFL3: tau1, v2, v3 ; REFL2: tau1, v2 ; GIVENS: cs, sn (columns k, k+1)
};

struct zring_hdr {
    int count[4];
    int end[4];
};

// Number of buffers of the reflector ring between the H-warp and the Z-warp.  The Float64 / ComplexF64 kernels use
// three (the H-warp may run two buffers = 32 bulge steps ahead): with two, the SASS-level profile showed the H-warp
// waiting ~900 cycles per buffer hand-over for the Z-warp's slow buffers (Z columns that had left L2); measured gain of
// the third buffer: 6 % of stage B at n = 64 (both kinds).  An SM has slots for 64 named barriers, i.e. 9 CTAs with
// 7 barriers each; the double-double kinds keep two buffers.
template <class T, int CPL> struct ring_bufs { static constexpr int value = 2; };
template <int CPL> struct ring_bufs<cx<double>, CPL> { static constexpr int value = 3; };
#ifndef GS_RING_F64_2
#define GS_RING_F64_2 3   // 64x64 Float64 (8 CTAs per SM x 7 barriers fit)
#endif
template <> struct ring_bufs<double, 2> { static constexpr int value = GS_RING_F64_2; };
#ifndef GS_RING_F64_1
#define GS_RING_F64_1 3   // 32x32 Float64 at 9 CTAs per SM (measured: 12 CTAs x 2 buffers 6.72 ms, 9 x 3 buffers 6.60 ms per 16384)
#endif
template <> struct ring_bufs<double, 1> { static constexpr int value = GS_RING_F64_1; };

// named barriers with immediate ids (a register id would make ptxas reserve all 16 barriers per CTA):
// full[b] = 1 + b, empty[b] = 1 + RB + b
template <int ID, int CNT = 64> GS_DEV void bar_sync_imm() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(CNT) : "memory"); }
template <int ID, int CNT = 64> GS_DEV void bar_arrive_imm() { asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(CNT) : "memory"); }
template <int RB, int CNT = 64> GS_DEV void named_bar_sync(int id) {
    if constexpr (RB == 2) {
        switch (id) {
            case 1: bar_sync_imm<1, CNT>(); break;
            case 2: bar_sync_imm<2, CNT>(); break;
            case 3: bar_sync_imm<3, CNT>(); break;
            default: bar_sync_imm<4, CNT>(); break;
        }
    } else {
        switch (id) {
            case 1: bar_sync_imm<1, CNT>(); break;
            case 2: bar_sync_imm<2, CNT>(); break;
            case 3: bar_sync_imm<3, CNT>(); break;
            case 4: bar_sync_imm<4, CNT>(); break;
            case 5: bar_sync_imm<5, CNT>(); break;
            default: bar_sync_imm<6, CNT>(); break;
        }
    }
}
template <int RB, int CNT = 64> GS_DEV void named_bar_arrive(int id) {
    if constexpr (RB == 2) {
        switch (id) {
            case 1: bar_arrive_imm<1, CNT>(); break;
            case 2: bar_arrive_imm<2, CNT>(); break;
            case 3: bar_arrive_imm<3, CNT>(); break;
            default: bar_arrive_imm<4, CNT>(); break;
        }
    } else {
        switch (id) {
            case 1: bar_arrive_imm<1, CNT>(); break;
            case 2: bar_arrive_imm<2, CNT>(); break;
            case 3: bar_arrive_imm<3, CNT>(); break;
            case 4: bar_arrive_imm<4, CNT>(); break;
            case 5: bar_arrive_imm<5, CNT>(); break;
            default: bar_arrive_imm<6, CNT>(); break;
        }
    }
}

// typed shared-memory access through 32-bit shared-space byte addresses (no generic->shared conversion and no
// 64-bit address arithmetic inside the step loops)
template <class T> GS_DEV T lds_e(uint32_t a);
template <> GS_DEV double lds_e<double>(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
template <> GS_DEV cx<double> lds_e<cx<double>>(uint32_t a) {
    cx<double> v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.re), "=d"(v.im) : "r"(a));
    return v;
}
template <> GS_DEV dd_t lds_e<dd_t>(uint32_t a) {
    dd_t v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.hi), "=d"(v.lo) : "r"(a));
    return v;
}
template <> GS_DEV cx<dd_t> lds_e<cx<dd_t>>(uint32_t a) {
    cx<dd_t> v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.re.hi), "=d"(v.re.lo) : "r"(a));
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.im.hi), "=d"(v.im.lo) : "r"(a + 16));
    return v;
}
template <class T> GS_DEV void sts_e(uint32_t a, const T& v);
template <> GS_DEV void sts_e<dd_t>(uint32_t a, const dd_t& v) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v.hi), "d"(v.lo) : "memory");
}
template <> GS_DEV void sts_e<cx<dd_t>>(uint32_t a, const cx<dd_t>& v) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v.re.hi), "d"(v.re.lo) : "memory");
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a + 16), "d"(v.im.hi), "d"(v.im.lo) : "memory");
}
template <> GS_DEV void sts_e<double>(uint32_t a, const double& v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}
template <> GS_DEV void sts_e<cx<double>>(uint32_t a, const cx<double>& v) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v.re), "d"(v.im) : "memory");
}

// predicated shared-memory accesses (guaranteed branch-free inside the step loops)
GS_DEV void sts_c64_if(uint32_t a, const cx<double>& v, bool p) {
    asm volatile("{ .reg .pred q; setp.ne.b32 q, %3, 0; @q st.shared.v2.f64 [%0], {%1, %2}; }" ::"r"(a), "d"(v.re), "d"(v.im),
                 "r"((int)p)
                 : "memory");
}
GS_DEV void lds_c64_if(cx<double>& v, uint32_t a, bool p) {
    asm volatile("{ .reg .pred q; setp.ne.b32 q, %3, 0; @q ld.shared.v2.f64 {%0, %1}, [%2]; }"
                 : "+d"(v.re), "+d"(v.im)
                 : "r"(a), "r"((int)p));
}
GS_DEV void sts_2i_if(uint32_t a, int x, int y, bool p) {
    asm volatile("{ .reg .pred q; setp.ne.b32 q, %3, 0; @q st.shared.v2.b32 [%0], {%1, %2}; }" ::"r"(a), "r"(x), "r"(y),
                 "r"((int)p)
                 : "memory");
}

GS_DEV void sts_f64_if(uint32_t a, double v, bool p) {
    asm volatile("{ .reg .pred q; setp.ne.b32 q, %2, 0; @q st.shared.f64 [%0], %1; }" ::"r"(a), "d"(v), "r"((int)p) : "memory");
}
GS_DEV void lds_f64_if(double& v, uint32_t a, bool p) {
    asm volatile("{ .reg .pred q; setp.ne.b32 q, %2, 0; @q ld.shared.f64 %0, [%1]; }" : "+d"(v) : "r"(a), "r"((int)p));
}

// LOG = true: the three-stage path.  There is no Z-warp and no ring; every transformation applied from the right goes to
// the global-memory log (qrlog.cuh) that stage C replays on Z.
template <class T, int CPL, bool LOG = false> struct FastSolver {
    typedef typename etraits<T>::real R;
    typedef cx<R> C;
    static constexpr bool CPLX = etraits<T>::is_complex;
    typedef zop_t<CPLX, R> ZOp;

    static constexpr int EX = CPLX ? 2 : 3;   // rows stored below the diagonal in each packed column
    static constexpr int RB = ring_bufs<T, CPL>::value;  // ring buffers (2 or 3)
    static constexpr int BAR_FULL0 = 1, BAR_EMPTY0 = 1 + RB;
    // Complex double-double at CPL = 3 (65 <= n <= 96): the three columns / rows a lane owns go to three H-warps — the
    // driver (slot 0: everything the single H-warp did, for its own slot) and NH - 1 helpers that run the step loop of a
    // sweep for their slot and are parked at a barrier otherwise (see sweep_steps / helper_loop).
    static constexpr int NH = (etraits<T>::is_complex && sizeof(R) == 16 && CPL >= GS_CDD_NH_MINCPL && !LOG) ? CPL : 1;
    // ... and the three rows of Z a lane owns to NH Z-warps (all of them consume every record of the ring)
    static constexpr int NZW = NH;
    static constexpr int RING_THREADS = 32 * (1 + NZW);   // the driver and the Z-warps meet at the ring's named barriers
    static constexpr int BAR_HSTEP = 7, BAR_HCMD = 8;
    int zslot;
    struct HCmd {
        int op, k0, istart, iend;
        C v0, v1;
    };
    enum { HCMD_SWEEP = 1, HCMD_EXIT = 2 };
    HCmd* hcmd;
    int rbase;                                           // first record of the buffer being filled: (sidx % RB) * cap
    int n, ldz, lane, cap;
    T* H;            // packed upper Hessenberg (+ bulge slots), shared memory
    T* Z;            // Schur vectors, global memory (column-major, leading dimension ldz)
    C* sW;
    ZOp* ring;        // [2][cap]
    zring_hdr* hdr;
    bool wantZ;
    LogWriter<R> lg;   // LOG only
    unsigned* stp;
    // producer state (uniform across the H-warp)
    int sidx, cnt;
#ifdef GS_QR_PROFILE
    long long prof[4];
#endif

    __host__ __device__ static int colbase(int j) { return ((j - 1) * (j + 2 * EX)) / 2; }   // offset of column j
    __host__ __device__ static int packed_elems(int n) { return (n * (n + 1)) / 2 + EX * n; }
#define HH(i, j) H[colbase(j) + (i)-1]
#define ZZ(i, j) Z[((i)-1) + (size_t)((j)-1) * ldz]

    // ---- decision rules (same as BatchedSolver::split_test_c / split_test_r / first_column_r, packed accessor) ----
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
            R rs = q_rcp(aa + ab);
            if (ba * (ab * rs) <= r_max(smallnum, ulp * (bb * (aa * rs)))) return true;
        }
        return false;
    }
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
            R rs = q_rcp(aa + bb);   // s = aa + bb as the reference has it (src/GenericSchur.jl:586)
            if (ba * (ab * rs) <= r_max(smallnum, eps * (bb * (aa * rs)))) return true;
        }
        return false;
    }
    GS_DEV void first_column_r(int m, const R& r1r, const R& r1i, const R& r2r, const R& r2i, R& v0, R& v1, R& v2) {
        R hmm = HH(m, m);
        R H21s = HH(m + 1, m);
        R s = q_rcp(r_abs(hmm - r2r) + r_abs(r2i) + r_abs(H21s));
        H21s = H21s * s;
        v0 = H21s * HH(m, m + 1) + (hmm - r1r) * ((hmm - r2r) * s) - r1i * (r2i * s);
        v1 = H21s * (hmm + HH(m + 1, m + 1) - r1r - r2r);
        v2 = H21s * HH(m + 2, m + 1);
        s = q_rcp(r_abs(v0) + r_abs(v1) + r_abs(v2));
        v0 = v0 * s;
        v1 = v1 * s;
        v2 = v2 * s;
    }

    // ------------------------------------------------------------------------------------------------
    // producer side of the Z ring
    // ------------------------------------------------------------------------------------------------
    GS_DEV void begin_buffer() {
        if constexpr (LOG) return;
        rbase = (sidx % RB) * cap;
        if (!wantZ) return;
#ifdef GS_QR_PROFILE
        const long long tb0 = clock64();
#endif
        if (sidx >= RB) named_bar_sync<RB, RING_THREADS>(BAR_EMPTY0 + sidx % RB);
#ifdef GS_QR_PROFILE
        prof[1] += clock64() - tb0;
#endif
        cnt = 0;
    }
    GS_DEV void publish(int end) {
        if constexpr (LOG) return;
        if (!wantZ) return;
        const int b = sidx % RB;
        if (lane == 0) {
            hdr->count[b] = cnt;
            hdr->end[b] = end;
        }
        __syncwarp();
        named_bar_arrive<RB, RING_THREADS>(BAR_FULL0 + b);
        sidx += 1;
    }
    GS_DEV void ensure_space(int m) {
        if constexpr (LOG) return;
        if (!wantZ) return;
        if (cnt + m > cap) {
            publish(0);
            begin_buffer();
        }
    }
    GS_DEV ZOp* slot() { return ring + rbase + cnt; }
    // make room for m ops once, so that the step loop can push without checking
    GS_DEV void reserve_ops(int) {}   // (the ring is small: every push checks for space, see push_*)
    GS_DEV void push_refl_c(int k, const C& tau1, const C& v2) {
        if constexpr (LOG) {
            lg.put_hdr(LOG_REFL, k, 1, k, r_const<R>(0.0), r_const<R>(0.0));
            lg.put4(tau1.re, tau1.im, v2.re, v2.im);
            return;
        }
        if (!wantZ) return;
        if (cnt == cap) {
            publish(0);
            begin_buffer();
        }
        if (lane == 0) {
            ZOp* e = ring + rbase + cnt;
            if constexpr (CPLX) {
                e->op = ZOP_REFL;
                e->k = k;
                e->a[0] = tau1.re;
                e->a[1] = tau1.im;
                e->a[2] = v2.re;
                e->a[3] = v2.im;
            }
        }
        cnt += 1;
    }
    GS_DEV void push_r(int op, int k, const R& a0, const R& a1, const R& a2) {
        if constexpr (LOG) {
            emit_r(op, k, a0, a1, a2);
            return;
        }
        if (!wantZ) return;
        if (cnt == cap) {
            publish(0);
            begin_buffer();
        }
        if (lane == 0) {
            ZOp* e = ring + rbase + cnt;
            if constexpr (!CPLX) {
                e->op = op;
                e->k = k;
                e->a[0] = a0;
                e->a[1] = a1;
                e->a[2] = a2;
            }
        }
        cnt += 1;
    }
    GS_DEV void finish_ring() {
        if constexpr (LOG) return;
        if (!wantZ) return;
        publish(1);
        // balance the outstanding "empty" arrivals of the last RB buffers
        const int last = sidx - 1;
        for (int c = (last - (RB - 1) < 0 ? 0 : last - (RB - 1)); c <= last; ++c) named_bar_sync<RB, RING_THREADS>(BAR_EMPTY0 + c % RB);
    }

    // ================================================================================================
    // complex single shift
    // ================================================================================================
    GS_DEV void emit_refl_c(int k, const C& tau1, const C& v2) {
        if constexpr (LOG) {
            push_refl_c(k, tau1, v2);
            return;
        }
        if (!wantZ) return;
        ensure_space(1);
        if (lane == 0) {
            ZOp* e = slot();
            e->op = ZOP_REFL;
            e->k = k;
            e->k2 = k;
            e->a[0] = tau1.re;
            e->a[1] = tau1.im;
            e->a[2] = v2.re;
            e->a[3] = v2.im;
        }
        cnt += 1;
    }
    GS_DEV void emit_scale_c(int j0, int j1, const C& t) {
        if constexpr (LOG) {
            if (j1 >= j0) lg.put_hdr(LOG_SCALE, j0, 0, j1, t.re, t.im);
            return;
        }
        if (!wantZ || j1 < j0) return;
        ensure_space(1);
        if (lane == 0) {
            ZOp* e = slot();
            e->op = ZOP_SCALE;
            e->k = j0;
            e->k2 = j1;
            e->a[0] = t.re;
            e->a[1] = t.im;
            e->a[2] = r_const<R>(0.0);
            e->a[3] = r_const<R>(0.0);
        }
        cnt += 1;
    }

    // ---- barriers among the NH H-warps (a plain __syncwarp when there is one) ----
    GS_DEV static void hbar() {
        if constexpr (NH > 1) asm volatile("bar.sync %0, %1;" ::"n"(BAR_HSTEP), "n"(32 * NH) : "memory");
        else __syncwarp();
    }
    GS_DEV static void cbar() {
        if constexpr (NH > 1) asm volatile("bar.sync %0, %1;" ::"n"(BAR_HCMD), "n"(32 * NH) : "memory");
    }

    // The step loop of a single-shift sweep (src/GenericSchur.jl:426-482) for ONE slot: column / row j = lane + 1 + 32 slot
    // of every lane.  All NH warps execute it in lock step — the reflector is formed redundantly by each (same inputs from
    // shared memory, same result), the shared entries are written by the driver only, the two barriers of a step order the
    // left update, the right update and the next reflector's inputs exactly as the __syncwarp's of the one-warp loop do.
    // (one copy of the loop for driver and helpers: `driver` is a warp-uniform run-time flag — two instantiations would
    // put two copies of the double-double arithmetic into the instruction caches)
    __device__ __noinline__ unsigned sweep_steps(bool driver, int slot, int k0, int istart, int iend, C v0, C v1) {
        const R zero = r_const<R>(0.0), one = r_const<R>(1.0);
        constexpr uint32_t ES = (uint32_t)sizeof(T);
        const uint32_t hb = smem_u32(H);
        const int j = lane + 1 + 32 * slot;
        const bool valid = j <= n;
        const int jj = valid ? j : -(1 << 28), ii = valid ? j : (1 << 28);
        const uint32_t ca = hb + ES * (uint32_t)(colbase(valid ? j : 1) - 1), ib = ES * (uint32_t)(valid ? j : 1);
        uint32_t ak = hb + ES * (uint32_t)(colbase(k0) - 1);   // column k
        C cL = (jj >= k0) ? lds_e<T>(ca + ES * k0) : mk_cx<R>(zero, zero);   // H[k, j] of the owned column j >= k
        C dR = (ii < k0) ? lds_e<T>(ak + ib) : mk_cx<R>(zero, zero);         // H[i, k] of the owned row i < k
        unsigned napplied = 0;
        for (int k = k0; k <= iend - 1; ++k) {
            const uint32_t kb = ES * (uint32_t)k;
            const uint32_t ak1 = ak + ES * (uint32_t)(k + EX);          // column k+1
            const uint32_t akm = ak - ES * (uint32_t)(k - 1 + EX);      // column k-1
            if (k > k0) {
                v0 = lds_e<T>(akm + kb);
                v1 = lds_e<T>(akm + kb + ES);
            }
            const C tau1 = reflector_cplx2(v0, v1);
            napplied += 1;
            const C v2 = v1, v2c = cconj(v1), tau1c = cconj(tau1);
            const R tau2 = (tau1 * v2).re;
            if (driver) push_refl_c(k, tau1, v2);
            // ---- left update of rows k, k+1: the owned column j >= k ----
            if (jj >= k) {
                const uint32_t a = ca + kb;
                const C b = lds_e<T>(a + ES);
                const C ss = e_axty<true>(tau1c, cL, tau2, b);
                sts_e<T>(a, cL - ss);
                cL = e_bsv<true>(b, ss, v2);
                if (jj <= k + 1) sts_e<T>(a + ES, cL);   // diagonal columns publish their carry
            }
            hbar();
            // ---- right update of columns k, k+1: the owned row i <= min(k+2, iend) ----
            const int jmax = (k + 2 < iend) ? k + 2 : iend;
            if (ii <= jmax) {
                const C d = (ii >= k) ? lds_e<T>(ak + ib) : dR;
                const C e = lds_e<T>(ak1 + ib);
                const C ss = e_axty<true>(tau1, d, tau2, e);
                sts_e<T>(ak + ib, d - ss);
                dR = e_bsv<true>(e, ss, v2c);
                if (ii >= k + 1) sts_e<T>(ak1 + ib, dR);
            }
            if (driver && lane == 0 && k > k0) {
                sts_e<T>(akm + kb, v0);
                sts_e<T>(akm + kb + ES, mk_cx<R>(zero, zero));
            }
            hbar();
            // column k+1's running entry H[k+1,k+1] was just rewritten by its own lane as row k+1
            if (jj == k + 1) cL = dR;
            ak = ak1;
            if (k == k0 && k0 > istart) {
                // late start (src/GenericSchur.jl:461-482): flush the carries, the driver rescales in shared memory, reload
                if (j >= k + 2 && j <= n) HH(k + 1, j) = cL;
                if (j <= k) HH(j, k + 1) = dR;
                hbar();
                if (driver) {
                    C t = mk_cx<R>(one, zero) - tau1;
                    const R rat = q_rcp(c_abs_q(t));
                    t = mk_cx<R>(t.re * rat, t.im * rat);
                    const C tc = cconj(t);
                    if (lane == 0) {
                        HH(k0 + 1, k0) = HH(k0 + 1, k0) * tc;
                        if (k0 + 2 <= iend) HH(k0 + 2, k0 + 1) = HH(k0 + 2, k0 + 1) * t;
                    }
                    __syncwarp();
                    for (int q = k0; q <= iend; ++q) {
                        if (q == k0 + 1) continue;
                        for (int c = q + 1 + lane; c <= n; c += 32) HH(q, c) = HH(q, c) * t;
                        for (int r = 1 + lane; r <= q - 1; r += 32) HH(r, q) = HH(r, q) * tc;
                        __syncwarp();
                    }
                    emit_scale_c(k0, k0, tc);
                    emit_scale_c(k0 + 2, iend, tc);
                }
                hbar();
                cL = (j >= k + 1 && j <= n) ? HH(k + 1, j) : mk_cx<R>(zero, zero);
                dR = (j < k + 1) ? HH(j, k + 1) : mk_cx<R>(zero, zero);
            }
        }
        // ---- flush the carries left after the last step (k = iend-1) ----
        if (j >= iend + 1 && j <= n) HH(iend, j) = cL;
        if (j <= iend - 1) HH(j, iend) = dR;
        hbar();
        return napplied;
    }

    // helpers: wait for the driver's command, run the step loop for the slot, repeat
    GS_DEV void helper_loop(int slot) {
        for (;;) {
            cbar();
            const int op = hcmd->op;
            if (op == HCMD_EXIT) break;
            const int k0 = hcmd->k0, is = hcmd->istart, ie = hcmd->iend;
            const C v0 = hcmd->v0, v1 = hcmd->v1;
            sweep_steps(false, slot, k0, is, ie, v0, v1);
        }
    }
    GS_DEV void helpers_exit() {
        if constexpr (NH > 1) {
            if (lane == 0) hcmd->op = HCMD_EXIT;
            __syncwarp();
            cbar();
        }
    }

    GS_DEV void sweep_complex(const C& shift, int istart, int iend) {
        const R zero = r_const<R>(0.0), one = r_const<R>(1.0);
        const R ulp = rtraits<R>::eps();
        int istart1 = 0;
        for (int base = iend - 1; base >= istart + 1 && !istart1; base -= 32) {
            int mm = base - lane;
            bool hit = false;
            if (mm >= istart + 1) {
                C h11 = HH(mm, mm), h22 = HH(mm + 1, mm + 1);
                C h11s = h11 - shift;
                R h21 = HH(mm + 1, mm).re;
                const R rs = q_rcp(abs1(h11s) + r_abs(h21));
                h11s = mk_cx<R>(h11s.re * rs, h11s.im * rs);
                h21 = h21 * rs;
                R h10 = HH(mm, mm - 1).re;
                hit = r_abs(h10) * r_abs(h21) <= ulp * (abs1(h11s) * (abs1(h11) + abs1(h22)));
            }
            unsigned m = __ballot_sync(0xffffffffu, hit);
            if (m) istart1 = base - (__ffs(m) - 1);
        }
        if (!istart1) istart1 = istart;
        const int k0 = istart1;
        C v0, v1;
        {
            C h11s = HH(k0, k0) - shift;
            R h21 = HH(k0 + 1, k0).re;
            const R rs = q_rcp(abs1(h11s) + r_abs(h21));
            v0 = mk_cx<R>(h11s.re * rs, h11s.im * rs);
            v1 = mk_cx<R>(h21 * rs, zero);
        }
        reserve_ops(iend - k0 + 4);
        if constexpr (NH > 1) {
            // several H-warps: hand the sweep to the helpers, run the driver's own slot
            if (lane == 0) {
                hcmd->op = HCMD_SWEEP;
                hcmd->k0 = k0;
                hcmd->istart = istart;
                hcmd->iend = iend;
                hcmd->v0 = v0;
                hcmd->v1 = v1;
            }
            __syncwarp();
            cbar();
            stp[1] += sweep_steps(true, 0, k0, istart, iend, v0, v1);
            tail_fix_c(iend);
            return;
        }
        // All shared-memory traffic of the step loop goes through 32-bit byte addresses:
        //   &H(i, j) = hb + ES * (colbase(j) - 1 + i).   jj / ii are the owned column / row (sentinels when > n).
        constexpr uint32_t ES = (uint32_t)sizeof(T);
        const uint32_t hb = smem_u32(H);
        uint32_t ca[CPL], ib[CPL];
        int jj[CPL], ii[CPL];
        C cL[CPL], dR[CPL];   // cL = H[k, j] for the owned column j >= k;  dR = H[i, k] for the owned row i < k
        uint32_t ak = hb + ES * (uint32_t)(colbase(k0) - 1);   // column k
#pragma unroll
        for (int s = 0; s < CPL; ++s) {
            const int j = lane + 1 + 32 * s;
            const bool valid = j <= n;
            jj[s] = valid ? j : -(1 << 28);
            ii[s] = valid ? j : (1 << 28);
            ca[s] = hb + ES * (uint32_t)(colbase(valid ? j : 1) - 1);
            ib[s] = ES * (uint32_t)(valid ? j : 1);
            cL[s] = (jj[s] >= k0) ? lds_e<T>(ca[s] + ES * k0) : mk_cx<R>(zero, zero);
            dR[s] = (ii[s] < k0) ? lds_e<T>(ak + ib[s]) : mk_cx<R>(zero, zero);
        }
        unsigned napplied = 0;
        for (int k = k0; k <= iend - 1; ++k) {
            const uint32_t kb = ES * (uint32_t)k;
            const uint32_t ak1 = ak + ES * (uint32_t)(k + EX);          // column k+1
            const uint32_t akm = ak - ES * (uint32_t)(k - 1 + EX);      // column k-1
            if (k > k0) {
                v0 = lds_e<T>(akm + kb);
                v1 = lds_e<T>(akm + kb + ES);
            }
            const C tau1 = reflector_cplx2(v0, v1);
            napplied += 1;
            const C v2 = v1, v2c = cconj(v1), tau1c = cconj(tau1);
            const R tau2 = (tau1 * v2).re;
            push_refl_c(k, tau1, v2);
            // ---- left update of rows k, k+1: owned columns j >= k ----
#pragma unroll
            for (int s = 0; s < CPL; ++s) {
                if (jj[s] >= k) {
                    const uint32_t a = ca[s] + kb;
                    const C b = lds_e<T>(a + ES);
                    const C ss = e_axty<(CPL >= 2)>(tau1c, cL[s], tau2, b);
                    sts_e<T>(a, cL[s] - ss);
                    cL[s] = e_bsv<(CPL >= 2)>(b, ss, v2);
                    if (jj[s] <= k + 1) sts_e<T>(a + ES, cL[s]);   // diagonal columns publish their carry
                }
            }
            __syncwarp();
            // ---- right update of columns k, k+1: owned rows i <= min(k+2, iend) ----
            const int jmax = (k + 2 < iend) ? k + 2 : iend;
#pragma unroll
            for (int s = 0; s < CPL; ++s) {
                if (ii[s] <= jmax) {
                    const C d = (ii[s] >= k) ? lds_e<T>(ak + ib[s]) : dR[s];
                    const C e = lds_e<T>(ak1 + ib[s]);
                    const C ss = e_axty<(CPL >= 2)>(tau1, d, tau2, e);
                    sts_e<T>(ak + ib[s], d - ss);
                    dR[s] = e_bsv<(CPL >= 2)>(e, ss, v2c);
                    if (ii[s] >= k + 1) sts_e<T>(ak1 + ib[s], dR[s]);
                }
            }
            if (lane == 0 && k > k0) {
                sts_e<T>(akm + kb, v0);
                sts_e<T>(akm + kb + ES, mk_cx<R>(zero, zero));
            }
            __syncwarp();
            // column k+1's running entry H[k+1,k+1] was just rewritten by its own lane as row k+1
#pragma unroll
            for (int s = 0; s < CPL; ++s)
                if (jj[s] == k + 1) cL[s] = dR[s];
            ak = ak1;

            if (k == k0 && k0 > istart) {
                // late start (src/GenericSchur.jl:461-482): flush the carries, rescale in shared memory, reload
#pragma unroll
                for (int s = 0; s < CPL; ++s) {
                    const int j = lane + 1 + 32 * s;
                    if (j >= k + 2 && j <= n) HH(k + 1, j) = cL[s];
                    if (j <= k) HH(j, k + 1) = dR[s];
                }
                __syncwarp();
                C t = mk_cx<R>(one, zero) - tau1;
                R at = c_abs(t);
                t = mk_cx<R>(t.re / at, t.im / at);
                const C tc = cconj(t);
                if (lane == 0) {
                    HH(k0 + 1, k0) = HH(k0 + 1, k0) * tc;
                    if (k0 + 2 <= iend) HH(k0 + 2, k0 + 1) = HH(k0 + 2, k0 + 1) * t;
                }
                __syncwarp();
                for (int j = k0; j <= iend; ++j) {
                    if (j == k0 + 1) continue;
                    for (int c = j + 1 + lane; c <= n; c += 32) HH(j, c) = HH(j, c) * t;
                    for (int r = 1 + lane; r <= j - 1; r += 32) HH(r, j) = HH(r, j) * tc;
                    __syncwarp();
                }
                emit_scale_c(k0, k0, tc);
                emit_scale_c(k0 + 2, iend, tc);
#pragma unroll
                for (int s = 0; s < CPL; ++s) {
                    const int j = lane + 1 + 32 * s;
                    cL[s] = (j >= k + 1 && j <= n) ? HH(k + 1, j) : mk_cx<R>(zero, zero);
                    dR[s] = (j < k + 1) ? HH(j, k + 1) : mk_cx<R>(zero, zero);
                }
            }
        }
        stp[1] += napplied;
        // ---- flush the carries left after the last step (k = iend-1) ----
#pragma unroll
        for (int s = 0; s < CPL; ++s) {
            const int j = lane + 1 + 32 * s;
            if (j >= iend + 1 && j <= n) HH(iend, j) = cL[s];
            if (j <= iend - 1) HH(j, iend) = dR[s];
        }
        __syncwarp();
        tail_fix_c(iend);
    }

    // make the tail sub-diagonal real (src/GenericSchur.jl:486-500); the H-warp (driver) alone
    GS_DEV void tail_fix_c(int iend) {
        const R zero = r_const<R>(0.0);
        C t = HH(iend, iend - 1);
        if (t.im != zero) {
            R rt = c_abs_q(t);
            const R rrt = q_rcp(rt);
            t = mk_cx<R>(t.re * rrt, t.im * rrt);
            const C tc = cconj(t);
            for (int c = iend + 1 + lane; c <= n; c += 32) HH(iend, c) = HH(iend, c) * tc;
            for (int r = 1 + lane; r <= iend - 1; r += 32) HH(r, iend) = HH(r, iend) * t;
            emit_scale_c(iend, iend, t);
            __syncwarp();
            if (lane == 0) HH(iend, iend - 1) = mk_cx<R>(rt, zero);
        }
        __syncwarp();
    }

    // ================================================================================================
    // complex single shift, software-pipelined (ComplexF64).
    //
    // The 2x2 diagonal block the bulge sits on (and the sub-diagonal entry below it) lives in registers, replicated
    // in every lane, so the chain  reflector k -> block update -> reflector k+1  never touches shared memory and
    // needs no lane hand-off.  Every other entry of rows k, k+1 (columns >= k+2: "left items") and of columns
    // k, k+1 (rows <= k-1: "right items") is a bulk item: lane l owns the indices l+1+32s and holds one running
    // entry per index (H[k, j] of its column while j >= k+2, H[j, k] of its row once j <= k-1).  One loop iteration
    // applies reflector k to the bulk items AND forms reflector k+1 from the register block; the two instruction
    // streams are independent and sit in one basic block, so the long sqrt / reciprocal chain of the reflector is
    // overlapped with the bulk arithmetic.  One __syncwarp per step orders the shared-memory traffic.
    // Same arithmetic per entry as src/GenericSchur.jl:426-459 (left: :442-446, right: :448-452).
    // ================================================================================================
    GS_DEV void late_start_step(int k0, int iend, C v0, C v1) {
        // the first step of a sweep that starts inside the active block (src/GenericSchur.jl:461-482), in shared memory
        const R zero = r_const<R>(0.0), one = r_const<R>(1.0);
        const C tau1 = reflector_cplx2(v0, v1);
        const C v2 = v1, v2c = cconj(v1), tau1c = cconj(tau1);
        const R tau2 = (tau1 * v2).re;
        push_refl_c(k0, tau1, v2);
        for (int j = k0 + lane; j <= n; j += 32) {
            const C a = HH(k0, j), b = HH(k0 + 1, j);
            const C ss = tau1c * a + tau2 * b;
            HH(k0, j) = a - ss;
            HH(k0 + 1, j) = b - ss * v2;
        }
        __syncwarp();
        const int jmax = (k0 + 2 < iend) ? k0 + 2 : iend;
        for (int i = 1 + lane; i <= jmax; i += 32) {
            const C d = HH(i, k0), e = HH(i, k0 + 1);
            const C ss = tau1 * d + tau2 * e;
            HH(i, k0) = d - ss;
            HH(i, k0 + 1) = e - ss * v2c;
        }
        __syncwarp();
        C t = mk_cx<R>(one, zero) - tau1;
        R at = c_abs(t);
        t = mk_cx<R>(t.re / at, t.im / at);
        const C tc = cconj(t);
        if (lane == 0) {
            HH(k0 + 1, k0) = HH(k0 + 1, k0) * tc;
            if (k0 + 2 <= iend) HH(k0 + 2, k0 + 1) = HH(k0 + 2, k0 + 1) * t;
        }
        __syncwarp();
        for (int j = k0; j <= iend; ++j) {
            if (j == k0 + 1) continue;
            for (int c = j + 1 + lane; c <= n; c += 32) HH(j, c) = HH(j, c) * t;
            for (int r = 1 + lane; r <= j - 1; r += 32) HH(r, j) = HH(r, j) * tc;
            __syncwarp();
        }
        emit_scale_c(k0, k0, tc);
        emit_scale_c(k0 + 2, iend, tc);
    }

    GS_DEV static double flip_if(double x, unsigned m) {   // x with its sign bit xor-ed by m (0 or 0x80000000)
        return __hiloint2double(__double2hiint(x) ^ (int)m, __double2loint(x));
    }

    GS_DEV void sweep_complex_pipelined(const C& shift, int istart, int iend) {
        static_assert(sizeof(R) == 8, "ComplexF64 only");
#ifdef GS_QR_PROFILE
        const long long tp0 = clock64();
#endif
        const R zero = 0.0;
        const R ulp = rtraits<R>::eps();
        int istart1 = 0;
        for (int base = iend - 1; base >= istart + 1 && !istart1; base -= 32) {
            int mm = base - lane;
            bool hit = false;
            if (mm >= istart + 1) {
                const C h11 = HH(mm, mm), h22 = HH(mm + 1, mm + 1);
                const C h11s = h11 - shift;
                const R h21 = HH(mm + 1, mm).re;
                const R rs = q_rcp(abs1(h11s) + r_abs(h21));
                const R h10 = HH(mm, mm - 1).re;
                hit = r_abs(h10) * r_abs(h21 * rs) <=
                      ulp * ((r_abs(h11s.re * rs) + r_abs(h11s.im * rs)) * (abs1(h11) + abs1(h22)));
            }
            unsigned m = __ballot_sync(0xffffffffu, hit);
            if (m) istart1 = base - (__ffs(m) - 1);
        }
        if (!istart1) istart1 = istart;
        const int k0 = istart1;
        C v0, v1;
        {
            C h11s = HH(k0, k0) - shift;
            R h21 = HH(k0 + 1, k0).re;
            R rs = q_rcp(abs1(h11s) + r_abs(h21));
            v0 = mk_cx<R>(h11s.re * rs, h11s.im * rs);
            v1 = mk_cx<R>(h21 * rs, zero);
        }
        int kf = k0;
        bool store_sub = false;   // does step kf write (beta, 0) into column kf-1 ?
        unsigned napplied = 0;
        if (k0 > istart) {
            late_start_step(k0, iend, v0, v1);
            napplied = 1;
            kf = k0 + 1;
            store_sub = true;
            if (kf <= iend - 1) {
                v0 = HH(kf, kf - 1);
                v1 = HH(kf + 1, kf - 1);
            }
        }
#ifdef GS_QR_PROFILE
        const long long tp1 = clock64();
        (void)tp0;
#endif
        if (kf > iend - 1) {
            // the late-start step was the only one: make the tail sub-diagonal real (src/GenericSchur.jl:486-500)
            stp[1] += napplied;
            C t = HH(iend, iend - 1);
            if (t.im != zero) {
                R rt = c_abs(t);
                t = mk_cx<R>(t.re / rt, t.im / rt);
                const C tc = cconj(t);
                for (int cidx = iend + 1 + lane; cidx <= n; cidx += 32) HH(iend, cidx) = HH(iend, cidx) * tc;
                for (int r = 1 + lane; r <= iend - 1; r += 32) HH(r, iend) = HH(r, iend) * t;
                emit_scale_c(iend, iend, t);
                __syncwarp();
                if (lane == 0) HH(iend, iend - 1) = mk_cx<R>(rt, zero);
            }
            __syncwarp();
            return;
        }
        constexpr uint32_t ES = (uint32_t)sizeof(T);
        const uint32_t hb = smem_u32(H);
        // Bulk items.  Index j = lane+1+32s is a LEFT item (column j, rows k, k+1) while j >= k+2 and a RIGHT item
        // (row j, columns k, k+1) once j <= k-1.  Right items are held CONJUGATED: conj(ss) = conj(tau1) conj(x) +
        // tau2 conj(y), conj(x') = conj(x) - conj(ss), conj(y') = conj(y) - conj(ss) v2 — the left item's formulas —
        // so one instruction stream with fixed coefficients serves both roles; only the imaginary sign of what is
        // loaded / stored differs (xor mask sg).  Entries enter and leave the register block through shared memory.
        uint32_t ca[CPL], ib[CPL];
        int jl[CPL];
        C c[CPL];
        uint32_t ak = hb + ES * (uint32_t)(colbase(kf) - 1);   // column k, "row 0"
#pragma unroll
        for (int s = 0; s < CPL; ++s) {
            const int j = lane + 1 + 32 * s;
            const bool valid = j <= n;
            jl[s] = valid ? j : -(1 << 28);
            ca[s] = hb + ES * (uint32_t)(colbase(valid ? j : 1) - 1);
            ib[s] = ES * (uint32_t)(valid ? j : 1);
            c[s] = mk_cx<R>(zero, zero);
            if (jl[s] >= kf + 2) c[s] = lds_e<T>(ca[s] + ES * kf);
            else if (j <= kf - 1) c[s] = cconj(lds_e<T>(ak + ib[s]));
        }
        // register block: d00 d01 / d10 d11 = H[k..k+1, k..k+1], e1 = H[k+2, k+1] (0 past the active block).  The first
        // column is carried in registers, the second one is picked up from shared memory at the top of each step.
        C d00 = HH(kf, kf), d10 = HH(kf + 1, kf);
        C tau1 = reflector_cplx2(v0, v1);
        R beta = v0.re;
        C v2 = v1;
        R tau2 = tau1.re * v2.re - tau1.im * v2.im;   // Re(tau1 v2); tau1 v2 is real because the reflector's tail is
        C f_sub, f_diag;   // H[iend, iend-1], H[iend, iend] after the last step
        const int capz = (wantZ && !LOG) ? cap : 0x7fffffff;
        const uint32_t ring32 = LOG ? 0u : smem_u32(ring);
        if constexpr (LOG) lg.put_hdr(LOG_REFL, kf, iend - kf, iend, zero, zero);
#ifdef GS_QR_PROFILE
        const long long tp2 = clock64();
        prof[2] += tp2 - tp1;
#endif
        for (int k = kf;; ++k) {
            if constexpr (!LOG) {
                if (cnt == capz) {
                    publish(0);
                    begin_buffer();
                }
            }
            __syncwarp();
            const uint32_t kb = ES * (uint32_t)k;
            const uint32_t ak1 = ak + ES * (uint32_t)(k + EX);        // column k+1
            // ---- loads: column k+1 of the block; second entry of every bulk item ----
            const C d01 = lds_e<T>(ak1 + kb), d11 = lds_e<T>(ak1 + kb + ES);
            // the sub-diagonal is real (stage A, every bulge step and the end-of-sweep fix-up write (beta, 0)): only Re is read
            R e1 = zero;
            if (k + 2 <= iend) e1 = lds_e<R>(ak1 + kb + 2 * ES);
            // Row k itself (j == k) is a right item too: its two entries are rows k of the block after the left update (a00,
            // a01, known to every lane) — the owner lane takes them over from the chain below, stores H[k, k] and carries
            // H[k, k+1] into the next step, so the chain does not have to finish that row.
            bool act[CPL], own2[CPL], isD[CPL];
            unsigned sg[CPL];
            uint32_t sa[CPL];
            C y[CPL];
#pragma unroll
            for (int s = 0; s < CPL; ++s) {
                const int j = lane + 1 + 32 * s;
                const bool isL = jl[s] >= k + 2;
                const bool isR = j <= k;
                isD[s] = j == k;
                act[s] = isL || isR;
                own2[s] = jl[s] == k + 2;
                sg[s] = isR ? 0x80000000u : 0u;
                sa[s] = isR ? ak + ib[s] : ca[s] + kb;                  // inactive lanes: a harmless address
                const uint32_t ya = isR ? ak1 + ib[s] : sa[s] + ES;
                y[s] = lds_e<T>(ya);
                y[s].im = flip_if(y[s].im, sg[s]);
            }
            const C tau1c = cconj(tau1), v2c = cconj(v2);
            // ---- chain: rows k, k+1 of columns k, k+1 (left), then columns k, k+1 of rows k..k+2 (right) ----
            const C ss0 = tau1c * d00 + tau2 * d10;
            const C a00 = d00 - ss0, a10 = e_fnma(ss0, v2, d10);
            const C ss1 = tau1c * d01 + tau2 * d11;
            const C a01 = d01 - ss1, a11 = e_fnma(ss1, v2, d11);
#pragma unroll
            for (int s = 0; s < CPL; ++s) {
                c[s].re = isD[s] ? a00.re : c[s].re;
                c[s].im = isD[s] ? -a00.im : c[s].im;
                y[s].re = isD[s] ? a01.re : y[s].re;
                y[s].im = isD[s] ? -a01.im : y[s].im;
            }
            const C sr1 = tau1 * a10 + tau2 * a11;
            // H[k+1, k] = a10 - sr1 is the head of the next reflector, i.e. the serial chain of the sweep.  v2 is the last
            // thing the previous reflector produces, so the expression is regrouped to need it once, at the very end:
            // a10 - sr1 = (1 - tau1) a10 - tau2 a11 = [(1 - tau1) d10 - tau2 d11] - v2 [(1 - tau1) ss0 - tau2 ss1].
            // Measured on B200: -18 % stage-B time for n <= 32 (CPL = 1); +8 % for n = 64 (CPL = 2), where the extra live values
            // collide with the 168-register cap that six CTAs per SM impose — so the regrouping is used for CPL = 1 only.
#ifndef GS_QR_REGROUP_MAXCPL
#define GS_QR_REGROUP_MAXCPL 1
#endif
            C n_v0;
            if constexpr (CPL <= GS_QR_REGROUP_MAXCPL) {
                const C omt = mk_cx<R>(1.0 - tau1.re, -tau1.im);
                const C pv = omt * d10 - tau2 * d11, qv = omt * ss0 - tau2 * ss1;
                n_v0 = e_fnma(v2, qv, pv);
            } else {
                n_v0 = a10 - sr1;
            }
            const C n_d00 = e_fnma(sr1, v2c, a11);   // H[k+1, k], H[k+1, k+1]
            const R n_v1 = -tau2 * e1;                                         // H[k+2, k], the bulge: real
            const C n_d10 = mk_cx<R>(fma(n_v1, v2.re, e1), -n_v1 * v2.im);    // H[k+2, k+1] = e1 + n_v1 conj(v2)
            const bool last = (k == iend - 1);
            {
                const bool l0 = lane == 0;
                if constexpr (LOG) {
                    const bool okl = lg.reserve();
                    stg_2f64_if(lg.slot(), tau1.re, tau1.im, okl && l0);
                    stg_2f64_if(lg.slot() + 16, v2.re, v2.im, okl && l0);
                    if (okl) lg.advance();
                } else {
                    const uint32_t re = ring32 + (uint32_t)sizeof(ZOp) * (uint32_t)(rbase + cnt);
                    sts_2i_if(re, (int)ZOP_REFL, k, l0 && wantZ);
                    sts_c64_if(re + 16, tau1, l0 && wantZ);
                    sts_c64_if(re + 32, v2, l0 && wantZ);
                }
                const uint32_t akm = ak - ES * (uint32_t)(k - 1 + EX);
                const bool sub = l0 && (k > kf || store_sub);
                sts_c64_if(akm + kb, mk_cx<R>(beta, zero), sub);
                sts_c64_if(akm + kb + ES, mk_cx<R>(zero, zero), sub);
            }
            cnt += 1;
            // ---- reflector k+1 (straight-line; the general routine only for out-of-range / degenerate input) ----
            C t1n, v2n;
            R betan, tau2n;
            bool ok;
            {
                const double a = n_v0.re, b = n_v0.im, cc = n_v1;
                // the bulge entry n_v1 = -tau2 e1 is real (tau2 and the sub-diagonal are)
                const double q = fma(a, a, fma(b, b, cc * cc));
                // q in [2^-900, 2^900]; tail and Im(alpha) not all exactly zero (src/householder.jl:67-69)
                const unsigned tz = ((unsigned)(__double2hiint(cc) | __double2hiint(b)) << 1) |
                                    (unsigned)(__double2loint(cc) | __double2loint(b));
                ok = q_exp_in(q, 1023u - 900u, 1023u + 900u) && (tz != 0u);
                double yr0;
                asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(yr0) : "d"(q));
                // The seed of 1/|alpha - beta|^2 is taken from the unrefined norm so that the second MUFU overlaps the
                // refinement of the first; its one cubic Newton step below uses the final denominator.
                const double amb0 = a + copysign(q * yr0, a);
                double y0;
                asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(fma(amb0, amb0, b * b)));
                const double qy = q * yr0;
                const double e = fma(-qy, yr0, 1.0);
                const double cf = fma(e, 0.375, 0.5);
                const double yr = fma(yr0 * e, cf, yr0);   // 1/sqrt(q), one cubic step: 2^-22 -> below 1 ulp
                const double nrm = q * yr;                 // sqrt(q) within ~1 ulp
                betan = -copysign(nrm, a);
                const double rb = -copysign(yr, a);        // 1/beta
                t1n = mk_cx<R>((betan - a) * rb, -b * rb);
                tau2n = -cc * rb;                          // Re(tau v2) = -x2 / beta for a real tail x2
                const double amb = a - betan;              // |amb| >= |beta|: no cancellation
                const double den = fma(amb, amb, b * b);
                const double e2 = fma(-den, y0, 1.0);
                const double rm = fma(y0, fma(e2, e2, e2), y0);   // y0 (1 + e + e^2)
                const double tr = amb * rm, ti = -b * rm;
                v2n = mk_cx<R>(cc * tr, cc * ti);
            }
            // ---- bulk items: reflector k on the owned columns (left) / rows (right, conjugated) ----
#pragma unroll
            for (int s = 0; s < CPL; ++s) {
                const C x = c[s];
                const C ss = tau1c * x + tau2 * y[s];
                C st = x - ss;
                st.im = flip_if(st.im, sg[s]);
                sts_c64_if(sa[s], st, act[s]);
                const C nc = e_fnma(ss, v2, y[s]);
                c[s].re = act[s] ? nc.re : c[s].re;
                c[s].im = act[s] ? nc.im : c[s].im;
                sts_c64_if(sa[s] + ES, nc, own2[s]);   // column k+2 enters the register block next step
            }
            if (last) {
                f_sub = n_v0;
                f_diag = n_d00;
                break;
            }
            if (!ok) {
                v0 = n_v0;
                v1 = mk_cx<R>(n_v1, zero);
                t1n = reflector_cplx2_generic<R>(v0, v1);
                betan = v0.re;
                v2n = v1;
                tau2n = t1n.re * v2n.re - t1n.im * v2n.im;
            }
            tau1 = t1n;
            v2 = v2n;
            tau2 = tau2n;
            beta = betan;
            d00 = n_d00;
            d10 = n_d10;
            ak = ak1;
        }
#ifdef GS_QR_PROFILE
        prof[3] += clock64() - tp2;
#endif
        stp[1] += napplied + (unsigned)(iend - kf);
        // ---- write back what is still in registers after the last step (k = iend-1); the unit-modulus factor that makes
        //      H[iend, iend-1] real (src/GenericSchur.jl:486-500) is applied on the way: row iend right of the diagonal
        //      gets conj(t), column iend above it gets t — both are a multiplication by conj(t) of what the lanes hold ----
        C tph = mk_cx<R>(1.0, zero);
        R fsr = f_sub.re;
        const bool fix = f_sub.im != zero;
        if (fix) {
            fsr = c_abs_q(f_sub);
            const R ri = q_rcp(fsr);
            tph = mk_cx<R>(f_sub.re * ri, f_sub.im * ri);
            emit_scale_c(iend, iend, tph);
        }
        const C tphc = cconj(tph);
        __syncwarp();
#pragma unroll
        for (int s = 0; s < CPL; ++s) {
            const int j = lane + 1 + 32 * s;
            C v = c[s];
            if (fix) v = v * tphc;
            if (j >= iend + 1 && j <= n) HH(iend, j) = v;
            if (j <= iend - 1) HH(j, iend) = cconj(v);   // includes H[iend-1, iend], carried by the owner of row iend-1
        }
        if (lane == 0) {
            HH(iend, iend - 1) = mk_cx<R>(fsr, zero);
            HH(iend, iend) = f_diag;
        }
        __syncwarp();
    }

    GS_DEV int qr_complex(int maxiter, unsigned* st) {
        stp = st;
#ifdef GS_QR_PROFILE
        prof[0] = prof[1] = prof[2] = prof[3] = 0;
        const long long tq0 = clock64();
#endif
        const R zero = r_const<R>(0.0), half = r_const<R>(0.5), threeq = r_const<R>(0.75);
        const R ulp = rtraits<R>::eps();
        const R smallnum = r_safemin<R>() * (r_const<R>((double)n) / ulp);
        const int maxinner = 30 * n;
        int istart = 1, iend = n, it = 0;
        sidx = 0;
        cnt = 0;
        begin_buffer();
        while (iend >= 1) {
            istart = 1;
            for (int its = 0; its <= maxinner; ++its) {
                it += 1;
                if (it > maxiter) {
                    st[3] = it;
                    finish_ring();
                    return iend;
                }
                int found = 0;
                for (int base = iend - 1; base >= istart && !found; base -= 32) {
                    int c = base - lane;
                    bool hit = (c >= istart) ? split_test_c(c, smallnum, ulp) : false;
                    unsigned m = __ballot_sync(0xffffffffu, hit);
                    if (m) found = base - (__ffs(m) - 1);
                }
                if (found) istart = found + 1;
                __syncwarp();
                if (istart > 1 && lane == 0) HH(istart, istart - 1) = mk_cx<R>(zero, zero);
                __syncwarp();
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
                    t = HH(iend, iend);
                    C u = c_sqrt_q(HH(iend - 1, iend)) * c_sqrt_q(HH(iend, iend - 1));
                    R s = abs1(u);
                    if (s != zero) {
                        C x = half * (HH(iend - 1, iend - 1) - t);
                        R sx = abs1(x);
                        s = r_max(s, sx);
                        const R rs = q_rcp(s);
                        C xs = mk_cx<R>(x.re * rs, x.im * rs), us = mk_cx<R>(u.re * rs, u.im * rs);
                        C y = s * c_sqrt_q(xs * xs + us * us);
                        if (sx > zero) {
                            const R rsx = q_rcp(sx);
                            if ((x.re * rsx) * y.re + (x.im * rsx) * y.im < zero) y = -y;
                        }
                        t = t - u * c_div_q(u, x + y);
                    }
                }
                st[0] += 1;
                if constexpr (sizeof(R) == 8) sweep_complex_pipelined(t, istart, iend);
                else sweep_complex(t, istart, iend);
                if constexpr (LOG) {
                    if (lg.ovf) return LOG_OVERFLOW_RC;   // the fused kernel redoes this matrix
                }
                // hand the sweep's reflectors to the Z-warp
                publish(0);
                begin_buffer();
            }
        }
        st[3] = it;
#ifdef GS_QR_PROFILE
        prof[0] = clock64() - tq0;
        st[0] = (unsigned)(prof[0] >> 6);   // total
        st[2] = (unsigned)(prof[1] >> 6);   // waiting for the Z-warp to release a ring buffer
        st[3] = (unsigned)(prof[3] >> 6);   // step loops
        // st[1] keeps the number of steps
#endif
        finish_ring();
        return 0;
    }

    // ================================================================================================
    // real double shift
    // ================================================================================================
    GS_DEV void emit_r(int op, int k, const R& a0, const R& a1, const R& a2) {
        if constexpr (LOG) {
            if (op == ZOP_GIVENS) {
                lg.put_hdr(LOG_GIVENS, k, 0, k, a0, a1);
            } else {
                lg.put_hdr(op == ZOP_REFL3 ? LOG_REFL3 : LOG_REFL2, k, 1, k, r_const<R>(0.0), r_const<R>(0.0));
                lg.put4(a0, a1, a2, r_const<R>(0.0));
            }
            return;
        }
        if (!wantZ) return;
        ensure_space(1);
        if (lane == 0) {
            ZOp* e = slot();
            e->op = op;
            e->k = k;
            e->a[0] = a0;
            e->a[1] = a1;
            e->a[2] = a2;
        }
        cnt += 1;
    }

    GS_DEV void sweep_real(const R& r1r, const R& r1i, const R& r2r, const R& r2i, int istart, int iend) {
        const R zero = r_const<R>(0.0), one = r_const<R>(1.0);
        const R eps = rtraits<R>::eps();
        int mx = 0;
        for (int base = iend - 2; base >= istart + 1 && !mx; base -= 32) {
            int m = base - lane;
            bool hit = false;
            if (m >= istart + 1) {
                R a0, a1, a2;
                first_column_r(m, r1r, r1i, r2r, r2i, a0, a1, a2);
                hit = r_abs(HH(m, m - 1)) * (r_abs(a1) + r_abs(a2)) <=
                      eps * r_abs(a0) * (r_abs(HH(m - 1, m - 1)) + r_abs(HH(m, m)) + r_abs(HH(m + 1, m + 1)));
            }
            unsigned msk = __ballot_sync(0xffffffffu, hit);
            if (msk) mx = base - (__ffs(msk) - 1);
        }
        if (!mx) mx = istart;
        R v0, v1, v2;
        first_column_r(mx, r1r, r1i, r2r, r2i, v0, v1, v2);
        reserve_ops(iend - mx + 4);
        // byte-address arithmetic as in sweep_complex: &H(i, j) = hb + ES * (colbase(j) - 1 + i)
        constexpr uint32_t ES = (uint32_t)sizeof(T);
        const uint32_t hb = smem_u32(H);
        uint32_t ca[CPL], ib[CPL];
        int jj[CPL], ii[CPL];
        // carries: (cL1, cL2) = H[k, j], H[k+1, j] for owned columns j >= k;  (dR1, dR2) = H[i, k], H[i, k+1], rows i < k
        R cL1[CPL], cL2[CPL], dR1[CPL], dR2[CPL];
        uint32_t ak = hb + ES * (uint32_t)(colbase(mx) - 1);
        {
            const uint32_t ak1 = ak + ES * (uint32_t)(mx + EX);
#pragma unroll
            for (int s = 0; s < CPL; ++s) {
                const int j = lane + 1 + 32 * s;
                const bool valid = j <= n;
                jj[s] = valid ? j : -(1 << 28);
                ii[s] = valid ? j : (1 << 28);
                ca[s] = hb + ES * (uint32_t)(colbase(valid ? j : 1) - 1);
                ib[s] = ES * (uint32_t)(valid ? j : 1);
                const bool c = jj[s] >= mx, r = ii[s] < mx;
                cL1[s] = c ? lds_e<T>(ca[s] + ES * mx) : zero;
                cL2[s] = c ? lds_e<T>(ca[s] + ES * (mx + 1)) : zero;
                dR1[s] = r ? lds_e<T>(ak + ib[s]) : zero;
                dR2[s] = r ? lds_e<T>(ak1 + ib[s]) : zero;
            }
        }
        unsigned napplied = 0;
        int k = mx;
        for (; k <= iend - 2; ++k) {   // three-row reflectors
            const uint32_t kb = ES * (uint32_t)k;
            const uint32_t ak1 = ak + ES * (uint32_t)(k + EX);          // column k+1
            const uint32_t ak2 = ak1 + ES * (uint32_t)(k + 1 + EX);     // column k+2
            const uint32_t akm = ak - ES * (uint32_t)(k - 1 + EX);      // column k-1
            if (k > mx) {
                v0 = lds_e<T>(akm + kb);
                v1 = lds_e<T>(akm + kb + ES);
                v2 = lds_e<T>(akm + kb + 2 * ES);
            }
            const R tau1 = reflector_real_small(v0, v1, v2, 3);
            napplied += 1;
            const R tau2 = tau1 * v1, tau3 = tau1 * v2;
            push_r(ZOP_REFL3, k, tau1, v1, v2);
#pragma unroll
            for (int s = 0; s < CPL; ++s) {
                if (jj[s] >= k) {
                    const uint32_t a = ca[s] + kb;
                    const R b = lds_e<T>(a + 2 * ES);
                    const R ss = cL1[s] + v1 * cL2[s] + v2 * b;
                    sts_e<T>(a, cL1[s] - ss * tau1);
                    cL1[s] = cL2[s] - ss * tau2;
                    cL2[s] = b - ss * tau3;
                    if (jj[s] <= k + 2) {
                        sts_e<T>(a + ES, cL1[s]);
                        sts_e<T>(a + 2 * ES, cL2[s]);
                    }
                }
            }
            __syncwarp();
            const int jmax = (k + 3 < iend) ? k + 3 : iend;
#pragma unroll
            for (int s = 0; s < CPL; ++s) {
                if (ii[s] <= jmax) {
                    R d1 = dR1[s], d2 = dR2[s];
                    if (ii[s] >= k) {
                        d1 = lds_e<T>(ak + ib[s]);
                        d2 = lds_e<T>(ak1 + ib[s]);
                    }
                    const R e = lds_e<T>(ak2 + ib[s]);
                    const R ss = d1 + v1 * d2 + v2 * e;
                    sts_e<T>(ak + ib[s], d1 - ss * tau1);
                    dR1[s] = d2 - ss * tau2;
                    dR2[s] = e - ss * tau3;
                    if (ii[s] >= k + 1) {
                        sts_e<T>(ak1 + ib[s], dR1[s]);
                        sts_e<T>(ak2 + ib[s], dR2[s]);
                    }
                }
            }
            if (lane == 0) {
                if (k > mx) {
                    sts_e<T>(akm + kb, v0);
                    sts_e<T>(akm + kb + ES, zero);
                    sts_e<T>(akm + kb + 2 * ES, zero);
                } else if (mx > istart) {
                    sts_e<T>(akm + kb, lds_e<T>(akm + kb) * (one - tau1));
                }
            }
            __syncwarp();
            // columns k+1, k+2 were rewritten by the right update in rows k+1, k+2: refresh their carries
#pragma unroll
            for (int s = 0; s < CPL; ++s) {
                if (jj[s] == k + 1 || jj[s] == k + 2) {
                    cL1[s] = lds_e<T>(ca[s] + kb + ES);
                    cL2[s] = lds_e<T>(ca[s] + kb + 2 * ES);
                }
            }
            ak = ak1;
        }
        // ---- last step: two-row reflector at k = iend-1; everything is written back ----
        {
            v0 = HH(k, k - 1);
            v1 = HH(k + 1, k - 1);
            v2 = zero;
            const R tau1 = reflector_real_small(v0, v1, v2, 2);
            napplied += 1;
            const R tau2 = tau1 * v1;
            emit_r(ZOP_REFL2, k, tau1, v1, zero);
#pragma unroll
            for (int s = 0; s < CPL; ++s) {
                const int j = lane + 1 + 32 * s;
                if (j >= k && j <= n) {
                    const R ss = cL1[s] + v1 * cL2[s];
                    HH(k, j) = cL1[s] - ss * tau1;
                    HH(k + 1, j) = cL2[s] - ss * tau2;
                }
            }
            __syncwarp();
#pragma unroll
            for (int s = 0; s < CPL; ++s) {
                const int i = lane + 1 + 32 * s;
                if (i <= iend) {
                    R d1, d2;
                    if (i >= k) {
                        d1 = HH(i, k);
                        d2 = HH(i, k + 1);
                    } else {
                        d1 = dR1[s];
                        d2 = dR2[s];
                    }
                    const R ss = d1 + v1 * d2;
                    HH(i, k) = d1 - ss * tau1;
                    HH(i, k + 1) = d2 - ss * tau2;
                }
            }
            if (lane == 0) {
                HH(k, k - 1) = v0;
                HH(k + 1, k - 1) = zero;
            }
            __syncwarp();
        }
        stp[1] += napplied;
    }


    // ================================================================================================
    // real double shift, software-pipelined (Float64) — the same construction as sweep_complex_pipelined: the 3x3
    // diagonal block (plus the sub-diagonal entry below it) is replicated in registers, every other entry of rows
    // k..k+2 (columns >= k+3) and of columns k..k+2 (rows <= k-1) is a bulk item with two running entries per owned
    // index, and one loop iteration applies reflector k to the bulk items while it forms reflector k+1 from the
    // register block.  Arithmetic per entry as in src/GenericSchur.jl:906-925.
    // ================================================================================================
    GS_DEV void sweep_real_pipelined(const R& r1r, const R& r1i, const R& r2r, const R& r2i, int istart, int iend) {
        static_assert(sizeof(T) == 8, "Float64 only");
        const R zero = 0.0, one = 1.0;
        const R eps = rtraits<R>::eps();
        int mx = 0;
        for (int base = iend - 2; base >= istart + 1 && !mx; base -= 32) {
            int m = base - lane;
            bool hit = false;
            if (m >= istart + 1) {
                R a0, a1, a2;
                first_column_r(m, r1r, r1i, r2r, r2i, a0, a1, a2);
                hit = r_abs(HH(m, m - 1)) * (r_abs(a1) + r_abs(a2)) <=
                      eps * r_abs(a0) * (r_abs(HH(m - 1, m - 1)) + r_abs(HH(m, m)) + r_abs(HH(m + 1, m + 1)));
            }
            unsigned msk = __ballot_sync(0xffffffffu, hit);
            if (msk) mx = base - (__ffs(msk) - 1);
        }
        if (!mx) mx = istart;
        R v0, v1, v2;
        first_column_r(mx, r1r, r1i, r2r, r2i, v0, v1, v2);
        constexpr uint32_t ES = (uint32_t)sizeof(T);
        const uint32_t hb = smem_u32(H);
        uint32_t ca[CPL], ib[CPL];
        int jl[CPL];
        R c1[CPL], c2[CPL];   // running entries: left item H[k, j], H[k+1, j];  right item H[j, k], H[j, k+1]
        uint32_t ak = hb + ES * (uint32_t)(colbase(mx) - 1);   // column k, "row 0"
        {
            const uint32_t ak1 = ak + ES * (uint32_t)(mx + EX);
#pragma unroll
            for (int s = 0; s < CPL; ++s) {
                const int j = lane + 1 + 32 * s;
                const bool valid = j <= n;
                jl[s] = valid ? j : -(1 << 28);
                ca[s] = hb + ES * (uint32_t)(colbase(valid ? j : 1) - 1);
                ib[s] = ES * (uint32_t)(valid ? j : 1);
                c1[s] = zero;
                c2[s] = zero;
                if (jl[s] >= mx + 3) {
                    c1[s] = lds_e<T>(ca[s] + ES * mx);
                    c2[s] = lds_e<T>(ca[s] + ES * (mx + 1));
                } else if (j <= mx - 1) {
                    c1[s] = lds_e<T>(ak + ib[s]);
                    c2[s] = lds_e<T>(ak1 + ib[s]);
                }
            }
        }
        // register block: columns k and k+1 of rows k..k+2 are carried, column k+2 is picked up at the top of each step
        R b00 = HH(mx, mx), b10 = HH(mx + 1, mx), b20 = HH(mx + 2, mx);
        R b01 = HH(mx, mx + 1), b11 = HH(mx + 1, mx + 1), b21 = HH(mx + 2, mx + 1);
        R tau1 = reflector_real_small(v0, v1, v2, 3);
        R beta = v0;
        R tau2 = tau1 * v1, tau3 = tau1 * v2;
        R L10 = zero, L20 = zero, L11 = zero, L21 = zero, L12 = zero, L22 = zero, L01 = zero, L02 = zero;
        const int capz = (wantZ && !LOG) ? cap : 0x7fffffff;
        const uint32_t ring32 = LOG ? 0u : smem_u32(ring);
        if constexpr (LOG) lg.put_hdr(LOG_REFL3, mx, iend - 1 - mx, iend, zero, zero);
#ifdef GS_QR_PROFILE
        const long long tp2 = clock64();
#endif
        for (int k = mx;; ++k) {
            if constexpr (!LOG) {
                if (cnt == capz) {
                    publish(0);
                    begin_buffer();
                }
            }
            __syncwarp();
            const uint32_t kb = ES * (uint32_t)k;
            const uint32_t ak1 = ak + ES * (uint32_t)(k + EX);        // column k+1
            const uint32_t ak2 = ak1 + ES * (uint32_t)(k + 1 + EX);   // column k+2
            // ---- loads: column k+2 of the block (rows k..k+3); third entry of every bulk item ----
            const R b02 = lds_e<T>(ak2 + kb), b12 = lds_e<T>(ak2 + kb + ES), b22 = lds_e<T>(ak2 + kb + 2 * ES);
            R e3 = zero;
            if (k + 3 <= iend) e3 = lds_e<T>(ak2 + kb + 3 * ES);
            bool act[CPL], own3[CPL];
            uint32_t sa[CPL];
            R y[CPL];
#pragma unroll
            for (int s = 0; s < CPL; ++s) {
                const int j = lane + 1 + 32 * s;
                const bool isL = jl[s] >= k + 3;
                const bool isR = j <= k - 1;
                act[s] = isL || isR;
                own3[s] = jl[s] == k + 3;
                sa[s] = isR ? ak + ib[s] : ca[s] + kb;                  // inactive lanes: a harmless address
                const uint32_t ya = isR ? ak2 + ib[s] : sa[s] + 2 * ES;
                lds_f64_if(c1[s], sa[s], j == k - 1);                   // row k-1 left the register block
                lds_f64_if(c2[s], ak1 + ib[s], j == k - 1);
                y[s] = lds_e<T>(ya);
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
            // f10, f20, f30 = H[k+1..k+3, k] feed the next reflector: the serial chain.  tau1, tau2 = tau1 v1, tau3 = tau1 v2 come
            // early out of the previous reflector (from 1/beta), v1 and v2 last (from 1/(alpha - beta)); the expressions
            // are regrouped so that v enters once: with r_i = (1 - tau1) b_i0 - tau2 b_i1 - tau3 b_i2 (row i of the block
            // times the right reflector's first column) and sx = r_0 + v1 r_1 + v2 r_2:
            //   f10 = r_1 - tau2 sx,   f20 = r_2 - tau3 sx,   f30 = -tau3 e3.
            // Measured on B200: the Float64 kernels run 8-12 CTAs per SM and are closer to issue-bound than chain-bound — the
            // ten extra FMAs cost more than the shorter chain gains (32x32: +4 %, 64x64: +9 % stage-B time), so this is OFF.
#ifndef GS_QR_REGROUP_REAL_MAXCPL
#define GS_QR_REGROUP_REAL_MAXCPL 0
#endif
            R f10, f20, f30;
            if constexpr (CPL <= GS_QR_REGROUP_REAL_MAXCPL) {
                const R omt = one - tau1;
                const R r0 = fma(-tau3, b02, fma(-tau2, b01, omt * b00));
                const R r1 = fma(-tau3, b12, fma(-tau2, b11, omt * b10));
                const R r2 = fma(-tau3, b22, fma(-tau2, b21, omt * b20));
                const R sx = fma(v2, r2, fma(v1, r1, r0));
                f10 = fma(-tau2, sx, r1);
                f20 = fma(-tau3, sx, r2);
                f30 = -(tau3 * e3);
            } else {
                f10 = fma(-t1, tau1, a10);
                f20 = fma(-t2, tau1, a20);
                f30 = -t3 * tau1;
            }
            const R f11 = fma(-t1, tau2, a11), f12 = fma(-t1, tau3, a12);   // row k+1
            const R f21 = fma(-t2, tau2, a21), f22 = fma(-t2, tau3, a22);   // row k+2
            const R f31 = -t3 * tau2, f32 = fma(-t3, tau3, e3);             // row k+3
            const R t0 = fma(v2, a02, fma(v1, a01, a00));
            const R f00 = fma(-t0, tau1, a00), f01 = fma(-t0, tau2, a01), f02 = fma(-t0, tau3, a02);   // row k: final
            const bool last = (k == iend - 2);
            {
                const bool l0 = lane == 0;
                if constexpr (LOG) {
                    const bool okl = lg.reserve();
                    stg_2f64_if(lg.slot(), tau1, v1, okl && l0);
                    stg_2f64_if(lg.slot() + 16, v2, zero, okl && l0);
                    if (okl) lg.advance();
                } else {
                    const uint32_t re = ring32 + (uint32_t)sizeof(ZOp) * (uint32_t)(rbase + cnt);
                    sts_2i_if(re, (int)ZOP_REFL3, k, l0 && wantZ);
                    sts_f64_if(re + 8, tau1, l0 && wantZ);
                    sts_f64_if(re + 16, v1, l0 && wantZ);
                    sts_f64_if(re + 24, v2, l0 && wantZ);
                }
                const uint32_t akm = ak - ES * (uint32_t)(k - 1 + EX);
                const bool sub = l0 && (k > mx);
                sts_f64_if(akm + kb, beta, sub);
                sts_f64_if(akm + kb + ES, zero, sub);
                sts_f64_if(akm + kb + 2 * ES, zero, sub);
                if (l0 && k == mx && mx > istart) sts_e<T>(akm + kb, lds_e<T>(akm + kb) * (one - tau1));
                sts_f64_if(ak + kb, f00, l0);
                sts_f64_if(ak1 + kb, f01, l0);
                sts_f64_if(ak2 + kb, f02, l0);
            }
            cnt += 1;
            // ---- reflector k+1 from (f10, f20, f30), straight-line (general routine only for out-of-range input) ----
            R t1n, v1n, v2n, betan, tau2n, tau3n;
            bool ok;
            {
                const double q = fma(f10, f10, fma(f20, f20, f30 * f30));
                const unsigned tz = ((unsigned)(__double2hiint(f20) | __double2hiint(f30)) << 1) |
                                    (unsigned)(__double2loint(f20) | __double2loint(f30));
                ok = q_exp_in(q, 1023u - 900u, 1023u + 900u) && (tz != 0u);
                double yr0;
                asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(yr0) : "d"(q));
                // seed of 1/(alpha - beta) from the unrefined norm: the second MUFU overlaps the refinement of the first
                double y0;
                asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(f10 + copysign(q * yr0, f10)));
                const double qy = q * yr0;
                const double e = fma(-qy, yr0, 1.0);
                const double cf = fma(e, 0.375, 0.5);
                const double yr = fma(yr0 * e, cf, yr0);   // 1/sqrt(q), one cubic step
                betan = -copysign(q * yr, f10);
                const double rb = -copysign(yr, f10);      // 1/beta
                t1n = (betan - f10) * rb;
                tau2n = -f20 * rb;                         // tau v1 = -x1 / beta,  tau v2 = -x2 / beta
                tau3n = -f30 * rb;
                const double amb = f10 - betan;            // |amb| >= |beta|: no cancellation
                const double e2 = fma(-amb, y0, 1.0);
                const double tt = fma(y0, fma(e2, e2, e2), y0);
                v1n = f20 * tt;
                v2n = f30 * tt;
            }
            // ---- bulk items ----
#pragma unroll
            for (int s = 0; s < CPL; ++s) {
                const R ss = fma(v2, y[s], fma(v1, c2[s], c1[s]));
                sts_f64_if(sa[s], fma(-ss, tau1, c1[s]), act[s]);
                const R n1 = fma(-ss, tau2, c2[s]), n2 = fma(-ss, tau3, y[s]);
                c1[s] = act[s] ? n1 : c1[s];
                c2[s] = act[s] ? n2 : c2[s];
                sts_f64_if(sa[s] + ES, n1, own3[s]);        // column k+3 enters the register block next step
                sts_f64_if(sa[s] + 2 * ES, n2, own3[s]);
            }
            if (last) {
                L10 = f10; L20 = f20; L11 = f11; L21 = f21; L12 = f12; L22 = f22; L01 = f01; L02 = f02;
                break;
            }
            if (!ok) {
                R w0 = f10, w1 = f20, w2 = f30;
                t1n = reflector_real_small(w0, w1, w2, 3);
                betan = w0;
                v1n = w1;
                v2n = w2;
                tau2n = t1n * v1n;
                tau3n = t1n * v2n;
            }
            tau1 = t1n;
            v1 = v1n;
            v2 = v2n;
            tau2 = tau2n;
            tau3 = tau3n;
            beta = betan;
            b00 = f11;
            b10 = f21;
            b20 = f31;
            b01 = f12;
            b11 = f22;
            b21 = f32;
            ak = ak1;
        }
#ifdef GS_QR_PROFILE
        prof[3] += clock64() - tp2;
#endif
        // ---- last step: the two-row reflector at k = iend-1 (src/GenericSchur.jl:927-946), entirely on registers:
        //      the lanes already hold rows iend-1, iend of their columns (left items) / columns iend-1, iend of their
        //      rows (right items), the block and row iend-2 are replicated; everything is written back afterwards ----
        {
            const int k = iend - 1;
            R w0 = L10, w1 = L20, w2 = zero;
            const R t1 = reflector_real_small(w0, w1, w2, 2);
            const R t2 = t1 * w1;
            emit_r(ZOP_REFL2, k, t1, w1, zero);
            // left on the block's columns
            const R sa0 = L11 + w1 * L21, sa1 = L12 + w1 * L22;
            const R g11 = L11 - sa0 * t1, g21 = L21 - sa0 * t2, g12 = L12 - sa1 * t1, g22 = L22 - sa1 * t2;
            // right on rows iend-2, iend-1, iend
            const R sr0 = L01 + w1 * L02, sr1 = g11 + w1 * g12, sr2 = g21 + w1 * g22;
            __syncwarp();
#pragma unroll
            for (int s = 0; s < CPL; ++s) {
                const int j = lane + 1 + 32 * s;
                const R ss = c1[s] + w1 * c2[s];
                const R n1 = c1[s] - ss * t1, n2 = c2[s] - ss * t2;
                if (j >= iend + 1 && j <= n) {
                    HH(iend - 1, j) = n1;
                    HH(iend, j) = n2;
                }
                if (j <= iend - 3) {
                    HH(j, iend - 1) = n1;
                    HH(j, iend) = n2;
                }
            }
            if (lane == 0) {
                HH(k, k - 1) = w0;
                HH(k + 1, k - 1) = zero;
                HH(k - 1, k) = L01 - sr0 * t1;
                HH(k - 1, k + 1) = L02 - sr0 * t2;
                HH(k, k) = g11 - sr1 * t1;
                HH(k, k + 1) = g12 - sr1 * t2;
                HH(k + 1, k) = g21 - sr2 * t1;
                HH(k + 1, k + 1) = g22 - sr2 * t2;
            }
            __syncwarp();
        }
        stp[1] += (unsigned)(iend - mx);
    }

    GS_DEV int qr_real(int maxiter, unsigned* st) {
        stp = st;
#ifdef GS_QR_PROFILE
        prof[0] = prof[1] = prof[2] = prof[3] = 0;
        const long long tq0 = clock64();
#endif
        const R zero = r_const<R>(0.0);
        const R eps = rtraits<R>::eps();
        const R smallnum = rtraits<R>::floatmin() * (r_const<R>((double)n) / eps);
        const R threeq = r_const<R>(0.75), m7_16 = r_const<R>(-0.4375);
        int istart = 1, iend = n, iwcur = n, iter = 0;
        sidx = 0;
        cnt = 0;
        begin_buffer();
        while (iend >= 1) {
            istart = 1;
            int iterqr = 0;
            while (true) {
                iter += 1;
                if (iter > maxiter) {
                    st[3] = iter;
                    finish_ring();
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
                __syncwarp();
                if (istart > 1 && lane == 0) HH(istart, istart - 1) = zero;
                __syncwarp();
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
                    const R rs = q_rcp(s);
                    H11 = H11 * rs;
                    H12 = H12 * rs;
                    H21 = H21 * rs;
                    H22 = H22 * rs;
                    R tr = (H11 + H22) * r_const<R>(0.5);
                    R d = (H11 - tr) * (H22 - tr) - H12 * H21;
                    R rtd = q_sqrt(r_abs(d));
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
                if constexpr (sizeof(T) == 8) sweep_real_pipelined(r1r, r1i, r2r, r2i, istart, iend);
                else sweep_real(r1r, r1i, r2r, r2i, istart, iend);
                if constexpr (LOG) {
                    if (lg.ovf) return LOG_OVERFLOW_RC;
                }
                publish(0);
                begin_buffer();
            }
            if (istart >= iend) {
                if (lane == 0) sW[iwcur - 1] = mk_cx<R>(HH(iend, iend), zero);
                iwcur -= 1;
            } else if (istart + 1 == iend) {
                R a = HH(iend - 1, iend - 1), b = HH(iend - 1, iend), c = HH(iend, iend - 1), d = HH(iend, iend);
                R cs, sn;
                C w1, w2;
                gs2x2(a, b, c, d, cs, sn, w1, w2);
                if (lane == 0) {
                    sW[iwcur - 1] = w2;
                    sW[iwcur - 2] = w1;
                }
                iwcur -= 2;
                emit_r(ZOP_GIVENS, iend - 1, cs, sn, zero);
                __syncwarp();
                for (int j = istart + lane; j <= n; j += 32) {
                    R a1 = HH(iend - 1, j), a2 = HH(iend, j);
                    HH(iend - 1, j) = cs * a1 + sn * a2;
                    HH(iend, j) = -sn * a1 + cs * a2;
                }
                __syncwarp();
                for (int r = 1 + lane; r <= iend; r += 32) {
                    R a1 = HH(r, iend - 1), a2 = HH(r, iend);
                    HH(r, iend - 1) = a1 * cs + a2 * sn;
                    HH(r, iend) = -a1 * sn + a2 * cs;
                }
                __syncwarp();
                if (lane == 0) {
                    HH(iend - 1, iend - 1) = a;
                    HH(iend - 1, iend) = b;
                    HH(iend, iend - 1) = c;
                    HH(iend, iend) = d;
                    if (iend > 2) HH(iend - 1, iend - 2) = zero;
                }
                __syncwarp();
            }
            iend = istart - 1;
        }
        st[3] = iter;
#ifdef GS_QR_PROFILE
        prof[0] = clock64() - tq0;
        st[0] = (unsigned)(prof[0] >> 6);
        st[2] = (unsigned)(prof[1] >> 6);
        st[3] = (unsigned)(prof[3] >> 6);
#endif
        finish_ring();
        return 0;
    }

    // ================================================================================================
    // consumer: the Z-warp.  Z lives in global memory (L2-resident while the matrix is being worked on); lane l owns
    // rows l+1 (+32 s).  Runs of consecutive reflectors are streamed with carry registers and a one-column prefetch.
    // ================================================================================================
    GS_DEV void z_consumer() {
        const R zero = r_const<R>(0.0);
        for (int cidx = 0;; ++cidx) {
            const int b = cidx % RB;
            named_bar_sync<RB, RING_THREADS>(BAR_FULL0 + b);
            const int count = hdr->count[b];
            const int end = hdr->end[b];
            const ZOp* ops = ring + b * cap;
            int i = 0;
            while (i < count) {
                const int op = ops[i].op;
                if constexpr (CPLX && sizeof(R) == 8 && GS_ZRUN > 0) {
                    if (op == ZOP_REFL) {
                        // A run of up to ZRUN reflectors on consecutive columns k, k+1, ...: all ZRUN+1 columns are
                        // fetched from L2 up front (one latency per run instead of one per reflector), the reflectors
                        // are applied from registers, then the columns go back.
                        constexpr int ZRUN = GS_ZRUN > 0 ? GS_ZRUN : 1;
                        const int k = ops[i].k;
                        int m = 1;
                        while (m < ZRUN && i + m < count && ops[i + m].op == ZOP_REFL && ops[i + m].k == k + m) m += 1;
                        C z[ZRUN + 1][CPL];
#pragma unroll
                        for (int t = 0; t <= ZRUN; ++t) {
                            if (t <= m) {
#pragma unroll
                                for (int s = 0; s < CPL; ++s) {
                                    const int r = lane + 1 + 32 * s;
                                    if (r <= n) z[t][s] = ZZ(r, k + t);
                                }
                            }
                        }
#pragma unroll
                        for (int t = 0; t < ZRUN; ++t) {
                            if (t < m) {
                                const C tau1 = mk_cx<R>(ops[i + t].a[0], ops[i + t].a[1]);
                                const C v2 = mk_cx<R>(ops[i + t].a[2], ops[i + t].a[3]);
                                const C v2c = cconj(v2);
                                const R tau2 = tau1.re * v2.re - tau1.im * v2.im;
#pragma unroll
                                for (int s = 0; s < CPL; ++s) {
                                    const C ss = e_axty<(CPL >= 2)>(tau1, z[t][s], tau2, z[t + 1][s]);
                                    z[t][s] = z[t][s] - ss;
                                    z[t + 1][s] = e_fnma(ss, v2c, z[t + 1][s]);
                                }
                            }
                        }
#pragma unroll
                        for (int t = 0; t <= ZRUN; ++t) {
                            if (t <= m) {
#pragma unroll
                                for (int s = 0; s < CPL; ++s) {
                                    const int r = lane + 1 + 32 * s;
                                    if (r <= n) ZZ(r, k + t) = z[t][s];
                                }
                            }
                        }
                        i += m;
                    } else {   // ZOP_SCALE
                        const C t = mk_cx<R>(ops[i].a[0], ops[i].a[1]);
                        for (int j = ops[i].k; j <= ops[i].k2; ++j) {
#pragma unroll
                            for (int s = 0; s < CPL; ++s) {
                                const int r = lane + 1 + 32 * s;
                                if (r <= n) ZZ(r, j) = ZZ(r, j) * t;
                            }
                        }
                        i += 1;
                    }
                } else if constexpr (CPLX && NH > 1) {
                    // one row per lane: this Z-warp's slot (the other rows belong to the other Z-warps)
                    const int r = lane + 1 + 32 * zslot;
                    const bool rv = r <= n;
                    if (op == ZOP_REFL) {
                        int k = ops[i].k;
                        C z0 = rv ? ZZ(r, k) : mk_cx<R>(zero, zero);
                        C zn = rv ? ZZ(r, k + 1) : mk_cx<R>(zero, zero);
                        while (i < count && ops[i].op == ZOP_REFL && ops[i].k == k) {
                            const C tau1 = mk_cx<R>(ops[i].a[0], ops[i].a[1]);
                            const C v2 = mk_cx<R>(ops[i].a[2], ops[i].a[3]);
                            const C v2c = cconj(v2);
                            const R tau2 = (tau1 * v2).re;
                            const C zp = (rv && k + 2 <= n) ? ZZ(r, k + 2) : mk_cx<R>(zero, zero);   // prefetch
                            if (rv) {
                                const C z1 = zn;
                                const C ss = e_axty<true>(tau1, z0, tau2, z1);
                                ZZ(r, k) = z0 - ss;
                                z0 = e_bsv<true>(z1, ss, v2c);
                                zn = zp;
                            }
                            k += 1;
                            i += 1;
                        }
                        if (rv) ZZ(r, k) = z0;
                    } else {   // ZOP_SCALE
                        const C t = mk_cx<R>(ops[i].a[0], ops[i].a[1]);
                        if (rv)
                            for (int j = ops[i].k; j <= ops[i].k2; ++j) ZZ(r, j) = ZZ(r, j) * t;
                        i += 1;
                    }
                } else if constexpr (CPLX) {
                    if (op == ZOP_REFL) {
                        int k = ops[i].k;
                        C z0[CPL], zn[CPL];
#pragma unroll
                        for (int s = 0; s < CPL; ++s) {
                            const int r = lane + 1 + 32 * s;
                            z0[s] = (r <= n) ? ZZ(r, k) : mk_cx<R>(zero, zero);
                            zn[s] = (r <= n) ? ZZ(r, k + 1) : mk_cx<R>(zero, zero);
                        }
                        while (i < count && ops[i].op == ZOP_REFL && ops[i].k == k) {
                            const C tau1 = mk_cx<R>(ops[i].a[0], ops[i].a[1]);
                            const C v2 = mk_cx<R>(ops[i].a[2], ops[i].a[3]);
                            const C v2c = cconj(v2);
                            const R tau2 = (tau1 * v2).re;
                            C zp[CPL];   // prefetch of column k+2 for the next reflector of the run
#pragma unroll
                            for (int s = 0; s < CPL; ++s) {
                                const int r = lane + 1 + 32 * s;
                                zp[s] = (r <= n && k + 2 <= n) ? ZZ(r, k + 2) : mk_cx<R>(zero, zero);
                            }
#pragma unroll
                            for (int s = 0; s < CPL; ++s) {
                                const int r = lane + 1 + 32 * s;
                                if (r <= n) {
                                    const C z1 = zn[s];
                                    const C ss = e_axty<(CPL >= 2)>(tau1, z0[s], tau2, z1);
                                    ZZ(r, k) = z0[s] - ss;
                                    z0[s] = e_bsv<(CPL >= 2)>(z1, ss, v2c);
                                    zn[s] = zp[s];
                                }
                            }
                            k += 1;
                            i += 1;
                        }
#pragma unroll
                        for (int s = 0; s < CPL; ++s) {
                            const int r = lane + 1 + 32 * s;
                            if (r <= n) ZZ(r, k) = z0[s];
                        }
                    } else {   // ZOP_SCALE
                        const C t = mk_cx<R>(ops[i].a[0], ops[i].a[1]);
                        for (int j = ops[i].k; j <= ops[i].k2; ++j) {
#pragma unroll
                            for (int s = 0; s < CPL; ++s) {
                                const int r = lane + 1 + 32 * s;
                                if (r <= n) ZZ(r, j) = ZZ(r, j) * t;
                            }
                        }
                        i += 1;
                    }
                } else {
                    if (op == ZOP_REFL3) {
                        int k = ops[i].k;
                        R z1[CPL], z2[CPL], zn[CPL];
#pragma unroll
                        for (int s = 0; s < CPL; ++s) {
                            const int r = lane + 1 + 32 * s;
                            z1[s] = (r <= n) ? ZZ(r, k) : zero;
                            z2[s] = (r <= n) ? ZZ(r, k + 1) : zero;
                            zn[s] = (r <= n) ? ZZ(r, k + 2) : zero;
                        }
                        while (i < count && ops[i].op == ZOP_REFL3 && ops[i].k == k) {
                            const R tau1 = ops[i].a[0], v2 = ops[i].a[1], v3 = ops[i].a[2];
                            const R tau2 = tau1 * v2, tau3 = tau1 * v3;
                            R zp[CPL];
#pragma unroll
                            for (int s = 0; s < CPL; ++s) {
                                const int r = lane + 1 + 32 * s;
                                zp[s] = (r <= n && k + 3 <= n) ? ZZ(r, k + 3) : zero;
                            }
#pragma unroll
                            for (int s = 0; s < CPL; ++s) {
                                const int r = lane + 1 + 32 * s;
                                if (r <= n) {
                                    const R z3 = zn[s];
                                    const R ss = z1[s] + v2 * z2[s] + v3 * z3;
                                    ZZ(r, k) = z1[s] - ss * tau1;
                                    z1[s] = z2[s] - ss * tau2;
                                    z2[s] = z3 - ss * tau3;
                                    zn[s] = zp[s];
                                }
                            }
                            k += 1;
                            i += 1;
                        }
#pragma unroll
                        for (int s = 0; s < CPL; ++s) {
                            const int r = lane + 1 + 32 * s;
                            if (r <= n) {
                                ZZ(r, k) = z1[s];
                                ZZ(r, k + 1) = z2[s];
                            }
                        }
                    } else if (op == ZOP_REFL2) {
                        const int k = ops[i].k;
                        const R tau1 = ops[i].a[0], v2 = ops[i].a[1];
                        const R tau2 = tau1 * v2;
#pragma unroll
                        for (int s = 0; s < CPL; ++s) {
                            const int r = lane + 1 + 32 * s;
                            if (r <= n) {
                                const R a = ZZ(r, k), b2 = ZZ(r, k + 1);
                                const R ss = a + v2 * b2;
                                ZZ(r, k) = a - ss * tau1;
                                ZZ(r, k + 1) = b2 - ss * tau2;
                            }
                        }
                        i += 1;
                    } else {   // ZOP_GIVENS
                        const int k = ops[i].k;
                        const R cs = ops[i].a[0], sn = ops[i].a[1];
#pragma unroll
                        for (int s = 0; s < CPL; ++s) {
                            const int r = lane + 1 + 32 * s;
                            if (r <= n) {
                                const R a1 = ZZ(r, k), a2 = ZZ(r, k + 1);
                                ZZ(r, k) = a1 * cs + a2 * sn;
                                ZZ(r, k + 1) = -a1 * sn + a2 * cs;
                            }
                        }
                        i += 1;
                    }
                }
            }
            named_bar_arrive<RB, RING_THREADS>(BAR_EMPTY0 + b);
            if (end) break;
        }
    }
#undef HH
#undef ZZ
};

// =================================================================================================
// Stage B kernel: QR iteration on an upper Hessenberg matrix.  One matrix per 64-thread CTA (H-warp + Z-warp).
//   in:  A_b = H_b (upper Hessenberg), Z_b = Q_b (or the caller's Z with GSCHUR_FLAG_HESS_INPUT), scratch (scale info)
//   out: A_b <- T_b, Z_b <- Z_b * (accumulated reflectors), w, info, stats
// =================================================================================================
// The Z-warp's code lives in its own (non-inlined) function: its register allocation — the run of ZRUN+1 columns held
// in registers — must not leak into the H-warp's step loop, which is compiled at the same per-thread register cap.
template <class T, int CPL>
__device__ __noinline__ void z_consumer_entry(int n, int lane, int ldz, T* Z, typename FastSolver<T, CPL>::ZOp* ring,
                                              zring_hdr* hdr, int cap, int zslot = 0) {
    FastSolver<T, CPL> F;
    F.zslot = zslot;
    F.n = n;
    F.lane = lane;
    F.ldz = ldz;
    F.Z = Z;
    F.ring = ring;
    F.hdr = hdr;
    F.cap = cap;
    F.wantZ = true;
    F.z_consumer();
}

template <class T, int CPL> struct fast_smem_layout {
    typedef typename etraits<T>::real R;
    typedef smem_layout<T> L;
    typedef FastSolver<T, CPL> FS;
    typedef zop_t<etraits<T>::is_complex, typename etraits<T>::real> ZOp;
    // A short ring (16 reflectors per buffer; two buffers, three for ComplexF64): the H-warp publishes every 16 bulge
    // steps.  Together with dropping the eigenvalue array for complex kinds (their eigenvalues are diag(T)) a 64x64
    // ComplexF64 matrix needs 37.6 KB of shared memory: six CTAs per SM (the limit is 37.8 KB).
    __host__ __device__ static int cap(int) { return 16; }
    __host__ __device__ static size_t off_w(int n) { return L::up16((size_t)FS::packed_elems(n) * sizeof(T)); }
    __host__ __device__ static size_t off_ring(int n) {
        return off_w(n) + (etraits<T>::is_complex ? 0 : L::up16((size_t)n * 2 * sizeof(R)));
    }
    __host__ __device__ static size_t off_hdr(int n) {
        return off_ring(n) + L::up16((size_t)ring_bufs<T, CPL>::value * (size_t)cap(n) * sizeof(ZOp));
    }
    __host__ __device__ static size_t off_cmd(int n) { return off_hdr(n) + L::up16(sizeof(zring_hdr)); }
    __host__ __device__ static size_t bytes(int n) { return off_cmd(n) + (FS::NH > 1 ? L::up16(sizeof(typename FS::HCmd)) : 0); }
};

// register budget: small real tiles are occupancy-bound (cap at 64 registers -> 16 CTAs/SM), the others shared-memory-bound
template <class T, int CPL> struct qr_min_blocks {
    static constexpr int value = (!etraits<T>::is_complex && CPL == 1 && sizeof(T) == 8) ? GS_QR_MINB_F64_1
                                 : (!etraits<T>::is_complex && CPL == 2 && sizeof(T) == 8) ? GS_QR_MINB_F64_2
                                 : (etraits<T>::is_complex && sizeof(T) == 16 && CPL == 2) ? 6   // 64x64 ComplexF64: six CTAs/SM fit in shared memory
                                 : (etraits<T>::is_complex && sizeof(T) == 16 && CPL == 1) ? GS_QR_MINB_C64_1   // n <= 32: register-bound
                                 : 1;
};

// threads per CTA: the H-warp and the Z-warp; the complex double-double kernel at CPL = 3 has three of each
template <class T, int CPL> struct qr_threads {
    static constexpr int value = 32 * (FastSolver<T, CPL>::NH + FastSolver<T, CPL>::NZW);
};

template <class T, int CPL>
__global__ void __launch_bounds__((qr_threads<T, CPL>::value), (qr_min_blocks<T, CPL>::value)) gschur_qr_kernel(BatchedParams p) {
    typedef typename etraits<T>::real R;
    typedef cx<R> C;
    constexpr bool CPLX = etraits<T>::is_complex;
    typedef FastSolver<T, CPL> FS;
    constexpr int NT = qr_threads<T, CPL>::value;
    constexpr int NW = NT / 32;
    typedef fast_smem_layout<T, CPL> FL;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = p.n;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    __shared__ long long s_next;
    __shared__ int s_info;
    __shared__ unsigned s_stats[4];
    __shared__ unsigned s_w0;
    const bool wantZ = (p.Z != nullptr);
    // Role assignment.  Warp schedulers are picked by the hardware warp slot (slot mod 4) and a 64-thread CTA takes two
    // consecutive slots, so with a fixed "warp 0 = H-warp" every H-warp of the SM would sit on two of the four
    // schedulers (measured: the FP64 pipes of those two saturate while the other two idle).  The H role therefore
    // alternates between the CTA's two warps with bit 2 of the slot number: H-warps land on all four schedulers.
    int warp = tid >> 5;
    if (!(p.flags & F_FIXED_ROLES)) {
        if (tid == 0) {
            unsigned w;
            asm volatile("mov.u32 %0, %%warpid;" : "=r"(w));
            s_w0 = w;
        }
        __syncthreads();
        warp ^= (int)((s_w0 >> 2) & 1u);   // role index: 0 = H-warp, 1 = Z-warp (2, 3: helper H-warps; 4, 5: more Z-warps: interchangeable)
    }

    FS F;
    F.n = n;
    F.lane = lane;
    F.cap = FL::cap(n);
    F.H = reinterpret_cast<T*>(smem_raw);
    F.sW = reinterpret_cast<C*>(smem_raw + FL::off_w(n));
    F.ring = reinterpret_cast<typename FS::ZOp*>(smem_raw + FL::off_ring(n));
    F.hdr = reinterpret_cast<zring_hdr*>(smem_raw + FL::off_hdr(n));
    F.hcmd = reinterpret_cast<typename FS::HCmd*>(smem_raw + FL::off_cmd(n));
    F.wantZ = wantZ;
    F.ldz = p.ldz;
    T* H = F.H;
    const R zero = r_const<R>(0.0);

    for (;;) {
        if (tid == 0) s_next = (long long)atomicAdd(p.counter, 1ULL);
        __syncthreads();
        long long b = s_next;
        __syncthreads();
        if (p.list) {   // redo pass of the three-stage path: work item i is matrix list[i]
            if (b >= (long long)*p.list_count) break;
            b = p.list[b];
        } else if (b >= p.batch) {
            break;
        }
        T* gA = reinterpret_cast<T*>(p.A) + b * p.strideA;
        F.Z = wantZ ? reinterpret_cast<T*>(p.Z) + b * p.strideZ : nullptr;

        // ---- load the Hessenberg part into packed storage; bulge slots start at zero ----
        int bad = 0;
        for (int j = 1 + warp; j <= n; j += NW) {
            const int cb = FS::colbase(j);
            for (int i = 1 + lane; i <= j + FS::EX; i += 32) {
                T v = e_zero<T>();
                if (i <= j + 1 && i <= n) v = gA[(i - 1) + (size_t)(j - 1) * p.lda];
                H[cb + i - 1] = v;
                if constexpr (CPLX) {
                    if (i == j + 1 && (p.flags & F_CHECK_SUBDIAG) && v.im != zero) bad = 1;
                }
            }
        }
        bad = __syncthreads_or(bad);
        int info = 0;
        if (bad) info = -4;
        if (info == 0) {
            const int maxiter = p.maxiter > 0 ? p.maxiter : 100 * n;
            if (warp == 0) {
                unsigned st[4] = {0u, 0u, 0u, 0u};
                int rc;
                if constexpr (CPLX) rc = F.qr_complex(maxiter, st);
                else rc = F.qr_real(maxiter, st);
                F.helpers_exit();
                if (lane == 0) {
                    s_info = rc;
                    s_stats[0] = st[0];
                    s_stats[1] = st[1];
                    s_stats[2] = st[2];
                    s_stats[3] = st[3];
                }
            } else if (warp == 1) {
                if (wantZ) z_consumer_entry<T, CPL>(n, lane, F.ldz, F.Z, F.ring, F.hdr, F.cap, 0);
            } else {
                if constexpr (FS::NH > 1) {
                    // roles 2 .. NH: helper H-warps (slots 1 .. NH-1); roles NH+1 .. : Z-warps (slots 1 .. NH-1)
                    if (warp <= FS::NH) F.helper_loop(warp - 1);
                    else if (wantZ) z_consumer_entry<T, CPL>(n, lane, F.ldz, F.Z, F.ring, F.hdr, F.cap, warp - FS::NH);
                }
            }
            __syncthreads();
            info = s_info;
        } else if (tid == 0) {
            s_stats[0] = s_stats[1] = s_stats[2] = s_stats[3] = 0u;
        }
        __syncthreads();

        // ---- unscale (src/GenericSchur.jl:367-370, 830-833) with the factors stage A recorded ----
        bool scaled = false;
        R cscale = r_const<R>(1.0), anrm = r_const<R>(1.0);
        if (p.scratch) {
            const double* sc = p.scratch + 8 * b;
            scaled = sc[0] != 0.0;
            if constexpr (rtraits<R>::ndoubles == 1) {
                cscale = sc[1];
                anrm = sc[3];
            } else {
                cscale = mk_dd(sc[1], sc[2]);
                anrm = mk_dd(sc[3], sc[4]);
            }
        }
        if (scaled) {
            const int total = FS::packed_elems(n);
            safescale_apply<T, R, NT>(cscale, anrm, [&](R mul) {
                for (int e = tid; e < total; e += NT) H[e] = e_scale(H[e], mul);
                if (!CPLX)
                    for (int e = tid; e < n; e += NT) F.sW[e] = mk_cx<R>(F.sW[e].re * mul, F.sW[e].im * mul);
            });
            __syncthreads();
        }
        // ---- store T (exact zeros below the (quasi-)triangle), w, info, stats ----
        for (int j = 1 + warp; j <= n; j += NW) {
            const int cb = FS::colbase(j);
            for (int i = 1 + lane; i <= n; i += 32) {
                const bool keep = CPLX ? (i <= j) : (i <= j + 1);
                gA[(i - 1) + (size_t)(j - 1) * p.lda] = keep ? H[cb + i - 1] : e_zero<T>();
            }
        }
        C* gw = reinterpret_cast<C*>(p.w) + b * (long long)n;
        for (int e = tid; e < n; e += NT) {
            if constexpr (CPLX) gw[e] = H[FS::colbase(e + 1) + e];
            else gw[e] = F.sW[e];
        }
        if (tid == 0) {
            if (p.info) p.info[b] = info;
            if (p.stats) {
                p.stats[4 * b + 0] = s_stats[0];
                p.stats[4 * b + 1] = s_stats[1];
                p.stats[4 * b + 2] = s_stats[2];
                p.stats[4 * b + 3] = s_stats[3];
            }
        }
        __syncthreads();
    }
}

// Two launches on the caller's stream: stage A (unless the input is already Hessenberg), then stage B.
template <class T, int CPL> int launch_fast(const BatchedParams& p_in, int dev_sms, cudaStream_t stream, std::string* err) {
    BatchedParams p = p_in;
    cudaError_t e;
    double* scratch = nullptr;
    unsigned long long* counter2 = nullptr;
    const bool hess_input = (p.flags & F_HESS_INPUT) != 0;
    if (!hess_input) {
        e = cudaMallocAsync((void**)&scratch, (size_t)p.batch * 8 * sizeof(double) + 16, stream);
        if (e != cudaSuccess) {
            *err = std::string("cudaMallocAsync: ") + cudaGetErrorString(e);
            return -2;
        }
        counter2 = reinterpret_cast<unsigned long long*>(scratch + (size_t)p.batch * 8);
        e = cudaMemsetAsync(counter2, 0, sizeof(unsigned long long), stream);
        p.scratch = scratch;
        stage_timing_mark(0, stream);
        int rc = launch_stage_a<T>(p, dev_sms, stream, err);
        if (rc) {
            cudaFreeAsync(scratch, stream);
            return rc;
        }
        p.counter = counter2;   // stage B gets its own work-queue head
    }
    stage_timing_mark(1, stream);
    auto kern = gschur_qr_kernel<T, CPL>;
    size_t smem = fast_smem_layout<T, CPL>::bytes(p.n);
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int per_sm = 0;
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, qr_threads<T, CPL>::value, smem);
    int rc = 0;
    if (e != cudaSuccess) {
        *err = std::string("qr kernel setup: ") + cudaGetErrorString(e);
        rc = -2;
    } else if (per_sm < 1) {
        *err = "qr kernel does not fit on an SM";
        rc = -3;
    } else {
        static const int cap_per_sm = std::getenv("GSCHUR_QR_CTAS_PER_SM") ? std::atoi(std::getenv("GSCHUR_QR_CTAS_PER_SM")) : 0;
        if (cap_per_sm > 0 && cap_per_sm < per_sm) per_sm = cap_per_sm;   // profiling knob: occupancy sweep
        long long grid = (long long)per_sm * dev_sms;
        if (grid > p.batch) grid = p.batch;
        if (std::getenv("GSCHUR_QR_FIXED_ROLES")) p.flags |= F_FIXED_ROLES;   // profiling knob
        kern<<<(unsigned)grid, qr_threads<T, CPL>::value, smem, stream>>>(p);
        note_launch();
        stage_timing_mark(2, stream);
        e = cudaGetLastError();
        if (e != cudaSuccess) {
            *err = std::string("qr kernel launch: ") + cudaGetErrorString(e);
            rc = -2;
        }
    }
    if (scratch) cudaFreeAsync(scratch, stream);
    return rc;
}

}  // namespace gs

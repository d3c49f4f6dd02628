// Stage C of the three-stage path, second generation: replay of the reflector log on Z with the rows of Z in REGISTERS.
//
// One thread per row of Z (rows are independent: no serial chain between threads, no barrier on the data path).  The
// first generation (gschur_zreplay_kernel in qr3.cuh) kept Z in shared memory: 64 KB for a 64x64 ComplexF64 matrix,
// three CTAs of one or two warps per SM — every warp waiting on its own load -> FMA -> store round trip.  Here a thread
// keeps its row in registers: all of it for Float64 (64 doubles) and for n <= 32, the 32 trailing columns for 64x64
// ComplexF64 (the leading 32 columns stay in shared memory, 32 KB: sweeps start at the top but most of them reach
// the trailing columns, which see 3/4 of all column updates).  Register files do not take run-time indices, so a run of
// reflectors is applied by code that is fully unrolled over the column POSITION and entered block-wise (8 positions per
// uniform branch); a position outside the run is skipped by a warp-uniform predicate.
// Per reflector and row: 14 (complex) / 7 (real) FP64 operations and two broadcast loads of the record; no traffic
// for Z itself.  The arithmetic per entry is that of src/GenericSchur.jl:455-459 (complex), :920-925, :940-945, :687
// (real), in the same order and grouping as the first generation: Z is bit-identical.
#pragma once
#include "qrlog.cuh"

namespace gs {

// f(c) for every c in [c0, c1] ∩ [0, NP), c a compile-time constant inside f after unrolling
template <int NP, class F> GS_DEV void for_cols(int c0, int c1, F&& f) {
#pragma unroll
    for (int blk = 0; blk < (NP + 7) / 8; ++blk) {
        if (c0 < 8 * blk + 8 && c1 >= 8 * blk) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int c = 8 * blk + q;
                if (c < NP) {
                    if (c >= c0 && c <= c1) f(c);
                }
            }
        }
    }
}

// A run of reflectors at positions [j0, j1] (position j: record at shared address rb + 32 j): apply(j, addr) applies the
// record at addr with j a compile-time constant.  A block of BW positions is ONE basic block (the record loads are
// scheduled ahead of the arithmetic; only the dependency through the row entries is serial): a position outside the run
// reads the identity record at `ident` (tau = 0) instead of being branched around.  (A second, select-free copy of the
// block for the interior of a run was measured: fewer instructions, no faster — the kernel is bound by the dependent
// FP64 chain of a row and by instruction fetch of the unrolled code, not by issue slots.)
template <int NP, int BW, class A> GS_DEV void for_run(int j0, int j1, uint32_t rb, uint32_t ident, A&& apply) {
#pragma unroll
    for (int blk = 0; blk < (NP + BW - 1) / BW; ++blk) {
        if (j0 < BW * blk + BW && j1 >= BW * blk) {
#pragma unroll
            for (int q = 0; q < BW; ++q) {
                const int c = BW * blk + q;
                if (c < NP) apply(c, ((unsigned)(c - j0) <= (unsigned)(j1 - j0)) ? rb + 32u * (uint32_t)c : ident);
            }
        }
    }
}

GS_DEV cx<double> zr_lds2(uint32_t a) {
    cx<double> v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.re), "=d"(v.im) : "r"(a));
    return v;
}
GS_DEV double zr_lds1(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
// volatile at the PTX level: ptxas keeps these where they are written (it sinks plain loads down to their first use,
// across the position branches, which serialises load latency and arithmetic)
GS_DEV cx<double> zr_ldv2(uint32_t a) {
    cx<double> v;
    asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.re), "=d"(v.im) : "r"(a));
    return v;
}
GS_DEV double zr_ldv1(uint32_t a) {
    double v;
    asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
GS_DEV void zr_sts2(uint32_t a, const cx<double>& v) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v.re), "d"(v.im) : "memory");
}
GS_DEV void zr_cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

// ---- ComplexF64: columns [0, NS) in shared memory (layout [column][row], N rows per column), [NS, N) in registers ----
struct RecC {
    cx<double> tau1, v2;
};
struct RecR {
    double tau1, v2, v3;
};
template <int N, int NREG> struct ZRegC {
    typedef cx<double> T;
    static constexpr int NS = N - NREG;
    double zr[NREG], zi[NREG];
    T a;            // while k - 1 < NS: the running entry Z[r, k] of the shared-memory part
    int run, k;     // reflectors left in the current run; 1-based first column of the next one
    uint32_t zs;    // shared byte address of Z[r, 0]
    uint32_t ident; // shared byte address of an all-zero record
    static constexpr uint32_t CS = 16u * N;

    GS_DEV static void refl(const T& tau1, const T& v2, const T& a_, const T& b_, T& out_a, T& out_b) {
        const double tau2 = tau1.re * v2.re - tau1.im * v2.im;
        T ss;
        ss.re = fma(tau1.re, a_.re, fma(-tau1.im, a_.im, tau2 * b_.re));
        ss.im = fma(tau1.re, a_.im, fma(tau1.im, a_.re, tau2 * b_.im));
        out_a = a_ - ss;
        out_b = e_fnma_cjb(ss, v2, b_);
    }

    GS_DEV void init(uint32_t zs_, uint32_t ident_) {
        run = 0;
        k = 0;
        a = mk_cx<double>(0.0, 0.0);
        zs = zs_;
        ident = ident_;
    }

    // m reflectors, the first one on columns k, k+1 (1-based), records at shared address ra
    GS_DEV void segment(int m, uint32_t ra) {
        int t = 0;
        if constexpr (NS > 0) {
            int ms = NS - k;
            if (ms > m) ms = m;
            if (ms > 0) {
                uint32_t za = zs + CS * (uint32_t)k;   // column k+1 (1-based) = 0-based k
                T tau1 = zr_lds2(ra), v2 = zr_lds2(ra + 16), b = zr_lds2(za);
                for (; t < ms; ++t) {
                    // the next record / entry are fetched before this reflector's store (shared-memory stores order loads)
                    const bool more = t + 1 < ms;
                    const uint32_t rn = ra + 32u * (uint32_t)(more ? t + 1 : t);
                    const uint32_t zn = more ? za + CS : za;
                    const T tau1n = zr_lds2(rn), v2n = zr_lds2(rn + 16), bn = zr_lds2(zn);
                    T oa, ob;
                    refl(tau1, v2, a, b, oa, ob);
                    zr_sts2(za - CS, oa);
                    a = ob;
                    tau1 = tau1n;
                    v2 = v2n;
                    b = bn;
                    za = zn;
                }
            }
            if (t < 0) t = 0;
            if (t < m && k - 1 + t == NS - 1) {   // the reflector that straddles the two parts
                const T tau1 = zr_lds2(ra + 32u * (uint32_t)t), v2 = zr_lds2(ra + 32u * (uint32_t)t + 16);
                T oa, ob;
                refl(tau1, v2, a, mk_cx<double>(zr[0], zi[0]), oa, ob);
                zr_sts2(zs + CS * (uint32_t)(NS - 1), oa);
                zr[0] = ob.re;
                zi[0] = ob.im;
                t += 1;
            }
        }
        if (t < m) {
            const int j0 = k - 1 + t - NS, j1 = j0 + (m - t) - 1;
            const uint32_t rb = ra + 32u * (uint32_t)t - 32u * (uint32_t)j0;
            for_run<NREG - 1, 4>(j0, j1, rb, ident, [&](int j, uint32_t ad) {
                const T tau1 = zr_lds2(ad), v2 = zr_lds2(ad + 16);
                T oa, ob;
                refl(tau1, v2, mk_cx<double>(zr[j], zi[j]), mk_cx<double>(zr[j + 1], zi[j + 1]), oa, ob);
                zr[j] = oa.re;
                zi[j] = oa.im;
                zr[j + 1] = ob.re;
                zi[j + 1] = ob.im;
            });
        }
    }

    // records [0, cnt) of the page at shared address pg
    GS_DEV void page(uint32_t pg, int cnt) {
        int i = 0;
        while (i < cnt) {
            const uint32_t ra = pg + 32u * (uint32_t)i;
            if (run == 0) {
                int op, kk, count, k2;
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(op), "=r"(kk), "=r"(count), "=r"(k2) : "r"(ra));
                i += 1;
                if (op == LOG_REFL) {
                    run = count;
                    k = kk;
                    if constexpr (NS > 0) {
                        if (k - 1 < NS) a = zr_lds2(zs + CS * (uint32_t)(k - 1));
                    }
                } else {   // LOG_SCALE: columns kk..k2 times t
                    const T t = zr_lds2(ra + 16);
                    if constexpr (NS > 0) {
                        const int je = k2 < NS ? k2 : NS;
                        for (int j = kk; j <= je; ++j) {
                            const uint32_t za = zs + CS * (uint32_t)(j - 1);
                            zr_sts2(za, zr_lds2(za) * t);
                        }
                    }
                    for_cols<NREG>(kk - 1 - NS, k2 - 1 - NS, [&](int j) {
                        const T v = mk_cx<double>(zr[j], zi[j]) * t;
                        zr[j] = v.re;
                        zi[j] = v.im;
                    });
                }
            } else {
                const int m = (run < cnt - i) ? run : (cnt - i);
                segment(m, ra);
                i += m;
                k += m;
                run -= m;
                if constexpr (NS > 0) {
                    if (run == 0 && k - 1 < NS) zr_sts2(zs + CS * (uint32_t)(k - 1), a);
                }
            }
        }
    }
};

// ---- Float64: the whole row in registers ----
template <int N> struct ZRegR {
    double z[N + 1];   // z[N] is a dummy third entry for a two-row reflector on the last two columns
    int run, k;
    uint32_t ident;

    GS_DEV void init(uint32_t ident_) {
        run = 0;
        k = 0;
        ident = ident_;
        z[N] = 0.0;
    }

    GS_DEV void page(uint32_t pg, int cnt) {
        int i = 0;
        while (i < cnt) {
            const uint32_t ra = pg + 32u * (uint32_t)i;
            if (run == 0) {
                int op, kk, count, k2;
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(op), "=r"(kk), "=r"(count), "=r"(k2) : "r"(ra));
                (void)k2;
                i += 1;
                if (op == LOG_REFL3) {
                    run = count;
                    k = kk;
                } else if (op == LOG_REFL2) {
                    // One payload record {tau1, v2, 0, 0} follows (possibly on the next page): a three-row reflector with
                    // v3 = 0 (src/GenericSchur.jl:940-945 — the sum and both updates are bit-identical, the third entry
                    // is left as it is)
                    run = 1;
                    k = kk;
                } else {   // LOG_GIVENS (cs, sn) on columns kk, kk+1
                    const cx<double> g = zr_lds2(ra + 16);
                    const double c = g.re, s = g.im;
                    for_cols<N - 1>(kk - 1, kk - 1, [&](int j) {
                        const double a1 = z[j], a2 = z[j + 1];
                        z[j] = a1 * c + a2 * s;
                        z[j + 1] = -a1 * s + a2 * c;
                    });
                }
            } else {
                const int m = (run < cnt - i) ? run : (cnt - i);
                const int j0 = k - 1, j1 = j0 + m - 1;
                const uint32_t rb = ra - 32u * (uint32_t)j0;
                for_run<N - 1, 4>(j0, j1, rb, ident, [&](int j, uint32_t ad) {
                    const cx<double> r = zr_lds2(ad);
                    const double tau1 = r.re, v2 = r.im, v3 = zr_lds1(ad + 16);
                    const double tau2 = tau1 * v2, tau3 = tau1 * v3;
                    const double z1 = z[j], z2 = z[j + 1], z3 = z[j + 2];
                    const double ss = z1 + v2 * z2 + v3 * z3;                    // src/GenericSchur.jl:920-925
                    z[j] = z1 - ss * tau1;
                    z[j + 1] = z2 - ss * tau2;
                    z[j + 2] = z3 - ss * tau3;
                });
                i += m;
                k += m;
                run -= m;
            }
        }
    }
};

// STG: records per staging buffer (a page, or a fraction of one where the shared memory decides the occupancy)
template <class T, int N, int NREG, int STG = LOG_PAGE_REC> struct zreg_layout {
    static constexpr int NS = N - NREG;
    static constexpr int PAGE_BYTES = LOG_PAGE_REC * 32;
    static constexpr int STG_BYTES = STG * 32;
    static_assert(LOG_PAGE_REC % STG == 0, "a staging buffer holds a whole fraction of a page");
    __host__ __device__ static constexpr size_t off_pages() { return (size_t)NS * N * sizeof(T); }
    __host__ __device__ static constexpr size_t off_ident() { return off_pages() + 2 * (size_t)STG_BYTES; }
    __host__ __device__ static constexpr size_t bytes() { return off_ident() + 32; }
};

// N threads per CTA (one per row, rows >= n idle), one matrix per CTA at a time
template <class T, int N, int NREG, int MINB, int STG = LOG_PAGE_REC>
__global__ void __launch_bounds__(N, MINB) gschur_zreg_kernel(BatchedParams p) {
    typedef zreg_layout<T, N, NREG, STG> ZL;
    constexpr bool CX = etraits<T>::is_complex;
    constexpr int NS = ZL::NS;
    constexpr int PB = ZL::PAGE_BYTES;
    constexpr int SB = ZL::STG_BYTES;
    constexpr int SPP = LOG_PAGE_REC / STG;   // staging buffers per page
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = p.n;
    const int tid = threadIdx.x;
    const uint32_t zs32 = smem_u32(smem_raw);
    const uint32_t pg32 = smem_u32(smem_raw + ZL::off_pages());
    const uint32_t id32 = smem_u32(smem_raw + ZL::off_ident());
    const bool act = tid < n;
    if (tid < 4) reinterpret_cast<double*>(smem_raw + ZL::off_ident())[tid] = 0.0;   // the identity record (visible after the first barrier)
    for (long long b = blockIdx.x; b < p.batch; b += gridDim.x) {
        const int* row = p.log_table + b * (long long)(2 + p.log_maxp);
        const int nrec = row[0];
        if (nrec <= 0 || row[1] != 0) continue;      // nothing logged, or the fused kernel redoes this matrix
        const int npages = (nrec + STG - 1) / STG;   // staging buffers to go through
        T* gZ = reinterpret_cast<T*>(p.Z) + b * p.strideZ + (act ? tid : 0);
        auto fetch_page = [&](int pgi) {
            const unsigned char* src = p.log_pool + (size_t)row[2 + pgi / SPP] * PB + (size_t)(pgi % SPP) * SB;
            const uint32_t dst = pg32 + (uint32_t)((pgi & 1) * SB);
            for (int o = tid * 16; o < SB; o += N * 16) zr_cp_async16(dst + o, src + o);
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        fetch_page(0);
        if constexpr (CX) {
            ZRegC<N, NREG> RP;
            RP.init(zs32 + 16u * (uint32_t)tid, id32);
            if constexpr (NS > 0) {
                for (int c = 0; c < NS && c < n; ++c) zr_sts2(RP.zs + RP.CS * (uint32_t)c, gZ[(size_t)c * p.ldz]);
            }
#pragma unroll
            for (int j = 0; j < NREG; ++j) {
                T v = mk_cx<double>(0.0, 0.0);
                if (NS + j < n) v = gZ[(size_t)(NS + j) * p.ldz];
                RP.zr[j] = v.re;
                RP.zi[j] = v.im;
            }
            for (int pgi = 0; pgi < npages; ++pgi) {
                if (pgi + 1 < npages) {
                    fetch_page(pgi + 1);
                    asm volatile("cp.async.wait_group 1;" ::: "memory");
                } else {
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                }
                __syncthreads();
                const int cnt = (nrec - pgi * STG < STG) ? nrec - pgi * STG : STG;
                RP.page(pg32 + (uint32_t)((pgi & 1) * SB), cnt);
                __syncthreads();
            }
            if (act) {
                if constexpr (NS > 0) {
                    for (int c = 0; c < NS && c < n; ++c) gZ[(size_t)c * p.ldz] = zr_lds2(RP.zs + RP.CS * (uint32_t)c);
                }
#pragma unroll
                for (int j = 0; j < NREG; ++j)
                    if (NS + j < n) gZ[(size_t)(NS + j) * p.ldz] = mk_cx<double>(RP.zr[j], RP.zi[j]);
            }
        } else {
            static_assert(CX || NS == 0, "the Float64 replay keeps the whole row in registers");
            ZRegR<N> RP;
            RP.init(id32);
#pragma unroll
            for (int j = 0; j < N; ++j) {
                double v = 0.0;
                if (j < n) v = reinterpret_cast<double*>(gZ)[(size_t)j * p.ldz];
                RP.z[j] = v;
            }
            for (int pgi = 0; pgi < npages; ++pgi) {
                if (pgi + 1 < npages) {
                    fetch_page(pgi + 1);
                    asm volatile("cp.async.wait_group 1;" ::: "memory");
                } else {
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                }
                __syncthreads();
                const int cnt = (nrec - pgi * STG < STG) ? nrec - pgi * STG : STG;
                RP.page(pg32 + (uint32_t)((pgi & 1) * SB), cnt);
                __syncthreads();
            }
            if (act) {
#pragma unroll
                for (int j = 0; j < N; ++j)
                    if (j < n) reinterpret_cast<double*>(gZ)[(size_t)j * p.ldz] = RP.z[j];
            }
        }
    }
}

}  // namespace gs

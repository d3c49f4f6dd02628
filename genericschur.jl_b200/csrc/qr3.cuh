// Three-stage Schur path for Float64 / ComplexF64 matrices with n <= 64:
//   stage A (gehrd*.cuh)   scale -> Hessenberg -> Q in place,
//   stage B (this file)    QR iteration on H only — ONE warp per matrix, no Z-warp, every right-hand transformation
//                          logged to global memory (qrlog.cuh),
//   stage C (this file)    replay of the log on Z = Q with the whole of Z in shared memory, one thread per row.
// Why: the fused stage B of fastqr.cuh streams Z through L2 once per bulge step (20 MB of L2 traffic per 64x64
// ComplexF64 matrix, 103x the algorithmic bytes) and pays for the Z-warp with registers and named-barrier hand-overs on
// the H-warp's side.  The log is 0.23 MB per matrix, written and read once; stage C touches Z in shared memory only and
// has no serial chain (the rows of Z are independent), so it runs at the shared-memory / FP64 throughput of the SM.
// Every decision rule and the arithmetic per entry are unchanged (src/GenericSchur.jl:194-335, 374-504, 513-699,
// 837-952): T, w, info and stats are bit-identical to the fused kernel's, Z differs at most by the order-preserving
// regrouping of nothing — the same operations are applied to every row in the same order.
#pragma once
#include "fastqr.cuh"
#include "chainqr.cuh"
#include "ownqr.cuh"
#include "zreg.cuh"

namespace gs {

// ---------------------------------------------------------------------------------------------------------------
// Stage B: QR iteration with logging.  32 threads per CTA.
// ---------------------------------------------------------------------------------------------------------------
template <class T, int CPL> struct qrlog_min_blocks {
    static constexpr bool CPLX = etraits<T>::is_complex;
    static constexpr int value = CPLX ? (CPL == 1 ? 12 : 6) : (CPL == 1 ? 16 : 11);
};

template <class T, int CPL>
__global__ void __launch_bounds__(32, qrlog_min_blocks<T, CPL>::value) gschur_qrlog_kernel(BatchedParams p) {
    typedef typename etraits<T>::real R;
    typedef cx<R> C;
    constexpr bool CPLX = etraits<T>::is_complex;
    constexpr int NT = 32;
    typedef FastSolver<T, CPL, true> FS;
    typedef fast_smem_layout<T, CPL> FL;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = p.n;
    const int lane = threadIdx.x;
    const bool wantZ = (p.Z != nullptr);
    FS F;
    F.n = n;
    F.lane = lane;
    F.cap = 0;
    F.H = reinterpret_cast<T*>(smem_raw);
    F.sW = reinterpret_cast<C*>(smem_raw + FL::off_w(n));
    F.ring = nullptr;
    F.hdr = nullptr;
    F.wantZ = wantZ;
    F.ldz = p.ldz;
    F.Z = nullptr;
    T* H = F.H;
    const R zero = r_const<R>(0.0);

    for (;;) {
        long long b = 0;
        if (lane == 0) b = (long long)atomicAdd(p.counter, 1ULL);
        b = __shfl_sync(0xffffffffu, b, 0);
        if (b >= p.batch) break;
        T* gA = reinterpret_cast<T*>(p.A) + b * p.strideA;

        // ---- load the Hessenberg part into packed storage (four columns in flight per lane); bulge slots start at zero ----
        int bad = 0;
        for (int j0 = 1; j0 <= n; j0 += 4) {
            T v[4][CPL + 1];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int j = j0 + q;
#pragma unroll
                for (int s = 0; s <= CPL; ++s) {
                    const int i = 1 + lane + 32 * s;
                    v[q][s] = e_zero<T>();
                    if (j <= n && i <= j + 1 && i <= n) v[q][s] = gA[(i - 1) + (size_t)(j - 1) * p.lda];
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int j = j0 + q;
                if (j > n) continue;
                const int cb = FS::colbase(j);
#pragma unroll
                for (int s = 0; s <= CPL; ++s) {
                    const int i = 1 + lane + 32 * s;
                    if (i <= j + FS::EX) H[cb + i - 1] = v[q][s];
                    if constexpr (CPLX) {
                        if (i == j + 1 && i <= n && (p.flags & F_CHECK_SUBDIAG) && v[q][s].im != zero) bad = 1;
                    }
                }
            }
        }
        bad = __any_sync(0xffffffffu, bad);
        __syncwarp();
        F.lg.init(p, b, lane, wantZ);
        int info = 0;
        unsigned st[4] = {0u, 0u, 0u, 0u};
        if (bad) {
            info = -4;
        } else {
            const int maxiter = p.maxiter > 0 ? p.maxiter : 100 * n;
            if constexpr (CPLX) info = F.qr_complex(maxiter, st);
            else info = F.qr_real(maxiter, st);
            if (F.lg.ovf) info = LOG_OVERFLOW_RC;
        }
        F.lg.finish();
        __syncwarp();
        if (info == LOG_OVERFLOW_RC) {
            // leave H (and Q) untouched in global memory: the fused kernel redoes this matrix after stage C
            if (lane == 0) {
                const unsigned idx = atomicAdd(p.redo_count, 1u);
                p.redo_list[idx] = b;
            }
            continue;
        }

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
                for (int e = lane; e < total; e += NT) H[e] = e_scale(H[e], mul);
                if (!CPLX)
                    for (int e = lane; e < n; e += NT) F.sW[e] = mk_cx<R>(F.sW[e].re * mul, F.sW[e].im * mul);
            });
            __syncwarp();
        }
        // ---- store T (exact zeros below the (quasi-)triangle), w, info, stats ----
        for (int j = 1; j <= n; ++j) {
            const int cb = FS::colbase(j);
            for (int i = 1 + lane; i <= n; i += 32) {
                const bool keep = CPLX ? (i <= j) : (i <= j + 1);
                gA[(i - 1) + (size_t)(j - 1) * p.lda] = keep ? H[cb + i - 1] : e_zero<T>();
            }
        }
        C* gw = reinterpret_cast<C*>(p.w) + b * (long long)n;
        for (int e = lane; e < n; e += NT) {
            if constexpr (CPLX) gw[e] = H[FS::colbase(e + 1) + e];
            else gw[e] = F.sW[e];
        }
        if (lane == 0) {
            if (p.info) p.info[b] = info;
            if (p.stats) {
                p.stats[4 * b + 0] = st[0];
                p.stats[4 * b + 1] = st[1];
                p.stats[4 * b + 2] = st[2];
                p.stats[4 * b + 3] = st[3];
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Stage C: replay of the reflector log on Z.  One CTA per matrix, one thread per row of Z, Z in shared memory
// (leading dimension n: a warp reads 32 consecutive rows of one column — conflict-free), log pages double-buffered
// through cp.async.  Per reflector and row: one shared-memory load, one store, 12 (complex) / 6 (real) FP64 operations.
// ---------------------------------------------------------------------------------------------------------------
GS_DEV void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
GS_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> GS_DEV void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <class T> struct zreplay_layout {
    typedef typename etraits<T>::real R;
    static constexpr int PAGE_BYTES = LOG_PAGE_REC * 4 * (int)sizeof(R);
    __host__ __device__ static size_t off_pages(int n) { return ((size_t)n * n * sizeof(T) + 15) & ~(size_t)15; }
    __host__ __device__ static size_t bytes(int n) { return off_pages(n) + 2 * (size_t)PAGE_BYTES; }
};

// RPT = rows of Z per thread: independent rows interleave in one thread's instruction stream, which hides the
// dependent-issue latency of the carried entry (4 FP64 operations deep per reflector) whatever the scheduler does.
template <class T, int RPT> struct ZReplay;

// ---- ComplexF64 ----
template <int RPT> struct ZReplay<cx<double>, RPT> {
    typedef cx<double> T;
    int run, k;
    T a[RPT];
    GS_DEV void init() {
        run = 0;
        k = 0;
#pragma unroll
        for (int r = 0; r < RPT; ++r) a[r] = mk_cx<double>(0.0, 0.0);
    }
    // apply records [0, cnt) of the page at shared address `pg` to the rows at shared byte addresses zr[] (of Z[r, 1]);
    // cs = column stride in bytes
    GS_DEV void page(uint32_t pg, int cnt, const uint32_t (&zr)[RPT], int nr, uint32_t cs) {
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
#pragma unroll
                    for (int r = 0; r < RPT; ++r) a[r] = lds_e<T>(zr[r] + cs * (uint32_t)(k - 1));
                } else {   // LOG_SCALE: columns kk..k2 times t
                    const T t = lds_e<T>(ra + 16);
                    for (int j = kk; j <= k2; ++j) {
#pragma unroll
                        for (int r = 0; r < RPT; ++r) {
                            const uint32_t za = zr[r] + cs * (uint32_t)(j - 1);
                            if (r < nr) sts_e<T>(za, lds_e<T>(za) * t);
                        }
                    }
                }
            } else {
                const int m = (run < cnt - i) ? run : (cnt - i);
                uint32_t zo = cs * (uint32_t)k;               // byte offset of column k+1 (1-based)
                uint32_t rr = ra;
                for (int t = 0; t < m; ++t) {
                    const T tau1 = lds_e<T>(rr), v2 = lds_e<T>(rr + 16);
                    T b[RPT];
#pragma unroll
                    for (int r = 0; r < RPT; ++r) b[r] = lds_e<T>(zr[r] + zo);
                    const double tau2 = tau1.re * v2.re - tau1.im * v2.im;
#pragma unroll
                    for (int r = 0; r < RPT; ++r) {
                        // ss = tau1 a + tau2 b;  Z[r,k] = a - ss;  a <- b - ss conj(v2)      (src/GenericSchur.jl:455-459)
                        T ss;
                        ss.re = fma(tau1.re, a[r].re, fma(-tau1.im, a[r].im, tau2 * b[r].re));
                        ss.im = fma(tau1.re, a[r].im, fma(tau1.im, a[r].re, tau2 * b[r].im));
                        if (r < nr) sts_e<T>(zr[r] + zo - cs, a[r] - ss);
                        a[r] = e_fnma_cjb(ss, v2, b[r]);
                    }
                    zo += cs;
                    rr += 32u;
                }
                i += m;
                k += m;
                run -= m;
                if (run == 0) {
#pragma unroll
                    for (int r = 0; r < RPT; ++r) if (r < nr) sts_e<T>(zr[r] + cs * (uint32_t)(k - 1), a[r]);
                }
            }
        }
    }
};

// ---- Float64 ----
template <int RPT> struct ZReplay<double, RPT> {
    typedef double T;
    int run, k;
    double z1[RPT], z2[RPT];
    GS_DEV void init() {
        run = 0;
        k = 0;
#pragma unroll
        for (int r = 0; r < RPT; ++r) z1[r] = z2[r] = 0.0;
    }
    GS_DEV void page(uint32_t pg, int cnt, const uint32_t (&zr)[RPT], int nr, uint32_t cs) {
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
#pragma unroll
                    for (int r = 0; r < RPT; ++r) {
                        z1[r] = lds_e<T>(zr[r] + cs * (uint32_t)(k - 1));
                        z2[r] = lds_e<T>(zr[r] + cs * (uint32_t)k);
                    }
                } else if (op == LOG_REFL2) {
                    // header + one payload record (which may sit on the next page): a pending run of length -1
                    run = -1;
                    k = kk;
                } else {   // LOG_GIVENS (cs, sn) on columns kk, kk+1
                    const double c = lds_e<T>(ra + 16), s = lds_e<T>(ra + 24);
#pragma unroll
                    for (int r = 0; r < RPT; ++r) {
                        const uint32_t za = zr[r] + cs * (uint32_t)(kk - 1);
                        const double a1 = lds_e<T>(za), a2 = lds_e<T>(za + cs);
                        if (r < nr) sts_e<T>(za, a1 * c + a2 * s);
                        if (r < nr) sts_e<T>(za + cs, -a1 * s + a2 * c);
                    }
                }
            } else if (run < 0) {   // the payload of a two-row reflector
                const double tau1 = lds_e<T>(ra), v2 = lds_e<T>(ra + 8);
                const double tau2 = tau1 * v2;
#pragma unroll
                for (int r = 0; r < RPT; ++r) {
                    const uint32_t za = zr[r] + cs * (uint32_t)(k - 1);
                    const double x = lds_e<T>(za), y = lds_e<T>(za + cs);
                    const double ss = x + v2 * y;
                    if (r < nr) sts_e<T>(za, x - ss * tau1);
                    if (r < nr) sts_e<T>(za + cs, y - ss * tau2);
                }
                i += 1;
                run = 0;
            } else {
                const int m = (run < cnt - i) ? run : (cnt - i);
                uint32_t zo = cs * (uint32_t)(k + 1);         // byte offset of column k+2 (1-based)
                uint32_t rr = ra;
                for (int t = 0; t < m; ++t) {
                    const double tau1 = lds_e<T>(rr), v2 = lds_e<T>(rr + 8), v3 = lds_e<T>(rr + 16);
                    double z3[RPT];
#pragma unroll
                    for (int r = 0; r < RPT; ++r) z3[r] = lds_e<T>(zr[r] + zo);
                    const double tau2 = tau1 * v2, tau3 = tau1 * v3;
#pragma unroll
                    for (int r = 0; r < RPT; ++r) {
                        const double ss = z1[r] + v2 * z2[r] + v3 * z3[r];           // src/GenericSchur.jl:920-925
                        if (r < nr) sts_e<T>(zr[r] + zo - 2 * cs, z1[r] - ss * tau1);
                        z1[r] = z2[r] - ss * tau2;
                        z2[r] = z3[r] - ss * tau3;
                    }
                    zo += cs;
                    rr += 32u;
                }
                i += m;
                k += m;
                run -= m;
                if (run == 0) {
#pragma unroll
                    for (int r = 0; r < RPT; ++r) {
                        if (r < nr) sts_e<T>(zr[r] + cs * (uint32_t)(k - 1), z1[r]);
                        if (r < nr) sts_e<T>(zr[r] + cs * (uint32_t)k, z2[r]);
                    }
                }
            }
        }
    }
};

// NT threads per CTA, RPT rows per thread: NT * RPT >= n
template <class T, int NT, int RPT> __global__ void __launch_bounds__(NT) gschur_zreplay_kernel(BatchedParams p) {
    typedef zreplay_layout<T> ZL;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = p.n;
    const int tid = threadIdx.x;
    T* Zs = reinterpret_cast<T*>(smem_raw);
    const uint32_t zs32 = smem_u32(Zs);
    const uint32_t pg32 = smem_u32(smem_raw + ZL::off_pages(n));
    constexpr int PB = ZL::PAGE_BYTES;
    for (long long b = blockIdx.x; b < p.batch; b += gridDim.x) {
        const int* row = p.log_table + b * (long long)(2 + p.log_maxp);
        const int nrec = row[0];
        if (nrec <= 0 || row[1] != 0) continue;      // nothing logged, or the fused kernel redoes this matrix
        const int npages = (nrec + LOG_PAGE_REC - 1) / LOG_PAGE_REC;
        T* gZ = reinterpret_cast<T*>(p.Z) + b * p.strideZ;
        auto fetch_page = [&](int pgi) {
            const unsigned char* src = p.log_pool + (size_t)row[2 + pgi] * PB;
            const uint32_t dst = pg32 + (uint32_t)((pgi & 1) * PB);
            for (int o = tid * 16; o < PB; o += NT * 16) cp_async16(dst + o, src + o);
            cp_async_commit();
        };
        fetch_page(0);
        // ---- Z -> shared memory (columns are contiguous in global memory; leading dimension n on chip) ----
        for (int e = tid; e < n * n; e += NT) {
            const int i = e % n, j = e / n;
            Zs[e] = gZ[i + (size_t)j * p.ldz];
        }
        ZReplay<T, RPT> RP;
        RP.init();
        // rows tid, tid + NT, ...: nr of them exist; the others read row 0 and never store
        const bool act = tid < n;
        int nr = 0;
        uint32_t zr[RPT];
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int rowi = tid + NT * r;
            nr += rowi < n ? 1 : 0;
            zr[r] = zs32 + (uint32_t)sizeof(T) * (uint32_t)(rowi < n ? rowi : 0);
        }
        const uint32_t cs = (uint32_t)sizeof(T) * (uint32_t)n;
        for (int pgi = 0; pgi < npages; ++pgi) {
            if (pgi + 1 < npages) {
                fetch_page(pgi + 1);
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();
            const int cnt = (nrec - pgi * LOG_PAGE_REC < LOG_PAGE_REC) ? nrec - pgi * LOG_PAGE_REC : LOG_PAGE_REC;
            if (act) RP.page(pg32 + (uint32_t)((pgi & 1) * PB), cnt, zr, nr, cs);
            __syncthreads();
        }
        for (int e = tid; e < n * n; e += NT) {
            const int i = e % n, j = e / n;
            gZ[i + (size_t)j * p.ldz] = Zs[e];
        }
        __syncthreads();
    }
}

// stage C variants by size: kernel, threads per CTA, dynamic shared memory
struct StageCKernel {
    void (*fn)(BatchedParams);
    int threads;
    size_t smem;
    const char* name;
};
template <class T> StageCKernel stage_c_select(int n) {
    StageCKernel k;
    constexpr bool CX = etraits<T>::is_complex;
    // GSCHUR_ZREG=0: first-generation replay (Z in shared memory); default: rows of Z in registers (zreg.cuh)
    const char* zsel = std::getenv("GSCHUR_ZREG");
    if (!(zsel && zsel[0] == '0')) {
        if (n <= 32) {
            k.fn = gschur_zreg_kernel<T, 32, 32, (CX ? 12 : 16)>;
            k.threads = 32;
            k.smem = zreg_layout<T, 32, 32>::bytes();
        } else {
            constexpr int NREG = CX ? 32 : 64;   // ComplexF64: the leading 32 columns stay in shared memory
            // ... 32 KB of it: half-page staging buffers (2 x 2 KB) let a sixth CTA onto the SM (registers allow six)
            constexpr int STG = CX ? LOG_PAGE_REC / 2 : LOG_PAGE_REC;
            k.fn = gschur_zreg_kernel<T, 64, NREG, 6, STG>;
            k.threads = 64;
            k.smem = zreg_layout<T, 64, NREG, STG>::bytes();
        }
        k.name = "zreg";
        return k;
    }
    k.smem = zreplay_layout<T>::bytes(n);
    k.name = "zreplay";
    if (n <= 32) {
        k.fn = gschur_zreplay_kernel<T, 32, 1>;
        k.threads = 32;
    } else if (CX) {
        // 64 x 64 ComplexF64: Z is 64 KB, three CTAs per SM — one warp with two rows per thread each (GSCHUR_ZRPT=1: two warps)
        const char* e = std::getenv("GSCHUR_ZRPT");
        if (e && e[0] == '1') {
            k.fn = gschur_zreplay_kernel<T, 64, 1>;
            k.threads = 64;
        } else {
            k.fn = gschur_zreplay_kernel<T, 32, 2>;
            k.threads = 32;
        }
    } else {
        k.fn = gschur_zreplay_kernel<T, 64, 1>;
        k.threads = 64;
    }
    return k;
}

// ---------------------------------------------------------------------------------------------------------------
// Host side: A -> B(log) -> C -> redo, in sub-batches sized by the log-pool budget.
// ---------------------------------------------------------------------------------------------------------------
inline size_t log_pool_budget_bytes() {
    size_t budget = (size_t)24 << 30;
    if (const char* e = std::getenv("GSCHUR_LOG_POOL_MB")) {
        const long long v = std::atoll(e);
        if (v >= 16) budget = (size_t)v << 20;
    }
    size_t fr = 0, tot = 0;
    if (cudaMemGetInfo(&fr, &tot) == cudaSuccess && fr / 2 < budget) budget = fr / 2;
    return budget;
}

// stage B variants: kernel, dynamic shared memory, matrices per (one-warp) CTA
struct StageBKernel {
    void (*fn)(BatchedParams);
    size_t smem;
    int per_cta;
    const char* name;
    int threads = 32;
};
template <class T, int CPL> StageBKernel stage_b_select(int n) {
    StageBKernel k;
    k.fn = gschur_qrlog_kernel<T, CPL>;
    k.smem = fast_smem_layout<T, CPL>::off_ring(n);
    k.per_cta = 1;
    k.name = "qrlog";
    // GSCHUR_CHAIN = 1 (default): owner-computes sweeps (ownqr.cuh); 32 / 16: lanes per matrix of the chain kernels
    // (chainqr.cuh); 0: first-generation logging kernel
    const char* sel = std::getenv("GSCHUR_CHAIN");
    const int lpm = sel ? std::atoi(sel) : 1;
    constexpr bool CX = std::is_same<T, cx<double>>::value;
    if (lpm == 32) {
        k.fn = gschur_chain_kernel<T, 32, CPL, (CX ? (CPL == 1 ? 12 : 6) : (CPL == 1 ? 16 : 12))>;
        k.smem = chain_layout<T, 32, CPL>::bytes(n);
        k.per_cta = 1;
        k.name = "chain32";
    } else if (lpm == 1) {   // owner-computes sweeps (ownqr.cuh)
        k.fn = gschur_chain_kernel<T, 32, CPL, (CX ? (CPL == 1 ? 12 : 6) : (CPL == 1 ? 16 : 12)), 1>;
        k.smem = chain_layout<T, 32, CPL>::bytes(n);
        k.per_cta = 1;
        k.name = "own";
    } else if (lpm == 16) {
        k.fn = gschur_chain_kernel<T, 16, 2 * CPL, (CX ? (CPL == 1 ? 10 : 3) : (CPL == 1 ? 12 : 6))>;
        k.smem = chain_layout<T, 16, 2 * CPL>::bytes(n);
        k.per_cta = 2;
        k.name = "chain16";
    }
    return k;
}

template <class T, int CPL> int launch_fast3(const BatchedParams& p_in, int dev_sms, cudaStream_t stream, std::string* err) {
    typedef typename etraits<T>::real R;
    constexpr bool CPLX = etraits<T>::is_complex;
    constexpr size_t PB = zreplay_layout<T>::PAGE_BYTES;
    const int n = p_in.n;
    const bool hess_input = (p_in.flags & F_HESS_INPUT) != 0;
    const bool wantZ = p_in.Z != nullptr;
    cudaError_t e = cudaSuccess;
#define F3_TRY(expr, what)                                                   \
    do {                                                                     \
        e = (expr);                                                          \
        if (e != cudaSuccess) {                                              \
            *err = std::string(what) + ": " + cudaGetErrorString(e);         \
            rc = -2;                                                         \
            goto done;                                                       \
        }                                                                    \
    } while (0)
    int rc = 0;
    // log geometry
    const long long exp_rec = log_expected_records(CPLX, n);
    const long long exp_pages = (exp_rec + LOG_PAGE_REC - 1) / LOG_PAGE_REC;
    const bool tiny = std::getenv("GSCHUR_LOG_TEST_TINY") != nullptr;   // test knob: force log overflows (redo path)
    const int maxp = tiny ? 2 : (int)(4 * exp_pages + 4);            // per-matrix cap: ~4x a random matrix
    const long long pages_per = (5 * exp_pages) / 4 + 1;             // pool share per matrix
    const size_t per_matrix = wantZ ? (size_t)pages_per * PB + (size_t)(2 + maxp) * sizeof(int) : 0;
    long long sub = p_in.batch;
    if (wantZ) {
        const size_t budget = log_pool_budget_bytes();
        const long long fit = (long long)(budget / per_matrix);
        const long long floor_sub = 4096;
        if (sub > fit) sub = fit < floor_sub ? floor_sub : fit;
        if (sub > p_in.batch) sub = p_in.batch;
    }
    double* scratch = nullptr;          // per matrix: 8 doubles of scale info; then the counters
    unsigned char* pool = nullptr;
    int* table = nullptr;
    long long* redo = nullptr;
    unsigned long long* ctr = nullptr;  // [0] stage A queue is p.counter; ctr[0] stage B queue, ctr[1] redo queue, then two u32
    unsigned pool_pages = 0;
    const size_t ctr_bytes = 64;
    F3_TRY(cudaMallocAsync((void**)&scratch, (size_t)sub * 8 * sizeof(double) + ctr_bytes, stream), "cudaMallocAsync(scratch)");
    ctr = reinterpret_cast<unsigned long long*>(scratch + (size_t)sub * 8);
    F3_TRY(cudaMallocAsync((void**)&redo, (size_t)sub * sizeof(long long), stream), "cudaMallocAsync(redo list)");
    if (wantZ) {
        long long pp = sub * pages_per;
        const long long min_pp = (long long)maxp * (sub < 64 ? sub : 64);   // small batches: room for a few hard matrices
        if (pp < min_pp) pp = min_pp;
        if (tiny) pp = 8;
        pool_pages = (unsigned)pp;
        F3_TRY(cudaMallocAsync((void**)&pool, (size_t)pool_pages * PB, stream), "cudaMallocAsync(log pool)");
        F3_TRY(cudaMallocAsync((void**)&table, (size_t)sub * (2 + maxp) * sizeof(int), stream), "cudaMallocAsync(log table)");
    }
    {
        const StageBKernel sb = stage_b_select<T, CPL>(n);
        auto kB = sb.fn;
        const StageCKernel sc = stage_c_select<T>(n);
        auto kC = sc.fn;
        auto kR = gschur_qr_kernel<T, CPL>;
        const size_t smemB = sb.smem;
        const size_t smemC = sc.smem;
        const size_t smemR = fast_smem_layout<T, CPL>::bytes(n);
        int perB = 0, perR = 0;
        F3_TRY(cudaFuncSetAttribute(kB, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemB), "qrlog kernel setup");
        F3_TRY(cudaFuncSetAttribute(kB, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared), "qrlog kernel setup");
        F3_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perB, kB, sb.threads, smemB), "qrlog kernel occupancy");
        F3_TRY(cudaFuncSetAttribute(kC, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemC), "zreplay kernel setup");
        F3_TRY(cudaFuncSetAttribute(kC, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared), "zreplay kernel setup");
        F3_TRY(cudaFuncSetAttribute(kR, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemR), "qr kernel setup");
        F3_TRY(cudaFuncSetAttribute(kR, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared), "qr kernel setup");
        F3_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perR, kR, qr_threads<T, CPL>::value, smemR), "qr kernel occupancy");
        if (perB < 1 || perR < 1) {
            *err = "qr kernels do not fit on an SM";
            rc = -3;
            goto done;
        }
        static const int cap_per_sm = std::getenv("GSCHUR_QR_CTAS_PER_SM") ? std::atoi(std::getenv("GSCHUR_QR_CTAS_PER_SM")) : 0;
        if (cap_per_sm > 0 && cap_per_sm < perB) perB = cap_per_sm;   // profiling knob: occupancy sweep
        for (long long b0 = 0; b0 < p_in.batch; b0 += sub) {
            const long long cn = (p_in.batch - b0 < sub) ? p_in.batch - b0 : sub;
            BatchedParams p = p_in;
            p.batch = cn;
            p.A = (char*)p_in.A + (size_t)b0 * p_in.strideA * sizeof(T);
            if (wantZ) p.Z = (char*)p_in.Z + (size_t)b0 * p_in.strideZ * sizeof(T);
            p.w = (char*)p_in.w + (size_t)b0 * n * sizeof(cx<R>);
            if (p_in.info) p.info = p_in.info + b0;
            if (p_in.stats) p.stats = p_in.stats + 4 * b0;
            p.log_pool = pool;
            p.log_pages = pool_pages;
            p.log_table = table;
            p.log_maxp = maxp;
            p.redo_list = redo;
            p.log_next = reinterpret_cast<unsigned*>(ctr + 2);
            p.redo_count = reinterpret_cast<unsigned*>(ctr + 2) + 1;
            F3_TRY(cudaMemsetAsync(ctr, 0, ctr_bytes, stream), "cudaMemsetAsync(counters)");
            F3_TRY(cudaMemsetAsync(p_in.counter, 0, sizeof(unsigned long long), stream), "cudaMemsetAsync(counter)");
            stage_timing_mark(0, stream);
            if (!hess_input) {
                p.scratch = scratch;
                int rca = launch_stage_a<T>(p, dev_sms, stream, err);
                if (rca) {
                    rc = rca;
                    goto done;
                }
            } else {
                p.scratch = nullptr;
            }
            stage_timing_mark(1, stream);
            p.counter = ctr;
            long long grid = (long long)perB * dev_sms;
            if (grid > (cn + sb.per_cta - 1) / sb.per_cta) grid = (cn + sb.per_cta - 1) / sb.per_cta;
            kB<<<(unsigned)grid, sb.threads, smemB, stream>>>(p);
            note_launch();
            stage_timing_mark(2, stream);
            if (wantZ) {
                kC<<<(unsigned)(cn < 0x7fffffffLL ? cn : 0x7fffffffLL), sc.threads, smemC, stream>>>(p);
                note_launch();
            }
            // redo pass: matrices whose log overflowed (none for ordinary input: every CTA exits at once)
            if (!wantZ) {
                stage_timing_mark(3, stream);
                continue;
            }
            p.counter = ctr + 1;
            p.list = redo;
            p.list_count = p.redo_count;
            grid = (long long)perR * dev_sms;
            if (grid > cn) grid = cn;
            if (grid > 2 * dev_sms) grid = 2 * dev_sms;
            kR<<<(unsigned)grid, qr_threads<T, CPL>::value, smemR, stream>>>(p);
            note_launch();
            stage_timing_mark(3, stream);
            F3_TRY(cudaGetLastError(), "three-stage kernel launch");
        }
    }
done:
    if (table) cudaFreeAsync(table, stream);
    if (pool) cudaFreeAsync(pool, stream);
    if (redo) cudaFreeAsync(redo, stream);
    if (scratch) cudaFreeAsync(scratch, stream);
    return rc;
#undef F3_TRY
}

}  // namespace gs

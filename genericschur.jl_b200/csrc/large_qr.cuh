// Regime (2), part 2: QR iteration on ONE large upper Hessenberg Float64 matrix.
//
// The reference's sweep (doubleShiftQR!, src/GenericSchur.jl:837-952) chases one 3x3 bulge down the whole matrix,
// touching 3 rows x n columns and 3 columns x n rows per step — 15 million strictly sequential steps at n = 4096.
// Here the same 3x3 double-shift bulges (same reflector, same first-column formula, same deflation criterion) are
//   * chased NB at a time as a chain (bulge b trails bulge b-1 by 4 columns; all advance in lock-step, one warp per
//     bulge, on disjoint rows in the left phase and disjoint columns in the right phase), with shift pairs taken
//     from the eigenvalues of the trailing 2 NB x 2 NB block of the active window (small-bulge multishift);
//   * confined to a diagonal window held in shared memory while the product U of their reflectors is accumulated;
//   * applied to everything outside the window — the rest of H's rows and columns and all of Z — as FP64 DMMA GEMMs
//     with U (dgemm.cuh).
// Active blocks of order <= 128 are finished by the batched kernel (fastqr.cuh) and back-transformed with GEMMs.
// Deflation uses the reference's test (Ahues-Tisseur with its `aa + bb`, src/GenericSchur.jl:563-600).
#pragma once
#include <algorithm>
#include <chrono>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "batched.cuh"
#include "large_gehrd.cuh"

namespace gs {

constexpr int LQ_NB = 12;       // bulges per chain (2 NB shifts per sweep)
constexpr int LQ_W = 112;       // maximum window order (H window + U in shared memory: 2 x 112 x 113 x 8 B = 198 KB)
constexpr int LQ_LDW = 113;     // shared-memory leading dimension (odd: conflict-free rows and columns)
constexpr int LQ_UW = LQ_NB;    // warps that apply the reflectors to U: one per bulge (reflector), its lanes stride over the rows of U
constexpr int LQ_MAXC = 4;      // chains of NB bulges in flight per sweep (each on its own SM)
constexpr int LQ_DELTA = 192;   // column distance between consecutive chains (>= window + steps per window)
constexpr int LQ_SMALL = 128;   // active blocks up to this order go to the batched kernel

struct LqScan {
    int istart;   // 1-based start of the active block ending at iend
};

// Largest k in (1, iend] whose sub-diagonal H[k,k-1] is negligible by the reference's criterion; istart = k (or 1).
// One CTA; the split entry is set to zero as the reference does (src/GenericSchur.jl:604-606).
__global__ void __launch_bounds__(1024) lq_scan_kernel(double* H, int n, int iend, LqScan* out) {
    __shared__ int best;
    if (threadIdx.x == 0) best = 1;
    __syncthreads();
    const double eps = 2.220446049250313e-16;
    const double smallnum = 2.2250738585072014e-308 * ((double)n / eps);
#define HL(i, j) H[((size_t)(i)-1) + ((size_t)(j)-1) * n]
    for (int base = iend; base >= 2; base -= 1024) {
        const int k = base - threadIdx.x;
        bool hit = false;
        if (k >= 2) {
            const double h = fabs(HL(k, k - 1));
            if (h < smallnum) hit = true;
            else {
                const double Hkk = HL(k, k), Hk1 = HL(k - 1, k - 1);
                double t = fabs(Hk1) + fabs(Hkk);
                if (t == 0.0) {
                    if (k > 2) t += fabs(HL(k - 1, k - 2));
                    if (k + 1 <= n) t += fabs(HL(k + 1, k));
                }
                if (h <= t * eps) {
                    const double o = fabs(HL(k - 1, k));
                    const double ab = fmax(h, o), ba = fmin(h, o);
                    const double d1 = fabs(Hkk), d2 = fabs(Hk1 - Hkk);
                    const double aa = fmax(d1, d2), bb = fmin(d1, d2);
                    const double s = aa + bb;
                    if (ba * (ab / s) <= fmax(smallnum, eps * (bb * (aa / s)))) hit = true;
                }
            }
        }
        if (hit) atomicMax(&best, k);
        __syncthreads();
        if (best > 1) break;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        out->istart = best;
        if (best > 1) HL(best, best - 1) = 0.0;
    }
#undef HL
}

// Chase the chain through one window.  Bulge b at time tau sits at column p = L + tau - 4 b (1-based); it is active
// while L <= p <= I-1.  The window is rows/columns wlo..whi of H.  U (wsz x wsz, ld LQ_W, global) receives the
// accumulated orthogonal factor: H_window <- U' H_window U.
// Warps 0..NB-1 chase one bulge each (left phase | barrier | right phase | barrier); warp NB+q applies reflector q of
// the step to U (all rows, strided over its lanes) off the bulge warps' critical path, picking the parameters up from
// a double-buffered slot in shared memory at the same two barriers.  (First version: four U warps, one row of U per
// lane, the 12 reflectors of a step one after the other, and run-time loops in the phases: 138 us per window of 62
// steps — 4300 cycles per step, the serial bottleneck of the whole large-matrix path.)
struct lq_refl {
    double tau1, v1, v2;
    int pc, nr;      // window-local column, reflector order (0 = inactive)
};
struct LqChaseArgs {
    int wlo[LQ_MAXC], whi[LQ_MAXC];     // window of each chain (whi < wlo: chain inactive in this time block)
    double* U[LQ_MAXC];
};
__global__ void __launch_bounds__(32 * (LQ_NB + LQ_UW)) lq_chase_kernel(double* H, int n, int L, int I, const double* shifts_all,
                                                                       int tau0_all, int M, LqChaseArgs args) {
    const int chain = blockIdx.x;
    const int wlo = args.wlo[chain], whi = args.whi[chain];
    if (whi < wlo) return;
    double* Uout = args.U[chain];
    const double* shifts = shifts_all + 4 * LQ_NB * chain;
    const int tau0 = tau0_all - LQ_DELTA * chain;     // chain c runs LQ_DELTA columns behind chain c-1
    extern __shared__ double sm[];
    double* Hw = sm;                       // LQ_W x LQ_LDW
    double* Uw = sm + LQ_W * LQ_LDW;       // LQ_W x LQ_LDW
    lq_refl* slots = reinterpret_cast<lq_refl*>(Uw + LQ_W * LQ_LDW);   // [2][LQ_NB]
    const int tid = threadIdx.x, lane = tid & 31, b = tid >> 5;
    const bool is_bulge = b < LQ_NB;
    const int wsz = whi - wlo + 1;
#define HW(i, j) Hw[((i)-wlo) + ((j)-wlo) * LQ_LDW]          // global 1-based indices
#define UW(i, j) Uw[(i) + (j)*LQ_LDW]                        // local 0-based indices
    for (int e = tid; e < wsz * wsz; e += blockDim.x) {
        const int i = e % wsz, j = e / wsz;
        Hw[i + j * LQ_LDW] = H[(size_t)(wlo - 1 + i) + (size_t)(wlo - 1 + j) * n];
        Uw[i + j * LQ_LDW] = (i == j) ? 1.0 : 0.0;
    }
    __syncthreads();
    double r1r = 0, r1i = 0, r2r = 0, r2i = 0;
    if (is_bulge) {
        r1r = shifts[4 * b + 0];
        r1i = shifts[4 * b + 1];
        r2r = shifts[4 * b + 2];
        r2i = shifts[4 * b + 3];
    }
    // A phase touches at most LQ_W columns (rows): LQ_IT strided passes per lane, fully unrolled — all loads of a phase are
    // issued before the first dependent FMA (a run-time loop serialises 3 loads -> 5 FMAs -> 3 stores per pass).
    constexpr int LQ_IT = (LQ_W + 31) / 32;
    for (int tau = tau0; tau < tau0 + M; ++tau) {
        lq_refl* slot = slots + (tau & 1) * LQ_NB;
        int p = 0, nr = 0;
        bool active = false;
        double v0 = 0.0, v1 = 0.0, v2 = 0.0, tau1 = 0.0, tau2 = 0.0, tau3 = 0.0;
        if (is_bulge) {
            p = L + tau - 4 * b;
            active = (p >= L && p <= I - 1);
            nr = (I - p + 1 < 3) ? I - p + 1 : 3;
            if (active) {
                if (p == L) {
                    // first column of (H - s1)(H - s2), src/GenericSchur.jl:855-863
                    const double hmm = HW(L, L);
                    double H21s = HW(L + 1, L);
                    double s = fabs(hmm - r2r) + fabs(r2i) + fabs(H21s);
                    H21s = H21s / s;
                    v0 = H21s * HW(L, L + 1) + (hmm - r1r) * ((hmm - r2r) / s) - r1i * (r2i / s);
                    v1 = H21s * (hmm + HW(L + 1, L + 1) - r1r - r2r);
                    v2 = (nr == 3) ? H21s * HW(L + 2, L + 1) : 0.0;
                    s = fabs(v0) + fabs(v1) + fabs(v2);
                    v0 /= s;
                    v1 /= s;
                    v2 /= s;
                } else {
                    v0 = HW(p, p - 1);
                    v1 = HW(p + 1, p - 1);
                    v2 = (nr == 3) ? HW(p + 2, p - 1) : 0.0;
                }
                tau1 = reflector_real_small(v0, v1, v2, nr);
                tau2 = tau1 * v1;
                tau3 = tau1 * v2;
                // ---- left phase: rows p..p+nr-1, columns p..whi (disjoint rows across bulges) ----
                {
                    const bool r3 = nr == 3;
                    double* hp = &HW(p, p + lane);
                    const int left = whi - p - lane;          // columns p+lane+32 i with 32 i <= left
                    double a[LQ_IT], bb[LQ_IT], c[LQ_IT];
#pragma unroll
                    for (int i = 0; i < LQ_IT; ++i) {
                        const bool in = 32 * i <= left;
                        a[i] = in ? hp[32 * i * LQ_LDW] : 0.0;
                        bb[i] = in ? hp[32 * i * LQ_LDW + 1] : 0.0;
                        c[i] = (in && r3) ? hp[32 * i * LQ_LDW + 2] : 0.0;
                    }
#pragma unroll
                    for (int i = 0; i < LQ_IT; ++i) {
                        const double ss = a[i] + v1 * bb[i] + v2 * c[i];
                        if (32 * i <= left) {
                            hp[32 * i * LQ_LDW] = a[i] - ss * tau1;
                            hp[32 * i * LQ_LDW + 1] = bb[i] - ss * tau2;
                            if (r3) hp[32 * i * LQ_LDW + 2] = c[i] - ss * tau3;
                        }
                    }
                }
                if (lane == 0 && p > L) {
                    HW(p, p - 1) = v0;
                    HW(p + 1, p - 1) = 0.0;
                    if (nr == 3) HW(p + 2, p - 1) = 0.0;
                }
            }
            if (lane == 0) {
                slot[b].tau1 = tau1;
                slot[b].v1 = v1;
                slot[b].v2 = v2;
                slot[b].pc = p - wlo;
                slot[b].nr = active ? nr : 0;
            }
        }
        __syncthreads();
        if (is_bulge) {
            // ---- right phase: columns p..p+nr-1, rows wlo..min(p+3, I) (disjoint columns across bulges) ----
            if (active) {
                const bool r3 = nr == 3;
                const int rmax = (p + 3 < I) ? p + 3 : I;
                double* hp = &HW(wlo + lane, p);
                const int left = rmax - wlo - lane;           // rows wlo+lane+32 i with 32 i <= left
                double a[LQ_IT], bb[LQ_IT], c[LQ_IT];
#pragma unroll
                for (int i = 0; i < LQ_IT; ++i) {
                    const bool in = 32 * i <= left;
                    a[i] = in ? hp[32 * i] : 0.0;
                    bb[i] = in ? hp[32 * i + LQ_LDW] : 0.0;
                    c[i] = (in && r3) ? hp[32 * i + 2 * LQ_LDW] : 0.0;
                }
#pragma unroll
                for (int i = 0; i < LQ_IT; ++i) {
                    const double ss = a[i] + v1 * bb[i] + v2 * c[i];
                    if (32 * i <= left) {
                        hp[32 * i] = a[i] - ss * tau1;
                        hp[32 * i + LQ_LDW] = bb[i] - ss * tau2;
                        if (r3) hp[32 * i + 2 * LQ_LDW] = c[i] - ss * tau3;
                    }
                }
            }
        } else {
            // ---- U <- U G: warp NB + q applies reflector q of this step to all rows of U (disjoint columns across warps) ----
            const int q = b - LQ_NB;
            const int rn = slot[q].nr;
            if (rn != 0) {
                const bool r3 = rn == 3;
                const double t1 = slot[q].tau1, w1 = slot[q].v1, w2 = slot[q].v2;
                const double t2 = t1 * w1, t3 = t1 * w2;
                double* up = &UW(lane, slot[q].pc);
                const int left = wsz - 1 - lane;
                double a[LQ_IT], bb[LQ_IT], c[LQ_IT];
#pragma unroll
                for (int i = 0; i < LQ_IT; ++i) {
                    const bool in = 32 * i <= left;
                    a[i] = in ? up[32 * i] : 0.0;
                    bb[i] = in ? up[32 * i + LQ_LDW] : 0.0;
                    c[i] = (in && r3) ? up[32 * i + 2 * LQ_LDW] : 0.0;
                }
#pragma unroll
                for (int i = 0; i < LQ_IT; ++i) {
                    const double ss = a[i] + w1 * bb[i] + w2 * c[i];
                    if (32 * i <= left) {
                        up[32 * i] = a[i] - ss * t1;
                        up[32 * i + LQ_LDW] = bb[i] - ss * t2;
                        if (r3) up[32 * i + 2 * LQ_LDW] = c[i] - ss * t3;
                    }
                }
            }
        }
        __syncthreads();
    }
    for (int e = tid; e < wsz * wsz; e += blockDim.x) {
        const int i = e % wsz, j = e / wsz;
        H[(size_t)(wlo - 1 + i) + (size_t)(wlo - 1 + j) * n] = Hw[i + j * LQ_LDW];
        Uout[i + j * LQ_W] = Uw[i + j * LQ_LDW];
    }
#undef HW
#undef UW
}

// One launch applies the accumulated factors of all chains of a time block to everything outside their windows,
// in place, on the FP64 tensor cores:
//   left  strips: H[wlo:whi, c0:c0+63] <- U' * H[...]         (64 columns per CTA)
//   right strips: H[r0:r0+63, wlo:whi] <- H[...] * U  (rows above the window),  Z[r0:r0+63, wlo:whi] <- Z[...] * U
// A CTA stages U (<= 112 x 112) and its strip in shared memory (leading dimensions = 4 mod 16: the DMMA fragment
// loads are conflict-free), so the update is safe in place, then each of its 8 warps owns 8 columns (left) or
// 8 rows (right) of the strip: 14 m8n8 accumulator tiles per warp.
constexpr int LQ_LDU = 116, LQ_LDT_L = 116, LQ_LDT_R = 68;
struct LqApplyArgs {
    int C, mode;                         // mode bits: 1 = left strips, 2 = right strips of H (top), 4 = right strips of Z
    int wlo[LQ_MAXC], wsz[LQ_MAXC];      // wsz <= 0: chain inactive
    int lc0[LQ_MAXC], lc1[LQ_MAXC];      // left strips cover columns lc0..lc1 (1-based, inclusive)
    const double* U[LQ_MAXC];
};
__host__ __device__ inline int lq_strips(int count) { return count > 0 ? (count + 63) / 64 : 0; }

__global__ void __launch_bounds__(256) lq_apply_kernel(double* H, double* Z, int n, LqApplyArgs a) {
    extern __shared__ double sm[];
    double* Us = sm;                              // [k + m * LQ_LDU]
    double* Ts = sm + LQ_W * LQ_LDU;              // strip tile
    // ---- decode blockIdx -> (chain, type, strip) ----
    int rem = blockIdx.x, chain = -1, type = -1, strip = 0;
    for (int c = 0; c < a.C && chain < 0; ++c) {
        if (a.wsz[c] <= 0) continue;
        const int nl = (a.mode & 1) ? lq_strips(a.lc1[c] - a.lc0[c] + 1) : 0;
        const int nt = (a.mode & 2) ? lq_strips(a.wlo[c] - 1) : 0;
        const int nz = ((a.mode & 4) && Z) ? lq_strips(n) : 0;
        if (rem < nl) { chain = c; type = 0; strip = rem; }
        else if (rem < nl + nt) { chain = c; type = 1; strip = rem - nl; }
        else if (rem < nl + nt + nz) { chain = c; type = 2; strip = rem - nl - nt; }
        else rem -= nl + nt + nz;
    }
    if (chain < 0) return;
    const int wlo = a.wlo[chain], wsz = a.wsz[chain];
    const double* U = a.U[chain];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, gid = lane >> 2, tig = lane & 3;
    const int K = (wsz + 3) & ~3;
    for (int e = tid; e < LQ_W * LQ_W; e += 256) {
        const int i = e % LQ_W, j = e / LQ_W;
        Us[i + j * LQ_LDU] = (i < wsz && j < wsz) ? U[i + j * LQ_W] : 0.0;
    }
    double acc[14][2];
#pragma unroll
    for (int t = 0; t < 14; ++t) acc[t][0] = acc[t][1] = 0.0;
    if (type == 0) {
        const int c0 = a.lc0[chain] + 64 * strip;                       // 1-based first column
        const int cnt = min(64, a.lc1[chain] - c0 + 1);
        double* C = H + (size_t)(wlo - 1) + (size_t)(c0 - 1) * n;
        for (int e = tid; e < LQ_W * 64; e += 256) {
            const int i = e % LQ_W, j = e / LQ_W;
            Ts[i + j * LQ_LDT_L] = (i < wsz && j < cnt) ? C[(size_t)i + (size_t)j * n] : 0.0;
        }
        __syncthreads();
        for (int k = 0; k < K; k += 4) {
            const double b = Ts[(k + tig) + (w * 8 + gid) * LQ_LDT_L];
#pragma unroll
            for (int mt = 0; mt < 14; ++mt) dmma8x8x4(acc[mt][0], acc[mt][1], Us[(k + tig) + (mt * 8 + gid) * LQ_LDU], b);
        }
#pragma unroll
        for (int mt = 0; mt < 14; ++mt)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int row = mt * 8 + gid, col = w * 8 + 2 * tig + q;
                if (row < wsz && col < cnt) C[(size_t)row + (size_t)col * n] = acc[mt][q];
            }
    } else {
        double* base = (type == 1) ? H : Z;
        const int rows_total = (type == 1) ? wlo - 1 : n;
        const int r0 = 64 * strip;                                      // 0-based first row
        const int cnt = min(64, rows_total - r0);
        double* C = base + (size_t)r0 + (size_t)(wlo - 1) * n;
        for (int e = tid; e < 64 * LQ_W; e += 256) {
            const int i = e % 64, j = e / 64;
            Ts[i + j * LQ_LDT_R] = (i < cnt && j < wsz) ? C[(size_t)i + (size_t)j * n] : 0.0;
        }
        __syncthreads();
        for (int k = 0; k < K; k += 4) {
            const double av = Ts[(w * 8 + gid) + (k + tig) * LQ_LDT_R];
#pragma unroll
            for (int nt = 0; nt < 14; ++nt) dmma8x8x4(acc[nt][0], acc[nt][1], av, Us[(k + tig) + (nt * 8 + gid) * LQ_LDU]);
        }
#pragma unroll
        for (int nt = 0; nt < 14; ++nt)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int row = w * 8 + gid, col = nt * 8 + 2 * tig + q;
                if (row < cnt && col < wsz) C[(size_t)row + (size_t)col * n] = acc[nt][q];
            }
    }
}

__global__ void lq_copy_block_kernel(const double* src, int lds, double* dst, int ldd, int rows, int cols) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < (size_t)rows * cols) {
        const int i = idx % rows, j = idx / rows;
        dst[(size_t)i + (size_t)j * ldd] = src[(size_t)i + (size_t)j * lds];
    }
}
__global__ void lq_eye_kernel(double* dst, int ld, int m) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < m * m) dst[(idx % m) + (size_t)(idx / m) * ld] = ((idx % m) == (idx / m)) ? 1.0 : 0.0;
}
// eigenvalues of a quasi-triangular T in standard form (2x2 blocks: a +- i sqrt|b| sqrt|c|)
__global__ void lq_eigs_kernel(const double* T, int n, double* w) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;   // 0-based
    if (j >= n) return;
#define TT(i, k) T[(size_t)(i) + (size_t)(k)*n]
    const bool sub_below = (j + 1 < n) && TT(j + 1, j) != 0.0;
    const bool sub_above = (j >= 1) && TT(j, j - 1) != 0.0;
    double re = TT(j, j), im = 0.0;
    if (sub_below) im = sqrt(fabs(TT(j, j + 1))) * sqrt(fabs(TT(j + 1, j)));
    else if (sub_above) im = -sqrt(fabs(TT(j - 1, j))) * sqrt(fabs(TT(j, j - 1)));
    w[2 * j] = re;
    w[2 * j + 1] = im;
#undef TT
}

inline void lq_copy(cudaStream_t s, const double* src, int lds, double* dst, int ldd, int rows, int cols) {
    if (rows <= 0 || cols <= 0) return;
    const size_t tot = (size_t)rows * cols;
    lq_copy_block_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(src, lds, dst, ldd, rows, cols);
    note_launch();
}

// streams, events and scratch of one lg_qr call: released on every way out (the error returns of LG_TRY included)
struct LqResources {
    cudaStream_t caller = nullptr;
    cudaStream_t* sA = nullptr;
    cudaStream_t* sB = nullptr;
    cudaEvent_t* evA = nullptr;
    cudaEvent_t* evB = nullptr;
    double** base = nullptr;
    ~LqResources() {
        if (sA && *sA) cudaStreamSynchronize(*sA);
        if (sB && *sB) cudaStreamSynchronize(*sB);
        for (int q = 0; q < 2; ++q) {
            if (evA && evA[q]) cudaEventDestroy(evA[q]);
            if (evB && evB[q]) cudaEventDestroy(evB[q]);
        }
        if (sB && *sB) cudaStreamDestroy(*sB);
        if (sA && *sA) cudaStreamDestroy(*sA);
        if (base && *base) cudaFreeAsync(*base, caller);
    }
};

struct LargeQrStats {
    long sweeps = 0, windows = 0, small_blocks = 0;
};

// H (n x n Hessenberg, zeros below the sub-diagonal), Z (n x n or null) on the device.  Returns 0, or k > 0 when the
// iteration limit is hit with the active block ending at row k.  w: 2n doubles (device).
inline int lg_qr(double* H, double* Z, int n, double* w, cudaStream_t s, std::string* err, LargeQrStats* stats) {
    const size_t wbuf = (size_t)LQ_SMALL * std::max(n, LQ_SMALL);
    double* base = nullptr;
    cudaStream_t sB = nullptr, sA = nullptr;
    cudaEvent_t evChase[2] = {nullptr, nullptr}, evFar[2] = {nullptr, nullptr};
    LqResources res;
    res.caller = s;
    res.sA = &sA;
    res.sB = &sB;
    res.evA = evChase;
    res.evB = evFar;
    res.base = &base;
    // scratch: U (LQ_SMALL^2), tmp (LQ_SMALL x n), blk (LQ_SMALL^2), zblk (LQ_SMALL^2), wblk, shifts, scan
    const size_t total = 5 * (size_t)LQ_SMALL * LQ_SMALL + wbuf + 2 * LQ_SMALL + 4 * LQ_NB * LQ_MAXC + 64 +
                         2 * (size_t)LQ_MAXC * LQ_W * LQ_W;
    LG_TRY(cudaMallocAsync((void**)&base, total * sizeof(double), s));
    double* dU = base;
    double* dBlk = dU + (size_t)LQ_SMALL * LQ_SMALL;
    double* dZb = dBlk + (size_t)LQ_SMALL * LQ_SMALL;
    double* dTmp = dZb + (size_t)LQ_SMALL * LQ_SMALL;
    double* dU1 = dTmp + wbuf;                               // second U buffer (windows alternate)
    double* dTmpA = dU1 + (size_t)LQ_SMALL * LQ_SMALL;       // scratch of the near-window GEMM (stream A)
    double* dWb = dTmpA + (size_t)LQ_SMALL * LQ_SMALL;
    double* dShift = dWb + 2 * LQ_SMALL;
    double* dUall = dShift + 4 * LQ_NB * LQ_MAXC;            // [2 parities][LQ_MAXC chains][LQ_W x LQ_W]
    LqScan* dScan = reinterpret_cast<LqScan*>(dUall + 2 * (size_t)LQ_MAXC * LQ_W * LQ_W);
    int* dInfo = reinterpret_cast<int*>(dScan + 1);
    unsigned long long* dCounter = reinterpret_cast<unsigned long long*>(dScan + 4);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t chase_smem = 2 * (size_t)LQ_W * LQ_LDW * sizeof(double) + 2 * LQ_NB * sizeof(lq_refl) + 16;
    LG_TRY(cudaFuncSetAttribute(lq_chase_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chase_smem));
    const size_t apply_smem = ((size_t)LQ_W * LQ_LDU + std::max(LQ_LDT_L * 64, LQ_LDT_R * LQ_W)) * sizeof(double);
    LG_TRY(cudaFuncSetAttribute(lq_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)apply_smem));
    cudaFuncSetAttribute(lq_apply_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    // every kernel of the QR pipeline asks for the same (maximum) shared-memory carve-out: the chase kernel needs
    // 198 KB, and an SM whose L1/shared split has to be reconfigured between kernels must drain first
    cudaFuncSetAttribute(dgemm_kernel<false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(dgemm_kernel<true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(dgemm_kernel<false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(lq_copy_block_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(lq_chase_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);

    auto run_small = [&](double* blk, int m, double* zb, double* wb, bool wantz) -> int {
        // batched kernel on one m x m Hessenberg block: blk <- T, zb <- Z_blk (in: identity), wb <- eigenvalues
        BatchedParams p{};
        p.A = blk;
        p.Z = wantz ? zb : nullptr;
        p.w = wb;
        p.tau = nullptr;
        p.strideA = (long long)m * m;
        p.strideZ = (long long)m * m;
        p.batch = 1;
        p.lda = m;
        p.ldz = m;
        p.n = m;
        p.scale = 0;
        p.maxiter = 0;
        p.mode = MODE_SCHUR;
        p.flags = F_HESS_INPUT;
        p.info = dInfo;
        p.stats = nullptr;
        p.counter = dCounter;
        p.scratch = nullptr;
        cudaMemsetAsync(dCounter, 0, sizeof(unsigned long long), s);
        std::string e2;
        int rc = launch_f64(p, sms, s, &e2);
        if (rc) *err = e2;
        return rc;
    };
    // pieces of "apply the orthogonal m x m factor Q (ld ldq) that transformed rows/cols lo..hi (1-based)":
    //   left : H[lo:hi, c0:c1] <- Q' H[lo:hi, c0:c1]      top : H[1:lo-1, lo:hi] <- H[...] Q      zz : Z[:, lo:hi] <- Z[...] Q
    auto left_part = [&](cudaStream_t st, const double* Q, int ldq, int lo, int hi, int c0, int c1, double* tmp) -> int {
        const int m = hi - lo + 1, nc = c1 - c0 + 1;
        if (nc <= 0) return 0;
        double* C = H + (size_t)(lo - 1) + (size_t)(c0 - 1) * n;
        LG_TRY(dgemm(st, true, false, m, nc, m, 1.0, Q, ldq, C, n, 0.0, tmp, m));
        lq_copy(st, tmp, m, C, n, m, nc);
        return 0;
    };
    auto top_part = [&](cudaStream_t st, const double* Q, int ldq, int lo, int hi, double* tmp) -> int {
        const int m = hi - lo + 1, nt = lo - 1;
        if (nt <= 0) return 0;
        double* C = H + (size_t)(lo - 1) * n;
        LG_TRY(dgemm(st, false, false, nt, m, m, 1.0, C, n, Q, ldq, 0.0, tmp, nt));
        lq_copy(st, tmp, nt, C, n, nt, m);
        return 0;
    };
    auto z_part = [&](cudaStream_t st, const double* Q, int ldq, int lo, int hi, double* tmp) -> int {
        const int m = hi - lo + 1;
        if (!Z) return 0;
        double* C = Z + (size_t)(lo - 1) * n;
        LG_TRY(dgemm(st, false, false, n, m, m, 1.0, C, n, Q, ldq, 0.0, tmp, n));
        lq_copy(st, tmp, n, C, n, n, m);
        return 0;
    };
    auto apply_outside = [&](const double* Q, int ldq, int lo, int hi) -> int {
        int rc = left_part(s, Q, ldq, lo, hi, hi + 1, n, dTmp);
        if (rc == 0) rc = top_part(s, Q, ldq, lo, hi, dTmp);
        if (rc == 0) rc = z_part(s, Q, ldq, lo, hi, dTmp);
        return rc;
    };
    // second stream: the far updates of window t run while window t+1 is being chased
    // The chase kernel needs a whole SM (198 KB of shared memory): its stream gets the highest priority so that the
    // block scheduler hands it the first SM the concurrently running GEMMs vacate.
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    LG_TRY(cudaStreamCreateWithPriority(&sB, cudaStreamNonBlocking, prio_lo));
    LG_TRY(cudaStreamCreateWithPriority(&sA, cudaStreamNonBlocking, prio_hi));
    LG_TRY(cudaStreamSynchronize(s));      // everything before this call has completed; from here on sA replaces s
    cudaStream_t s_caller = s;
    s = sA;
    for (int q = 0; q < 2; ++q) {
        LG_TRY(cudaEventCreateWithFlags(&evChase[q], cudaEventDisableTiming));
        LG_TRY(cudaEventCreateWithFlags(&evFar[q], cudaEventDisableTiming));
    }
    struct Win {
        int tau0, M;
        int wlo[LQ_MAXC], whi[LQ_MAXC];
    };
    std::vector<Win> wins;

    const bool dbg_time = std::getenv("GSCHUR_LQ_TIMING") != nullptr;
    double t_scan = 0, t_small = 0, t_shift = 0, t_enq = 0;
    auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    int iend = n;
    long sweeps_total = 0;
    int since_deflation = 0;
    long maxsweeps = 30L * n + 100;
    // development knobs (timing experiments only): cap the number of sweeps / skip one side of the pipeline
    const char* dbg_ms = std::getenv("GSCHUR_LQ_MAXSWEEPS");
    if (dbg_ms) maxsweeps = std::atol(dbg_ms);
    const char* dbg_skip = std::getenv("GSCHUR_LQ_SKIP");
    const int skip = dbg_skip ? std::atoi(dbg_skip) : 0;
    std::vector<double> hshift(2 * 2 * LQ_NB * LQ_MAXC), hpairs(4 * LQ_NB * LQ_MAXC);
    int rc_final = 0;
    while (iend >= 1) {
        double t0 = now();
        lq_scan_kernel<<<1, 1024, 0, s>>>(H, n, iend, dScan);
        note_launch();
        LqScan hs;
        LG_TRY(cudaMemcpyAsync(&hs, dScan, sizeof(LqScan), cudaMemcpyDeviceToHost, s));
        LG_TRY(cudaStreamSynchronize(s));
        const int istart = hs.istart;
        const int nw = iend - istart + 1;
        t_scan += now() - t0;
        t0 = now();
        if (nw <= LQ_SMALL) {
            // finish the block with the batched kernel, back-transform with GEMMs
            lq_copy(s, H + (size_t)(istart - 1) + (size_t)(istart - 1) * n, n, dBlk, nw, nw, nw);
            lq_eye_kernel<<<(nw * nw + 255) / 256, 256, 0, s>>>(dZb, nw, nw);
            note_launch();
            int rc = run_small(dBlk, nw, dZb, dWb, true);
            if (rc) return -2;
            int hinfo = 0;
            LG_TRY(cudaMemcpyAsync(&hinfo, dInfo, sizeof(int), cudaMemcpyDeviceToHost, s));
            LG_TRY(cudaStreamSynchronize(s));
            if (hinfo != 0) { rc_final = iend; break; }
            lq_copy(s, dBlk, nw, H + (size_t)(istart - 1) + (size_t)(istart - 1) * n, n, nw, nw);
            if (nw > 1) {
                rc = apply_outside(dZb, nw, istart, iend);
                if (rc) return rc;
            }
            if (stats) stats->small_blocks += 1;
            t_small += now() - t0;
            iend = istart - 1;
            since_deflation = 0;
            continue;
        }
        if (++sweeps_total > maxsweeps) { rc_final = iend; break; }
        since_deflation += 1;
        // ---- chains and shifts: C chains of NB bulges; 2 NB C shifts = eigenvalues of the trailing block ----
        int C = nw / 768;
        C = std::max(1, std::min(C, LQ_MAXC));
        const int ns = 2 * LQ_NB * C;
        const int npairs = LQ_NB * C;
        hshift.resize(2 * ns);
        hpairs.resize(4 * npairs);
        lq_copy(s, H + (size_t)(iend - ns) + (size_t)(iend - ns) * n, n, dBlk, ns, ns, ns);
        if (since_deflation % 6 == 0) {
            // exceptional shifts (in the spirit of src/GenericSchur.jl:614-627): perturb the trailing diagonal
            std::vector<double> hb((size_t)ns * ns);
            LG_TRY(cudaMemcpyAsync(hb.data(), dBlk, sizeof(double) * ns * ns, cudaMemcpyDeviceToHost, s));
            LG_TRY(cudaStreamSynchronize(s));
            for (int j = 0; j < ns; ++j) {
                const double sub = (j > 0) ? fabs(hb[j + (size_t)(j - 1) * ns]) : fabs(hb[1]);
                hshift[2 * j] = hb[j + (size_t)j * ns] + 0.75 * sub;
                hshift[2 * j + 1] = 0.0;
            }
        } else {
            int rc = run_small(dBlk, ns, nullptr, dWb, false);
            if (rc) return -2;
            LG_TRY(cudaMemcpyAsync(hshift.data(), dWb, sizeof(double) * 2 * ns, cudaMemcpyDeviceToHost, s));
            LG_TRY(cudaStreamSynchronize(s));
        }
        // pair the shifts: conjugate pairs stay together, real ones are paired up (a leftover real is doubled)
        {
            std::vector<std::complex<double>> cp, re;
            for (int j = 0; j < ns; ++j) {
                std::complex<double> z(hshift[2 * j], hshift[2 * j + 1]);
                if (z.imag() != 0.0) cp.push_back(z);
                else re.push_back(z);
            }
            int np = 0;
            for (size_t j = 0; j + 1 < cp.size() && np < npairs; j += 2, ++np) {
                std::complex<double> a = cp[j];
                hpairs[4 * np] = a.real();
                hpairs[4 * np + 1] = fabs(a.imag());
                hpairs[4 * np + 2] = a.real();
                hpairs[4 * np + 3] = -fabs(a.imag());
            }
            std::sort(re.begin(), re.end(), [](const std::complex<double>& x, const std::complex<double>& y) { return x.real() < y.real(); });
            for (size_t j = 0; j < re.size() && np < npairs; j += 2, ++np) {
                const double a = re[j].real(), b2 = (j + 1 < re.size()) ? re[j + 1].real() : re[j].real();
                hpairs[4 * np] = a;
                hpairs[4 * np + 1] = 0.0;
                hpairs[4 * np + 2] = b2;
                hpairs[4 * np + 3] = 0.0;
            }
            for (; np < npairs; ++np)
                for (int q = 0; q < 4; ++q) hpairs[4 * np + q] = hpairs[4 * (np - 1) + q];
        }
        LG_TRY(cudaMemcpyAsync(dShift, hpairs.data(), sizeof(double) * 4 * npairs, cudaMemcpyHostToDevice, s));
        t_shift += now() - t0;
        t0 = now();
        // ---- chase the chains from the top (L = istart) to the bottom (I = iend) ----
        const int L = istart, I = iend;
        const int tau_end = (I - 1 - L) + 4 * (LQ_NB - 1) + (C - 1) * LQ_DELTA;   // last time step with an active bulge
        const int Mfull = LQ_W - (4 * (LQ_NB - 1) + 6);
        wins.clear();
        for (int tau0 = 0; tau0 <= tau_end; tau0 += Mfull) {
            Win wn{};
            wn.tau0 = tau0;
            wn.M = std::min(Mfull, tau_end - tau0 + 1);
            bool any = false;
            for (int c = 0; c < LQ_MAXC; ++c) {
                wn.wlo[c] = 1;
                wn.whi[c] = 0;
                if (c >= C) continue;
                int pmin = 1 << 30, pmax = -1;
                for (int b = 0; b < LQ_NB; ++b) {
                    const int p0 = L + tau0 - 4 * b - c * LQ_DELTA, p1 = p0 + wn.M - 1;
                    if (p1 < L || p0 > I - 1) continue;
                    pmin = std::min(pmin, std::max(p0, L));
                    pmax = std::max(pmax, std::min(p1, I - 1));
                }
                if (pmax < 0) continue;
                wn.wlo[c] = std::max(L, pmin - 1);
                wn.whi[c] = std::min(I, pmax + 3);
                if (wn.whi[c] - wn.wlo[c] + 1 > LQ_W) { *err = "internal: chase window too large"; return -2; }
                any = true;
            }
            if (any) wins.push_back(wn);
        }
        const int nwin = (int)wins.size();
        for (int t = 0; t < nwin; ++t) {
            const Win& wn = wins[t];
            LqChaseArgs ca{};
            LqApplyArgs near{}, farl{}, topz{};
            near.C = farl.C = topz.C = C;
            near.mode = 1;
            farl.mode = 1;
            topz.mode = 2 | 4;
            int gnear = 0, gfarl = 0, gtopz = 0;
            for (int c = 0; c < LQ_MAXC; ++c) {
                double* U = dUall + ((size_t)(t & 1) * LQ_MAXC + c) * LQ_W * LQ_W;
                ca.wlo[c] = wn.wlo[c];
                ca.whi[c] = wn.whi[c];
                ca.U[c] = U;
                const int wsz = wn.whi[c] - wn.wlo[c] + 1;
                near.wlo[c] = farl.wlo[c] = topz.wlo[c] = wn.wlo[c];
                near.wsz[c] = farl.wsz[c] = topz.wsz[c] = (c < C) ? wsz : 0;
                near.U[c] = farl.U[c] = topz.U[c] = U;
                if (c >= C || wsz <= 0) continue;
                int nearhi = wn.whi[c];
                if (t + 1 < nwin && wins[t + 1].whi[c] >= wins[t + 1].wlo[c]) nearhi = std::min(n, std::max(nearhi, wins[t + 1].whi[c]));
                near.lc0[c] = wn.whi[c] + 1;
                near.lc1[c] = nearhi;
                farl.lc0[c] = nearhi + 1;
                farl.lc1[c] = n;
                gnear += lq_strips(near.lc1[c] - near.lc0[c] + 1);
                gfarl += lq_strips(farl.lc1[c] - farl.lc0[c] + 1);
                gtopz += lq_strips(wn.wlo[c] - 1) + (Z ? lq_strips(n) : 0);
            }
            // stream A: chase all chains (one CTA each), then the near columns the next windows are going to need
            if (!(skip & 2)) {
                lq_chase_kernel<<<C, 32 * (LQ_NB + LQ_UW), chase_smem, s>>>(H, n, L, I, dShift, wn.tau0, wn.M, ca);
                note_launch();
            }
            LG_TRY(cudaEventRecord(evChase[t & 1], s));
            if (t >= 1) LG_TRY(cudaStreamWaitEvent(s, evFar[(t - 1) & 1], 0));
            if (gnear > 0) {
                lq_apply_kernel<<<gnear, 256, apply_smem, s>>>(H, Z, n, near);
                note_launch();
            }
            // stream B: everything else outside the windows (left strips first, then the right strips: they overlap)
            LG_TRY(cudaStreamWaitEvent(sB, evChase[t & 1], 0));
            if (!(skip & 1)) {
                if (gfarl > 0) {
                    lq_apply_kernel<<<gfarl, 256, apply_smem, sB>>>(H, Z, n, farl);
                    note_launch();
                }
                if (gtopz > 0) {
                    lq_apply_kernel<<<gtopz, 256, apply_smem, sB>>>(H, Z, n, topz);
                    note_launch();
                }
            }
            LG_TRY(cudaEventRecord(evFar[t & 1], sB));
            if (stats) stats->windows += 1;
        }
        if (nwin > 0) LG_TRY(cudaStreamWaitEvent(s, evFar[(nwin - 1) & 1], 0));
        if (nwin > 1) LG_TRY(cudaStreamWaitEvent(s, evFar[(nwin - 2) & 1], 0));
        LG_TRY(cudaGetLastError());
        t_enq += now() - t0;
        if (stats) stats->sweeps += 1;
    }
    if (rc_final == 0 && w) {
        lq_eigs_kernel<<<(n + 255) / 256, 256, 0, s>>>(H, n, w);
        note_launch();
    }
    if (dbg_time) fprintf(stderr, "[lg_qr] scan+sync %.3f s, small blocks %.3f s, shifts %.3f s, window enqueue %.3f s\n", t_scan, t_small, t_shift, t_enq);
    LG_TRY(cudaStreamSynchronize(s));
    LG_TRY(cudaStreamSynchronize(sB));
    s = s_caller;
    return rc_final;   // `res` destroys the streams and events and returns the scratch to the pool
}

}  // namespace gs

/*
 * libgschur_cuda — C ABI of the B200 (sm_100a) Schur hot path.
 *
 * Drop-in boundary for RalphAS/GenericSchur.jl's `gschur!` / `schur!` / `eigvals!` / `hessenberg!`:
 * the reference has no FFI seam (it is pure Julia), so these entry points are what a Julia `ccall`
 * shim binds in place of the Julia methods cited at each declaration (paths relative to the reference
 * tree).  INTEGRATION.md shows the shim.  Plain pointers and sizes only; no torch / CUDA types.
 *
 * Element kinds (column-major matrices, Julia `Matrix` layout):
 *   GSCHUR_F64  Float64                      8 B / element
 *   GSCHUR_C64  ComplexF64 (re, im)         16 B
 *   GSCHUR_DD   double-double (hi, lo)      16 B   (DoubleFloats.Double64 / MultiFloats.Float64x2 layout)
 *   GSCHUR_CDD  Complex{double-double} (re.hi, re.lo, im.hi, im.lo)  32 B
 * Eigenvalues `w` are always complex: 2 doubles per value for F64/C64, 4 doubles for DD/CDD.
 *
 * Return convention: 0 = success; > 0 = number of matrices that did not converge (their info[] > 0; the
 * shim throws UnconvergedException("iteration limit $maxiter reached"), src/GenericSchur.jl:235,555);
 * < 0 = argument or CUDA error (gschur_cuda_last_error() has the text; the shim throws
 * DimensionMismatch / ArgumentError / ErrorException).
 *
 * There is no CPU fallback: every entry point fails with GSCHUR_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef GSCHUR_CUDA_H
#define GSCHUR_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSCHUR_F64 0
#define GSCHUR_C64 1
#define GSCHUR_DD 2
#define GSCHUR_CDD 3

/* flags */
#define GSCHUR_FLAG_DEVICE_PTRS 0x1u /* A, Z, w, info, stats are device pointers on the current device      */
#define GSCHUR_FLAG_HESS_INPUT 0x2u  /* A already upper Hessenberg: skip the reduction (gschur!(H::Hessenberg, Z)) */
#define GSCHUR_FLAG_CHECK_SUBDIAG 0x4u /* with HESS_INPUT, complex kinds: reject a non-real sub-diagonal (checksd) */

/* error codes (negative returns) */
#define GSCHUR_ERR_ARG (-1)      /* bad kind / n / leading dimension / stride / NULL pointer */
#define GSCHUR_ERR_CUDA (-2)     /* CUDA runtime failure or no usable device */
#define GSCHUR_ERR_SIZE (-3)     /* n not supported by any kernel for this kind */
#define GSCHUR_ERR_SUBDIAG (-4)  /* "algorithm assumes real subdiagonal" (src/GenericSchur.jl:206-210) */

#define GSCHUR_STATS_PER_MATRIX 4 /* sweeps, reflector applications, exceptional shifts, iterations */

int gschur_cuda_version(void);
int gschur_cuda_device_count(void);
/* thread-local text of the last error on this thread ("" if none) */
const char* gschur_cuda_last_error(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
uint64_t gschur_cuda_launch_count(void);
/* largest n the batched (one matrix per CTA / cluster) kernels take for `kind`; 0 if kind is invalid */
int gschur_cuda_max_batched_n(int kind);

/*
 * Batched Schur decomposition: for b in [0, batch): A_b = Z_b T_b Z_b^H.
 * Replaces gschur!(A::StridedMatrix{Complex{T}}; wantZ, scale)  src/GenericSchur.jl:350-372
 *      and gschur!(A::StridedMatrix{T<:AbstractFloat}; wantZ, scale) src/GenericSchur.jl:805-835
 * (and through them schur! src/pirates.jl:8-10 and eigvals! src/pirates.jl:17-27), one call per batch.
 *
 *   A      in: A_b (destroyed).  out: T_b — upper triangular (complex kinds) or quasi-upper-triangular in
 *          standard form (real kinds); entries below are written as exact zeros.
 *   lda, strideA   leading dimension (>= n) and distance between consecutive matrices, in elements.
 *   Z      out: Schur vectors, or NULL for wantZ = false (the eigvals! path).
 *   w      out: n*batch complex eigenvalues, w[b*n + j] pairs with T_b[j,j].
 *   scale  non-zero: apply _scale! (src/util.jl:14-29) before and undo it after, as the reference does.
 *   maxiter <= 0 selects the reference default 100*n.
 *   info   per matrix (NULL ok): 0 converged; k > 0: iteration limit hit with the active block ending at row k.
 *   stats  NULL or GSCHUR_STATS_PER_MATRIX uint32 per matrix.
 *   devices/ndev  host-pointer mode only: the batch is split in contiguous slices over these CUDA devices
 *          (NULL / 0 = device 0).  Ignored with GSCHUR_FLAG_DEVICE_PTRS.
 */
int gschur_cuda_batched(int kind, int n, int64_t batch,
                        void* A, int lda, int64_t strideA,
                        void* Z, int ldz, int64_t strideZ,
                        void* w, int scale, int maxiter,
                        int32_t* info, uint32_t* stats,
                        const int* devices, int ndev, uint32_t flags);

/*
 * Same, device-resident and asynchronous: all pointers are device pointers on the current device, the work is
 * enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream) and the call returns
 * without synchronising.  Return value is 0 or a negative error; convergence is reported through info[] only.
 */
int gschur_cuda_batched_async(int kind, int n, int64_t batch,
                              void* A, int lda, int64_t strideA,
                              void* Z, int ldz, int64_t strideZ,
                              void* w, int scale, int maxiter,
                              int32_t* info, uint32_t* stats,
                              void* stream, uint32_t flags);

/*
 * Measures the FP64 FMA peak of the current device with a DFMA-only micro-kernel (the roofline denominator of
 * the compute-bound batched kernels; MEASURED_PEAKS.json has no FP64 figure).  *tflops receives 2*FMA/s / 1e12,
 * *ms (NULL ok) the best kernel time.
 */
int gschur_cuda_measure_fp64_peak(double* tflops, double* ms);

/*
 * Measures the FP64 tensor-core (DMMA, mma.sync.m8n8k4.f64) peak of the current device with register-resident
 * operands: the BLAS3 denominator of the large-matrix path's GEMM updates.  *tflops, *ms (NULL ok) as above.
 */
int gschur_cuda_measure_dmma_peak(double* tflops, double* ms);

/*
 * Measures the L2 streaming bandwidth of the current device: a 38.8 MB buffer (resident in the 126 MB L2) is read
 * and written back 64 times with L1-bypassing 16-byte accesses, the access pattern of stage B's Z stream (DESIGN.md
 * section 6: the L2 ceiling behind the FP64 one).  *gbs receives (bytes read + bytes written) / s / 1e9, *ms (NULL
 * ok) the best kernel time.
 */
int gschur_cuda_measure_l2_bandwidth(double* gbs, double* ms);

/*
 * Per-kernel timing of the two-kernel batched path (stage A: scale + Hessenberg + Q; stage B: QR iteration), used by
 * bench.py for the per-kernel roofline.  enable: 1 / 0 switches CUDA-event recording on the launching stream on / off,
 * -1 leaves it unchanged; when both pointers are given the times (ms) of the most recent call are returned.
 */
int gschur_cuda_stage_timing(int enable, float* ms_stage_a, float* ms_stage_b);
/*
 * Three-value form for the three-stage path (n <= 64, Float64 / ComplexF64): stage A as above, stage B = QR iteration
 * on H with the reflector log, stage C = replay of the log on the Schur vectors (+ the redo pass, normally empty).
 * Times are summed over the sub-batches of the call.  The two-value form reports B + C as its stage B.
 */
int gschur_cuda_stage_timing3(int enable, float* ms_stage_a, float* ms_stage_b, float* ms_stage_c);

/*
 * Frees everything the library caches between calls: the per-device chunk buffers and pinned staging block of the
 * host-pointer pipeline, and the stream-ordered workspace pool of the batched kernels (scale records, reflector log).
 * The next call allocates again.  Returns 0.
 */
int gschur_cuda_release_workspace(void);

/*
 * Batched Householder reduction to Hessenberg form, A_b = Q_b H_b Q_b^H.
 * Replaces _hessenberg!(A) src/hessenberg.jl:3-17 (LinearAlgebra.hessenberg! src/pirates.jl:232) and
 * _materializeQ(H) src/hessenberg.jl:150-166.
 *   A    in: A_b; out: the factors exactly as the reference leaves them — H on and above the sub-diagonal,
 *        reflector tails below it (complex kinds: the sub-diagonal is real).
 *   tau  out: (n-1)*batch reflector scalars (element type of `kind`), stride n-1.
 *   Q    out: explicit Q_b, or NULL.
 */
int gschur_cuda_hessenberg_batched(int kind, int n, int64_t batch,
                                   void* A, int lda, int64_t strideA,
                                   void* tau,
                                   void* Q, int ldq, int64_t strideQ,
                                   const int* devices, int ndev, uint32_t flags);

/*
 * Eigenvectors from the Schur form, batched, ComplexF64 (kind 1) only.
 * Replaces geigvecs(S; left) = _geigvecs!(S.T, S.Z) / _gleigvecs!(S.T, S.Z) + _enormalize!
 * (src/vectors.jl:12-20, 45-131, 372-460; src/util.jl:128-461, 572-592) — the vectors eigen! returns
 * (src/pirates.jl:63-90) — for the T, Z that gschur_cuda_batched left on the device or the host.
 *   T    in: upper triangular Schur forms (not modified)      Z  in: Schur vectors (NULL: eigenvectors of T itself)
 *   V    out: n x n per matrix, column k = eigenvector of T[k,k]; unit 2-norm with the largest component real
 *        (flag 0x10: skip _enormalize!, i.e. the raw _geigvecs! scaling with max abs1 component 1)
 *   left != 0: left eigenvectors.   flags: GSCHUR_FLAG_DEVICE_PTRS.
 */
int gschur_cuda_eigvecs_batched(int kind, int n, int64_t batch, const void* T, int ldt, int64_t strideT, const void* Z,
                                int ldz, int64_t strideZ, void* V, int ldv, int64_t strideV, int left, uint32_t flags);
const char* gschur_cuda_eigvecs_last_error(void);

/*
 * Balancing pre-step and its back-transformation, triangularize post-step, batched,
 * Float64 / ComplexF64 (kinds 0, 1).  flags: GSCHUR_FLAG_DEVICE_PTRS.  Errors: gschur_cuda_balance_last_error().
 *
 * gschur_cuda_balance_batched       balance!(A; scale, permute) src/balance.jl:33-199: A is overwritten by the balanced
 *     matrix; per matrix D (n doubles, the diagonal similarity), ilo_ihi_trivial (3 ints: ilo, ihi, trivial flag of the
 *     Balancer) and perm (n ints: the exchange targets `sp`, 1-based, 0 where unused; prow = perm[0:ilo-1],
 *     pcol = perm[ihi:n]).  info (NULL ok): 0, or -5 for the reference's error("NaN encountered while balancing").
 * gschur_cuda_balance_apply_batched lmul!(B, V) (inverse = 0: right eigenvectors of the balanced matrix -> those of
 *     A) / ldiv!(B, V) (inverse = 1: left eigenvectors) src/balance.jl:203-260, V n x n per matrix, in place.
 * gschur_cuda_triangularize_batched triangularize(S::Schur{<:Real}) src/triang.jl:9-43: the complex upper triangular
 *     Schur form (Tc, Zc: n x n ComplexF64 each, leading dimension n; w: n eigenvalues) of a standardised real one.
 */
int gschur_cuda_balance_batched(int kind, int n, int64_t batch, void* A, int lda, int64_t strideA, double* D,
                                int32_t* ilo_ihi_trivial, int32_t* perm, int32_t* info, int scale, int permute, uint32_t flags);
int gschur_cuda_balance_apply_batched(int kind, int n, int64_t batch, void* V, int ldv, int64_t strideV, const double* D,
                                      const int32_t* ilo_ihi_trivial, const int32_t* perm, int inverse, uint32_t flags);
int gschur_cuda_triangularize_batched(int n, int64_t batch, const double* T, int ldt, int64_t strideT, const double* Z, int ldz,
                                      int64_t strideZ, void* Tc, void* Zc, void* w, uint32_t flags);
const char* gschur_cuda_balance_last_error(void);

/* ------------------------------------------------------------------------------------------------------------
 * Regime (2): ONE large Float64 matrix on one GPU (BASELINE config 4).  Host or device pointers
 * (GSCHUR_FLAG_DEVICE_PTRS), blocking.
 * ---------------------------------------------------------------------------------------------------------- */

/*
 * Blocked (compact-WY, DMMA-GEMM) Householder reduction A = Q H Q' of one n x n Float64 matrix.
 * Replaces _hessenberg!(A) src/hessenberg.jl:3-17 + _materializeQ src/hessenberg.jl:150-166 for large n.
 * A out: H on/above the sub-diagonal, reflector tails below; tau: n-1 (NULL ok); Q: n x n (NULL ok).
 */
int gschur_cuda_hessenberg_large(int n, double* A, int lda, double* tau, double* Q, int ldq, uint32_t flags);

/*
 * Schur decomposition of ONE large n x n Float64 matrix: blocked Hessenberg + Q, then small-bulge multishift QR
 * (chains of the reference's 3x3 double-shift bulges chased through diagonal windows, reflectors accumulated and
 * applied to the rest of H and to Z as DMMA GEMMs), active blocks of order <= 128 finished by the batched kernel.
 * Replaces gschur!(A::StridedMatrix{Float64}; wantZ, scale) src/GenericSchur.jl:805-835 for large n.
 *   A in/out: matrix -> quasi-triangular T (standard-form 2x2 blocks); Z out (NULL = wantZ false); w: n complex.
 *   info (NULL ok): 0 or k > 0 (iteration limit, active block ends at row k); stats3 (NULL ok): sweeps, windows,
 *   small blocks.  Returns 0, 1 (not converged) or a negative error.
 */
int gschur_cuda_large(int n, double* A, int lda, double* Z, int ldz, double* w, int scale, int* info,
                      long long* stats3, uint32_t flags);

/* the library's FP64 tensor-core (DMMA) GEMM, C = alpha op(A) op(B) + beta C, device pointers; exposed for tests */
int gschur_cuda_dgemm(int ta, int tb, int M, int N, int K, double alpha, const double* A, int lda,
                      const double* B, int ldb, double beta, double* C, int ldc);

/* text of the last error of the large-matrix entry points on this thread */
const char* gschur_cuda_large_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* GSCHUR_CUDA_H */

// Stage A of the two-kernel path for n <= 64: scale -> Householder Hessenberg reduction -> explicit Q.
//   out:  A_b <- H_b (upper Hessenberg, exact zeros below the sub-diagonal; complex: real sub-diagonal)
//         Z_b <- Q_b (if wanted)         scratch[8 b ..] <- (scaled?, cscale, anrm)
// One CTA per matrix, the matrix in shared memory with an odd leading dimension; reflector applications are
// thread-per-column (left) / thread-per-row (right) as in batched.cuh.  Q is generated in place over the stored
// reflectors (the LAPACK xORGHR/xORG2R scheme: shift the reflector columns right by one, accumulate backwards) so
// that the stage needs one n x n tile only and several CTAs fit per SM.  It computes what
// _materializeQ (src/hessenberg.jl:150-166) computes: Q = H_1 H_2 ... H_{n-1} applied to the identity.
// ComplexF64 tiles are staged by per-column TMA bulk copies (16-byte aligned columns); Float64 tiles by plain
// coalesced loads (an odd leading dimension of 8-byte elements cannot be a TMA destination).
#pragma once
#include "batched.cuh"

namespace gs {

template <class T> struct gehrd_smem_layout {
    typedef typename etraits<T>::real R;
    typedef smem_layout<T> L;
    __host__ __device__ static size_t off_tau(int n) { return L::up16((size_t)n * L::ld(n) * sizeof(T)); }
    __host__ __device__ static size_t off_red(int n) { return off_tau(n) + L::up16((size_t)n * sizeof(T)); }
    __host__ __device__ static size_t off_mbar(int n) { return off_red(n) + L::up16(32 * sizeof(R)); }
    __host__ __device__ static size_t bytes(int n) { return off_mbar(n) + 16; }
    // global-tile variant: only tau, the reduction scratch and the (unused) mbarrier live in shared memory
    __host__ __device__ static size_t up16tau(int n) { return L::up16((size_t)n * sizeof(T)); }
    __host__ __device__ static size_t bytes_gt(int n) { return up16tau(n) + L::up16(32 * sizeof(R)) + 16; }
};

// GT = true: the n x n tile does not fit in shared memory (ComplexF64 n > 118, complex double-double n > 83):
// the tile lives in global memory instead — in the caller's Z buffer (or in scratch when Z is not wanted), which
// is where Q has to end up anyway.  Same code, generic addressing; element-per-thread accesses of 16/32-byte
// elements use whole 32-byte sectors, so the strided (row) accesses cost no extra L2 traffic.
template <class T, int NT, bool GT>
__global__ void __launch_bounds__(NT) gehrd_q_kernel(BatchedParams p) {
    typedef typename etraits<T>::real R;
    constexpr bool CPLX = etraits<T>::is_complex;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typedef gehrd_smem_layout<T> GL;
    const int n = p.n;
    const int ld = GT ? (p.Z ? p.ldz : n) : smem_layout<T>::ld(n);
    const int tid = threadIdx.x;

    BatchedSolver<T, NT> S;
    S.n = n;
    S.ld = ld;
    S.tid = tid;
    S.lane = tid & 31;
    S.H = reinterpret_cast<T*>(smem_raw);
    S.Z = nullptr;
    S.sTau = reinterpret_cast<T*>(smem_raw + (GT ? 0 : GL::off_tau(n)));
    S.sW = nullptr;
    S.sRed = reinterpret_cast<R*>(smem_raw + (GT ? GL::up16tau(n) : GL::off_red(n)));
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw + (GT ? GL::up16tau(n) + GL::L::up16(32 * sizeof(R)) : GL::off_mbar(n)));
    __shared__ long long s_next;
    const bool wantZ = (p.Z != nullptr);
    T* H = S.H;
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
        if (GT) {
            // tile = this matrix's Z buffer (or its slice of the scratch tile array when Z is not wanted)
            H = wantZ ? gZ : reinterpret_cast<T*>(p.gt_tiles) + (size_t)blockIdx.x * n * n;
            S.H = H;
        }

        // ---- stage the tile ----
        const bool use_tma = !GT && (sizeof(T) % 16 == 0) && ((reinterpret_cast<uintptr_t>(gA) & 15) == 0) &&
                             (((size_t)p.lda * sizeof(T)) % 16 == 0);
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
        if (p.scale) scaled = S.scale_in(cscale, anrm);
        S.hessenberg();
        // ---- H out: the upper Hessenberg part with zeros below it (Schur requests), or the factored form — H above, the
        //      reflector tails below the sub-diagonal, tau separately — that hessenberg! returns (src/hessenberg.jl:3-17) ----
        const bool factors = p.mode == MODE_HESSENBERG;
        for (int e = tid; e < n * n; e += NT) {
            int i = e % n, j = e / n;
            gA[i + (size_t)j * p.lda] = (factors || i <= j + 1) ? H[i + (size_t)j * ld] : e_zero<T>();
        }
        if (factors && p.tau) {
            T* gtau = reinterpret_cast<T*>(p.tau) + b * (long long)(n > 1 ? n - 1 : 0);
            for (int e = tid; e < n - 1; e += NT) gtau[e] = S.sTau[e];
        }
        if (tid == 0 && p.scratch) {
            double* sc = p.scratch + 8 * b;
            sc[0] = scaled ? 1.0 : 0.0;
            if constexpr (rtraits<R>::ndoubles == 1) {
                sc[1] = cscale;
                sc[2] = 0.0;
                sc[3] = anrm;
                sc[4] = 0.0;
            } else {
                sc[1] = cscale.hi;
                sc[2] = cscale.lo;
                sc[3] = anrm.hi;
                sc[4] = anrm.lo;
            }
        }
        if (wantZ) {
            __syncthreads();
            // shift the reflector tails one column to the right (row r holds tails in columns 1..r-2)
            for (int r = 3 + tid; r <= n; r += NT)
                for (int j = r - 1; j >= 2; --j) AA(r, j) = AA(r, j - 1);
            __syncthreads();
            // first row and column of Q; clear what is on or above the diagonal of the trailing block
            for (int e = tid; e < n * n; e += NT) {
                int i = e % n + 1, j = e / n + 1;
                if (j == 1) AA(i, 1) = (i == 1) ? e_one<T>() : e_zero<T>();
                else if (i <= j) AA(i, j) = e_zero<T>();
            }
            __syncthreads();
            // backward accumulation on B = A(2:n, 2:n): reflector i has its tail in B(i+1:m, i), i.e. A(i+2:n, i+1)
            const int m = n - 1;
            for (int i = m; i >= 1; --i) {
                const T taui = S.sTau[i - 1];
                for (int j = i + 1 + tid; j <= m; j += NT) {     // column j of B = column j+1 of A
                    T w = AA(i + 1, j + 1);
                    for (int r = i + 1; r <= m; ++r) w = w + cconj(AA(r + 1, i + 1)) * AA(r + 1, j + 1);
                    w = taui * w;
                    AA(i + 1, j + 1) = AA(i + 1, j + 1) - w;
                    for (int r = i + 1; r <= m; ++r) AA(r + 1, j + 1) = AA(r + 1, j + 1) - AA(r + 1, i + 1) * w;
                }
                __syncthreads();
                for (int r = i + 1 + tid; r <= m; r += NT) AA(r + 1, i + 1) = -(taui * AA(r + 1, i + 1));
                if (tid == 0) AA(i + 1, i + 1) = e_one<T>() - taui;
                __syncthreads();
            }
            if (!GT)
                for (int e = tid; e < n * n; e += NT) {
                    int i = e % n, j = e / n;
                    gZ[i + (size_t)j * p.ldz] = H[i + j * ld];
                }
        }
        __syncthreads();
    }
#undef AA
}

template <class T, int NT, bool GT> int launch_gehrd_v(const BatchedParams& p, int dev_sms, cudaStream_t stream, std::string* err) {
    auto kern = gehrd_q_kernel<T, NT, GT>;
    size_t smem = GT ? gehrd_smem_layout<T>::bytes_gt(p.n) : gehrd_smem_layout<T>::bytes(p.n);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int per_sm = 0;
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem);
    if (e != cudaSuccess) {
        *err = std::string("gehrd kernel setup: ") + cudaGetErrorString(e);
        return -2;
    }
    if (per_sm < 1) {
        *err = "gehrd kernel does not fit on an SM";
        return -3;
    }
    long long grid = (long long)per_sm * dev_sms;
    if (grid > p.batch) grid = p.batch;
    BatchedParams q = p;
    T* tiles = nullptr;
    if (GT && !p.Z) {
        // no Z buffer to work in: one scratch tile per resident CTA
        e = cudaMallocAsync((void**)&tiles, (size_t)grid * p.n * p.n * sizeof(T), stream);
        if (e != cudaSuccess) {
            *err = std::string("cudaMallocAsync(tiles): ") + cudaGetErrorString(e);
            return -2;
        }
        q.gt_tiles = tiles;
    }
    kern<<<(unsigned)grid, NT, smem, stream>>>(q);
    note_launch();
    e = cudaGetLastError();
    if (tiles) cudaFreeAsync(tiles, stream);
    if (e != cudaSuccess) {
        *err = std::string("gehrd kernel launch: ") + cudaGetErrorString(e);
        return -2;
    }
    return 0;
}
template <class T, int NT> int launch_gehrd(const BatchedParams& p, int dev_sms, cudaStream_t stream, std::string* err) {
    if (gehrd_smem_layout<T>::bytes(p.n) <= 232448) return launch_gehrd_v<T, NT, false>(p, dev_sms, stream, err);
    return launch_gehrd_v<T, NT, true>(p, dev_sms, stream, err);
}

}  // namespace gs

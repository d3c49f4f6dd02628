// Host-side launch interface between the C ABI (gschur_api.cu) and the per-kind kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <string>

namespace gs {

struct BatchedParams {
    void* A;
    void* Z;      // nullptr: wantZ = false
    void* w;      // complex eigenvalues, n per matrix
    void* tau;    // hessenberg-only mode: (n-1) per matrix
    long long strideA, strideZ, batch;
    int lda, ldz, n;
    int scale, maxiter;
    int mode;     // 0: Schur; 1: Hessenberg factors (+Q in Z) only
    unsigned flags;
    int* info;
    unsigned* stats;
    unsigned long long* counter;
    double* scratch;   // two-kernel path: 8 doubles per matrix (scaled flag, cscale, anrm) from stage A to stage B
};

enum { MODE_SCHUR = 0, MODE_HESSENBERG = 1 };
enum { F_HESS_INPUT = 0x2u, F_CHECK_SUBDIAG = 0x4u, F_FIXED_ROLES = 0x100u };

// every kernel launch of the library is counted (gschur_cuda_launch_count)
void note_launch();
unsigned long long launch_counter();

// optional per-stage timing of the two-kernel path (bench.py's per-kernel roofline): CUDA events on the launching stream
void stage_timing_enable(bool on);
bool stage_timing_enabled();
void stage_timing_mark(int which, cudaStream_t s);      // which = 0: before stage A, 1: between, 2: after stage B
int stage_timing_read(float* ms_a, float* ms_b);        // synchronises on the last events; 0 = ok

// dynamic shared memory one CTA needs for an n x n matrix of `kind`
size_t batched_smem_bytes(int kind, int n);

// each returns 0 or a negative GSCHUR_ERR_* code with *err filled
int launch_f64(const BatchedParams& p, int dev_sms, cudaStream_t stream, std::string* err);
int launch_c64(const BatchedParams& p, int dev_sms, cudaStream_t stream, std::string* err);
int launch_dd(const BatchedParams& p, int dev_sms, cudaStream_t stream, std::string* err);
int launch_cdd(const BatchedParams& p, int dev_sms, cudaStream_t stream, std::string* err);

}  // namespace gs

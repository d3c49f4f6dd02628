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
    void* gt_tiles;   // stage A, global-tile variant without a Z buffer: one n x n scratch tile per resident CTA
    long long strideA, strideZ, batch;
    int lda, ldz, n;
    int scale, maxiter;
    int mode;     // 0: Schur; 1: Hessenberg factors (+Q in Z) only
    unsigned flags;
    int* info;
    unsigned* stats;
    unsigned long long* counter;
    double* scratch;   // two-kernel path: 8 doubles per matrix (scaled flag, cscale, anrm) from stage A to stage B
    // ---- reflector log of the three-stage path (stage B writes it, stage C replays it on Z; see qrlog.cuh) ----
    unsigned char* log_pool;      // log_pages pages of LOG_PAGE_REC records
    unsigned* log_next;           // page allocator (atomic)
    unsigned log_pages;
    int* log_table;               // per matrix: [0] record count, [1] status (0 ok, 1 log overflow), [2 ..] page ids
    int log_maxp;                 // page ids per matrix
    // matrices whose log overflowed are appended here and redone by the fused kernel (H and Q are left untouched)
    long long* redo_list;
    unsigned* redo_count;
    // when set, work item i of the queue is matrix list[i], i < *list_count (the redo pass)
    const long long* list;
    const unsigned* list_count;
};

enum { MODE_SCHUR = 0, MODE_HESSENBERG = 1 };
enum { F_HESS_INPUT = 0x2u, F_CHECK_SUBDIAG = 0x4u, F_FIXED_ROLES = 0x100u };
enum { LOG_PAGE_REC = 128 };   // records per log page (a record is 4 reals: 32 B for the Float64 kinds)

// every kernel launch of the library is counted (gschur_cuda_launch_count)
void note_launch();
unsigned long long launch_counter();

// optional per-stage timing of the two-kernel path (bench.py's per-kernel roofline): CUDA events on the launching stream
void stage_timing_enable(bool on);
bool stage_timing_enabled();
void stage_timing_begin_call();                          // forget the marks of the previous call
void stage_timing_mark(int which, cudaStream_t s);      // which = 0: before stage A, 1: after A, 2: after B, 3: after C (+ redo)
int stage_timing_read(float* ms_a, float* ms_b, float* ms_c);   // synchronises on the last events; sums over sub-batches; 0 = ok

// dynamic shared memory one CTA needs for an n x n matrix of `kind`
size_t batched_smem_bytes(int kind, int n);

// each returns 0 or a negative GSCHUR_ERR_* code with *err filled
int launch_f64(const BatchedParams& p, int dev_sms, cudaStream_t stream, std::string* err);
int launch_c64(const BatchedParams& p, int dev_sms, cudaStream_t stream, std::string* err);
int launch_dd(const BatchedParams& p, int dev_sms, cudaStream_t stream, std::string* err);
int launch_cdd(const BatchedParams& p, int dev_sms, cudaStream_t stream, std::string* err);

}  // namespace gs

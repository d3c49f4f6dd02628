// libgschur_cuda: C ABI (include/gschur_cuda.h) and host-side orchestration.
//  - device-pointer entry: one persistent kernel launch per call on the caller's stream;
//  - host-pointer entry: the batch is split in contiguous slices over the requested devices (one host
//    thread per device, no collective), each slice is pipelined in chunks over three streams so that
//    H2D, compute and D2H overlap.
// There is no CPU fallback anywhere in this file.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gschur_cuda.h"
#include "launch.h"

using namespace gs;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return fail(GSCHUR_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));      \
    } while (0)

size_t elem_size(int kind) { return kind == GSCHUR_F64 ? 8 : (kind == GSCHUR_CDD ? 32 : 16); }
size_t eig_size(int kind) { return kind >= GSCHUR_DD ? 32 : 16; }

constexpr size_t kMaxSmem = 232448;   // 227 KiB opt-in limit per CTA on sm_100

int max_batched_n(int kind) {
    static int cache[4] = {0, 0, 0, 0};
    if (kind < 0 || kind > 3) return 0;
    if (!cache[kind]) {
        int n = 1;
        while (gs::batched_smem_bytes(kind, n + 1) <= kMaxSmem) ++n;   // single-kernel path (Hessenberg-only requests)
        cache[kind] = n;
    }
    return cache[kind];
}

// ---- per-device state: work-queue counters -------------------------------------------------------------
struct DeviceState {
    unsigned long long* counters = nullptr;   // ring of queue heads
    std::atomic<unsigned> next{0};
    int sm_count = 0;
    bool ok = false;
};
constexpr int kCounterRing = 1024;
constexpr int kMaxDevices = 64;
DeviceState g_dev[kMaxDevices];
std::mutex g_dev_mu;

int device_state(int dev, DeviceState** out) {
    if (dev < 0 || dev >= kMaxDevices) return fail(GSCHUR_ERR_ARG, "device index out of range");
    std::lock_guard<std::mutex> lk(g_dev_mu);
    DeviceState& d = g_dev[dev];
    if (!d.ok) {
        cudaDeviceProp prop;
        CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
        if (prop.major != 10)
            return fail(GSCHUR_ERR_CUDA, "libgschur_cuda is built for sm_100a only; device is sm_" +
                                             std::to_string(prop.major) + std::to_string(prop.minor));
        d.sm_count = prop.multiProcessorCount;
        CUDA_TRY(cudaMalloc(&d.counters, kCounterRing * sizeof(unsigned long long)));
        // The stream-ordered allocations of the batched paths (scale info, reflector-log pool and page table: gigabytes
        // per call) are kept in the device's default pool between calls instead of being returned to the driver at
        // every synchronisation; gschur_cuda_release_workspace() trims the pool.
        cudaMemPool_t mp = nullptr;
        if (cudaDeviceGetDefaultMemPool(&mp, dev) == cudaSuccess && mp) {
            unsigned long long thr = ~0ULL;
            cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &thr);
        }
        d.ok = true;
    }
    *out = &d;
    return 0;
}

// ---- kernel launch: one translation unit per element kind (launch_<kind>.cu) ------------------------------
int launch_kind(int kind, const BatchedParams& p, int dev_sms, cudaStream_t stream) {
    std::string err;
    int rc;
    switch (kind) {
        case GSCHUR_F64: rc = gs::launch_f64(p, dev_sms, stream, &err); break;
        case GSCHUR_C64: rc = gs::launch_c64(p, dev_sms, stream, &err); break;
        case GSCHUR_DD: rc = gs::launch_dd(p, dev_sms, stream, &err); break;
        case GSCHUR_CDD: rc = gs::launch_cdd(p, dev_sms, stream, &err); break;
        default: return fail(GSCHUR_ERR_ARG, "bad kind");
    }
    if (rc) return fail(rc, err);
    return 0;
}

// two-kernel Schur path: lanes own up to 4 (Float64 kinds) / 3 (double-double kinds) columns
int max_schur_n(int kind) { return (kind == GSCHUR_F64 || kind == GSCHUR_C64) ? 128 : 96; }

int check_args(int kind, int n, int64_t batch, const void* A, int lda, int64_t strideA, const void* Z, int ldz,
               int64_t strideZ, const void* w, bool schur_mode = true) {
    if (kind < 0 || kind > 3) return fail(GSCHUR_ERR_ARG, "kind must be 0..3");
    if (n < 0 || batch < 0) return fail(GSCHUR_ERR_ARG, "n and batch must be non-negative");
    if (n == 0 || batch == 0) return 0;
    if (!A) return fail(GSCHUR_ERR_ARG, "A is NULL");
    if (lda < n) return fail(GSCHUR_ERR_ARG, "DimensionMismatch: lda < n");
    if (batch > 1 && strideA < (int64_t)lda * (n - 1) + n) return fail(GSCHUR_ERR_ARG, "strideA overlaps matrices");
    if (Z) {
        if (ldz < n) return fail(GSCHUR_ERR_ARG, "DimensionMismatch: ldz < n");
        if (batch > 1 && strideZ < (int64_t)ldz * (n - 1) + n) return fail(GSCHUR_ERR_ARG, "strideZ overlaps matrices");
    }
    if (!w) return fail(GSCHUR_ERR_ARG, "w is NULL");
    // Schur and Hessenberg-only requests share the stage A kernel and therefore the limit; the single-kernel path that
    // GSCHUR_FORCE_GENERIC selects holds the whole problem in shared memory and takes less
    (void)schur_mode;
    const int lim = std::getenv("GSCHUR_FORCE_GENERIC") ? std::min(max_schur_n(kind), max_batched_n(kind)) : max_schur_n(kind);
    if (n > lim)
        return fail(GSCHUR_ERR_SIZE, "n = " + std::to_string(n) + " exceeds the batched-kernel limit " + std::to_string(lim) +
                                         " for this kind");
    return 0;
}

// enqueue one device-resident batch on `stream` of the current device
int enqueue_device(int kind, int mode, int n, int64_t batch, void* A, int lda, int64_t strideA, void* Z, int ldz,
                   int64_t strideZ, void* w, void* tau, int scale, int maxiter, int32_t* info, uint32_t* stats,
                   cudaStream_t stream, uint32_t flags) {
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    DeviceState* ds = nullptr;
    int rc = device_state(dev, &ds);
    if (rc) return rc;
    gs::stage_timing_begin_call();
    unsigned slot = ds->next.fetch_add(1) % kCounterRing;
    unsigned long long* counter = ds->counters + slot;
    CUDA_TRY(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), stream));
    BatchedParams p{};
    p.A = A;
    p.Z = Z;
    p.w = w;
    p.tau = tau;
    p.strideA = strideA;
    p.strideZ = strideZ;
    p.batch = batch;
    p.lda = lda;
    p.ldz = ldz;
    p.n = n;
    p.scale = scale;
    p.maxiter = maxiter;
    p.mode = mode;
    p.flags = flags & (GSCHUR_FLAG_HESS_INPUT | GSCHUR_FLAG_CHECK_SUBDIAG);
    p.info = info;
    p.stats = stats;
    p.counter = counter;
    p.scratch = nullptr;
    return launch_kind(kind, p, ds->sm_count, stream);
}

// ---- host-pointer path: one worker per device, chunked 3-stream pipeline ---------------------------------
struct HostJob {
    int kind, mode, n, lda, ldz, scale, maxiter;
    int64_t strideA, strideZ;
    char *A, *Z, *w, *tau;
    int32_t* info;
    uint32_t* stats;
    uint32_t flags;
};

struct ChunkBuf {
    void *dA = nullptr, *dZ = nullptr, *dw = nullptr, *dtau = nullptr;
    int32_t* dinfo = nullptr;
    uint32_t* dstats = nullptr;
    cudaEvent_t evH2D = nullptr, evComp = nullptr, evD2H = nullptr;   // this buffer's chunk: input landed / kernels done / results out
    bool busy = false;                                                 // evD2H has been recorded in this call
    size_t capA = 0, capZ = 0, capw = 0, captau = 0, capn = 0, capst = 0;
    // pageable caller arrays: pinned staging for this buffer's chunk (A in / T out, Z) and the chunk waiting in it
    char *hA = nullptr, *hZ = nullptr;
    size_t hcapA = 0, hcapZ = 0;
    int64_t pend_c0 = 0, pend_cn = 0;
};

// A few host threads that copy between pageable caller memory and pinned staging (one memcpy thread moves 6-10 GB/s; the
// PCIe link takes 55).  The workers are created on first use and live for the life of the process (detached: nothing to
// join when the library is unloaded at exit).
class CopyPool {
public:
    static CopyPool& get() {
        static CopyPool* pool = new CopyPool();
        return *pool;
    }
    void copy(char* dst, const char* src, size_t bytes) {
        if (bytes == 0) return;
        if (workers_ == 0 || bytes < 4 * kSeg) {
            std::memcpy(dst, src, bytes);
            return;
        }
        std::lock_guard<std::mutex> job(job_mu_);   // one copy at a time (callers on several devices take turns)
        {
            std::lock_guard<std::mutex> lk(mu_);
            dst_ = dst;
            src_ = src;
            bytes_ = bytes;
            next_.store(0);
            active_ = workers_;
            gen_ += 1;
        }
        cv_.notify_all();
        work();
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [&] { return active_ == 0; });
    }

private:
    static constexpr size_t kSeg = (size_t)2 << 20;
    CopyPool() {
        int nt = 8;
        if (const char* e = std::getenv("GSCHUR_COPY_THREADS")) nt = std::atoi(e);
        const int hc = (int)std::thread::hardware_concurrency();
        if (hc > 0 && nt > hc) nt = hc;
        if (nt < 1) nt = 1;
        workers_ = nt - 1;   // the calling thread copies too
        for (int i = 0; i < workers_; ++i) std::thread([this] { loop(); }).detach();
    }
    void work() {
        for (;;) {
            const size_t o = next_.fetch_add(kSeg);
            if (o >= bytes_) break;
            std::memcpy(dst_ + o, src_ + o, bytes_ - o < kSeg ? bytes_ - o : kSeg);
        }
    }
    void loop() {
        unsigned long long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
            }
            work();
            std::lock_guard<std::mutex> lk(mu_);
            if (--active_ == 0) done_.notify_all();
        }
    }
    std::mutex job_mu_, mu_;
    std::condition_variable cv_, done_;
    char* dst_ = nullptr;
    const char* src_ = nullptr;
    size_t bytes_ = 0;
    std::atomic<size_t> next_{0};
    int workers_ = 0, active_ = 0;
    unsigned long long gen_ = 0;
};

cudaError_t grow_pinned(char*& p, size_t& cap, size_t need) {
    if (need <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaHostAlloc((void**)&p, need, cudaHostAllocDefault);
    if (e == cudaSuccess) cap = need;
    return e;
}

// Per-device staging buffers and streams are cached across calls (grow-only): cudaMalloc / cudaFree of gigabytes
// and stream creation would otherwise sit inside every end-to-end call.
constexpr int NBUF = 4;
struct DevicePipe {
    ChunkBuf buf[NBUF];
    // Three streams: copies in, kernels, copies out.  The kernels of successive chunks are SERIALISED on one stream:
    // both stages are persistent grids sized to fill the GPU, and letting the grids of two chunks compete for SMs made
    // the end-to-end time bimodal (measured: 410 ms or 1000 ms for the same call, depending on which grid's CTAs became
    // resident first).  Copies overlap the kernels through per-buffer events.
    cudaStream_t sH2D = nullptr, sComp = nullptr, sD2H = nullptr;
    // pinned staging for the small outputs (w, info, stats, tau) of a whole slice: a D2H copy into the caller's
    // pageable arrays would block the enqueue loop until the chunk's kernels finish and serialise the pipeline
    char* hstage = nullptr;
    size_t hcap = 0;
    std::mutex mu;     // one host-pointer call at a time per device
};
DevicePipe g_pipe[kMaxDevices];

template <class P> cudaError_t grow(P*& p, size_t& cap, size_t need) {
    if (need <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc((void**)&p, need);
    if (e == cudaSuccess) cap = need;
    return e;
}

int run_slice(const HostJob& J, int dev, int64_t b0, int64_t b1, std::string* err) {
#define SL_TRY(expr)                                                                  \
    do {                                                                              \
        cudaError_t e__ = (expr);                                                     \
        if (e__ != cudaSuccess) {                                                     \
            *err = std::string(#expr) + ": " + cudaGetErrorString(e__);               \
            rc_final = GSCHUR_ERR_CUDA;                                               \
            goto cleanup;                                                             \
        }                                                                             \
    } while (0)
    int rc_final = 0;
    // GSCHUR_PIPE_TRACE=1: one line per slice on stderr with host and device times of the pipeline (development aid)
    const bool trace = std::getenv("GSCHUR_PIPE_TRACE") != nullptr;
    cudaEvent_t tevK0 = nullptr, tevK1 = nullptr, tevD1 = nullptr;
    double t_begin = 0, t_enq = 0;
    auto now_s = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const int n = J.n;
    const size_t es = elem_size(J.kind), ws = eig_size(J.kind);
    const size_t mat = (size_t)n * n * es;
    const int64_t count = b1 - b0;
    // chunk size: aim at >= 16 chunks per slice but at least enough matrices to fill the GPU a few times over
    int nchunks = 8;
    if (const char* e = std::getenv("GSCHUR_PIPE_CHUNKS")) {   // tuning knob
        const int v = std::atoi(e);
        if (v >= 1 && v <= 4096) nchunks = v;
    }
    int64_t chunk = (count + nchunks - 1) / nchunks;
    const int64_t min_chunk = 2048;
    if (chunk < min_chunk) chunk = min_chunk;
    if (chunk > count) chunk = count;
    const bool wantZ = J.Z != nullptr;
    {
        // Device-memory budget of the NBUF chunk buffers (A and Z per matrix): GSCHUR_PIPE_BUDGET_MB (default 8 GiB),
        // never more than a quarter of what is free now — the kernels' own workspaces (reflector log pool) need room
        // too.  A larger slice simply flows through more chunks.
        size_t budget = (size_t)8 << 30;
        if (const char* e = std::getenv("GSCHUR_PIPE_BUDGET_MB")) {
            const long long v = std::atoll(e);
            if (v >= 1) budget = (size_t)v << 20;
        }
        if (cudaSetDevice(dev) == cudaSuccess) {
            size_t fr = 0, tot = 0;
            size_t have = 0;
            for (int i = 0; i < NBUF; ++i) have += g_pipe[dev].buf[i].capA + g_pipe[dev].buf[i].capZ;
            if (cudaMemGetInfo(&fr, &tot) == cudaSuccess && (fr + have) / 4 < budget) budget = (fr + have) / 4;
        }
        const size_t per_matrix = (size_t)NBUF * (wantZ ? 2 : 1) * mat + 64 * NBUF;
        int64_t fit = (int64_t)(budget / per_matrix);
        if (fit < 1) fit = 1;
        if (chunk > fit) chunk = fit;
    }
    const bool hess = J.mode == MODE_HESSENBERG;
    const bool zin = wantZ && (J.flags & GSCHUR_FLAG_HESS_INPUT) && !hess;
    const bool denseA = (J.lda == n) && (J.strideA == (int64_t)n * n);
    const bool denseZ = (J.ldz == n) && (J.strideZ == (int64_t)n * n);
    // Pageable caller memory (a Julia Array, an ordinary numpy array): cudaMemcpyAsync from / to it is staged by the driver
    // through one thread and synchronous with respect to the host, which serialises the three-stream pipeline (measured:
    // 52 k instead of 179 k matrices/s on 65536 x 64x64 ComplexF64).  Such slices go through the library's own pinned
    // staging, filled and drained by the copy pool, chunk by chunk around the device pipeline.  GSCHUR_HOST_STAGING=0: off.
    bool staged = false;
    {
        const char* hs = std::getenv("GSCHUR_HOST_STAGING");
        const char* hr = std::getenv("GSCHUR_HOST_REGISTER");
        const bool allowed = !(hs && hs[0] == '0') && !(hr && hr[0] == '1');
        if (allowed && denseA && (!wantZ || denseZ) && (size_t)count * mat >= ((size_t)8 << 20)) {
            auto pageable = [&](const char* ptr) {
                cudaPointerAttributes at;
                if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) {
                    cudaGetLastError();
                    return false;
                }
                return at.type == cudaMemoryTypeUnregistered;
            };
            staged = pageable(J.A + (size_t)b0 * mat) || (wantZ && pageable(J.Z + (size_t)b0 * mat));
        }
        if (staged) {
            // pinned staging of the NBUF chunks: GSCHUR_STAGE_BUDGET_MB (default 4 GiB)
            size_t budget = (size_t)4 << 30;
            if (const char* e = std::getenv("GSCHUR_STAGE_BUDGET_MB")) {
                const long long v = std::atoll(e);
                if (v >= 1) budget = (size_t)v << 20;
            }
            int64_t fit = (int64_t)(budget / ((size_t)NBUF * (wantZ ? 2 : 1) * mat));
            if (fit < 1) fit = 1;
            if (chunk > fit) chunk = fit;
        }
    }
    if (cudaSetDevice(dev) != cudaSuccess) {
        *err = "cudaSetDevice failed";
        return GSCHUR_ERR_CUDA;
    }
    {
        DeviceState* ds = nullptr;
        int rc = device_state(dev, &ds);
        if (rc) {
            *err = g_err;
            return rc;
        }
    }
    DevicePipe& P = g_pipe[dev];
    std::lock_guard<std::mutex> lk(P.mu);
    ChunkBuf* buf = P.buf;
    const size_t tau_per = es * (size_t)(n > 1 ? n - 1 : 0);
    const size_t off_w = 0;
    const size_t off_info = off_w + (hess ? tau_per : ws * n) * (size_t)count;
    const size_t off_stats = off_info + sizeof(int32_t) * (size_t)count;
    const size_t stage_bytes = off_stats + sizeof(uint32_t) * GSCHUR_STATS_PER_MATRIX * (size_t)count;
    if (stage_bytes > P.hcap) {
        if (P.hstage) cudaFreeHost(P.hstage);
        P.hstage = nullptr;
        P.hcap = 0;
        if (cudaMallocHost((void**)&P.hstage, stage_bytes) != cudaSuccess) {
            *err = "cudaMallocHost(staging) failed";
            return GSCHUR_ERR_CUDA;
        }
        P.hcap = stage_bytes;
    }
    char* hst = P.hstage;
    // Pageable caller memory (a Julia Array, an ordinary numpy array): cudaMemcpyAsync from / to it is staged by the driver
    // and synchronous with respect to the host, which serialises the three-stream pipeline.  GSCHUR_HOST_REGISTER=1
    // page-locks the slice's ranges of A and Z for the duration of the call; memory that is already pinned or registered
    // is left alone.  Off by default: measured on the B200 boxes (65536 x 64x64 ComplexF64, fresh arrays) page-locking
    // 8.6 GB costs as much as it saves (1.65 s against 1.45-1.63 s without; 0.36 s from pinned arrays) — a caller that
    // reuses its arrays should pin them once itself.
    void* reg_ptr[2] = {nullptr, nullptr};
    {
        const char* hr = std::getenv("GSCHUR_HOST_REGISTER");
        const bool want_reg = hr && hr[0] == '1';
        auto try_register = [&](char* base, bool dense, int64_t stride, int ld, int slot) {
            if (!want_reg || !base || !dense) return;
            (void)stride;
            (void)ld;
            char* lo = base + (size_t)b0 * mat;
            const size_t bytes = (size_t)count * mat;
            if (bytes < ((size_t)32 << 20)) return;     // small slices: the registration costs more than it saves
            cudaPointerAttributes at;
            if (cudaPointerGetAttributes(&at, lo) != cudaSuccess) {
                cudaGetLastError();
                return;
            }
            if (at.type != cudaMemoryTypeUnregistered) return;
            if (cudaHostRegister(lo, bytes, cudaHostRegisterDefault) == cudaSuccess) reg_ptr[slot] = lo;
            else cudaGetLastError();
        };
        try_register(J.A, denseA, J.strideA, J.lda, 0);
        try_register(J.Z, wantZ && denseZ, J.strideZ, J.ldz, 1);
    }
    if (!P.sH2D) SL_TRY(cudaStreamCreateWithFlags(&P.sH2D, cudaStreamNonBlocking));
    if (!P.sComp) SL_TRY(cudaStreamCreateWithFlags(&P.sComp, cudaStreamNonBlocking));
    if (!P.sD2H) SL_TRY(cudaStreamCreateWithFlags(&P.sD2H, cudaStreamNonBlocking));
    for (int i = 0; i < NBUF; ++i) {
        if (!buf[i].evH2D) SL_TRY(cudaEventCreateWithFlags(&buf[i].evH2D, cudaEventDisableTiming));
        if (!buf[i].evComp) SL_TRY(cudaEventCreateWithFlags(&buf[i].evComp, cudaEventDisableTiming));
        if (!buf[i].evD2H) SL_TRY(cudaEventCreateWithFlags(&buf[i].evD2H, cudaEventDisableTiming));
        buf[i].busy = false;
        SL_TRY(grow(buf[i].dA, buf[i].capA, mat * chunk));
        if (wantZ) SL_TRY(grow(buf[i].dZ, buf[i].capZ, mat * chunk));
        if (!hess) SL_TRY(grow(buf[i].dw, buf[i].capw, ws * n * chunk));
        if (hess) SL_TRY(grow(buf[i].dtau, buf[i].captau, es * (n > 1 ? n - 1 : 1) * chunk));
        SL_TRY(grow(buf[i].dinfo, buf[i].capn, sizeof(int32_t) * chunk));
        SL_TRY(grow(buf[i].dstats, buf[i].capst, sizeof(uint32_t) * GSCHUR_STATS_PER_MATRIX * chunk));
        if (staged) {
            // no page-locked memory to be had: leave the staging to the driver (slower, not an error)
            cudaError_t eh = grow_pinned(buf[i].hA, buf[i].hcapA, mat * chunk);
            if (eh == cudaSuccess && wantZ) eh = grow_pinned(buf[i].hZ, buf[i].hcapZ, mat * chunk);
            if (eh != cudaSuccess) {
                cudaGetLastError();
                staged = false;
            }
        }
    }
    {
        // Chunk schedule: full chunks in the middle, a quarter and a half chunk at either end — the first H2D and the last
        // D2H are the only copies that nothing overlaps, so they are kept short (when the slice is large enough to matter).
        std::vector<int64_t> sizes;
        {
            int64_t left = count;
            const bool taper = count >= 4 * chunk && chunk >= 4 * min_chunk && count <= 64 * chunk;
            if (taper) {
                sizes.push_back(chunk / 4);
                sizes.push_back(chunk / 2);
                left -= chunk / 4 + chunk / 2 + chunk / 2 + chunk / 4;
            }
            while (left > 0) {
                const int64_t c = left < chunk ? left : chunk;
                sizes.push_back(c);
                left -= c;
            }
            if (taper) {
                sizes.push_back(chunk / 2);
                sizes.push_back(chunk / 4);
            }
        }
        int64_t c0 = b0;
        auto drain = [&](ChunkBuf& B) {
            if (B.pend_cn <= 0) return;
            CopyPool::get().copy(J.A + (size_t)B.pend_c0 * mat, B.hA, mat * (size_t)B.pend_cn);
            if (wantZ) CopyPool::get().copy(J.Z + (size_t)B.pend_c0 * mat, B.hZ, mat * (size_t)B.pend_cn);
            B.pend_cn = 0;
        };
        for (int i = 0; i < NBUF; ++i) buf[i].pend_cn = 0;
        if (trace) {
            cudaEventCreate(&tevK0);
            cudaEventCreate(&tevK1);
            cudaEventCreate(&tevD1);
            t_begin = now_s();
        }
        for (int ci = 0; ci < (int)sizes.size(); c0 += sizes[ci], ++ci) {
            ChunkBuf& B = buf[ci % NBUF];
            const int64_t cn = sizes[ci];
            cudaStream_t s = P.sH2D;
            // H2D (after the previous chunk in this buffer has been copied out)
            if (staged) {
                if (B.busy) {   // the results of the chunk that used this buffer: staging -> caller
                    SL_TRY(cudaEventSynchronize(B.evD2H));
                    drain(B);
                }
                CopyPool::get().copy(B.hA, J.A + (size_t)c0 * mat, mat * cn);
                if (zin) CopyPool::get().copy(B.hZ, J.Z + (size_t)c0 * mat, mat * cn);
                SL_TRY(cudaMemcpyAsync(B.dA, B.hA, mat * cn, cudaMemcpyHostToDevice, s));
                if (zin) SL_TRY(cudaMemcpyAsync(B.dZ, B.hZ, mat * cn, cudaMemcpyHostToDevice, s));
                B.pend_c0 = c0;
                B.pend_cn = cn;
            } else if (B.busy) {
                SL_TRY(cudaStreamWaitEvent(P.sH2D, B.evD2H, 0));
            }
            if (staged) {
            } else if (denseA) {
                SL_TRY(cudaMemcpyAsync(B.dA, J.A + (size_t)c0 * J.strideA * es, mat * cn, cudaMemcpyHostToDevice, s));
            } else {
                for (int64_t b = 0; b < cn; ++b)
                    SL_TRY(cudaMemcpy2DAsync((char*)B.dA + b * mat, (size_t)n * es,
                                             J.A + (size_t)(c0 + b) * J.strideA * es, (size_t)J.lda * es,
                                             (size_t)n * es, n, cudaMemcpyHostToDevice, s));
            }
            if (zin && !staged) {
                if (denseZ) {
                    SL_TRY(cudaMemcpyAsync(B.dZ, J.Z + (size_t)c0 * J.strideZ * es, mat * cn, cudaMemcpyHostToDevice, s));
                } else {
                    for (int64_t b = 0; b < cn; ++b)
                        SL_TRY(cudaMemcpy2DAsync((char*)B.dZ + b * mat, (size_t)n * es,
                                                 J.Z + (size_t)(c0 + b) * J.strideZ * es, (size_t)J.ldz * es,
                                                 (size_t)n * es, n, cudaMemcpyHostToDevice, s));
                }
            }
            SL_TRY(cudaEventRecord(B.evH2D, P.sH2D));
            s = P.sComp;
            SL_TRY(cudaStreamWaitEvent(P.sComp, B.evH2D, 0));
            if (trace && ci == 0) cudaEventRecord(tevK0, P.sComp);
            int rc = enqueue_device(J.kind, J.mode, n, cn, B.dA, n, (int64_t)n * n, wantZ ? B.dZ : nullptr, n,
                                    (int64_t)n * n, B.dw, B.dtau, J.scale, J.maxiter, B.dinfo, B.dstats, s, J.flags);
            if (rc) {
                *err = g_err;
                rc_final = rc;
                goto cleanup;
            }
            // D2H
            SL_TRY(cudaEventRecord(B.evComp, P.sComp));
            s = P.sD2H;
            SL_TRY(cudaStreamWaitEvent(P.sD2H, B.evComp, 0));
            if (staged) {
                SL_TRY(cudaMemcpyAsync(B.hA, B.dA, mat * cn, cudaMemcpyDeviceToHost, s));
                if (wantZ) SL_TRY(cudaMemcpyAsync(B.hZ, B.dZ, mat * cn, cudaMemcpyDeviceToHost, s));
            } else if (denseA) {
                SL_TRY(cudaMemcpyAsync(J.A + (size_t)c0 * J.strideA * es, B.dA, mat * cn, cudaMemcpyDeviceToHost, s));
            } else {
                for (int64_t b = 0; b < cn; ++b)
                    SL_TRY(cudaMemcpy2DAsync(J.A + (size_t)(c0 + b) * J.strideA * es, (size_t)J.lda * es,
                                             (char*)B.dA + b * mat, (size_t)n * es, (size_t)n * es, n,
                                             cudaMemcpyDeviceToHost, s));
            }
            if (wantZ && !staged) {
                if (denseZ) {
                    SL_TRY(cudaMemcpyAsync(J.Z + (size_t)c0 * J.strideZ * es, B.dZ, mat * cn, cudaMemcpyDeviceToHost, s));
                } else {
                    for (int64_t b = 0; b < cn; ++b)
                        SL_TRY(cudaMemcpy2DAsync(J.Z + (size_t)(c0 + b) * J.strideZ * es, (size_t)J.ldz * es,
                                                 (char*)B.dZ + b * mat, (size_t)n * es, (size_t)n * es, n,
                                                 cudaMemcpyDeviceToHost, s));
                }
            }
            const size_t l0 = (size_t)(c0 - b0);     // index within the slice
            if (!hess)
                SL_TRY(cudaMemcpyAsync(hst + off_w + l0 * n * ws, B.dw, ws * n * cn, cudaMemcpyDeviceToHost, s));
            if (hess && n > 1)
                SL_TRY(cudaMemcpyAsync(hst + off_w + l0 * tau_per, B.dtau, tau_per * cn, cudaMemcpyDeviceToHost, s));
            if (J.info && !hess)
                SL_TRY(cudaMemcpyAsync(hst + off_info + l0 * sizeof(int32_t), B.dinfo, sizeof(int32_t) * cn,
                                       cudaMemcpyDeviceToHost, s));
            if (J.stats && !hess)
                SL_TRY(cudaMemcpyAsync(hst + off_stats + l0 * sizeof(uint32_t) * GSCHUR_STATS_PER_MATRIX, B.dstats,
                                       sizeof(uint32_t) * GSCHUR_STATS_PER_MATRIX * cn, cudaMemcpyDeviceToHost, s));
            SL_TRY(cudaEventRecord(B.evD2H, P.sD2H));
            B.busy = true;
        }
        if (trace) {
            cudaEventRecord(tevK1, P.sComp);
            cudaEventRecord(tevD1, P.sD2H);
            t_enq = now_s();
        }
        if (staged) {   // the chunks still in the staging buffers, oldest first
            const int nc = (int)sizes.size();
            for (int ci = nc > NBUF ? nc - NBUF : 0; ci < nc; ++ci) {
                ChunkBuf& B = buf[ci % NBUF];
                SL_TRY(cudaEventSynchronize(B.evD2H));
                drain(B);
            }
        }
    }
cleanup:
    if (P.sH2D) cudaStreamSynchronize(P.sH2D);
    if (P.sComp) cudaStreamSynchronize(P.sComp);
    if (P.sD2H) cudaStreamSynchronize(P.sD2H);
    if (trace && tevK0) {
        const double t_sync = now_s();
        float k_ms = 0, d_ms = 0;
        cudaEventElapsedTime(&k_ms, tevK0, tevK1);
        cudaEventElapsedTime(&d_ms, tevK0, tevD1);
        std::fprintf(stderr, "[gschur pipe] dev %d: enqueue %.1f ms, until sync %.1f ms; first kernel -> last kernel %.1f ms, -> last D2H %.1f ms\n",
                     dev, 1e3 * (t_enq - t_begin), 1e3 * (t_sync - t_begin), k_ms, d_ms);
        cudaEventDestroy(tevK0);
        cudaEventDestroy(tevK1);
        cudaEventDestroy(tevD1);
    }
    for (int i = 0; i < 2; ++i)
        if (reg_ptr[i]) cudaHostUnregister(reg_ptr[i]);
    if (rc_final == 0) {
        if (!hess) std::memcpy(J.w + (size_t)b0 * n * ws, hst + off_w, ws * n * (size_t)count);
        if (hess && n > 1) std::memcpy(J.tau + (size_t)b0 * tau_per, hst + off_w, tau_per * (size_t)count);
        if (J.info && !hess) std::memcpy(J.info + b0, hst + off_info, sizeof(int32_t) * (size_t)count);
        if (J.stats && !hess)
            std::memcpy(J.stats + (size_t)b0 * GSCHUR_STATS_PER_MATRIX, hst + off_stats,
                        sizeof(uint32_t) * GSCHUR_STATS_PER_MATRIX * (size_t)count);
    }
    return rc_final;
#undef SL_TRY
}

int run_host(const HostJob& J, int64_t batch, const int* devices, int ndev) {
    int dev0 = 0;
    std::vector<int> devs;
    if (devices && ndev > 0) devs.assign(devices, devices + ndev);
    else devs.push_back(dev0);
    int have = 0;
    CUDA_TRY(cudaGetDeviceCount(&have));
    for (int d : devs)
        if (d < 0 || d >= have) return fail(GSCHUR_ERR_ARG, "device " + std::to_string(d) + " not present");
    const int G = (int)devs.size();
    std::vector<int> rcs(G, 0);
    std::vector<std::string> errs(G);
    std::vector<std::thread> th;
    // info is needed to count failures even when the caller passed NULL
    std::vector<int32_t> info_local;
    HostJob JJ = J;
    if (!JJ.info && JJ.mode == MODE_SCHUR) {
        info_local.assign((size_t)batch, 0);
        JJ.info = info_local.data();
    }
    int prev = 0;
    cudaGetDevice(&prev);
    for (int g = 0; g < G; ++g) {
        int64_t b0 = batch * g / G, b1 = batch * (g + 1) / G;
        if (b1 <= b0) continue;
        if (G == 1) {
            rcs[g] = run_slice(JJ, devs[g], b0, b1, &errs[g]);
        } else {
            th.emplace_back([&, g, b0, b1]() { rcs[g] = run_slice(JJ, devs[g], b0, b1, &errs[g]); });
        }
    }
    for (auto& t : th) t.join();
    cudaSetDevice(prev);
    for (int g = 0; g < G; ++g)
        if (rcs[g]) return fail(rcs[g], "device " + std::to_string(devs[g]) + ": " + errs[g]);
    if (JJ.mode == MODE_SCHUR) {
        int64_t bad = 0;
        bool subdiag = false;
        for (int64_t b = 0; b < batch; ++b) {
            if (JJ.info[b] > 0) ++bad;
            if (JJ.info[b] == GSCHUR_ERR_SUBDIAG) subdiag = true;
        }
        if (subdiag) return fail(GSCHUR_ERR_SUBDIAG, "ArgumentError: algorithm assumes real subdiagonal");
        if (bad) g_err = "UnconvergedException: iteration limit reached for " + std::to_string(bad) + " matrices";
        return (int)(bad > 0x7fffffff ? 0x7fffffff : bad);
    }
    return 0;
}

}  // namespace

extern "C" {

int gschur_cuda_version(void) { return 100; }

int gschur_cuda_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        g_err = "cudaGetDeviceCount failed (no CUDA device / driver)";
        cudaGetLastError();
        return 0;
    }
    return n;
}

const char* gschur_cuda_last_error(void) { return g_err.c_str(); }

uint64_t gschur_cuda_launch_count(void) { return gs::launch_counter(); }

// per-stage device times (ms) of the most recent batched call while timing is enabled (summed over its sub-batches)
int gschur_cuda_stage_timing3(int enable, float* ms_stage_a, float* ms_stage_b, float* ms_stage_c) {
    if (enable >= 0) gs::stage_timing_enable(enable != 0);
    if (ms_stage_a && ms_stage_b) return gs::stage_timing_read(ms_stage_a, ms_stage_b, ms_stage_c);
    return 0;
}
// two-value form: stage B here is everything after stage A (QR iteration + Z replay)
int gschur_cuda_stage_timing(int enable, float* ms_stage_a, float* ms_stage_b) {
    float c = 0.f;
    int rc = gschur_cuda_stage_timing3(enable, ms_stage_a, ms_stage_b, &c);
    if (rc == 0 && ms_stage_a && ms_stage_b) *ms_stage_b += c;
    return rc;
}

int gschur_cuda_release_workspace(void) {
    // frees what the host-pointer pipeline caches per device (chunk buffers, pinned staging) and returns the
    // stream-ordered allocations of the batched paths (scale info, reflector log pools) to the driver
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess) return GSCHUR_ERR_CUDA;
    int prev = 0;
    cudaGetDevice(&prev);
    for (int d = 0; d < have && d < kMaxDevices; ++d) {
        DevicePipe& P = g_pipe[d];
        std::lock_guard<std::mutex> lk(P.mu);
        bool any = P.hstage != nullptr;
        for (int i = 0; i < NBUF; ++i) any = any || P.buf[i].hA || P.buf[i].hZ;
        for (int i = 0; i < NBUF; ++i) any = any || P.buf[i].dA || P.buf[i].dZ || P.buf[i].dw || P.buf[i].dtau || P.buf[i].dinfo || P.buf[i].dstats;
        if (cudaSetDevice(d) != cudaSuccess) continue;
        if (any) {
            cudaDeviceSynchronize();
            for (int i = 0; i < NBUF; ++i) {
                ChunkBuf& B = P.buf[i];
                if (B.dA) cudaFree(B.dA);
                if (B.dZ) cudaFree(B.dZ);
                if (B.dw) cudaFree(B.dw);
                if (B.dtau) cudaFree(B.dtau);
                if (B.dinfo) cudaFree(B.dinfo);
                if (B.dstats) cudaFree(B.dstats);
                B.dA = B.dZ = B.dw = B.dtau = nullptr;
                B.dinfo = nullptr;
                B.dstats = nullptr;
                B.capA = B.capZ = B.capw = B.captau = B.capn = B.capst = 0;
                if (B.hA) cudaFreeHost(B.hA);
                if (B.hZ) cudaFreeHost(B.hZ);
                B.hA = B.hZ = nullptr;
                B.hcapA = B.hcapZ = 0;
            }
            if (P.hstage) cudaFreeHost(P.hstage);
            P.hstage = nullptr;
            P.hcap = 0;
        }
        cudaMemPool_t mp = nullptr;
        if (cudaDeviceGetDefaultMemPool(&mp, d) == cudaSuccess && mp) cudaMemPoolTrimTo(mp, 0);
    }
    cudaSetDevice(prev);
    return 0;
}

int gschur_cuda_max_batched_n(int kind) {
    if (kind < 0 || kind > 3) return 0;
    return max_schur_n(kind);
}

int gschur_cuda_batched_async(int kind, int n, int64_t batch, void* A, int lda, int64_t strideA, void* Z, int ldz,
                              int64_t strideZ, void* w, int scale, int maxiter, int32_t* info, uint32_t* stats,
                              void* stream, uint32_t flags) {
    g_err.clear();
    int rc = check_args(kind, n, batch, A, lda, strideA, Z, ldz, strideZ, w);
    if (rc) return rc;
    if (n == 0 || batch == 0) return 0;
    return enqueue_device(kind, MODE_SCHUR, n, batch, A, lda, strideA, Z, ldz, strideZ, w, nullptr, scale, maxiter,
                          info, stats, (cudaStream_t)stream, flags);
}

int gschur_cuda_batched(int kind, int n, int64_t batch, void* A, int lda, int64_t strideA, void* Z, int ldz,
                        int64_t strideZ, void* w, int scale, int maxiter, int32_t* info, uint32_t* stats,
                        const int* devices, int ndev, uint32_t flags) {
    g_err.clear();
    int rc = check_args(kind, n, batch, A, lda, strideA, Z, ldz, strideZ, w);
    if (rc) return rc;
    if (n == 0 || batch == 0) return 0;
    if (gschur_cuda_device_count() < 1) return fail(GSCHUR_ERR_CUDA, "no CUDA device available (there is no CPU fallback)");
    if (flags & GSCHUR_FLAG_DEVICE_PTRS) {
        // the return value counts failures even when the caller does not want the per-matrix codes
        int32_t* dinfo = info;
        if (!dinfo) CUDA_TRY(cudaMalloc((void**)&dinfo, sizeof(int32_t) * (size_t)batch));
        rc = enqueue_device(kind, MODE_SCHUR, n, batch, A, lda, strideA, Z, ldz, strideZ, w, nullptr, scale, maxiter,
                            dinfo, stats, (cudaStream_t)0, flags);
        std::vector<int32_t> h((size_t)batch);
        cudaError_t ce = cudaSuccess;
        if (!rc) ce = cudaStreamSynchronize((cudaStream_t)0);
        if (!rc && ce == cudaSuccess) ce = cudaMemcpy(h.data(), dinfo, sizeof(int32_t) * batch, cudaMemcpyDeviceToHost);
        if (!info) cudaFree(dinfo);
        if (rc) return rc;
        CUDA_TRY(ce);
        int64_t bad = 0;
        for (int64_t b = 0; b < batch; ++b) {
            if (h[b] == GSCHUR_ERR_SUBDIAG) return fail(GSCHUR_ERR_SUBDIAG, "ArgumentError: algorithm assumes real subdiagonal");
            if (h[b] > 0) ++bad;
        }
        return (int)bad;
    }
    HostJob J;
    J.kind = kind;
    J.mode = MODE_SCHUR;
    J.n = n;
    J.lda = lda;
    J.ldz = ldz;
    J.scale = scale;
    J.maxiter = maxiter;
    J.strideA = strideA;
    J.strideZ = strideZ;
    J.A = (char*)A;
    J.Z = (char*)Z;
    J.w = (char*)w;
    J.tau = nullptr;
    J.info = info;
    J.stats = stats;
    J.flags = flags;
    return run_host(J, batch, devices, ndev);
}

int gschur_cuda_hessenberg_batched(int kind, int n, int64_t batch, void* A, int lda, int64_t strideA, void* tau,
                                   void* Q, int ldq, int64_t strideQ, const int* devices, int ndev, uint32_t flags) {
    g_err.clear();
    int dummy = 0;
    int rc = check_args(kind, n, batch, A, lda, strideA, Q, ldq, strideQ, &dummy, false);
    if (rc) return rc;
    if (n == 0 || batch == 0) return 0;
    if (n > 1 && !tau) return fail(GSCHUR_ERR_ARG, "tau is NULL");
    if (gschur_cuda_device_count() < 1) return fail(GSCHUR_ERR_CUDA, "no CUDA device available (there is no CPU fallback)");
    if (flags & GSCHUR_FLAG_DEVICE_PTRS) {
        rc = enqueue_device(kind, MODE_HESSENBERG, n, batch, A, lda, strideA, Q, ldq, strideQ, nullptr, tau, 0, 0,
                            nullptr, nullptr, (cudaStream_t)0, 0);
        if (rc) return rc;
        CUDA_TRY(cudaStreamSynchronize((cudaStream_t)0));
        return 0;
    }
    HostJob J;
    J.kind = kind;
    J.mode = MODE_HESSENBERG;
    J.n = n;
    J.lda = lda;
    J.ldz = ldq;
    J.scale = 0;
    J.maxiter = 0;
    J.strideA = strideA;
    J.strideZ = strideQ;
    J.A = (char*)A;
    J.Z = (char*)Q;
    J.w = nullptr;
    J.tau = (char*)tau;
    J.info = nullptr;
    J.stats = nullptr;
    J.flags = 0;
    return run_host(J, batch, devices, ndev);
}

}  // extern "C"

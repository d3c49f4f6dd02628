// shared-memory sizing query used by the ABI's argument checks
#include <atomic>
#include <mutex>
#include <vector>
#include "batched.cuh"
namespace gs {
static std::atomic<unsigned long long> g_launch_counter{0};
void note_launch() { g_launch_counter.fetch_add(1); }
unsigned long long launch_counter() { return g_launch_counter.load(); }

// Per-stage timing of the batched paths: CUDA events on the launching stream, one group of four marks per
// (sub-)batch — before stage A, after A, after B, after C (+ redo) — summed over the groups of the most recent call.
// Development / bench aid: single caller at a time (guarded by a mutex); the events belong to the device that was current
// when timing was switched on, and launches on any other device (the multi-device host path) are not marked.
static bool g_stage_timing = false;
static int g_stage_dev = -1;
static std::mutex g_ev_mu;
struct MarkGroup { cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr}; };
static std::vector<MarkGroup> g_marks;
static size_t g_marks_used = 0;
void stage_timing_enable(bool on) {
    g_stage_timing = on;
    g_stage_dev = -1;
    if (on) cudaGetDevice(&g_stage_dev);
}
bool stage_timing_enabled() { return g_stage_timing; }
void stage_timing_begin_call() {
    if (!g_stage_timing) return;
    std::lock_guard<std::mutex> lk(g_ev_mu);
    g_marks_used = 0;
}
void stage_timing_mark(int which, cudaStream_t s) {
    if (!g_stage_timing || which < 0 || which > 3) return;
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess || dev != g_stage_dev) return;
    std::lock_guard<std::mutex> lk(g_ev_mu);
    if (which == 0) {
        if (g_marks_used == g_marks.size()) g_marks.emplace_back();
        g_marks_used += 1;
    }
    if (g_marks_used == 0) return;
    MarkGroup& g = g_marks[g_marks_used - 1];
    if (!g.ev[which]) cudaEventCreate(&g.ev[which]);
    cudaEventRecord(g.ev[which], s);
    if (which == 2) {   // the fused path has no stage C: its mark 3 coincides with mark 2
        if (!g.ev[3]) cudaEventCreate(&g.ev[3]);
        cudaEventRecord(g.ev[3], s);
    }
}
int stage_timing_read(float* ms_a, float* ms_b, float* ms_c) {
    std::lock_guard<std::mutex> lk(g_ev_mu);
    if (g_marks_used == 0) return -1;
    float a = 0, b = 0, c = 0;
    for (size_t i = 0; i < g_marks_used; ++i) {
        MarkGroup& g = g_marks[i];
        for (int k = 0; k < 4; ++k)
            if (!g.ev[k]) return -1;
        if (cudaEventSynchronize(g.ev[3]) != cudaSuccess) return -2;
        float x = 0, y = 0, z = 0;
        if (cudaEventElapsedTime(&x, g.ev[0], g.ev[1]) != cudaSuccess) return -2;
        if (cudaEventElapsedTime(&y, g.ev[1], g.ev[2]) != cudaSuccess) return -2;
        if (cudaEventElapsedTime(&z, g.ev[2], g.ev[3]) != cudaSuccess) return -2;
        a += x;
        b += y;
        c += z;
    }
    *ms_a = a;
    *ms_b = b;
    if (ms_c) *ms_c = c;
    return 0;
}

size_t batched_smem_bytes(int kind, int n) {
    switch (kind) {
        case 0: return smem_layout<double>::bytes(n);
        case 1: return smem_layout<cx<double>>::bytes(n);
        case 2: return smem_layout<dd_t>::bytes(n);
        case 3: return smem_layout<cx<dd_t>>::bytes(n);
    }
    return (size_t)-1;
}
}  // namespace gs

// shared-memory sizing query used by the ABI's argument checks
#include <atomic>
#include "batched.cuh"
namespace gs {
static std::atomic<unsigned long long> g_launch_counter{0};
void note_launch() { g_launch_counter.fetch_add(1); }
unsigned long long launch_counter() { return g_launch_counter.load(); }

static bool g_stage_timing = false;
static cudaEvent_t g_ev[3] = {nullptr, nullptr, nullptr};
void stage_timing_enable(bool on) { g_stage_timing = on; }
bool stage_timing_enabled() { return g_stage_timing; }
void stage_timing_mark(int which, cudaStream_t s) {
    if (!g_stage_timing || which < 0 || which > 2) return;
    if (!g_ev[which]) cudaEventCreate(&g_ev[which]);
    cudaEventRecord(g_ev[which], s);
}
int stage_timing_read(float* ms_a, float* ms_b) {
    if (!g_ev[0] || !g_ev[1] || !g_ev[2]) return -1;
    if (cudaEventSynchronize(g_ev[2]) != cudaSuccess) return -2;
    if (cudaEventElapsedTime(ms_a, g_ev[0], g_ev[1]) != cudaSuccess) return -2;
    if (cudaEventElapsedTime(ms_b, g_ev[1], g_ev[2]) != cudaSuccess) return -2;
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

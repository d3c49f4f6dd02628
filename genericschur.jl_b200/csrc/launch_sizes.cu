// shared-memory sizing query used by the ABI's argument checks
#include <atomic>
#include "batched.cuh"
namespace gs {
static std::atomic<unsigned long long> g_launch_counter{0};
void note_launch() { g_launch_counter.fetch_add(1); }
unsigned long long launch_counter() { return g_launch_counter.load(); }

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

// Kernel instantiations for element kind f64 (one translation unit per kind keeps builds parallel).
// n <= 64 goes to the warp-specialised kernel (fastqr.cuh); larger n to the generic block-synchronous one.
#include <cstdlib>
#include "fastqr.cuh"
namespace gs {
int launch_f64(const BatchedParams& p, int dev_sms, cudaStream_t stream, std::string* err) {
    static const bool force_generic = std::getenv("GSCHUR_FORCE_GENERIC") != nullptr;
    if (!force_generic && p.mode == MODE_SCHUR) {
        if (p.n <= 32) return launch_fast<double, 1>(p, dev_sms, stream, err);
        if (p.n <= 64) return launch_fast<double, 2>(p, dev_sms, stream, err);
    }
    return launch_t<double>(p, dev_sms, stream, err);
}
}  // namespace gs

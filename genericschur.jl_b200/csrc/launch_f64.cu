// Kernel instantiations for element kind f64 (one translation unit per kind keeps builds parallel).
// Schur requests go to the two-kernel path (gehrd.cuh + fastqr.cuh, n <= 128); Hessenberg-only requests and
// anything forced by GSCHUR_FORCE_GENERIC to the block-synchronous single-kernel path (batched.cuh).
#include <cstdlib>
#include "fastqr.cuh"
namespace gs {
int launch_f64(const BatchedParams& p, int dev_sms, cudaStream_t stream, std::string* err) {
    static const bool force_generic = std::getenv("GSCHUR_FORCE_GENERIC") != nullptr;
    if (!force_generic && p.mode == MODE_SCHUR) {
        if (p.n <= 32) return launch_fast<double, 1>(p, dev_sms, stream, err);
        if (p.n <= 64) return launch_fast<double, 2>(p, dev_sms, stream, err);
        if (p.n <= 96) return launch_fast<double, 3>(p, dev_sms, stream, err);
        if (p.n <= 128) return launch_fast<double, 4>(p, dev_sms, stream, err);
    }
    return launch_t<double>(p, dev_sms, stream, err);
}
}  // namespace gs

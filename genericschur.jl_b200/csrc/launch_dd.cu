// Kernel instantiations for element kind dd (one translation unit per kind keeps builds parallel).
// Schur requests go to the two-kernel path (gehrd.cuh + fastqr.cuh, n <= 96); Hessenberg-only requests to the
// stage A kernel (gehrd.cuh, factor output); anything forced by GSCHUR_FORCE_GENERIC to the block-synchronous
// single-kernel path (batched.cuh).
#include <cstdlib>
#include "fastqr.cuh"
namespace gs {
int launch_dd(const BatchedParams& p, int dev_sms, cudaStream_t stream, std::string* err) {
    static const bool force_generic = std::getenv("GSCHUR_FORCE_GENERIC") != nullptr;
    if (!force_generic && p.mode == MODE_SCHUR) {
        if (p.n <= 32) return launch_fast<dd_t, 1>(p, dev_sms, stream, err);
        if (p.n <= 64) return launch_fast<dd_t, 2>(p, dev_sms, stream, err);
        if (p.n <= 96) return launch_fast<dd_t, 3>(p, dev_sms, stream, err);
    }
    // Hessenberg-only requests: the stage A kernel in its factor-output mode (same size limits as the Schur path)
    if (!force_generic && p.mode == MODE_HESSENBERG) return launch_gehrd<dd_t, 64>(p, dev_sms, stream, err);
    return launch_t<dd_t>(p, dev_sms, stream, err);
}
}  // namespace gs

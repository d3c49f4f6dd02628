// Kernel instantiations for element kind f64 (one translation unit per kind keeps builds parallel).
// Schur requests go to the three-stage path (qr3.cuh, n <= 64; GSCHUR_QR=fused forces the fused stage B) or the
// two-kernel path (gehrd.cuh + fastqr.cuh, n <= 128); Hessenberg-only requests to the
// stage A kernel (gehrd.cuh, factor output); anything forced by GSCHUR_FORCE_GENERIC to the block-synchronous
// single-kernel path (batched.cuh).
#include <cstdlib>
#include "qr3.cuh"
namespace gs {
int launch_f64(const BatchedParams& p, int dev_sms, cudaStream_t stream, std::string* err) {
    static const bool force_generic = std::getenv("GSCHUR_FORCE_GENERIC") != nullptr;
    // read on every call: the tests switch paths with monkeypatch.setenv
    const char* qrsel = std::getenv("GSCHUR_QR");
    const bool fused = qrsel && qrsel[0] == 'f';
    if (!force_generic && p.mode == MODE_SCHUR) {
        if (!fused && p.n <= 32) return launch_fast3<double, 1>(p, dev_sms, stream, err);
        if (!fused && p.n <= 64) return launch_fast3<double, 2>(p, dev_sms, stream, err);
        if (p.n <= 32) return launch_fast<double, 1>(p, dev_sms, stream, err);
        if (p.n <= 64) return launch_fast<double, 2>(p, dev_sms, stream, err);
        if (p.n <= 96) return launch_fast<double, 3>(p, dev_sms, stream, err);
        if (p.n <= 128) return launch_fast<double, 4>(p, dev_sms, stream, err);
    }
    // Hessenberg-only requests: the stage A kernel in its factor-output mode (same size limits as the Schur path)
    if (!force_generic && p.mode == MODE_HESSENBERG) return launch_gehrd<double, 64>(p, dev_sms, stream, err);
    return launch_t<double>(p, dev_sms, stream, err);
}
}  // namespace gs

// Kernel instantiations for element kind f64 (one translation unit per kind keeps builds parallel).
#include "batched.cuh"
namespace gs {
int launch_f64(const BatchedParams& p, int dev_sms, cudaStream_t stream, std::string* err) {
    return launch_t<double>(p, dev_sms, stream, err);
}
}  // namespace gs

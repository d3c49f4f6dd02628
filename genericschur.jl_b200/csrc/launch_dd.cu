// Kernel instantiations for element kind dd (one translation unit per kind keeps builds parallel).
#include "batched.cuh"
namespace gs {
int launch_dd(const BatchedParams& p, int dev_sms, cudaStream_t stream, std::string* err) {
    return launch_t<dd_t>(p, dev_sms, stream, err);
}
}  // namespace gs

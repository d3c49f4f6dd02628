// Development harness: compiles ONLY the ComplexF64 / Float64 n <= 64 three-stage kernels (fast rebuilds, optional
// -DGS_QR_PROFILE clock64 instrumentation) and times the stages on random matrices.
//   nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -I genericschur.jl_b200/csrc \
//        scripts/dev/qr3_dev.cu genericschur.jl_b200/csrc/launch_sizes.cu -o scripts/dev/qr3_dev
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>
#include <complex>
#include <cmath>
#include "qr3.cuh"
using namespace gs;
#ifndef DEV_REAL
typedef cx<double> ET;
#else
typedef double ET;
#endif
int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 64;
    const long long batch = argc > 2 ? atoll(argv[2]) : 16384;
    const int reps = argc > 3 ? atoi(argv[3]) : 3;
    const size_t el = (size_t)n * n * batch;
    std::vector<ET> hA(el);
    std::mt19937_64 rng(1234);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    double* pa = reinterpret_cast<double*>(hA.data());
    for (size_t i = 0; i < el * (sizeof(ET) / 8); ++i) pa[i] = U(rng);
    ET *dA0, *dA, *dZ;
    cx<double>* dw;
    int* dinfo;
    unsigned* dstats;
    unsigned long long* dctr;
    cudaMalloc(&dA0, el * sizeof(ET));
    cudaMalloc(&dA, el * sizeof(ET));
    cudaMalloc(&dZ, el * sizeof(ET));
    cudaMalloc(&dw, (size_t)n * batch * 16);
    cudaMalloc(&dinfo, batch * 4);
    cudaMalloc(&dstats, batch * 16);
    cudaMalloc(&dctr, 64);
    cudaMemcpy(dA0, hA.data(), el * sizeof(ET), cudaMemcpyHostToDevice);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    cudaMemPool_t mp;
    cudaDeviceGetDefaultMemPool(&mp, 0);
    unsigned long long thr = ~0ULL;
    cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &thr);
    stage_timing_enable(true);
    for (int r = 0; r < reps; ++r) {
        cudaMemcpy(dA, dA0, el * sizeof(ET), cudaMemcpyDeviceToDevice);
        cudaMemset(dctr, 0, 64);
        BatchedParams p{};
        p.A = dA; p.Z = dZ; p.w = dw; p.strideA = p.strideZ = (long long)n * n; p.batch = batch; p.lda = p.ldz = n; p.n = n;
        p.scale = 1; p.maxiter = 0; p.mode = MODE_SCHUR; p.info = dinfo; p.stats = dstats; p.counter = dctr;
        std::string err;
        stage_timing_begin_call();
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        int rc = (n <= 32) ? launch_fast3<ET, 1>(p, prop.multiProcessorCount, 0, &err) : launch_fast3<ET, 2>(p, prop.multiProcessorCount, 0, &err);
        cudaEventRecord(e1);
        cudaError_t ce = cudaDeviceSynchronize();
        float ms = 0, a = 0, b = 0, c = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        stage_timing_read(&a, &b, &c);
        printf("rc=%d %s %s | n=%d batch=%lld: total %.2f ms  A %.2f  B %.2f  C %.2f  -> %.0f matrices/s\n", rc, err.c_str(),
               cudaGetErrorString(ce), n, batch, ms, a, b, c, batch / ms * 1e3);
    }
    // ---- residuals of a few matrices on the host (plain Float64): ||A - Z T Z'||_F / (n eps ||A||_F), ||Z'Z - I||_F / (n eps) ----
    {
        const int nchk = batch < 4 ? (int)batch : 4;
        const size_t m = (size_t)n * n;
        std::vector<ET> T(m), Z(m);
        typedef std::complex<double> cd;
        for (int c = 0; c < nchk; ++c) {
            const long long b = (c == 0) ? 0 : (c == 1 ? batch - 1 : (batch / 2 + c));
            cudaMemcpy(T.data(), dA + b * m, m * sizeof(ET), cudaMemcpyDeviceToHost);
            cudaMemcpy(Z.data(), dZ + b * m, m * sizeof(ET), cudaMemcpyDeviceToHost);
            auto get = [&](const ET* M, int i, int j) -> cd {
#ifndef DEV_REAL
                return cd(M[i + (size_t)j * n].re, M[i + (size_t)j * n].im);
#else
                return cd(M[i + (size_t)j * n], 0.0);
#endif
            };
            std::vector<cd> ZT(m), R(m);
            double na = 0, nr = 0, no = 0, low = 0;
            for (int i = 0; i < n; ++i)
                for (int j = 0; j < n; ++j) {
                    cd s = 0;
                    for (int k = 0; k < n; ++k) s += get(Z.data(), i, k) * get(T.data(), k, j);
                    ZT[i + (size_t)j * n] = s;
                }
            for (int i = 0; i < n; ++i)
                for (int j = 0; j < n; ++j) {
                    cd s = 0, o = 0;
                    for (int k = 0; k < n; ++k) {
                        s += ZT[i + (size_t)k * n] * std::conj(get(Z.data(), j, k));
                        o += std::conj(get(Z.data(), k, i)) * get(Z.data(), k, j);
                    }
                    const cd a = get(hA.data() + b * m, i, j);
                    na += std::norm(a);
                    nr += std::norm(a - s);
                    no += std::norm(o - (i == j ? 1.0 : 0.0));
#ifndef DEV_REAL
                    if (i > j) low += std::abs(get(T.data(), i, j));
#else
                    if (i > j + 1) low += std::abs(get(T.data(), i, j));
#endif
                }
            const double eps = 2.220446049250313e-16;
            printf("matrix %lld: backward %.3f  orth %.3f  below-triangle %.1e\n", b, std::sqrt(nr / na) / (n * eps), std::sqrt(no) / (n * eps), low);
        }
    }
    std::vector<unsigned> st(batch * 4);
    std::vector<int> info(batch);
    cudaMemcpy(st.data(), dstats, batch * 16, cudaMemcpyDeviceToHost);
    cudaMemcpy(info.data(), dinfo, batch * 4, cudaMemcpyDeviceToHost);
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    long long bad = 0;
    for (long long i = 0; i < batch; ++i) { s0 += st[4 * i]; s1 += st[4 * i + 1]; s2 += st[4 * i + 2]; s3 += st[4 * i + 3]; bad += info[i] != 0; }
#ifdef GS_QR_PROFILE
    printf("profile (cycles per matrix): total %.0f, step loops %.0f, sweep prologue %.0f; steps %.0f -> %.1f cycles/step in loops, %.1f overall; bad=%lld\n",
           64 * s0 / batch, 64 * s3 / batch, 64 * s2 / batch, s1 / batch, 64 * s3 / s1, 64 * s0 / s1, bad);
#else
    printf("stats per matrix: sweeps %.1f steps %.1f exceptional %.2f iterations %.1f; bad=%lld\n", s0 / batch, s1 / batch, s2 / batch, s3 / batch, bad);
#endif
    return 0;
}

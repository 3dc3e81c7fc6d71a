// Diagnostics: measures the FP64 FMA throughput of the device the roofline is quoted against
// (MEASURED_PEAKS.json has no FP64 figure; SURVEY.md section 8(d) asks for a DFMA-chain microbenchmark).
#include "../../include/dwbc_b200.h"
#include "cuda_compat.h"

#ifndef DWBC_CPU_EMUL
namespace {
__global__ void dfma_chain_kernel(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-9, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
}  // namespace

extern "C" int dwbc_probe_fp64(int device, double seconds, double *tflops) {
    if (!tflops) return DWBC_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return DWBC_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return DWBC_ERR_CUDA;
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 14;
    double *out = nullptr;
    if (cudaMalloc(&out, sizeof(double) * blocks * threads) != cudaSuccess) return DWBC_ERR_CUDA;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    dfma_chain_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-7);  // warm-up
    cudaDeviceSynchronize();
    double best = 0.0, elapsed = 0.0;
    int reps = 0;
    while (elapsed < seconds * 1e3 || reps < 3) {
        cudaEventRecord(e0);
        dfma_chain_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-7);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double tf = 2.0 * 8.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
        best = tf > best ? tf : best;
        elapsed += ms;
        reps++;
        if (reps > 2000) break;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *tflops = best;
    return cudaGetLastError() == cudaSuccess ? DWBC_OK : DWBC_ERR_CUDA;
}
#else
extern "C" int dwbc_probe_fp64(int, double, double *tflops) {
    if (tflops) *tflops = 0.0;
    return DWBC_ERR_CUDA;  // there is no device in the emulation build
}
#endif

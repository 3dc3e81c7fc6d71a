// Microbenchmark (diagnostics only): FP64 throughput of mma.sync m8n8k4 / m16n8k8 (DMMA) against the DFMA pipe, and a
// register-fragment GEMM of the size the WBC kernels use (40x40x40 from shared memory), to decide whether the small dense
// products should move to DMMA (BASELINE.json north_star: "only if ncu shows they beat the FP64 pipe").
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/dmma_probe tools/dmma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double *d, const double *a, const double *b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

__global__ void k_dfma(double *out, int iters, double a, double b) {
    double x[8];
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 1e-9 + i;
    for (int i = 0; i < iters; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) x[j] = fma(x[j], a, b);
    double s = 0;
    for (int i = 0; i < 8; i++) s += x[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dmma884(double *out, int iters, double a, double b) {
    double c[8][2];
    for (int i = 0; i < 8; i++) c[i][0] = threadIdx.x * 1e-9 + i, c[i][1] = i;
    for (int i = 0; i < iters; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) dmma884(c[j][0], c[j][1], a, b);
    double s = 0;
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dmma1688(double *out, int iters, double a, double b) {
    double c[4][4], av[4] = {a, a, a, a}, bv[2] = {b, b};
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) c[i][j] = threadIdx.x * 1e-9 + i + j;
    for (int i = 0; i < iters; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) dmma1688(c[j], av, bv);
    double s = 0;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) s += c[i][j];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// C (40x40) += A (40x40, column-major ld 41) * B (40x40, column-major ld 41), all in shared memory, 128 threads, repeated.
// scalar: 2x2 register tiles (the product's tgemm); dmma: each warp owns 8-row bands x all 5 column tiles.
constexpr int N = 40, LD = 41;
__global__ void k_gemm_scalar(double *out, int reps) {
    __shared__ double A[LD * N], B[LD * N], C[LD * N];
    for (int e = threadIdx.x; e < LD * N; e += blockDim.x) A[e] = 1e-3 * (e % 7), B[e] = 1e-3 * (e % 5), C[e] = 0;
    __syncthreads();
    for (int r = 0; r < reps; r++) {
        for (int tile = threadIdx.x; tile < 20 * 20; tile += blockDim.x) {
            const int tj = tile / 20, ti = tile - tj * 20;
            double c00 = 0, c01 = 0, c10 = 0, c11 = 0;
#pragma unroll 2
            for (int k = 0; k < N; k++) {
                const double a0 = A[ti + k * LD], a1 = A[ti + 20 + k * LD], b0 = B[k + tj * LD], b1 = B[k + (tj + 20) * LD];
                c00 = fma(a0, b0, c00); c01 = fma(a0, b1, c01); c10 = fma(a1, b0, c10); c11 = fma(a1, b1, c11);
            }
            C[ti + tj * LD] += c00; C[ti + (tj + 20) * LD] += c01; C[ti + 20 + tj * LD] += c10; C[ti + 20 + (tj + 20) * LD] += c11;
        }
        __syncthreads();
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = C[threadIdx.x];
}
__global__ void k_gemm_dmma(double *out, int reps) {
    __shared__ double A[LD * N], B[LD * N], C[LD * N];
    for (int e = threadIdx.x; e < LD * N; e += blockDim.x) A[e] = 1e-3 * (e % 7), B[e] = 1e-3 * (e % 5), C[e] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    for (int r = 0; r < reps; r++) {
        // 25 output tiles of 8x8; a warp takes a row band (5 tiles sharing the A fragment) at a time
        for (int bi = warp; bi < 5; bi += nw) {
            double c[5][2];
#pragma unroll
            for (int j = 0; j < 5; j++) c[j][0] = c[j][1] = 0.0;
#pragma unroll 2
            for (int k0 = 0; k0 < N; k0 += 4) {
                const double a = A[bi * 8 + g + (k0 + t) * LD];
#pragma unroll
                for (int j = 0; j < 5; j++) {
                    const double b = B[k0 + t + (j * 8 + g) * LD];
                    dmma884(c[j][0], c[j][1], a, b);
                }
            }
#pragma unroll
            for (int j = 0; j < 5; j++) {
                C[bi * 8 + g + (j * 8 + 2 * t) * LD] += c[j][0];
                C[bi * 8 + g + (j * 8 + 2 * t + 1) * LD] += c[j][1];
            }
        }
        __syncthreads();
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = C[threadIdx.x];
}

template <class F>
double time_ms(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    f();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int i = 0; i < 5; i++) {
        cudaEventRecord(e0);
        f();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    return best;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 1 << 13;
    double *out;
    cudaMalloc(&out, sizeof(double) * blocks * 256);
    double ms = time_ms([&] { k_dfma<<<blocks, threads>>>(out, iters, 0.999999, 1e-7); });
    printf("{\"dfma_tflops\": %.2f, ", 2.0 * 8 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12);
    ms = time_ms([&] { k_dmma884<<<blocks, threads>>>(out, iters, 0.999999, 1e-7); });
    printf("\"dmma_m8n8k4_tflops\": %.2f, ", 2.0 * 8 * 256.0 * iters * (double)blocks * (threads / 32) / (ms * 1e-3) / 1e12);
    ms = time_ms([&] { k_dmma1688<<<blocks, threads>>>(out, iters, 0.999999, 1e-7); });
    printf("\"dmma_m16n8k8_tflops\": %.2f, ", 2.0 * 4 * 1024.0 * iters * (double)blocks * (threads / 32) / (ms * 1e-3) / 1e12);
    const int gb = p.multiProcessorCount * 5, reps = 2000;
    ms = time_ms([&] { k_gemm_scalar<<<gb, 128>>>(out, reps); });
    printf("\"smem_gemm40_scalar2x2_tflops\": %.2f, ", 2.0 * N * N * N * (double)reps * gb / (ms * 1e-3) / 1e12);
    ms = time_ms([&] { k_gemm_dmma<<<gb, 128>>>(out, reps); });
    printf("\"smem_gemm40_dmma_tflops\": %.2f, ", 2.0 * N * N * N * (double)reps * gb / (ms * 1e-3) / 1e12);
    ms = time_ms([&] { k_gemm_dmma<<<gb, 160>>>(out, reps); });
    printf("\"smem_gemm40_dmma_5warps_tflops\": %.2f}\n", 2.0 * N * N * N * (double)reps * gb / (ms * 1e-3) / 1e12);
    return cudaGetLastError() != cudaSuccess;
}

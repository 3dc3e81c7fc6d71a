// TEST INFRASTRUCTURE -- NOT part of the product and never loaded by it.
//
// A tiny host-side stand-in for the CUDA execution model, so that the *same* kernel sources under
// libdwbc_b200/csrc can be compiled with g++ (-DDWBC_CPU_EMUL) and checked against the oracle in the
// GPU-less build container.  Every CUDA thread of a block becomes a std::thread; __syncwarp /
// __syncthreads are barriers; warp shuffles go through a per-warp exchange slot.  It only supports what
// the kernels use (1-D grids/blocks, full-mask warp primitives, dynamic shared memory) and is orders of
// magnitude slower than a GPU: tests run a handful of states through it.
//
// The product library (libdwbc_b200.so, built by nvcc) does not contain any of this; if CUDA is
// missing the product fails loudly (libdwbc_b200/_lib.py).
#pragma once
#ifndef DWBC_CPU_EMUL
#error "cuda_emul.h is only for the emulation build"
#endif
#include <pthread.h>
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ thread_local /* only `extern __shared__` is used; mapped through emul_smem() */

struct dim3 {
    unsigned x = 1, y = 1, z = 1;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3_ {
    unsigned x, y, z;
};
typedef void *cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize };

namespace emul {
// Blocking barrier (futex wait, no spinning): a block is 128 host threads on a handful of cores, and std::barrier's
// spin-then-yield phase turned every __syncthreads into a scheduler storm (sys time 3x the user time, run time at the mercy of
// whatever else the host was doing).
struct Barrier {
    pthread_barrier_t b;
    explicit Barrier(int n) { pthread_barrier_init(&b, nullptr, (unsigned)n); }
    ~Barrier() { pthread_barrier_destroy(&b); }
    Barrier(const Barrier &) = delete;
    void arrive_and_wait() { pthread_barrier_wait(&b); }
};
struct WarpCtx {
    Barrier bar;
    unsigned long long slot[32];
    explicit WarpCtx(int n) : bar(n) {}
};
struct BlockCtx {
    Barrier bar;
    std::vector<std::unique_ptr<WarpCtx>> warps;
    std::vector<char> smem;
    BlockCtx(int nthreads, size_t smem_bytes) : bar(nthreads), smem(smem_bytes + 64) {
        for (int w = 0; w < (nthreads + 31) / 32; w++) warps.emplace_back(new WarpCtx(std::min(32, nthreads - 32 * w)));
    }
};
struct ThreadCtx {
    uint3_ threadIdx, blockIdx;
    dim3 blockDim, gridDim;
    BlockCtx *block = nullptr;
    WarpCtx *warp = nullptr;
    int lane = 0;
};
inline thread_local ThreadCtx tctx;

template <class F>
void launch(dim3 grid, dim3 block, size_t smem, F &&body) {
    for (unsigned b = 0; b < grid.x; b++) {
        BlockCtx bc((int)block.x, smem);
        std::vector<std::thread> th;
        for (unsigned t = 0; t < block.x; t++) {
            th.emplace_back([&, t]() {
                tctx.threadIdx = {t, 0, 0};
                tctx.blockIdx = {b, 0, 0};
                tctx.blockDim = block;
                tctx.gridDim = grid;
                tctx.block = &bc;
                tctx.warp = bc.warps[t / 32].get();
                tctx.lane = (int)(t % 32);
                body();
            });
        }
        for (auto &x : th) x.join();
    }
}
template <class T>
inline T shfl(T v, int src) {
    static_assert(sizeof(T) <= 8, "shfl payload");
    WarpCtx *w = tctx.warp;
    unsigned long long raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    w->slot[tctx.lane] = raw;
    w->bar.arrive_and_wait();
    raw = w->slot[src & 31];
    w->bar.arrive_and_wait();
    T out;
    std::memcpy(&out, &raw, sizeof(T));
    return out;
}
}  // namespace emul

#define threadIdx (emul::tctx.threadIdx)
#define blockIdx (emul::tctx.blockIdx)
#define blockDim (emul::tctx.blockDim)
#define gridDim (emul::tctx.gridDim)

inline void __syncwarp(unsigned = 0xffffffffu) { emul::tctx.warp->bar.arrive_and_wait(); }
inline void __syncthreads() { emul::tctx.block->bar.arrive_and_wait(); }
template <class T>
inline T __shfl_sync(unsigned, T v, int src, int = 32) { return emul::shfl(v, src); }
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return emul::shfl(v, emul::tctx.lane ^ m); }
template <class T>
inline T __shfl_down_sync(unsigned, T v, int d, int = 32) { return emul::shfl(v, emul::tctx.lane + d > 31 ? emul::tctx.lane : emul::tctx.lane + d); }
inline unsigned __ballot_sync(unsigned, int pred) {
    unsigned bit = pred ? (1u << emul::tctx.lane) : 0u, acc = 0;
    for (int i = 0; i < 32; i++) acc |= emul::shfl(bit, i);
    return acc;
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, !pred) == 0; }
inline long long clock64() { return 0; }
inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline int __ffs(unsigned v) { return v ? __builtin_ctz(v) + 1 : 0; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline double __longlong_as_double(long long v) {
    double d;
    std::memcpy(&d, &v, 8);
    return d;
}
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
inline double __ldg(const double *p) { return *p; }
inline int __ldg(const int *p) { return *p; }

inline char *emul_smem() {
    uintptr_t p = (uintptr_t)emul::tctx.block->smem.data();
    return (char *)((p + 15) & ~(uintptr_t)15);
}

// ---- runtime API stand-ins ----
inline cudaError_t cudaMalloc(void **p, size_t n) {
    *p = std::calloc(1, n ? n : 1);
    return *p ? cudaSuccess : 2;
}
inline cudaError_t cudaFree(void *p) {
    std::free(p);
    return cudaSuccess;
}
inline cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
inline cudaError_t cudaFreeHost(void *p) { return cudaFree(p); }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) {
    std::memcpy(d, s, n);
    return cudaSuccess;
}
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) {
    std::memcpy(d, s, n);
    return cudaSuccess;
}
inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dpitch, const void *s, size_t spitch, size_t width, size_t height, cudaMemcpyKind,
                                     cudaStream_t = nullptr) {
    for (size_t r = 0; r < height; r++) std::memcpy((char *)d + r * dpitch, (const char *)s + r * spitch, width);
    return cudaSuccess;
}
inline cudaError_t cudaMemset(void *d, int v, size_t n) {
    std::memset(d, v, n);
    return cudaSuccess;
}
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr) { return cudaMemset(d, v, n); }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDevice(int *d) {
    *d = 0;
    return cudaSuccess;
}
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
typedef int cudaEvent_t;
inline cudaError_t cudaEventCreate(cudaEvent_t *e) {
    *e = 0;
    return cudaSuccess;
}
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) {
    *s = nullptr;
    return cudaSuccess;
}
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) {
    *ms = 0.f;
    return cudaSuccess;
}
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
template <class K>
inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return cudaSuccess; }
struct cudaDeviceProp {
    int multiProcessorCount = 4;
    size_t sharedMemPerBlockOptin = 227 * 1024;
    char name[64] = "cpu-emulation";
};
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
    *p = cudaDeviceProp();
    return cudaSuccess;
}

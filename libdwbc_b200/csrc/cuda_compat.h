// Thin portability seam used by every kernel source in this directory.
//
// Normal build (nvcc, sm_100a): plain CUDA.  With -DDWBC_CPU_EMUL (tests only, see
// tests/emul/cuda_emul.h) the same sources compile with g++ so the kernel logic can be checked
// against the oracle in a container without a GPU.  The emulation is never part of the product.
#pragma once

#ifdef DWBC_CPU_EMUL
#include "cuda_emul.h"
#define DWBC_DYN_SMEM(ptr) unsigned char *ptr = (unsigned char *)emul_smem()
#define DWBC_LAUNCH(kernel, grid, block, smem, stream, ...) \
    emul::launch(dim3(grid), dim3(block), (size_t)(smem), [&]() { kernel(__VA_ARGS__); })
#define DWBC_UNROLL(n)
#else
#include <cuda_runtime.h>
#define DWBC_DYN_SMEM(ptr)                                  \
    extern __shared__ __align__(16) unsigned char _dyn_smem[]; \
    unsigned char *ptr = _dyn_smem
#define DWBC_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__)
#define DWBC_PRAGMA(x) _Pragma(#x)
#define DWBC_UNROLL(n) DWBC_PRAGMA(unroll n)
#endif

#define DWBC_FULL 0xffffffffu

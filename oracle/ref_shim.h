/*
 * ref_shim.h — TEST INFRASTRUCTURE.  Lets g++ compile the reference's own
 * /root/reference/src/cuda_kernel.cu as HOST code, where it lies, so that the real
 * t1ha2_atonce (cuda_kernel.cu:196-246) and the real cuda_kmer_t1ha2 kernel body
 * (cuda_kernel.cu:250-321) can be executed on the CPU to pin the oracle.
 * Nothing from the reference is copied into this repository; see oracle/Makefile.
 */
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>

#define __device__
#define __global__
#define __shared__ thread_local
struct hgref_dim3 { unsigned x, y, z; };
extern thread_local hgref_dim3 blockIdx, blockDim, threadIdx;
static inline void __syncthreads() {}
using std::min;
using std::max;

/* g++ rejects `extern "C" static`, which nvcc accepts in the reference file; every
 * system header is already included above, so dropping the keyword from here on only
 * affects the reference translation unit (set by the Makefile for that file alone). */
#ifdef HGREF_DROP_STATIC
#define static
#endif

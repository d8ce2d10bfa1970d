// am_common.cuh -- shared helpers for the sm_100a kernels of libaccessmath_b200.so
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>

#define AM_OK 0
#define AM_ERR_CUDA 1
#define AM_ERR_ARG 2
#define AM_ERR_CAPACITY 3

#define AM_CUDA(call)                                                                         \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            fprintf(stderr, "[accessmath_b200] CUDA error %s at %s:%d: %s\n", #call, __FILE__, \
                    __LINE__, cudaGetErrorString(e_));                                        \
            return AM_ERR_CUDA;                                                               \
        }                                                                                     \
    } while (0)

static inline int am_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// words per row of a bit-packed mask, padded to a multiple of 4 words (128-bit loads)
static inline __host__ __device__ int am_words_per_row_impl(int width) { return (((width + 31) >> 5) + 3) & ~3; }

__device__ __forceinline__ int warp_incl_scan(int v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// Exclusive scan across a thread block (blockDim.x multiple of 32, <= 1024). `smem` holds 33 ints.
// Returns the exclusive prefix of v for this thread; *total receives the block sum.
__device__ __forceinline__ int block_excl_scan(int v, int* smem, int* total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    int inc = warp_incl_scan(v);
    __syncthreads();                       // protect smem reuse between consecutive calls
    if (lane == 31) smem[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int w = lane < nw ? smem[lane] : 0;
        int winc = warp_incl_scan(w);
        smem[lane] = winc - w;
        if (lane == 31) smem[32] = winc;
    }
    __syncthreads();
    *total = smem[32];
    return inc - v + smem[wid];
}

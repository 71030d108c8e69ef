// Shared helpers for libdeepsent (sm_100a).  Error convention: see include/deepsent.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include "../../include/deepsent.h"

namespace ds {

std::string& last_error();
int fail(const char* fmt, ...);
extern int g_debug[16];

#define DS_CUDA(expr)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess)                                                                 \
      return ds::fail("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
  } while (0)

#define DS_LAUNCH_CHECK()                                                                    \
  do {                                                                                       \
    ++ds::g_debug[15]; /* kernel-launch counter (ds_debug_get(15)) */                        \
    cudaError_t _e = cudaPeekAtLastError();                                                  \
    if (_e != cudaSuccess) {                                                                 \
      cudaGetLastError();                                                                    \
      return ds::fail("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
    }                                                                                        \
  } while (0)

#define DS_REQUIRE(cond, msg)                                              \
  do {                                                                     \
    if (!(cond)) return ds::fail("%s:%d %s (%s)", __FILE__, __LINE__, msg, #cond); \
  } while (0)

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace ds

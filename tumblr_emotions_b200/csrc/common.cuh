// Shared helpers for libdeepsent (sm_100a).  Error convention: see include/deepsent.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <utility>
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

// ---- programmatic dependent launch ----
// Every kernel of the library opens with pdl_trigger() (the next kernel of the stream may be scheduled as soon as all CTAs of
// this one have started) and executes pdl_wait() before its first access to global memory (blocks until the kernels it depends on
// have completed and their writes are visible), and every launch goes through ds::launch, which allows the overlap.  What
// overlaps is the launch latency and the data-independent prologue (barrier init, TMEM allocation, tensor-map prefetch) of kernel
// N+1 with the tail of kernel N; the data dependences are exactly those of a serial stream.  g_pdl (ds_dependent_launch, DS_PDL_*
// bits of include/deepsent.h) says which launches carry the attribute; without it the two instructions are no-ops.
extern int g_pdl;
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_wait(); pdl_trigger(); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_as(int pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}
template <typename... KArgs, typename... Args>
static inline cudaError_t launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  return launch_as(g_pdl & 2, kern, grid, block, smem, st, std::forward<Args>(args)...);
}

// halo-tile 3x3 kernel (conv_halo.cu), dispatched from ds_conv_bf16x3
bool conv3x3_halo_fits(int64_t batch, int64_t h, int64_t w, int64_t cin, int64_t n);
bool conv3x3_halo_pays(int64_t batch, int64_t h, int64_t w, int64_t cin, int64_t n);
int conv3x3_halo_launch(const uint16_t* a_hi, const uint16_t* a_lo, int64_t lda, int64_t batch, int64_t h, int64_t w, int64_t cin,
                        const uint16_t* bt_hi, const uint16_t* bt_lo, int64_t ldb, int64_t n, float* c, int64_t ldc,
                        const float* scale, const float* bias, double* stats, int flags, void* stream,
                        uint16_t* y_hi = nullptr, uint16_t* y_lo = nullptr, int64_t ldy = 0);
static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// ---- split-bf16: x ~= hi + lo with hi = bf16(x), lo = bf16(x - hi) (16 significant bits) ----
__device__ __forceinline__ uint32_t bf16_rn_bits(float x) {
  uint32_t r;
  asm("{\n\t.reg .b16 h;\n\tcvt.rn.bf16.f32 h, %1;\n\tmov.b32 %0, {h, 0};\n\t}" : "=r"(r) : "f"(x));
  return r;   // low 16 bits
}
__device__ __forceinline__ void split_bf16(float x, uint32_t& hi, uint32_t& lo) {
  hi = bf16_rn_bits(x);
  lo = bf16_rn_bits(x - __uint_as_float(hi << 16));
}
__device__ __forceinline__ float merge_bf16(uint32_t hi, uint32_t lo) {
  return __uint_as_float(hi << 16) + __uint_as_float(lo << 16);
}
// 4 consecutive values -> 8 bytes per plane
__device__ __forceinline__ void store4_split(uint16_t* hi, uint16_t* lo, const float v[4]) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) split_bf16(v[i], h[i], l[i]);
  *reinterpret_cast<uint2*>(hi) = make_uint2(h[0] | (h[1] << 16), h[2] | (h[3] << 16));
  *reinterpret_cast<uint2*>(lo) = make_uint2(l[0] | (l[1] << 16), l[2] | (l[3] << 16));
}
__device__ __forceinline__ void load4_split(const uint16_t* hi, const uint16_t* lo, float v[4]) {
  const uint2 h = *reinterpret_cast<const uint2*>(hi), l = *reinterpret_cast<const uint2*>(lo);
  v[0] = merge_bf16(h.x & 0xffffu, l.x & 0xffffu); v[1] = merge_bf16(h.x >> 16, l.x >> 16);
  v[2] = merge_bf16(h.y & 0xffffu, l.y & 0xffffu); v[3] = merge_bf16(h.y >> 16, l.y >> 16);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace ds

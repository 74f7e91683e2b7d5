// Shared host/device helpers of libmpl_b200 (B200 / sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdio>

#include "../../include/mpl_b200.h"

namespace mpl {

// thread-local error text behind mpl_last_error()
void set_error(const char* fmt, ...);

#define MPL_CUDA(expr)                                                                               \
  do {                                                                                               \
    cudaError_t e__ = (expr);                                                                        \
    if (e__ != cudaSuccess) {                                                                        \
      ::mpl::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return MPL_ERR_CUDA;                                                                           \
    }                                                                                                \
  } while (0)

#define MPL_LAUNCH_CHECK()                                                                                \
  do {                                                                                                    \
    cudaError_t e__ = cudaGetLastError();                                                                 \
    if (e__ != cudaSuccess) {                                                                             \
      ::mpl::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
      return MPL_ERR_CUDA;                                                                                \
    }                                                                                                     \
  } while (0)

#define MPL_TRY(expr)          \
  do {                         \
    int s__ = (expr);          \
    if (s__ != MPL_OK) return s__; \
  } while (0)

constexpr int kNumSMs = 148;  // B200

enum Act { ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2 };

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// nn.GELU() default: exact erf form (multiview_mpl.py:22,27)
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// erf-form GELU (nn.GELU() default, multiview_mpl.py:22,27) for results that are rounded to fp16 / bf16 right after:
// 0.5 x (1 + tanh(x (c0 + c1 x^2))) with (c0, c1) fitted to the ERF form (max |error| 2.7e-4 over all x, at |x| ~ 2 where
// the bf16 half-ulp is 4e-3; tanh.approx adds <= 2^-11 relative).  6 FMA-pipe instructions + one MUFU.
__device__ __forceinline__ float gelu_tanh_fit(float x) {
  const float x2 = x * x;
  const float u = x * fmaf(x2, 0.03470089f, 0.80015708f);
  float th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(u));
  const float hx = 0.5f * x;
  return fmaf(hx, th, hx);
}

// the same fit on a packed pair of fp16 values (5 packed FMA-pipe instructions + the MUFU pair)
__device__ __forceinline__ uint32_t gelu_tanh_fit_h2(uint32_t xu) {
  const __half2 x = *reinterpret_cast<const __half2*>(&xu);
  const __half2 x2 = __hmul2(x, x);
  const __half2 t = __hfma2(x2, __float2half2_rn(0.03470089f), __float2half2_rn(0.80015708f));
  const __half2 u = __hmul2(x, t);
  uint32_t thu, uu = *reinterpret_cast<const uint32_t*>(&u);
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(thu) : "r"(uu));
  const __half2 th = *reinterpret_cast<const __half2*>(&thu);
  const __half2 hx = __hmul2(x, __float2half2_rn(0.5f));
  const __half2 r = __hfma2(hx, th, hx);
  return *reinterpret_cast<const uint32_t*>(&r);
}
__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == ACT_GELU) return gelu_erf(x);
  if (act == ACT_RELU) return fmaxf(x, 0.0f);
  return x;
}

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

}  // namespace mpl

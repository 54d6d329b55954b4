// common.cuh -- shared host/device helpers for libmode_b200 (sm_100a only).
#pragma once
#include <algorithm>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mode_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libmode_b200 is written for sm_100a (B200) only"
#endif

namespace mode {

// thread-local error message + process-wide launch counter (defined in api.cu)
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define MODE_CHECK_ARG(cond, ...)                \
  do {                                           \
    if (!(cond)) {                               \
      ::mode::set_error(__VA_ARGS__);            \
      return MODE_EINVAL;                        \
    }                                            \
  } while (0)

#define MODE_CHECK_LAUNCH(name)                                                      \
  do {                                                                               \
    cudaError_t e__ = cudaGetLastError();                                            \
    if (e__ != cudaSuccess) {                                                        \
      ::mode::set_error("%s: CUDA error: %s", name, cudaGetErrorString(e__));        \
      return MODE_ECUDA;                                                             \
    }                                                                                \
    ::mode::count_launch();                                                          \
  } while (0)

#define MODE_CHECK_CUDA(expr, name)                                                  \
  do {                                                                               \
    cudaError_t e__ = (expr);                                                        \
    if (e__ != cudaSuccess) {                                                        \
      ::mode::set_error("%s: CUDA error: %s", name, cudaGetErrorString(e__));        \
      return MODE_ECUDA;                                                             \
    }                                                                                \
  } while (0)

constexpr int kNumSMs = 148;  // B200

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// cudaFuncSetAttribute is per device: launchers keep their high-water marks per (thread, device)
constexpr int kMaxDevices = 64;
static inline int current_device() {
  int d = 0;
  cudaGetDevice(&d);
  return (d >= 0 && d < kMaxDevices) ? d : 0;
}

__device__ __forceinline__ float bf16_bits_to_float(uint16_t b) { return __uint_as_float(((uint32_t)b) << 16); }
__device__ __forceinline__ uint16_t float_to_bf16_bits(float f) {
  __nv_bfloat16 h = __float2bfloat16_rn(f);
  return *reinterpret_cast<uint16_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// 16-bit storage formats of the tensor-core path: bf16 (BASELINE's named dtype) or fp16 (same speed, 3 more mantissa bits)
constexpr int kFmtBF16 = 0, kFmtFP16 = 1;
template <int FMT>
__device__ __forceinline__ void unpack2(uint32_t v, float& lo, float& hi) {
  if (FMT == kFmtBF16) {
    lo = __uint_as_float(v << 16);
    hi = __uint_as_float(v & 0xFFFF0000u);
  } else {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&v));
    lo = f.x, hi = f.y;
  }
}
template <int FMT>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  if (FMT == kFmtBF16) return pack_bf16x2(lo, hi);
  // fp16 storage saturates instead of overflowing to inf (|x| > 65504 -> +-65504; NaN stays NaN): one F2FP.SATFINITE, same cost
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint16_t float_to_h16_bits(float f, int fmt) {
  if (fmt == kFmtBF16) return float_to_bf16_bits(f);
  const __half h = __float2half_rn(f);
  return *reinterpret_cast<const uint16_t*>(&h);
}
__device__ __forceinline__ float h16_bits_to_float(uint16_t b, int fmt) {
  if (fmt == kFmtBF16) return bf16_bits_to_float(b);
  return __half2float(*reinterpret_cast<const __half*>(&b));
}

// packed fp32 pairs (sm_100 FFMA2 / FADD2): two lanes per instruction
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0, float c1) {
  asm("{\n\t.reg .b64 a, b, c, d;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\tmov.b64 c, {%6, %7};\n\tfma.rn.f32x2 d, a, b, c;\n\tmov.b64 {%0, %1}, d;\n\t}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}
__device__ __forceinline__ void fadd2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm("{\n\t.reg .b64 a, b, d;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\tadd.rn.f32x2 d, a, b;\n\tmov.b64 {%0, %1}, d;\n\t}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
// max(x, 0) on a packed 16-bit pair (ReLU after rounding == rounding after ReLU: rounding is monotonic and keeps 0)
template <int FMT>
__device__ __forceinline__ uint32_t relu2(uint32_t v) {
  if (FMT == kFmtBF16) {
    const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&v), __float2bfloat162_rn(0.f));
    return *reinterpret_cast<const uint32_t*>(&r);
  }
  const __half2 r = __hmax2(*reinterpret_cast<const __half2*>(&v), __float2half2_rn(0.f));
  return *reinterpret_cast<const uint32_t*>(&r);
}

// streaming 128-bit accesses: data touched once should not pollute L1
__device__ __forceinline__ uint4 ld_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_na_v4(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

}  // namespace mode

// Shared host/device helpers for the selftok_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

namespace stk {

// ---- host-side error plumbing (status codes mirror include/selftok_b200.h) ---------------------------------
void set_error(const std::string& msg);
#define STK_CUDA(expr)                                                                            \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      ::stk::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " @" + __FILE__ + ":" + \
                       std::to_string(__LINE__));                                                 \
      return -5;                                                                                  \
    }                                                                                             \
  } while (0)
#define STK_CHECK(cond, code, msg)                                                                \
  do {                                                                                            \
    if (!(cond)) {                                                                                \
      ::stk::set_error(std::string(msg) + " [" #cond "] @" + __FILE__ + ":" + std::to_string(__LINE__)); \
      return (code);                                                                              \
    }                                                                                             \
  } while (0)
#define STK_TRY(expr)                                                                             \
  do {                                                                                            \
    int _s = (expr);                                                                              \
    if (_s != 0) return _s;                                                                       \
  } while (0)

extern thread_local int64_t g_launch_count;   // bumped by every launch wrapper
inline void count_launch() { ++g_launch_count; }

// ---- device helpers ------------------------------------------------------------------------------------------
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

// torch.nn.GELU(approximate="tanh"): 0.5*x*(1+tanh(u)), u = sqrt(2/pi)*(x+0.044715*x^3)
__device__ __forceinline__ float gelu_tanh(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float inner = k0 * (x + k1 * x * x * x);
  return 0.5f * x * (1.0f + tanhf(inner));
}
// Same function as x / (1 + exp(-2u)) with the fast exp/divide intrinsics (~1e-6 relative): used by the tensor-core
// epilogues, whose operands are rounded to 16 bits right afterwards.
__device__ __forceinline__ float gelu_tanh_fast(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  const float u = k0 * (x + k1 * x * x * x);
  return __fdividef(x, 1.0f + __expf(-2.0f * u));
}
// GELU-tanh on a float4 with Blackwell's packed fp32 pipe (fma/mul/add .f32x2: two lanes per instruction):
// y = x / (1 + 2^(c2 * x * (1 + k1 x^2))),  c2 = -2 log2(e) sqrt(2/pi); ex2 / rcp stay scalar MUFU ops.
__device__ __forceinline__ float4 gelu_tanh_fast4(float4 v) {
  const float k1 = 0.044715f, c2 = -2.0f * 1.4426950408889634f * 0.7978845608028654f;
  float4 y;
  asm("{\n\t"
      ".reg .b64 xa, xb, ta, tb, ka, ca, one;\n\t"
      ".reg .f32 e0, e1, e2, e3;\n\t"
      "mov.b64 xa, {%4, %5};\n\t"
      "mov.b64 xb, {%6, %7};\n\t"
      "mov.b64 ka, {%8, %8};\n\t"
      "mov.b64 ca, {%9, %9};\n\t"
      "mov.b64 one, {%10, %10};\n\t"
      "mul.f32x2 ta, xa, xa;\n\t"            // x^2
      "mul.f32x2 tb, xb, xb;\n\t"
      "fma.rn.f32x2 ta, ta, ka, one;\n\t"    // 1 + k1 x^2
      "fma.rn.f32x2 tb, tb, ka, one;\n\t"
      "mul.f32x2 ta, ta, xa;\n\t"            // x (1 + k1 x^2)
      "mul.f32x2 tb, tb, xb;\n\t"
      "mul.f32x2 ta, ta, ca;\n\t"            // exponent (base 2)
      "mul.f32x2 tb, tb, ca;\n\t"
      "mov.b64 {e0, e1}, ta;\n\t"
      "mov.b64 {e2, e3}, tb;\n\t"
      "ex2.approx.ftz.f32 e0, e0;\n\t"
      "ex2.approx.ftz.f32 e1, e1;\n\t"
      "ex2.approx.ftz.f32 e2, e2;\n\t"
      "ex2.approx.ftz.f32 e3, e3;\n\t"
      "mov.b64 ta, {e0, e1};\n\t"
      "mov.b64 tb, {e2, e3};\n\t"
      "add.f32x2 ta, ta, one;\n\t"           // 1 + e
      "add.f32x2 tb, tb, one;\n\t"
      "mov.b64 {e0, e1}, ta;\n\t"
      "mov.b64 {e2, e3}, tb;\n\t"
      "rcp.approx.ftz.f32 e0, e0;\n\t"
      "rcp.approx.ftz.f32 e1, e1;\n\t"
      "rcp.approx.ftz.f32 e2, e2;\n\t"
      "rcp.approx.ftz.f32 e3, e3;\n\t"
      "mov.b64 ta, {e0, e1};\n\t"
      "mov.b64 tb, {e2, e3};\n\t"
      "mul.f32x2 ta, ta, xa;\n\t"            // x / (1 + e)
      "mul.f32x2 tb, tb, xb;\n\t"
      "mov.b64 {%0, %1}, ta;\n\t"
      "mov.b64 {%2, %3}, tb;\n\t"
      "}"
      : "=f"(y.x), "=f"(y.y), "=f"(y.z), "=f"(y.w)
      : "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "f"(k1), "f"(c2), "f"(1.0f));
  return y;
}
__device__ __forceinline__ float silu(float x) { return x / (1.0f + expf(-x)); }

enum Act { ACT_NONE = 0, ACT_GELU = 1, ACT_SILU = 2 };
__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == ACT_GELU) return gelu_tanh(x);
  if (act == ACT_SILU) return silu(x);
  return x;
}

// bf16 split: hi = rn(x), lo = rn(x - hi).  hi + lo carries ~16 mantissa bits of x.
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}
// 16-bit operand planes hold either bf16 (hi [+ lo] split) or, in the single-pass fp16 mode, IEEE half values; the
// storage type is __nv_bfloat16 in both cases (the bits are what the tensor core is told they are).
__device__ __forceinline__ void split16(float x, bool fp16, uint16_t& hi, uint16_t& lo) {
  if (fp16) {
    hi = __half_as_ushort(__float2half_rn(fminf(fmaxf(x, -65504.f), 65504.f)));   // saturate instead of overflowing to inf
    lo = 0;
  } else {
    __nv_bfloat16 h = __float2bfloat16_rn(x);
    hi = __bfloat16_as_ushort(h);
    lo = __bfloat16_as_ushort(__float2bfloat16_rn(x - __bfloat162float(h)));
  }
}

// two fp32 -> one packed 16-bit pair (lo in bits 0-15): IEEE half with saturation to +-65504 (one F2FP.SATFINITE instead of
// two clamps + convert), or bf16 round-to-nearest
__device__ __forceinline__ uint32_t pack2_sat16(float lo, float hi, bool fp16) {
  uint32_t r;
  if (fp16) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// bf16 residual plane of a pair: lo = rn(x - rn_bf16(x))
__device__ __forceinline__ uint32_t pack2_resid_bf16(float a, float b, uint32_t hi_pair) {
  const float ra = a - __uint_as_float(hi_pair << 16), rb = b - __uint_as_float(hi_pair & 0xffff0000u);
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(rb), "f"(ra));
  return r;
}

// packed fp32 pipe (Blackwell FFMA2 / FADD2: one issue slot for two lanes of arithmetic): (d0, d1) = (a0, a1) * (b0, b1) + (c0, c1)
__device__ __forceinline__ void ffma2(float a0, float a1, float b0, float b1, float c0, float c1, float& d0, float& d1) {
  asm("{\n\t.reg .b64 a, b, c;\n\t"
      "mov.b64 a, {%2, %3};\n\t"
      "mov.b64 b, {%4, %5};\n\t"
      "mov.b64 c, {%6, %7};\n\t"
      "fma.rn.f32x2 a, a, b, c;\n\t"
      "mov.b64 {%0, %1}, a;\n\t}"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}
__device__ __forceinline__ void fadd2(float a0, float a1, float b0, float b1, float& d0, float& d1) {
  asm("{\n\t.reg .b64 a, b;\n\t"
      "mov.b64 a, {%2, %3};\n\t"
      "mov.b64 b, {%4, %5};\n\t"
      "add.rn.f32x2 a, a, b;\n\t"
      "mov.b64 {%0, %1}, a;\n\t}"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}

}  // namespace stk

// Epilogue math shared by the GEMM kernels.
//
// GELU (nn.GELU(), exact-erf form, reference mem/modeling_finetune.py:62) is evaluated as
//     gelu(x) = x * Phi(x),   Phi(x) ~= 1 / (1 + exp(-x * (c1 + c3 u + c5 u^2))),   u = min(x^2, 36)
// with (c1, c3, c5) a minimax fit of the logit of the normal CDF:  |gelu error| <= 2.6e-5 and
// |gelu' error| <= 1.1e-4 over the whole real line (tools/fit_gelu.py) -- an eighth of the bf16 rounding
// of the stored activation -- for 2 MUFU + 8 FMA-pipe instructions per element.  The libdevice erff
// version (~40 instructions) made the fc1 epilogue the bottleneck of the whole GEMM (XU pipe at 129 %,
// profiles/r01_ncu_gemm_fc1_v1.csv).  Past |x| = 6 the exponent grows linearly, so Phi saturates to 0 / 1.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace memb {
namespace epi {

// -c * log2(e): the logistic is evaluated with ex2
constexpr float kG1 = -2.3011213283050083f;
constexpr float kG3 = -0.10677573333398353f;
constexpr float kG5 = 0.0010142644560782372f;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Phi(x) and, optionally, the polynomial state needed for the derivative
__device__ __forceinline__ float gelu_cdf(float x, float& u) {
  u = fminf(x * x, 36.0f);
  const float q = x * fmaf(u, fmaf(u, kG5, kG3), kG1);  // -p(x) * log2(e)
  return rcp_approx(1.0f + ex2_approx(q));
}
__device__ __forceinline__ float gelu_fwd(float x) {
  float u;
  return x * gelu_cdf(x, u);
}
// d/dx [x * sigma(p(x))] = s + x s (1 - s) p'(x),  p'(x) = c1 + 3 c3 u + 5 c5 u^2  (0 growth correction past the clamp
// is irrelevant: s (1 - s) < 1e-8 there)
__device__ __forceinline__ float gelu_grad(float x) {
  float u;
  const float s = gelu_cdf(x, u);
  constexpr float kLn2 = 0.69314718055994531f;
  const float dp = fmaf(u, fmaf(u, -5.0f * kG5 * kLn2, -3.0f * kG3 * kLn2), -kG1 * kLn2);
  return fmaf(fmaf(-s, s, s) * x, dp, s);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

}  // namespace epi
}  // namespace memb

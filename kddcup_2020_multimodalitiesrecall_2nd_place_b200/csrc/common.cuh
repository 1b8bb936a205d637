// Shared host/device helpers: error plumbing for the C ABI, 16-bit pack/unpack, warp reductions.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "../../include/mmrecall.h"

namespace mmr {

// ---- error state (thread-local text behind mmr_last_error) ----
char* last_error_buf();
mmr_status fail(mmr_status code, const char* fmt, ...);

#define MMR_CUDA_OK(expr)                                                                              \
  do {                                                                                                 \
    cudaError_t _e = (expr);                                                                           \
    if (_e != cudaSuccess)                                                                             \
      return ::mmr::fail(MMR_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                         __LINE__);                                                                    \
  } while (0)

#define MMR_TRY(expr)                 \
  do {                                \
    mmr_status _s = (expr);           \
    if (_s != MMR_OK) return _s;      \
  } while (0)

#define MMR_REQUIRE(cond, ...)                                    \
  do {                                                            \
    if (!(cond)) return ::mmr::fail(MMR_ERR_INVALID, __VA_ARGS__); \
  } while (0)

// Tuning knobs (mmr_set_tuning / environment at first use); see include/mmrecall.h for the indices.
int tuning(int knob);

// MMR_OK iff the *current* device is sm_100; cached per device.
mmr_status require_sm100();

// ---- kernel launch with programmatic dependent launch (see ptx.cuh: pdl_wait / pdl_launch_dependents) ----
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = tuning(MMR_TUNE_PDL) != 0 ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---- 16-bit element traits ----
struct BF16 {
  using T = __nv_bfloat16;
  using T2 = __nv_bfloat162;
  static constexpr int kFmt = MMR_DT_BF16;
  __device__ __forceinline__ static uint32_t pack(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
  __device__ __forceinline__ static float2 unpack(uint32_t u) {
    return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
  }
};
struct FP16 {
  using T = __half;
  using T2 = __half2;
  static constexpr int kFmt = MMR_DT_FP16;
  __device__ __forceinline__ static uint32_t pack(float a, float b) {
    __half2 v = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
  __device__ __forceinline__ static float2 unpack(uint32_t u) {
    return __half22float2(*reinterpret_cast<__half2*>(&u));
  }
};

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

// Activations of the hot path (fp32 in registers).
__device__ __forceinline__ float gelu_tanh_f(float x) {
  // pixelbert.py:326-328: 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3)))
  const float u = x * fmaf(x * x, 0.7978845608028654f * 0.044715f, 0.7978845608028654f);   // 3 FP ops instead of 4
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}
__device__ __forceinline__ float gelu_erf_f(float x) {
  // modeling.py:119: x * 0.5 * (1 + erf(x / sqrt(2))).  erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, i.e.
  // at fp32 rounding level) instead of erff(): ~12 instructions with one ex2 and one rcp on the SFU, which keeps the
  // FFN epilogue inside the time of its tile's MMAs (erff() made the LXMERT FFN-in GEMM epilogue-bound).
  const float z = fabsf(x) * 0.7071067811865475f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));   // MUFU.RCP (1 ulp), not the IEEE sequence
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float e = 1.0f - poly * t * __expf(-z * z);   // erf(|x| / sqrt 2)
  const float hx = 0.5f * x;
  return fmaf(copysignf(e, x), hx, hx);
}
__device__ __forceinline__ float tanh_precise_f(float x) { return tanhf(x); }

}  // namespace mmr

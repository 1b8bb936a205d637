// Row-wise HBM-bound kernels: LayerNorm and the fp32 -> 16-bit cast of the region features.
// One warp per row, 16-byte vector loads/stores, statistics in fp32 registers (two-pass over registers).
#include "common.cuh"
#include "ptx.cuh"

namespace mmr {

constexpr int kLnMaxVec = 12;  // rows up to 12 * 128 = 1536 floats (logit_fc LayerNorm, kdd_model.py:170)

// tf.contrib.layers.layer_norm (pixelbert.py:414-417) / BertLayerNorm(eps=1e-12) (modeling.py:266):
// y = (x - mean) * rsqrt(var_biased + eps) * gamma + beta
template <class E16>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, int M, int H, typename E16::T* __restrict__ out16,
                 int64_t ldo16, float* __restrict__ out32, int64_t ldo32, float scale, int accumulate) {
  pdl_wait();
  pdl_launch_dependents();
  const int warps_per_block = blockDim.x >> 5;
  const int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const int nv = H >> 7;  // float4 per lane
  const float* xr = x + int64_t(row) * ldx;
  float4 v[kLnMaxVec];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    if (i < nv) {
      v[i] = *reinterpret_cast<const float4*>(xr + (i * 32 + lane) * 4);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  const float mean = warp_sum(s) / float(H);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    if (i < nv) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / float(H) + eps);
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    if (i < nv) {
      const int c0 = (i * 32 + lane) * 4;
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c0));
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c0));
      float4 y;
      y.x = ((v[i].x - mean) * rstd * g.x + b.x) * scale;
      y.y = ((v[i].y - mean) * rstd * g.y + b.y) * scale;
      y.z = ((v[i].z - mean) * rstd * g.z + b.z) * scale;
      y.w = ((v[i].w - mean) * rstd * g.w + b.w) * scale;
      if (out32 != nullptr) {
        float4* op = reinterpret_cast<float4*>(out32 + int64_t(row) * ldo32 + c0);
        if (accumulate) {
          const float4 o = *op;
          y.x += o.x; y.y += o.y; y.z += o.z; y.w += o.w;
        }
        *op = y;
      }
      if (out16 != nullptr) {
        uint2 pk;
        pk.x = E16::pack(y.x, y.y);
        pk.y = E16::pack(y.z, y.w);
        *reinterpret_cast<uint2*>(out16 + int64_t(row) * ldo16 + c0) = pk;
      }
    }
  }
}

template <class E16>
__global__ void __launch_bounds__(256)
cast16_kernel(const float4* __restrict__ x, uint4* __restrict__ out, int64_t n8) {
  pdl_wait();
  pdl_launch_dependents();
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (; i < n8; i += stride) {
    float4 a = __ldcs(x + 2 * i);      // streaming: features are read exactly once
    float4 b = __ldcs(x + 2 * i + 1);
    if constexpr (E16::kFmt == MMR_DT_FP16) {   // saturate instead of overflowing to inf (NaN stays NaN)
      constexpr float kMax = 65504.0f;
      a.x = fminf(fmaxf(a.x, -kMax), kMax); a.y = fminf(fmaxf(a.y, -kMax), kMax);
      a.z = fminf(fmaxf(a.z, -kMax), kMax); a.w = fminf(fmaxf(a.w, -kMax), kMax);
      b.x = fminf(fmaxf(b.x, -kMax), kMax); b.y = fminf(fmaxf(b.y, -kMax), kMax);
      b.z = fminf(fmaxf(b.z, -kMax), kMax); b.w = fminf(fmaxf(b.w, -kMax), kMax);
    }
    uint4 o;
    o.x = E16::pack(a.x, a.y);
    o.y = E16::pack(a.z, a.w);
    o.z = E16::pack(b.x, b.y);
    o.w = E16::pack(b.z, b.w);
    out[i] = o;
  }
}

mmr_status layernorm(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps, int M, int H,
                     void* out16, int64_t ldo16, float* out32, int64_t ldo32, float scale, int accumulate,
                     int dtype, cudaStream_t stream) {
  MMR_TRY(require_sm100());
  MMR_REQUIRE(x && gamma && beta, "mmr_layernorm: null input");
  MMR_REQUIRE(out16 || out32, "mmr_layernorm: no output given");
  MMR_REQUIRE(M > 0, "mmr_layernorm: M=%d", M);
  MMR_REQUIRE(H % 128 == 0 && H <= kLnMaxVec * 128, "mmr_layernorm: H=%d must be a multiple of 128, <= %d", H,
              kLnMaxVec * 128);
  MMR_REQUIRE(ldx % 4 == 0 && (!out32 || ldo32 % 4 == 0) && (!out16 || ldo16 % 4 == 0),
              "mmr_layernorm: row strides must keep 16-byte (fp32) / 8-byte (16-bit) alignment");
  const int wpb = 8;
  const int grid = (M + wpb - 1) / wpb;
  if (dtype == MMR_DT_BF16) {
    (void)launch_pdl(layernorm_kernel<BF16>, dim3(grid), dim3(wpb * 32), 0, stream, x, ldx, gamma, beta, eps, M, H,
                                                          static_cast<BF16::T*>(out16), ldo16, out32, ldo32,
                                                          scale, accumulate);
  } else if (dtype == MMR_DT_FP16) {
    (void)launch_pdl(layernorm_kernel<FP16>, dim3(grid), dim3(wpb * 32), 0, stream, x, ldx, gamma, beta, eps, M, H,
                                                          static_cast<FP16::T*>(out16), ldo16, out32, ldo32,
                                                          scale, accumulate);
  } else {
    return fail(MMR_ERR_INVALID, "mmr_layernorm: bad dtype %d", dtype);
  }
  MMR_CUDA_OK(cudaGetLastError());
  return MMR_OK;
}

mmr_status cast16(const float* x, void* out16, int64_t n, int dtype, cudaStream_t stream) {
  MMR_TRY(require_sm100());
  MMR_REQUIRE(x && out16, "mmr_cast16: null pointer");
  MMR_REQUIRE(n > 0 && n % 8 == 0, "mmr_cast16: n=%lld must be a positive multiple of 8", (long long)n);
  MMR_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out16) & 15) == 0,
              "mmr_cast16: pointers must be 16-byte aligned");
  const int64_t n8 = n / 8;
  int grid = int((n8 + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  if (dtype == MMR_DT_BF16) {
    (void)launch_pdl(cast16_kernel<BF16>, dim3(grid), dim3(256), 0, stream, reinterpret_cast<const float4*>(x),
                                                  reinterpret_cast<uint4*>(out16), n8);
  } else if (dtype == MMR_DT_FP16) {
    (void)launch_pdl(cast16_kernel<FP16>, dim3(grid), dim3(256), 0, stream, reinterpret_cast<const float4*>(x),
                                                  reinterpret_cast<uint4*>(out16), n8);
  } else {
    return fail(MMR_ERR_INVALID, "mmr_cast16: bad dtype %d", dtype);
  }
  MMR_CUDA_OK(cudaGetLastError());
  return MMR_OK;
}

}  // namespace mmr

extern "C" mmr_status mmr_layernorm(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps,
                                    int M, int H, void* out16, int64_t ldo16, float* out32, int64_t ldo32,
                                    float scale, int accumulate, int dtype, void* stream) {
  return mmr::layernorm(x, ldx, gamma, beta, eps, M, H, out16, ldo16, out32, ldo32, scale, accumulate, dtype,
                        static_cast<cudaStream_t>(stream));
}
extern "C" mmr_status mmr_cast16(const float* x, void* out16, int64_t n, int dtype, void* stream) {
  return mmr::cast16(x, out16, n, dtype, static_cast<cudaStream_t>(stream));
}

// Attention for the [CLS] query row only: the last encoder block of every scorer.
//
// The reference computes the whole last block and then reads one row of it: pooled = tanh(W . sequence_output[:, 0])
// (imagebert_zk/pixelbert.py:258-266, imagebert_lds/src/pixelmodel.py:251-259, lxmert/src/lxrt/modeling.py:596-608
// through :925).  Every row of a post-LN block depends on the OTHER rows of its pair only through the keys and values
// of its attention (pixelbert.py:790-850), so for the last block the driver (model.cu) still projects K and V for
// all rows, but runs attention, the output projection + LayerNorm and the FFN for the B [CLS] rows alone.  This file
// is that attention: one CTA per pair, one warp per head, Sq = 1.
//
// Arithmetic mirrors attention_tc2.cu step by step so that the pruned and the full forward agree to fp32 rounding of
// the accumulation order: t = (q . k) / 8 * log2(e) + mask * log2(e); e = ex2(t - max); sum over the UNROUNDED e;
// P rounded to the 16-bit operand type; O = sum_j P_j V_j in fp32; O / sum rounded to 16 bit.
#include "common.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

namespace mmr {

constexpr int kClsHeadDim = 64;
constexpr int kClsMaxKeys = 128;
constexpr float kClsLog2e = 1.4426950408889634f;

__device__ __forceinline__ float cls_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// q: first query row of pair b at q + b * q_pair_stride (elements); k / v: key row j of pair b at
// k + (b * Sk + j) * ldkv.  out16: row b at out16 + b * ldo.  blockDim = 32 * heads.
template <class E16>
__global__ void __launch_bounds__(512)
cls_attention_kernel(const typename E16::T* __restrict__ q, int64_t q_pair_stride, const typename E16::T* __restrict__ k,
                     const typename E16::T* __restrict__ v, int64_t ldkv, const int32_t* __restrict__ key_mask,
                     typename E16::T* __restrict__ out16, int64_t ldo, int Sk) {
  __shared__ float s_p[16][kClsMaxKeys];
  pdl_wait();
  pdl_launch_dependents();
  const int b = blockIdx.x, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // this head's query: every lane holds all 64 values (32 packed pairs would cost 32 registers; read as 8 uint4)
  const uint4* q4 = reinterpret_cast<const uint4*>(q + int64_t(b) * q_pair_stride + h * kClsHeadDim);
  uint4 qv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) qv[i] = __ldg(q4 + i);
  // lane j scores keys j, j + 32, j + 64, j + 96
  float t[4];
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int j = c * 32 + lane;
    t[c] = -INFINITY;
    if (j < Sk) {
      const uint4* k4 = reinterpret_cast<const uint4*>(k + (int64_t(b) * Sk + j) * ldkv + h * kClsHeadDim);
      float dot = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint4 kv = __ldg(k4 + i);
        const uint32_t qa[4] = {qv[i].x, qv[i].y, qv[i].z, qv[i].w}, ka[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 a = E16::unpack(qa[e]), bb = E16::unpack(ka[e]);
          dot = fmaf(a.x, bb.x, dot);
          dot = fmaf(a.y, bb.y, dot);
        }
      }
      const float m = (key_mask == nullptr || __ldg(key_mask + int64_t(b) * Sk + j) != 0) ? 0.0f : -10000.0f * kClsLog2e;
      t[c] = fmaf(dot, 0.125f * kClsLog2e, m);
      mx = fmaxf(mx, t[c]);
    }
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int j = c * 32 + lane;
    if (j < Sk) {
      const float e = cls_ex2(t[c] - mx);
      sum += e;
      s_p[h][j] = E16::unpack(E16::pack(e, 0.f)).x;   // P as the tensor-core path rounds it
    }
  }
  sum = warp_sum(sum);
  __syncwarp();
  // O[d] for this lane's two dims: one 128-byte V row segment per key, coalesced over the warp
  float o0 = 0.f, o1 = 0.f;
  const uint32_t* v32 = reinterpret_cast<const uint32_t*>(v + int64_t(b) * Sk * ldkv + h * kClsHeadDim) + lane;
  const int64_t ld32 = ldkv / 2;
#pragma unroll 8
  for (int j = 0; j < Sk; ++j) {
    const float2 vv = E16::unpack(__ldg(v32 + int64_t(j) * ld32));
    const float p = s_p[h][j];
    o0 = fmaf(p, vv.x, o0);
    o1 = fmaf(p, vv.y, o1);
  }
  const float inv = 1.0f / sum;
  reinterpret_cast<uint32_t*>(out16 + int64_t(b) * ldo + h * kClsHeadDim)[lane] = E16::pack(o0 * inv, o1 * inv);
}

mmr_status cls_attention(const void* q, int64_t q_pair_stride, const void* k, const void* v, int64_t ldkv,
                         const int32_t* key_mask, void* out16, int64_t ldo, int B, int Sk, int heads, int dtype,
                         cudaStream_t stream) {
  MMR_TRY(require_sm100());
  MMR_REQUIRE(q && k && v && out16, "cls_attention: null pointer");
  MMR_REQUIRE(B > 0 && Sk > 0 && Sk <= kClsMaxKeys && heads > 0 && heads <= 16, "cls_attention: B=%d Sk=%d heads=%d", B, Sk,
              heads);
  MMR_REQUIRE(q_pair_stride % 8 == 0 && ldkv % 8 == 0 && ldo % 2 == 0 &&
                  ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v)) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(out16) & 3) == 0,
              "cls_attention: operands must be 16-byte aligned with row strides that keep it");
  if (dtype == MMR_DT_BF16)
    (void)launch_pdl(cls_attention_kernel<BF16>, dim3(B), dim3(32 * heads), 0, stream, static_cast<const BF16::T*>(q),
                     q_pair_stride, static_cast<const BF16::T*>(k), static_cast<const BF16::T*>(v), ldkv, key_mask,
                     static_cast<BF16::T*>(out16), ldo, Sk);
  else
    (void)launch_pdl(cls_attention_kernel<FP16>, dim3(B), dim3(32 * heads), 0, stream, static_cast<const FP16::T*>(q),
                     q_pair_stride, static_cast<const FP16::T*>(k), static_cast<const FP16::T*>(v), ldkv, key_mask,
                     static_cast<FP16::T*>(out16), ldo, Sk);
  MMR_CUDA_OK(cudaGetLastError());
  return MMR_OK;
}

}  // namespace mmr

extern "C" mmr_status mmr_cls_attention(const void* q, int64_t q_pair_stride, const void* k, const void* v, int64_t ldkv,
                                        const int32_t* key_mask, void* out16, int64_t ldo, int B, int Sk, int heads,
                                        int dtype, void* stream) {
  return mmr::cls_attention(q, q_pair_stride, k, v, ldkv, key_mask, out16, ldo, B, Sk, heads, dtype,
                            static_cast<cudaStream_t>(stream));
}

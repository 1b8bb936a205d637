// Kernels of the STRICT precision mode (mmr_config.precision = MMR_PRECISION_STRICT).
//
// The default path rounds every MMA operand to 16 bit once; on weights of trained magnitude that costs up to 3e-3 on
// the score against the reference's fp32 path (TF-1 / torch CPU kernels).  Strict mode keeps
// the same tcgen05 GEMM kernels and feeds them TWO-TERM operands: x = hi + lo with hi = round16(x), lo = round16(x - hi),
// and the three significant partial products as ONE GEMM over a concatenated K axis,
//     [ A_hi | A_lo | A_hi ] . [ W_hi | W_hi | W_lo ]^T = A_hi W_hi + A_lo W_hi + A_hi W_lo      (fp32 accumulate in TMEM)
// i.e. ~21 operand bits instead of 11 for 3x the MMA work.  Everything between the GEMMs stays fp32: activations are
// split on the fly by split3_kernel (which also applies the PRECISE GELU -- the fast path's tanh.approx is itself an
// 11-bit operation), attention runs in fp32 on the CUDA cores (1.2 % of the FLOPs), LayerNorm is the two-pass kernel of
// rowops.cu.  Same reference lines as the fast path: pixelbert.py:658-995, modeling.py:300-434.
#include "common.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

namespace mmr {

__device__ __forceinline__ float act_precise(float x, int act) {
  switch (act) {
    case MMR_ACT_RELU: return fmaxf(x, 0.f);
    case MMR_ACT_GELU_TANH: {
      // pixelbert.py:326-328, in the reference's operation order
      const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
      return x * 0.5f * (1.0f + tanhf(u));
    }
    case MMR_ACT_GELU_ERF: return x * 0.5f * (1.0f + erff(x * 0.7071067811865475f));   // modeling.py:119
    case MMR_ACT_TANH: return tanhf(x);
    default: return x;
  }
}

// out[r, :] = [hi | lo | hi] (weights = 0: an activation row) or [hi | hi | lo] (weights = 1) of act(x[r, :]), K wide each.
template <class E16>
__global__ void __launch_bounds__(256)
split3_kernel(const float* __restrict__ x, int64_t ldx, int rows, int K, typename E16::T* __restrict__ out, int64_t ldo,
              int act, int weights) {
  pdl_wait();
  pdl_launch_dependents();
  const int k8 = K >> 3;
  const int64_t n8 = int64_t(rows) * k8;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const int r = int(i / k8), c = int(i - int64_t(r) * k8) * 8;
    const float4 a = *reinterpret_cast<const float4*>(x + int64_t(r) * ldx + c);
    const float4 b = *reinterpret_cast<const float4*>(x + int64_t(r) * ldx + c + 4);
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float v0 = act_precise(v[2 * e], act), v1 = act_precise(v[2 * e + 1], act);
      hi[e] = E16::pack(v0, v1);
      const float2 h = E16::unpack(hi[e]);
      lo[e] = E16::pack(v0 - h.x, v1 - h.y);
    }
    const uint4 H = make_uint4(hi[0], hi[1], hi[2], hi[3]), L = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    typename E16::T* o = out + int64_t(r) * ldo + c;
    *reinterpret_cast<uint4*>(o) = H;
    *reinterpret_cast<uint4*>(o + K) = weights ? H : L;
    *reinterpret_cast<uint4*>(o + 2 * int64_t(K)) = weights ? L : H;
  }
}

mmr_status split3(const float* x, int64_t ldx, int rows, int K, void* out16, int64_t ldo, int act, int weights, int dtype,
                  cudaStream_t stream) {
  MMR_TRY(require_sm100());
  MMR_REQUIRE(x && out16 && rows > 0 && K > 0 && K % 8 == 0 && ldx % 4 == 0 && ldo % 8 == 0 &&
                  ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out16)) & 15) == 0,
              "split3: bad argument (rows=%d K=%d)", rows, K);
  const int64_t n8 = int64_t(rows) * (K / 8);
  int grid = int((n8 + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  if (dtype == MMR_DT_BF16)
    (void)launch_pdl(split3_kernel<BF16>, dim3(grid), dim3(256), 0, stream, x, ldx, rows, K, static_cast<BF16::T*>(out16), ldo,
                     act, weights);
  else
    (void)launch_pdl(split3_kernel<FP16>, dim3(grid), dim3(256), 0, stream, x, ldx, rows, K, static_cast<FP16::T*>(out16), ldo,
                     act, weights);
  MMR_CUDA_OK(cudaGetLastError());
  return MMR_OK;
}

__global__ void __launch_bounds__(256) act32_kernel(float* __restrict__ x, int64_t n, int act) {
  pdl_wait();
  pdl_launch_dependents();
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) x[i] = act_precise(x[i], act);
}
mmr_status act32(float* x, int64_t n, int act, cudaStream_t stream) {
  MMR_REQUIRE(x && n > 0, "act32: bad argument");
  int grid = int((n + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  (void)launch_pdl(act32_kernel, dim3(grid), dim3(256), 0, stream, x, n, act);
  MMR_CUDA_OK(cudaGetLastError());
  return MMR_OK;
}

// fp32 attention, one CTA per (pair, head): scores = q k^T / 8 + (1 - m) * -10000, softmax, context (pixelbert.py:790-850,
// modeling.py:325-352).  Pair b: query rows at q + b * q_pair + i * ldq (i < Sq), keys / values at k|v + b * kv_pair +
// j * ldkv (j < Sk), output rows at out + b * o_pair + i * ldo; head h = columns [64 h, 64 h + 64).
constexpr int kF32MaxKeys = 128;
constexpr int kF32Threads = 256;
__global__ void __launch_bounds__(kF32Threads)
attention_f32_kernel(const float* __restrict__ q, int64_t q_pair, int64_t ldq, const float* __restrict__ k,
                     const float* __restrict__ v, int64_t kv_pair, int64_t ldkv, const int32_t* __restrict__ key_mask,
                     float* __restrict__ out, int64_t o_pair, int64_t ldo, int Sq, int Sk, int heads) {
  extern __shared__ float sm[];
  float* sK = sm;                                   // [Sk][65]
  float* sV = sK + kF32MaxKeys * 65;                // [Sk][64]
  float* sMask = sV + kF32MaxKeys * 64;             // [Sk]
  float* sQ = sMask + kF32MaxKeys;                  // [8 warps][64]
  float* sP = sQ + 8 * 64;                          // [8 warps][128]
  pdl_wait();
  pdl_launch_dependents();
  const int b = blockIdx.x / heads, h = blockIdx.x - b * heads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < Sk * 16; i += kF32Threads) {
    const int j = i >> 4, c = (i & 15) * 4;
    const float4 kk = *reinterpret_cast<const float4*>(k + int64_t(b) * kv_pair + int64_t(j) * ldkv + h * 64 + c);
    const float4 vv = *reinterpret_cast<const float4*>(v + int64_t(b) * kv_pair + int64_t(j) * ldkv + h * 64 + c);
    sK[j * 65 + c] = kk.x; sK[j * 65 + c + 1] = kk.y; sK[j * 65 + c + 2] = kk.z; sK[j * 65 + c + 3] = kk.w;
    *reinterpret_cast<float4*>(sV + j * 64 + c) = vv;
  }
  for (int j = threadIdx.x; j < Sk; j += kF32Threads)
    sMask[j] = (key_mask == nullptr || key_mask[int64_t(b) * Sk + j] != 0) ? 0.0f : -10000.0f;
  __syncthreads();
  float* myQ = sQ + warp * 64;
  float* myP = sP + warp * kF32MaxKeys;
  for (int i = warp; i < Sq; i += kF32Threads / 32) {
    const float* qr = q + int64_t(b) * q_pair + int64_t(i) * ldq + h * 64;
    myQ[lane] = qr[lane];
    myQ[lane + 32] = qr[lane + 32];
    __syncwarp();
    float s[4], mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int j = c * 32 + lane;
      s[c] = -INFINITY;
      if (j < Sk) {
        float dot = 0.f;
#pragma unroll 16
        for (int d = 0; d < 64; ++d) dot = fmaf(myQ[d], sK[j * 65 + d], dot);
        s[c] = dot * 0.125f + sMask[j];
        mx = fmaxf(mx, s[c]);
      }
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int j = c * 32 + lane;
      if (j < Sk) {
        const float e = expf(s[c] - mx);
        myP[j] = e;
        sum += e;
      }
    }
    sum = warp_sum(sum);
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < Sk; ++j) {
      const float p = myP[j];
      o0 = fmaf(p, sV[j * 64 + lane], o0);
      o1 = fmaf(p, sV[j * 64 + lane + 32], o1);
    }
    const float inv = 1.0f / sum;
    float* orow = out + int64_t(b) * o_pair + int64_t(i) * ldo + h * 64;
    orow[lane] = o0 * inv;
    orow[lane + 32] = o1 * inv;
    __syncwarp();   // myQ / myP are rewritten by this warp's next row
  }
}

mmr_status attention_f32(const float* q, int64_t q_pair, int64_t ldq, const float* k, const float* v, int64_t kv_pair,
                         int64_t ldkv, const int32_t* key_mask, float* out, int64_t o_pair, int64_t ldo, int B, int Sq,
                         int Sk, int heads, cudaStream_t stream) {
  MMR_TRY(require_sm100());
  MMR_REQUIRE(q && k && v && out && B > 0 && Sq > 0 && Sk > 0 && Sk <= kF32MaxKeys && heads > 0,
              "attention_f32: bad argument (B=%d Sq=%d Sk=%d)", B, Sq, Sk);
  MMR_REQUIRE(ldkv % 4 == 0 && kv_pair % 4 == 0 &&
                  ((reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v)) & 15) == 0,
              "attention_f32: keys / values must be 16-byte aligned");
  const size_t smem = size_t(kF32MaxKeys * 65 + kF32MaxKeys * 64 + kF32MaxKeys + 8 * 64 + 8 * kF32MaxKeys) * 4;
  static bool configured = false;
  if (!configured) {
    MMR_CUDA_OK(cudaFuncSetAttribute(attention_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    configured = true;
  }
  (void)launch_pdl(attention_f32_kernel, dim3(B * heads), dim3(kF32Threads), smem, stream, q, q_pair, ldq, k, v, kv_pair, ldkv,
                   key_mask, out, o_pair, ldo, Sq, Sk, heads);
  MMR_CUDA_OK(cudaGetLastError());
  return MMR_OK;
}

}  // namespace mmr

extern "C" mmr_status mmr_split3(const float* x, int64_t ldx, int rows, int K, void* out16, int64_t ldo, int act,
                                 int weights, int dtype, void* stream) {
  return mmr::split3(x, ldx, rows, K, out16, ldo, act, weights, dtype, static_cast<cudaStream_t>(stream));
}
extern "C" mmr_status mmr_attention_f32(const float* q, int64_t q_pair, int64_t ldq, const float* k, const float* v,
                                        int64_t kv_pair, int64_t ldkv, const int32_t* key_mask, float* out, int64_t o_pair,
                                        int64_t ldo, int B, int Sq, int Sk, int heads, void* stream) {
  return mmr::attention_f32(q, q_pair, ldq, k, v, kv_pair, ldkv, key_mask, out, o_pair, ldo, B, Sq, Sk, heads,
                            static_cast<cudaStream_t>(stream));
}

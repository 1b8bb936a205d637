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

// -------------------------------------------------------------------------------------------------------------------
// The end of the last block in ONE kernel: LayerNorm of the FFN output (+ residual, already summed by the GEMM before),
// pooler tanh(W . x[CLS] + b) (pixelbert.py:258-266, pixelmodel.py:251-259) and the 2-way match head -- AM-softmax
// (model_triple.py:56-86) or linear + softmax (run_pretraining_predict_score.py:479-501) -- for the B [CLS] rows.
// A cluster of four CTAs owns eight rows: every CTA normalises the eight rows into shared memory (24 KB, redundantly:
// 6 K floats), computes ITS 192 pooler columns for them on the CUDA cores (four adjacent lanes = four output columns,
// each lane taking every fourth 8-element chunk of their weight rows streamed from L2; the inputs are broadcast from
// shared memory, and 8 rows x 4 columns per thread is what keeps that broadcast -- 4 passes per LDS.128 whatever the
// lanes read -- off the critical path: one thread = one column measured 30 us, bound by exactly that; 151 MFLOP in all,
// a tensor-core launch for 256 rows is latency, not throughput), reduces the head's three sums
// over its columns and sends them to the cluster's first
// CTA through distributed shared memory, which finishes the softmax.  Replaces three launches (LayerNorm, pooler GEMM,
// head: 6 + 18 + 5 us under ncu) of the [CLS] tail.
constexpr int kPhRows = 8;          // rows per cluster
constexpr int kPhCtas = 4;          // CTAs per cluster
constexpr int kPhCols = 768 / kPhCtas;   // 192 pooler columns per CTA
constexpr int kPhThreads = 256;      // 8 LayerNorm warps; 192 of the threads = 48 column quads x 4 interleaved K quarters

template <class E16>
__global__ void __cluster_dims__(kPhCtas, 1, 1) __launch_bounds__(kPhThreads)
cls_pool_head_kernel(const float* __restrict__ y32, const float* __restrict__ gamma, const float* __restrict__ beta,
                     const typename E16::T* __restrict__ Wp, const float* __restrict__ bp, int head_kind,
                     const float* __restrict__ hw, const float* __restrict__ hb, const int32_t* __restrict__ labels,
                     int B, float* __restrict__ pooled32, float* __restrict__ probs, float* __restrict__ logits) {
  __shared__ __align__(16) float xs[kPhRows][768];
  __shared__ float part[kPhThreads / 32][kPhRows][3];
  __shared__ float gather[kPhCtas][kPhRows][3];       // meaningful in the cluster's CTA 0
  pdl_wait();
  pdl_launch_dependents();
  cluster_arrive_relaxed();     // "this CTA runs": waited for below, before anything is stored into CTA 0's `gather`
  const uint32_t rank = cluster_ctarank();
  const int row0 = int(blockIdx.x / kPhCtas) * kPhRows;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // ---- LayerNorm of row row0 + warp (biased variance, eps 1e-12), rounded to the operand type like the x16 mirror the
  // tensor-core pooler reads
  if (warp < kPhRows) {
    const int r = row0 + warp;
    float4 v[6];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      v[i] = r < B ? *reinterpret_cast<const float4*>(y32 + int64_t(r) * 768 + (i * 32 + lane) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(s) * (1.0f / 768.0f);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / 768.0f) + 1e-12f);
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int c0 = (i * 32 + lane) * 4;
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c0));
      const float4 be = __ldg(reinterpret_cast<const float4*>(beta + c0));
      const float2 lo = E16::unpack(E16::pack((v[i].x - mean) * rstd * g.x + be.x, (v[i].y - mean) * rstd * g.y + be.y));
      const float2 hi = E16::unpack(E16::pack((v[i].z - mean) * rstd * g.z + be.z, (v[i].w - mean) * rstd * g.w + be.w));
      *reinterpret_cast<float4*>(&xs[warp][c0]) = make_float4(lo.x, lo.y, hi.x, hi.y);
    }
  }
  __syncthreads();
  // ---- pooler: lanes 4g .. 4g + 3 own the output columns 4g .. 4g + 3 of this CTA; lane quarter q takes the
  // 8-element chunks k8 = 4 i + q of the four weight rows (the four lanes read 64 contiguous bytes of a row; their
  // shared-memory reads hit four different bank groups), 8 rows x 4 columns of accumulators per thread, the weight
  // chunks of the next iteration in flight while this one is multiplied
  float sums[kPhRows][3];
#pragma unroll
  for (int r = 0; r < kPhRows; ++r) sums[r][0] = sums[r][1] = sums[r][2] = 0.f;
  if (threadIdx.x < kPhCols) {
    const int q = threadIdx.x & 3;
    const int j0 = int(rank) * kPhCols + (threadIdx.x >> 2) * 4;
    const uint4* wrow = reinterpret_cast<const uint4*>(Wp + int64_t(j0) * 768);   // rows j0 .. j0 + 3, 96 uint4 apart
    float acc[kPhRows][4];
#pragma unroll
    for (int r = 0; r < kPhRows; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
    uint4 wn[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) wn[c] = __ldg(wrow + c * 96 + q);
#pragma unroll 1
    for (int i = 0; i < 24; ++i) {
      uint4 w[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) w[c] = wn[c];
      if (i + 1 < 24) {
#pragma unroll
        for (int c = 0; c < 4; ++c) wn[c] = __ldg(wrow + c * 96 + 4 * (i + 1) + q);
      }
      const int k = (4 * i + q) * 8;
      float2 wf[4][4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        wf[c][0] = E16::unpack(w[c].x); wf[c][1] = E16::unpack(w[c].y);
        wf[c][2] = E16::unpack(w[c].z); wf[c][3] = E16::unpack(w[c].w);
      }
#pragma unroll
      for (int r = 0; r < kPhRows; ++r) {
        const float4 a = *reinterpret_cast<const float4*>(&xs[r][k]);
        const float4 b = *reinterpret_cast<const float4*>(&xs[r][k + 4]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float t = acc[r][c];
          t = fmaf(a.x, wf[c][0].x, t); t = fmaf(a.y, wf[c][0].y, t); t = fmaf(a.z, wf[c][1].x, t); t = fmaf(a.w, wf[c][1].y, t);
          t = fmaf(b.x, wf[c][2].x, t); t = fmaf(b.y, wf[c][2].y, t); t = fmaf(b.z, wf[c][3].x, t); t = fmaf(b.w, wf[c][3].y, t);
          acc[r][c] = t;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < kPhRows; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {            // the four K quarters of a column, in a fixed order
        acc[r][c] += __shfl_xor_sync(0xffffffffu, acc[r][c], 1);
        acc[r][c] += __shfl_xor_sync(0xffffffffu, acc[r][c], 2);
      }
    // lane quarter q finishes column j0 + q
    const int j = j0 + q;
    const float bj = __ldg(bp + j);
    const float h0 = __ldg(hw + j), h1 = __ldg(hw + 768 + j);       // head weights [2, 768]
#pragma unroll
    for (int r = 0; r < kPhRows; ++r) {
      const float pre = q == 0 ? acc[r][0] : (q == 1 ? acc[r][1] : (q == 2 ? acc[r][2] : acc[r][3]));
      const float p = tanhf(pre + bj);
      if (row0 + r < B) pooled32[int64_t(row0 + r) * 768 + j] = p;
      sums[r][0] = p * p;
      sums[r][1] = p * h0;
      sums[r][2] = p * h1;
    }
  }
  // ---- head sums over this CTA's columns -> cluster CTA 0
#pragma unroll
  for (int r = 0; r < kPhRows; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float t = warp_sum(sums[r][c]);
      if (lane == 0) part[warp][r][c] = t;
    }
  __syncthreads();
  cluster_wait();               // every CTA of the cluster has started: CTA 0's shared memory may be written
  if (threadIdx.x < kPhRows * 3) {
    const int r = threadIdx.x / 3, c = threadIdx.x % 3;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kPhThreads / 32; ++w) t += part[w][r][c];
    const uint32_t dst = mapa_u32(smem_u32(&gather[rank][r][c]), 0);
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(dst), "f"(t) : "memory");
  }
  cluster_sync_all();    // release / acquire at cluster scope: CTA 0 sees the four CTAs' partial sums
  if (rank == 0 && threadIdx.x < kPhRows && row0 + int(threadIdx.x) < B) {
    const int r = threadIdx.x, b = row0 + r;
    float ss = 0.f, d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int q = 0; q < kPhCtas; ++q) {     // fixed order: the bits do not depend on which CTA arrives first
      ss += gather[q][r][0];
      d0 += gather[q][r][1];
      d1 += gather[q][r][2];
    }
    float l0, l1;
    if (head_kind == 0) {
      // AM-softmax (model_triple.py:56-86): hw = column-normalised am_kernel, cosines, margin on the fed label, x 30
      const float inv = rsqrtf(fmaxf(ss, 1e-12f));
      float c0 = fminf(fmaxf(d0 * inv, -1.f), 1.f), c1 = fminf(fmaxf(d1 * inv, -1.f), 1.f);
      const int y = labels[b];
      const float g = y ? c1 : c0;
      const float m = g > 0.35f ? 0.35f : 0.f;
      if (y) c1 -= m; else c0 -= m;
      l0 = 30.f * c0;
      l1 = 30.f * c1;
    } else {
      l0 = d0 + hb[0];                       // run_pretraining_predict_score.py:491-492
      l1 = d1 + hb[1];
    }
    const float mx = fmaxf(l0, l1);
    const float e0 = expf(l0 - mx), e1 = expf(l1 - mx);
    const float inv = 1.0f / (e0 + e1);
    probs[2 * b] = e0 * inv;
    probs[2 * b + 1] = e1 * inv;
    if (logits != nullptr) {
      logits[2 * b] = l0;
      logits[2 * b + 1] = l1;
    }
  }
}

// head_kind 0: AM-softmax (hw = wn [2,768], labels); 1: linear (hw = W [2,768], hb [2]).
mmr_status cls_pool_head(const float* y32, const float* gamma, const float* beta, const void* Wp16, const float* bp,
                         int head_kind, const float* hw, const float* hb, const int32_t* labels, int B, float* pooled32,
                         float* probs, float* logits, int dtype, cudaStream_t stream) {
  MMR_TRY(require_sm100());
  MMR_REQUIRE(y32 && gamma && beta && Wp16 && bp && hw && pooled32 && probs && B > 0, "cls_pool_head: null argument");
  MMR_REQUIRE(head_kind == 0 ? labels != nullptr : hb != nullptr, "cls_pool_head: head parameters missing");
  const int clusters = (B + kPhRows - 1) / kPhRows;
  if (dtype == MMR_DT_BF16)
    (void)launch_pdl(cls_pool_head_kernel<BF16>, dim3(clusters * kPhCtas), dim3(kPhThreads), 0, stream, y32, gamma, beta,
                     static_cast<const BF16::T*>(Wp16), bp, head_kind, hw, hb, labels, B, pooled32, probs, logits);
  else
    (void)launch_pdl(cls_pool_head_kernel<FP16>, dim3(clusters * kPhCtas), dim3(kPhThreads), 0, stream, y32, gamma, beta,
                     static_cast<const FP16::T*>(Wp16), bp, head_kind, hw, hb, labels, B, pooled32, probs, logits);
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

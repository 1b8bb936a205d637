// Input-fusion ("embedding") and match-head kernels of the three scorers.  All are HBM/L2-bound row kernels:
// one warp per 768-wide output row, 16-byte vector loads, LayerNorm statistics in fp32 registers.
//
//   zk  (imagebert_zk/model_triple.py:178-195, pixelbert.py:493-621): label-text conv term (as gather-sums over
//       8 pre-multiplied tables, see model.cu), box FC, sum with the ReLU'd region projection; then
//       concat [word ; region] + token-type + position ([0..Lq-1] + [Lq]*R) + LayerNorm.
//   lds (imagebert_lds/src/pixelmodel.py:444-602): LayerNorm on the text rows only; raw region rows; the
//       "reshape4D" label rows.
//   lxmert (lxrt/modeling.py:269-297, 496-533): BertEmbeddings for query and label tokens, Conv2d(8,1,1) over the
//       label tokens, box FC + LayerNorm accumulated into the visual embedding.
//   heads: AM-softmax (model_triple.py:56-86), linear 2-way (run_pretraining_predict_score.py:479-501),
//       LayerNorm(1536) + linear 2-way (tasks/kdd_model.py:167-172), each followed by softmax.
#include "common.cuh"
#include "ptx.cuh"
#include "kernels.cuh"

namespace mmr {

constexpr int kH = 768;
constexpr int kNV = kH / 128;  // float4 per lane

struct Row {
  float4 v[kNV];
};

__device__ __forceinline__ int col_of(int i, int lane) { return (i * 32 + lane) * 4; }

__device__ __forceinline__ void row_zero(Row& r) {
#pragma unroll
  for (int i = 0; i < kNV; ++i) r.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}
__device__ __forceinline__ void row_load(Row& r, const float* __restrict__ p, int lane) {
#pragma unroll
  for (int i = 0; i < kNV; ++i) r.v[i] = __ldg(reinterpret_cast<const float4*>(p + col_of(i, lane)));
}
__device__ __forceinline__ void row_add(Row& r, const float* __restrict__ p, int lane) {
#pragma unroll
  for (int i = 0; i < kNV; ++i) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p + col_of(i, lane)));
    r.v[i].x += a.x; r.v[i].y += a.y; r.v[i].z += a.z; r.v[i].w += a.w;
  }
}
__device__ __forceinline__ void row_axpy(Row& r, float s, const Row& a) {
#pragma unroll
  for (int i = 0; i < kNV; ++i) {
    r.v[i].x = fmaf(s, a.v[i].x, r.v[i].x); r.v[i].y = fmaf(s, a.v[i].y, r.v[i].y);
    r.v[i].z = fmaf(s, a.v[i].z, r.v[i].z); r.v[i].w = fmaf(s, a.v[i].w, r.v[i].w);
  }
}
// y = (x - mean) * rsqrt(var + eps) * gamma + beta, in place (biased variance, eps = 1e-12)
__device__ __forceinline__ void row_layernorm(Row& r, const float* __restrict__ gamma,
                                              const float* __restrict__ beta, int lane) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kNV; ++i) s += (r.v[i].x + r.v[i].y) + (r.v[i].z + r.v[i].w);
  const float mean = warp_sum(s) * (1.0f / kH);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kNV; ++i) {
    const float a = r.v[i].x - mean, b = r.v[i].y - mean, c = r.v[i].z - mean, d = r.v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / kH) + 1e-12f);
#pragma unroll
  for (int i = 0; i < kNV; ++i) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + col_of(i, lane)));
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta + col_of(i, lane)));
    r.v[i].x = (r.v[i].x - mean) * rstd * g.x + b.x;
    r.v[i].y = (r.v[i].y - mean) * rstd * g.y + b.y;
    r.v[i].z = (r.v[i].z - mean) * rstd * g.z + b.z;
    r.v[i].w = (r.v[i].w - mean) * rstd * g.w + b.w;
  }
}
template <class E16>
__device__ __forceinline__ void row_store(const Row& r, typename E16::T* out16, float* out32, int lane) {
#pragma unroll
  for (int i = 0; i < kNV; ++i) {
    const int c = col_of(i, lane);
    if (out32 != nullptr) *reinterpret_cast<float4*>(out32 + c) = r.v[i];
    if (out16 != nullptr) {
      uint2 pk;
      pk.x = E16::pack(r.v[i].x, r.v[i].y);
      pk.y = E16::pack(r.v[i].z, r.v[i].w);
      *reinterpret_cast<uint2*>(out16 + c) = pk;
    }
  }
}

// ---------------------------------------------------------------------------------------------- zk
// t[b,r,:] = label_term + box_fc + feat   (model_triple.py:189-195), written as the 16-bit operand of the
// kdd_featureemb GEMM.  label_term = mean_w ReLU(bc1 + sum_k T_k[id[w-3+k]]) with T_k = E . Wc1[k] (SAME
// padding of the 8-tap kernel: 3 left, 4 right; out-of-range taps contribute nothing).
template <class E16>
__global__ void __launch_bounds__(256)
zk_region_sum_kernel(const float* __restrict__ feat32, const float* __restrict__ boxes5,
                     const int32_t* __restrict__ label_ids, const float* __restrict__ tables, int vocab,
                     const float* __restrict__ bc1, const float* __restrict__ Wb, const float* __restrict__ bb,
                     typename E16::T* __restrict__ out16, float* __restrict__ out32, int rows) {
  pdl_wait();
  pdl_launch_dependents();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  int ids[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) ids[t] = __ldg(label_ids + int64_t(row) * 8 + t);
  Row acc;
  row_zero(acc);
  for (int w = 0; w < 8; ++w) {
    Row c;
    row_load(c, bc1, lane);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int j = w - 3 + k;
      if (j >= 0 && j < 8) row_add(c, tables + (int64_t(k) * vocab + ids[j]) * kH, lane);
    }
#pragma unroll
    for (int i = 0; i < kNV; ++i) {
      acc.v[i].x += fmaxf(c.v[i].x, 0.f); acc.v[i].y += fmaxf(c.v[i].y, 0.f);
      acc.v[i].z += fmaxf(c.v[i].z, 0.f); acc.v[i].w += fmaxf(c.v[i].w, 0.f);
    }
  }
  Row out;
  row_load(out, feat32 + int64_t(row) * kH, lane);
  row_axpy(out, 0.125f, acc);  // mean over the 8 positions
  row_add(out, bb, lane);
#pragma unroll
  for (int d = 0; d < 5; ++d) {
    const float bx = __ldg(boxes5 + int64_t(row) * 5 + d);
    Row wrow;
    row_load(wrow, Wb + d * kH, lane);
    row_axpy(out, bx, wrow);
  }
  row_store<E16>(out, out16 ? out16 + int64_t(row) * kH : nullptr, out32 ? out32 + int64_t(row) * kH : nullptr, lane);
}

// ---- label term, computed once per DISTINCT label phrase of the batch
// The label term of a box depends on its 8 label token ids only, and a batch holds few distinct phrases (the detector
// has 33 classes — load_data_v4.py:34-38 — and every padded box carries the all-[PAD] phrase), while the term costs
// 48 gathered 3 KB table rows per box: 1.3 GB of L1/L2 traffic per 256-pair step, 71 us, when evaluated per box.
//   claim   one thread per box: hash of the 8 ids -> open-addressing table of {epoch, row}; the first box of a phrase
//           claims the slot (atomicCAS) and becomes the phrase's representative, the others compare ids and point at it
//   term    one CTA per representative (the others exit at once): warp w evaluates position w, the eight ReLU'd rows
//           are summed in position order — the same order, hence the same bits, as zk_region_sum_kernel
//   sum     zk_region_sum_rep_kernel: feat + term[rep] / 8 + box FC, per box
// Which box wins a slot is a race; the value it computes is not.  The table is never cleared: entries carry the
// forward's epoch and older ones count as empty.  The epoch lives in DEVICE memory (read by the claim kernel, moved
// on by the sum kernel, the last of the three): a forward replayed from a CUDA graph sees a fresh epoch too.
__global__ void __launch_bounds__(256)
zk_label_claim_kernel(const int32_t* __restrict__ label_ids, int rows, unsigned long long* __restrict__ tab,
                      uint32_t tab_mask, const uint32_t* __restrict__ epoch_ptr, int32_t* __restrict__ rep) {
  pdl_wait();
  pdl_launch_dependents();
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  const uint32_t epoch = *reinterpret_cast<const volatile uint32_t*>(epoch_ptr);
  const int4* ids4 = reinterpret_cast<const int4*>(label_ids);
  const int4 a = __ldg(ids4 + 2 * int64_t(row)), b = __ldg(ids4 + 2 * int64_t(row) + 1);
  uint32_t hsh = 2166136261u;
  const int32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int t = 0; t < 8; ++t) hsh = (hsh ^ uint32_t(w[t])) * 16777619u;
  hsh ^= hsh >> 15;
  uint32_t slot = hsh & tab_mask;
  const unsigned long long mine = (static_cast<unsigned long long>(epoch) << 32) | uint32_t(row);
  for (uint32_t probe = 0; probe <= tab_mask; ++probe) {
    unsigned long long v = *reinterpret_cast<volatile unsigned long long*>(tab + slot);
    if (uint32_t(v >> 32) != epoch) {            // empty for this forward: try to claim it
      const unsigned long long old = atomicCAS(tab + slot, v, mine);
      if (old == v) {
        rep[row] = row;
        return;
      }
      v = old;                                   // somebody of this forward got there first
    }
    const int owner = int(uint32_t(v));
    const int4 oa = __ldg(ids4 + 2 * int64_t(owner)), ob = __ldg(ids4 + 2 * int64_t(owner) + 1);
    if (oa.x == a.x && oa.y == a.y && oa.z == a.z && oa.w == a.w && ob.x == b.x && ob.y == b.y && ob.z == b.z &&
        ob.w == b.w) {
      rep[row] = owner;
      return;
    }
    slot = (slot + 1) & tab_mask;
  }
  rep[row] = row;   // table full (it is sized at <= 25 % load): evaluate the phrase for this box itself
}

__global__ void __launch_bounds__(256)
zk_label_term_kernel(const int32_t* __restrict__ label_ids, const float* __restrict__ tables, int vocab,
                     const float* __restrict__ bc1, const int32_t* __restrict__ rep, float* __restrict__ term32) {
  __shared__ float s_c[8][kH];
  pdl_wait();
  pdl_launch_dependents();
  const int row = blockIdx.x;
  if (__ldg(rep + row) != row) return;           // block-uniform
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int ids[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) ids[t] = __ldg(label_ids + int64_t(row) * 8 + t);
  Row c;
  row_load(c, bc1, lane);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int j = w - 3 + k;
    if (j >= 0 && j < 8) row_add(c, tables + (int64_t(k) * vocab + ids[j]) * kH, lane);
  }
#pragma unroll
  for (int i = 0; i < kNV; ++i)
    *reinterpret_cast<float4*>(&s_c[w][col_of(i, lane)]) =
        make_float4(fmaxf(c.v[i].x, 0.f), fmaxf(c.v[i].y, 0.f), fmaxf(c.v[i].z, 0.f), fmaxf(c.v[i].w, 0.f));
  __syncthreads();
  for (int col = threadIdx.x; col < kH; col += 256) {
    float acc = 0.f;
#pragma unroll
    for (int p = 0; p < 8; ++p) acc += s_c[p][col];      // position order 0..7, as in zk_region_sum_kernel
    term32[int64_t(row) * kH + col] = acc;
  }
}

template <class E16>
__global__ void __launch_bounds__(256)
zk_region_sum_rep_kernel(const float* __restrict__ feat32, const float* __restrict__ boxes5,
                         const int32_t* __restrict__ rep, const float* __restrict__ term32,
                         const float* __restrict__ Wb, const float* __restrict__ bb,
                         typename E16::T* __restrict__ out16, float* __restrict__ out32, int rows,
                         uint32_t* __restrict__ epoch_ptr) {
  pdl_wait();
  pdl_launch_dependents();
  if (blockIdx.x == 0 && threadIdx.x == 0) {   // claim and term kernels of this forward are complete (stream order)
    const uint32_t next = *epoch_ptr + 1u;
    *epoch_ptr = next == 0u ? 1u : next;       // 0 = "never written"
  }
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  Row acc;
  row_load(acc, term32 + int64_t(__ldg(rep + row)) * kH, lane);
  Row out;
  row_load(out, feat32 + int64_t(row) * kH, lane);
  row_axpy(out, 0.125f, acc);  // mean over the 8 positions
  row_add(out, bb, lane);
#pragma unroll
  for (int d = 0; d < 5; ++d) {
    const float bx = __ldg(boxes5 + int64_t(row) * 5 + d);
    Row wrow;
    row_load(wrow, Wb + d * kH, lane);
    row_axpy(out, bx, wrow);
  }
  row_store<E16>(out, out16 ? out16 + int64_t(row) * kH : nullptr, out32 ? out32 + int64_t(row) * kH : nullptr, lane);
}

// X0 = LN(concat(E[q], region) + Ttype[seg] + Pos[[0..Lq-1] + [Lq]*R]); also emits the key mask
// [j < len_query | j - Lq < num_boxes] (model_triple.py:198-201).
template <class E16>
__global__ void __launch_bounds__(256)
zk_embed_kernel(const int32_t* __restrict__ query_ids, const int32_t* __restrict__ segment_ids,
                const float* __restrict__ region32, const int32_t* __restrict__ len_query,
                const int32_t* __restrict__ num_boxes, const float* __restrict__ E, const float* __restrict__ T,
                const float* __restrict__ P, const float* __restrict__ gamma, const float* __restrict__ beta,
                int Lq, int R, int rows, typename E16::T* __restrict__ x16, float* __restrict__ x32,
                int32_t* __restrict__ key_mask) {
  pdl_wait();
  pdl_launch_dependents();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int S = Lq + R;
  const int b = row / S, s = row % S;
  Row x;
  if (s < Lq) {
    row_load(x, E + int64_t(__ldg(query_ids + b * Lq + s)) * kH, lane);
    row_add(x, P + int64_t(s) * kH, lane);
  } else {
    row_load(x, region32 + (int64_t(b) * R + (s - Lq)) * kH, lane);
    row_add(x, P + int64_t(Lq) * kH, lane);
  }
  row_add(x, T + int64_t(__ldg(segment_ids + row)) * kH, lane);
  row_layernorm(x, gamma, beta, lane);
  row_store<E16>(x, x16 + int64_t(row) * kH, x32 + int64_t(row) * kH, lane);
  if (lane == 0) key_mask[row] = (s < Lq) ? (s < __ldg(len_query + b)) : ((s - Lq) < __ldg(num_boxes + b));
}

// ---------------------------------------------------------------------------------------------- lds
template <class E16>
__global__ void __launch_bounds__(256)
lds_embed_kernel(const int32_t* __restrict__ query_ids, const int32_t* __restrict__ segment_ids,
                 const int32_t* __restrict__ label_ids, const float* __restrict__ region32,
                 const float* __restrict__ E, const float* __restrict__ T, const float* __restrict__ P,
                 const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ wl,
                 int Lq, int R, int rows, typename E16::T* __restrict__ x16, float* __restrict__ x32) {
  pdl_wait();
  pdl_launch_dependents();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int S = Lq + 2 * R;
  const int b = row / S, s = row % S;
  Row x;
  if (s < Lq) {
    // text = LN(E[q] + Ttype[seg] + Pos[s])  (pixelmodel.py:560-600)
    row_load(x, E + int64_t(__ldg(query_ids + b * Lq + s)) * kH, lane);
    row_add(x, T + int64_t(__ldg(segment_ids + b * Lq + s)) * kH, lane);
    row_add(x, P + int64_t(s) * kH, lane);
    row_layernorm(x, gamma, beta, lane);
  } else if (s < Lq + R) {
    row_load(x, region32 + (int64_t(b) * R + (s - Lq)) * kH, lane);  // raw: no type/pos/LN (pixelmodel.py:601)
  } else {
    // reshape4D quirk (pixelmodel.py:489-498): out[j] = sum_c E[id[floor(8j/H)]][(8j mod H) + c] * wl[c]
    const int r = s - Lq - R;
    float w8[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) w8[c] = __ldg(wl + c);
#pragma unroll
    for (int i = 0; i < kNV; ++i) {
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = col_of(i, lane) + e;
        const int tok = (8 * j) / kH, h0 = (8 * j) % kH;
        const float* src = E + int64_t(__ldg(label_ids + (int64_t(b) * R + r) * 8 + tok)) * kH + h0;
        const float4 a = __ldg(reinterpret_cast<const float4*>(src));
        const float4 c4 = __ldg(reinterpret_cast<const float4*>(src + 4));
        o[e] = a.x * w8[0] + a.y * w8[1] + a.z * w8[2] + a.w * w8[3] + c4.x * w8[4] + c4.y * w8[5] +
               c4.z * w8[6] + c4.w * w8[7];
      }
      x.v[i] = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
  row_store<E16>(x, x16 + int64_t(row) * kH, x32 + int64_t(row) * kH, lane);
}

// ---------------------------------------------------------------------------------------------- lxmert
// lang0 = LN(E[q] + Pos[s] + Ttype[0])  (modeling.py:283-297)
template <class E16>
__global__ void __launch_bounds__(256)
lx_lang_embed_kernel(const int32_t* __restrict__ query_ids, const int32_t* __restrict__ pair_map, int n_map,
                     const float* __restrict__ E, const float* __restrict__ T, const float* __restrict__ P,
                     const float* __restrict__ gamma, const float* __restrict__ beta, int Lq, int rows,
                     typename E16::T* __restrict__ x16, float* __restrict__ x32) {
  pdl_wait();
  pdl_launch_dependents();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int s = row % Lq;
  // output row group u embeds the query of pair pair_map[u] (one row group per DISTINCT query), or of pair u itself
  // (row groups beyond the n_map distinct queries are padding up to a whole GEMM tile: they repeat the last one)
  const int src_pair = pair_map != nullptr ? __ldg(pair_map + min(row / Lq, n_map - 1)) : row / Lq;
  Row x;
  row_load(x, E + int64_t(__ldg(query_ids + int64_t(src_pair) * Lq + s)) * kH, lane);
  row_add(x, P + int64_t(s) * kH, lane);
  row_add(x, T, lane);
  row_layernorm(x, gamma, beta, lane);
  row_store<E16>(x, x16 + int64_t(row) * kH, x32 + int64_t(row) * kH, lane);
}

// Language stream computed once per distinct query (mmr_inputs.lang_unique / lang_slot): the compact key mask of the
// representatives, and the expansion of the compact stream to all pairs before the cross-modality blocks.
__global__ void __launch_bounds__(256)
lx_gather_mask_kernel(const int32_t* __restrict__ mask, const int32_t* __restrict__ pair_map, int n_map, int Lq, int n,
                      int32_t* __restrict__ out) {
  pdl_wait();
  pdl_launch_dependents();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __ldg(mask + int64_t(__ldg(pair_map + min(i / Lq, n_map - 1))) * Lq + i % Lq);
}
template <class E16>
__global__ void __launch_bounds__(256)
lx_expand_rows_kernel(const float* __restrict__ src32, const typename E16::T* __restrict__ src16,
                      const int32_t* __restrict__ slot, int Lq, int rows, float* __restrict__ dst32,
                      typename E16::T* __restrict__ dst16) {
  pdl_wait();
  pdl_launch_dependents();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int64_t src = int64_t(__ldg(slot + row / Lq)) * Lq + row % Lq;
#pragma unroll
  for (int i = 0; i < kNV; ++i) {
    const int c = col_of(i, lane);
    *reinterpret_cast<float4*>(dst32 + int64_t(row) * kH + c) = __ldg(reinterpret_cast<const float4*>(src32 + src * kH + c));
    *reinterpret_cast<uint2*>(dst16 + int64_t(row) * kH + c) = __ldg(reinterpret_cast<const uint2*>(src16 + src * kH + c));
  }
}

// z[b,r,:] = bconv + sum_t wconv[t] * LN(E[id_t] + Pos[t] + Ttype[0])   (modeling.py:915, 526-527)
template <class E16>
__global__ void __launch_bounds__(256)
lx_label_z_kernel(const int32_t* __restrict__ label_ids, const float* __restrict__ E, const float* __restrict__ T,
                  const float* __restrict__ P, const float* __restrict__ gamma, const float* __restrict__ beta,
                  const float* __restrict__ wconv, const float* __restrict__ bconv, int rows,
                  typename E16::T* __restrict__ z16, float* __restrict__ z32) {
  pdl_wait();
  pdl_launch_dependents();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  Row z;
  const float b0 = __ldg(bconv);
#pragma unroll
  for (int i = 0; i < kNV; ++i) z.v[i] = make_float4(b0, b0, b0, b0);
  for (int t = 0; t < 8; ++t) {
    Row x;
    row_load(x, E + int64_t(__ldg(label_ids + int64_t(row) * 8 + t)) * kH, lane);
    row_add(x, P + int64_t(t) * kH, lane);
    row_add(x, T, lane);
    row_layernorm(x, gamma, beta, lane);
    row_axpy(z, __ldg(wconv + t), x);
  }
  row_store<E16>(z, z16 ? z16 + int64_t(row) * kH : nullptr, z32 ? z32 + int64_t(row) * kH : nullptr, lane);
}

// acc32[row,:] += scale * LN_b(box4 . Wb^T + bb)   (modeling.py:524-525, 530); Wb is torch [768,4]
__global__ void __launch_bounds__(256)
lx_box_ln_kernel(const float* __restrict__ boxes4, const float* __restrict__ Wb, const float* __restrict__ bb,
                 const float* __restrict__ gamma, const float* __restrict__ beta, float scale, int rows,
                 float* __restrict__ acc32) {
  pdl_wait();
  pdl_launch_dependents();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float4 bx = __ldg(reinterpret_cast<const float4*>(boxes4 + int64_t(row) * 4));
  Row y;
#pragma unroll
  for (int i = 0; i < kNV; ++i) {
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = col_of(i, lane) + e;
      const float4 wr = __ldg(reinterpret_cast<const float4*>(Wb + int64_t(j) * 4));
      o[e] = bx.x * wr.x + bx.y * wr.y + bx.z * wr.z + bx.w * wr.w + __ldg(bb + j);
    }
    y.v[i] = make_float4(o[0], o[1], o[2], o[3]);
  }
  row_layernorm(y, gamma, beta, lane);
  float* dst = acc32 + int64_t(row) * kH;
#pragma unroll
  for (int i = 0; i < kNV; ++i) {
    float4* p = reinterpret_cast<float4*>(dst + col_of(i, lane));
    float4 o = *p;
    o.x = fmaf(scale, y.v[i].x, o.x); o.y = fmaf(scale, y.v[i].y, o.y);
    o.z = fmaf(scale, y.v[i].z, o.z); o.w = fmaf(scale, y.v[i].w, o.w);
    *p = o;
  }
}

// ---------------------------------------------------------------------------------------------- heads
__device__ __forceinline__ void softmax2(float l0, float l1, float* out) {
  const float m = fmaxf(l0, l1);
  const float e0 = expf(l0 - m), e1 = expf(l1 - m);
  const float inv = 1.0f / (e0 + e1);
  out[0] = e0 * inv;
  out[1] = e1 * inv;
}

// AM-softmax head (model_triple.py:56-86).  wn = column-normalised am_kernel stored as [2,768].
__global__ void __launch_bounds__(256)
zk_head_kernel(const float* __restrict__ pooled, const float* __restrict__ wn, const int32_t* __restrict__ labels,
               int B, float* __restrict__ probs, float* __restrict__ logits) {
  pdl_wait();
  pdl_launch_dependents();
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const int lane = threadIdx.x & 31;
  Row x, w0, w1;
  row_load(x, pooled + int64_t(b) * kH, lane);
  row_load(w0, wn, lane);
  row_load(w1, wn + kH, lane);
  float ss = 0.f, d0 = 0.f, d1 = 0.f;
#pragma unroll
  for (int i = 0; i < kNV; ++i) {
    ss += x.v[i].x * x.v[i].x + x.v[i].y * x.v[i].y + x.v[i].z * x.v[i].z + x.v[i].w * x.v[i].w;
    d0 += x.v[i].x * w0.v[i].x + x.v[i].y * w0.v[i].y + x.v[i].z * w0.v[i].z + x.v[i].w * w0.v[i].w;
    d1 += x.v[i].x * w1.v[i].x + x.v[i].y * w1.v[i].y + x.v[i].z * w1.v[i].z + x.v[i].w * w1.v[i].w;
  }
  ss = warp_sum(ss); d0 = warp_sum(d0); d1 = warp_sum(d1);
  if (lane == 0) {
    const float inv = rsqrtf(fmaxf(ss, 1e-12f));            // tf.nn.l2_normalize(dim=1), eps 1e-12
    float c0 = fminf(fmaxf(d0 * inv, -1.f), 1.f);           // clip_by_value(-1, 1)
    float c1 = fminf(fmaxf(d1 * inv, -1.f), 1.f);
    const int y = labels[b];
    const float g = y ? c1 : c0;
    const float m = g > 0.35f ? 0.35f : 0.f;                // margin only when the fed label's cosine > m
    if (y) c1 -= m; else c0 -= m;
    softmax2(30.f * c0, 30.f * c1, probs + 2 * b);
    if (logits != nullptr) { logits[2 * b] = 30.f * c0; logits[2 * b + 1] = 30.f * c1; }
  }
}

// logits = x . W^T + b (W [2,768]); optional LayerNorm over `width` (<= 1536) first (LXMERT logit_fc.2/3).
__global__ void __launch_bounds__(256)
linear_head_kernel(const float* __restrict__ x, int width, const float* __restrict__ ln_gamma,
                   const float* __restrict__ ln_beta, const float* __restrict__ W, const float* __restrict__ bias,
                   int B, float* __restrict__ probs, float* __restrict__ logits) {
  pdl_wait();
  pdl_launch_dependents();
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const int lane = threadIdx.x & 31;
  const int nv = width >> 7;
  float4 v[12];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    if (i < nv) {
      v[i] = __ldg(reinterpret_cast<const float4*>(x + int64_t(b) * width + col_of(i, lane)));
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  if (ln_gamma != nullptr) {
    const float mean = warp_sum(s) / float(width);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      if (i < nv) {
        const float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        q += (a * a + bb * bb) + (c * c + d * d);
      }
    }
    const float rstd = rsqrtf(warp_sum(q) / float(width) + 1e-12f);
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      if (i < nv) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(ln_gamma + col_of(i, lane)));
        const float4 be = __ldg(reinterpret_cast<const float4*>(ln_beta + col_of(i, lane)));
        v[i].x = (v[i].x - mean) * rstd * g.x + be.x; v[i].y = (v[i].y - mean) * rstd * g.y + be.y;
        v[i].z = (v[i].z - mean) * rstd * g.z + be.z; v[i].w = (v[i].w - mean) * rstd * g.w + be.w;
      }
    }
  }
  float d0 = 0.f, d1 = 0.f;
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    if (i < nv) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(W + col_of(i, lane)));
      const float4 c = __ldg(reinterpret_cast<const float4*>(W + width + col_of(i, lane)));
      d0 += v[i].x * a.x + v[i].y * a.y + v[i].z * a.z + v[i].w * a.w;
      d1 += v[i].x * c.x + v[i].y * c.y + v[i].z * c.z + v[i].w * c.w;
    }
  }
  d0 = warp_sum(d0); d1 = warp_sum(d1);
  if (lane == 0) {
    softmax2(d0 + bias[0], d1 + bias[1], probs + 2 * b);
    if (logits != nullptr) { logits[2 * b] = d0 + bias[0]; logits[2 * b + 1] = d1 + bias[1]; }
  }
}

// ---------------------------------------------------------------------------------------------- launchers
static inline int blocks_for(int rows) { return (rows + 7) / 8; }

#define MMR_DISPATCH16(dtype, CALL)                                        \
  do {                                                                     \
    if ((dtype) == MMR_DT_BF16) { using E16 = BF16; CALL; }                \
    else { using E16 = FP16; CALL; }                                       \
  } while (0)

mmr_status zk_region_sum(const float* feat32, const float* boxes5, const int32_t* label_ids, const float* tables,
                         int vocab, const float* bc1, const float* Wb, const float* bb, void* out16, int rows,
                         int dtype, cudaStream_t st, float* out32) {
  MMR_DISPATCH16(dtype, ((void)launch_pdl(zk_region_sum_kernel<E16>, dim3(blocks_for(rows)), dim3(256), 0, st, 
                            feat32, boxes5, label_ids, tables, vocab, bc1, Wb, bb,
                            static_cast<typename E16::T*>(out16), out32, rows)));
  MMR_CUDA_OK(cudaGetLastError());
  return MMR_OK;
}

mmr_status zk_label_terms(const int32_t* label_ids, const float* tables, int vocab, const float* bc1,
                          unsigned long long* tab, uint32_t tab_mask, const uint32_t* epoch_dev, int32_t* rep,
                          float* term32, int rows, cudaStream_t st) {
  (void)launch_pdl(zk_label_claim_kernel, dim3((rows + 255) / 256), dim3(256), 0, st, label_ids, rows, tab, tab_mask,
                   epoch_dev, rep);
  MMR_CUDA_OK(cudaGetLastError());
  (void)launch_pdl(zk_label_term_kernel, dim3(rows), dim3(256), 0, st, label_ids, tables, vocab, bc1,
                   static_cast<const int32_t*>(rep), term32);
  MMR_CUDA_OK(cudaGetLastError());
  return MMR_OK;
}
mmr_status zk_region_sum_rep(const float* feat32, const float* boxes5, const int32_t* rep, const float* term32,
                             const float* Wb, const float* bb, void* out16, int rows, uint32_t* epoch_dev, int dtype,
                             cudaStream_t st, float* out32) {
  MMR_DISPATCH16(dtype, ((void)launch_pdl(zk_region_sum_rep_kernel<E16>, dim3(blocks_for(rows)), dim3(256), 0, st,
                            feat32, boxes5, rep, term32, Wb, bb, static_cast<typename E16::T*>(out16), out32, rows,
                            epoch_dev)));
  MMR_CUDA_OK(cudaGetLastError());
  return MMR_OK;
}

mmr_status zk_embed(const int32_t* query_ids, const int32_t* segment_ids, const float* region32,
                    const int32_t* len_query, const int32_t* num_boxes, const float* E, const float* T,
                    const float* P, const float* gamma, const float* beta, int Lq, int R, int B, void* x16,
                    float* x32, int32_t* key_mask, int dtype, cudaStream_t st) {
  const int rows = B * (Lq + R);
  MMR_DISPATCH16(dtype, ((void)launch_pdl(zk_embed_kernel<E16>, dim3(blocks_for(rows)), dim3(256), 0, st, 
                            query_ids, segment_ids, region32, len_query, num_boxes, E, T, P, gamma, beta, Lq, R,
                            rows, static_cast<typename E16::T*>(x16), x32, key_mask)));
  MMR_CUDA_OK(cudaGetLastError());
  return MMR_OK;
}

mmr_status lds_embed(const int32_t* query_ids, const int32_t* segment_ids, const int32_t* label_ids,
                     const float* region32, const float* E, const float* T, const float* P, const float* gamma,
                     const float* beta, const float* wl, int Lq, int R, int B, void* x16, float* x32, int dtype,
                     cudaStream_t st) {
  const int rows = B * (Lq + 2 * R);
  MMR_DISPATCH16(dtype, ((void)launch_pdl(lds_embed_kernel<E16>, dim3(blocks_for(rows)), dim3(256), 0, st, 
                            query_ids, segment_ids, label_ids, region32, E, T, P, gamma, beta, wl, Lq, R, rows,
                            static_cast<typename E16::T*>(x16), x32)));
  MMR_CUDA_OK(cudaGetLastError());
  return MMR_OK;
}

mmr_status lx_lang_embed(const int32_t* query_ids, const float* E, const float* T, const float* P,
                         const float* gamma, const float* beta, int Lq, int B, void* x16, float* x32, int dtype,
                         cudaStream_t st, const int32_t* pair_map, int n_map) {
  const int rows = B * Lq;   // B = row groups written (with pair_map: the n_map distinct queries, padded by repetition)
  MMR_DISPATCH16(dtype, ((void)launch_pdl(lx_lang_embed_kernel<E16>, dim3(blocks_for(rows)), dim3(256), 0, st, 
                            query_ids, pair_map, n_map, E, T, P, gamma, beta, Lq, rows,
                            static_cast<typename E16::T*>(x16), x32)));
  MMR_CUDA_OK(cudaGetLastError());
  return MMR_OK;
}
mmr_status lx_gather_mask(const int32_t* mask, const int32_t* pair_map, int n_map, int Lq, int groups, int32_t* out,
                          cudaStream_t st) {
  const int n = groups * Lq;
  (void)launch_pdl(lx_gather_mask_kernel, dim3((n + 255) / 256), dim3(256), 0, st, mask, pair_map, n_map, Lq, n, out);
  MMR_CUDA_OK(cudaGetLastError());
  return MMR_OK;
}
mmr_status lx_expand_rows(const float* src32, const void* src16, const int32_t* slot, int Lq, int B, float* dst32,
                          void* dst16, int dtype, cudaStream_t st) {
  const int rows = B * Lq;
  MMR_DISPATCH16(dtype, ((void)launch_pdl(lx_expand_rows_kernel<E16>, dim3(blocks_for(rows)), dim3(256), 0, st, src32,
                            static_cast<const typename E16::T*>(src16), slot, Lq, rows, dst32,
                            static_cast<typename E16::T*>(dst16))));
  MMR_CUDA_OK(cudaGetLastError());
  return MMR_OK;
}

mmr_status lx_label_z(const int32_t* label_ids, const float* E, const float* T, const float* P,
                      const float* gamma, const float* beta, const float* wconv, const float* bconv, int rows,
                      void* z16, int dtype, cudaStream_t st, float* z32) {
  MMR_DISPATCH16(dtype, ((void)launch_pdl(lx_label_z_kernel<E16>, dim3(blocks_for(rows)), dim3(256), 0, st, 
                            label_ids, E, T, P, gamma, beta, wconv, bconv, rows,
                            static_cast<typename E16::T*>(z16), z32)));
  MMR_CUDA_OK(cudaGetLastError());
  return MMR_OK;
}

mmr_status lx_box_ln(const float* boxes4, const float* Wb, const float* bb, const float* gamma, const float* beta,
                     float scale, int rows, float* acc32, cudaStream_t st) {
  (void)launch_pdl(lx_box_ln_kernel, dim3(blocks_for(rows)), dim3(256), 0, st, boxes4, Wb, bb, gamma, beta, scale, rows, acc32);
  MMR_CUDA_OK(cudaGetLastError());
  return MMR_OK;
}

mmr_status zk_head(const float* pooled, const float* wn, const int32_t* labels, int B, float* probs, float* logits,
                   cudaStream_t st) {
  (void)launch_pdl(zk_head_kernel, dim3(blocks_for(B)), dim3(256), 0, st, pooled, wn, labels, B, probs, logits);
  MMR_CUDA_OK(cudaGetLastError());
  return MMR_OK;
}

mmr_status linear_head(const float* x, int width, const float* ln_gamma, const float* ln_beta, const float* W,
                       const float* bias, int B, float* probs, float* logits, cudaStream_t st) {
  MMR_REQUIRE(width % 128 == 0 && width <= 1536, "linear_head: width %d unsupported", width);
  (void)launch_pdl(linear_head_kernel, dim3(blocks_for(B)), dim3(256), 0, st, x, width, ln_gamma, ln_beta, W, bias, B, probs, logits);
  MMR_CUDA_OK(cudaGetLastError());
  return MMR_OK;
}

}  // namespace mmr

extern "C" mmr_status mmr_am_softmax_head(const float* pooled, const float* wn, const int32_t* labels, int B,
                                          float* probs, float* logits, void* stream) {
  MMR_TRY(mmr::require_sm100());
  MMR_REQUIRE(pooled && wn && labels && probs && B > 0, "mmr_am_softmax_head: null argument");
  return mmr::zk_head(pooled, wn, labels, B, probs, logits, static_cast<cudaStream_t>(stream));
}
extern "C" mmr_status mmr_linear_head(const float* x, int width, const float* ln_gamma, const float* ln_beta,
                                      const float* W, const float* bias, int B, float* probs, float* logits,
                                      void* stream) {
  MMR_TRY(mmr::require_sm100());
  MMR_REQUIRE(x && W && bias && probs && B > 0, "mmr_linear_head: null argument");
  MMR_REQUIRE((ln_gamma == nullptr) == (ln_beta == nullptr), "mmr_linear_head: give both LayerNorm vectors or none");
  return mmr::linear_head(x, width, ln_gamma, ln_beta, W, bias, B, probs, logits, static_cast<cudaStream_t>(stream));
}

// Box normalisation of the loaders (load_data_v4.py:142-145; lxmert utils.py:31) on the device.  The reference divides
// float32 boxes by a Python list of ints, i.e. in float64, and stores float32: reproduced with double division.
__global__ void boxes_normalize_kernel(const float* __restrict__ b4, const int32_t* __restrict__ hh,
                                       const int32_t* __restrict__ ww, int64_t n_slots, int R, int with_area,
                                       float* __restrict__ out) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_slots) return;
  const int64_t rec = i / R;
  const double h = double(hh[rec]), w = double(ww[rec]);
  const float4 b = *reinterpret_cast<const float4*>(b4 + 4 * i);
  const int od = with_area ? 5 : 4;
  float* o = out + od * i;
  o[0] = float(double(b.x) / h);
  o[1] = float(double(b.y) / w);
  o[2] = float(double(b.z) / h);
  o[3] = float(double(b.w) / w);
  if (with_area) o[4] = float(double((b.z - b.x) * (b.w - b.y)) / (w * h));
}
extern "C" mmr_status mmr_boxes_normalize(const float* boxes4, const int32_t* image_h, const int32_t* image_w, int64_t n,
                                          int max_boxes, int with_area, float* out, void* stream) {
  MMR_TRY(mmr::require_sm100());
  MMR_REQUIRE(boxes4 && image_h && image_w && out && n > 0 && max_boxes > 0, "mmr_boxes_normalize: bad argument");
  const int64_t slots = n * max_boxes;
  boxes_normalize_kernel<<<unsigned((slots + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      boxes4, image_h, image_w, slots, max_boxes, with_area, out);
  MMR_CUDA_OK(cudaGetLastError());
  return MMR_OK;
}

// Internal (C++) launcher declarations shared by the model driver; the C ABI wraps a subset of these.
#pragma once
#include "common.cuh"

namespace mmr {

mmr_status gemm(const void* A16, int64_t lda, const void* W16, int64_t ldw, int M, int N, int K,
                const float* bias, const float* residual, int64_t ldr, void* out16, int64_t ldo16,
                float* out32, int64_t ldo32, int act, int dtype, cudaStream_t stream, bool single_cta_only = false);
// single_cta_only: always the 128-row single-CTA kernel, whatever M -- the [CLS]-row tail of the last block uses it so
// that the kernel (hence the bits of every row) does not depend on the batch size
mmr_status layernorm(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps, int M, int H,
                     void* out16, int64_t ldo16, float* out32, int64_t ldo32, float scale, int accumulate,
                     int dtype, cudaStream_t stream);
mmr_status attention(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                     const int32_t* key_mask, void* out16, int64_t ldo, int B, int Sq, int Sk, int heads,
                     int dtype, cudaStream_t stream);
// tcgen05 / TMEM attention (attention_tc.cu); same contract as attention()
bool attention_tc_eligible(const void* out16, int64_t ldo);
mmr_status attention_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                        const int32_t* key_mask, void* out16, int64_t ldo, int B, int Sq, int Sk, int heads, int dtype,
                        cudaStream_t stream);
// tcgen05 / TMEM attention pipelined four items deep per SM, P kept in TMEM (attention_tc2.cu); same contract
bool attention_tc2_eligible(const void* out16, int64_t ldo);
mmr_status attention_tc2(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                         const int32_t* key_mask, void* out16, int64_t ldo, int B, int Sq, int Sk, int heads, int dtype,
                         cudaStream_t stream);
// two attention problems in one launch of the tcgen05 kernel (attention_tc2.cu)
struct AttentionArgs {
  const void *q, *k, *v;
  int64_t ldq, ldk, ldv;
  const int32_t* key_mask;
  void* out16;
  int64_t ldo;
  int B, Sq, Sk;
};
mmr_status attention_pair(const AttentionArgs& a, const AttentionArgs& b, int heads, int dtype, cudaStream_t stream);
mmr_status cast16(const float* x, void* out16, int64_t n, int dtype, cudaStream_t stream);
// out = LN(A . W^T + bias + residual) for N = 768 in one kernel (gemm_ln_sm100.cu); residual may alias out32.
bool gemm_ln_eligible(int M, int N, int K, int dtype);
// Row-statistics exchange table of the fused kernel: owned by the CALLER (one per mmr_handle, carved out of its
// workspace; a per-call scratch for the standalone operator), never shared between handles, streams or graphs.
struct LnTable {
  void* stats = nullptr;      // [m_tiles][6][256] 16-byte {mean, tag, M2, tag} entries
  uint32_t* epoch = nullptr;  // [0] tag of the next launch, [1] CTAs finished
  int m_tiles = 0;
};
size_t gemm_ln_table_bytes(int M);
mmr_status gemm_ln_table_init(void* mem_256_aligned, int M, LnTable* out, cudaStream_t stream);
mmr_status gemm_ln(const void* A16, int64_t lda, const void* W16, int64_t ldw, int M, int K, const float* bias,
                   const float* residual, int64_t ldr, const float* gamma, const float* beta, float eps, void* out16,
                   int64_t ldo16, float* out32, int64_t ldo32, int dtype, const LnTable& table, cudaStream_t stream);

// the same with a second weight / bias / gamma / beta set for the rows from split_row on (a multiple of 256)
mmr_status gemm_ln_2w(const void* A16, int64_t lda, const void* W16, const void* W16b, int64_t ldw, int M, int K,
                      const float* bias, const float* biasb, const float* residual, int64_t ldr, const float* gamma,
                      const float* gammab, const float* beta, const float* betab, int split_row, float eps, void* out16,
                      int64_t ldo16, float* out32, int64_t ldo32, int dtype, const LnTable& table, cudaStream_t stream);
// same contract, one CTA pair per 256-row block and all 768 columns (gemm_lnrow_sm100.cu); reached through gemm_ln()
mmr_status gemm_lnrow(const void* A16, int64_t lda, const void* W16, int64_t ldw, int M, int K, const float* bias,
                      const float* residual, int64_t ldr, const float* gamma, const float* beta, float eps, void* out16,
                      int64_t ldo16, float* out32, int64_t ldo32, int dtype, cudaStream_t stream);

mmr_status zk_region_sum(const float* feat32, const float* boxes5, const int32_t* label_ids, const float* tables,
                         int vocab, const float* bc1, const float* Wb, const float* bb, void* out16, int rows,
                         int dtype, cudaStream_t st, float* out32 = nullptr);
// label term once per distinct label phrase of the batch (claim + term kernels), then the per-box sum
mmr_status zk_label_terms(const int32_t* label_ids, const float* tables, int vocab, const float* bc1,
                          unsigned long long* tab, uint32_t tab_mask, const uint32_t* epoch_dev, int32_t* rep,
                          float* term32, int rows, cudaStream_t st);
mmr_status zk_region_sum_rep(const float* feat32, const float* boxes5, const int32_t* rep, const float* term32,
                             const float* Wb, const float* bb, void* out16, int rows, uint32_t* epoch_dev, int dtype,
                             cudaStream_t st, float* out32 = nullptr);
mmr_status zk_embed(const int32_t* query_ids, const int32_t* segment_ids, const float* region32,
                    const int32_t* len_query, const int32_t* num_boxes, const float* E, const float* T,
                    const float* P, const float* gamma, const float* beta, int Lq, int R, int B, void* x16,
                    float* x32, int32_t* key_mask, int dtype, cudaStream_t st);
mmr_status lds_embed(const int32_t* query_ids, const int32_t* segment_ids, const int32_t* label_ids,
                     const float* region32, const float* E, const float* T, const float* P, const float* gamma,
                     const float* beta, const float* wl, int Lq, int R, int B, void* x16, float* x32, int dtype,
                     cudaStream_t st);
mmr_status lx_lang_embed(const int32_t* query_ids, const float* E, const float* T, const float* P,
                         const float* gamma, const float* beta, int Lq, int B, void* x16, float* x32, int dtype,
                         cudaStream_t st, const int32_t* pair_map = nullptr, int n_map = 0);
mmr_status lx_gather_mask(const int32_t* mask, const int32_t* pair_map, int n_map, int Lq, int groups, int32_t* out,
                          cudaStream_t st);
mmr_status lx_expand_rows(const float* src32, const void* src16, const int32_t* slot, int Lq, int B, float* dst32,
                          void* dst16, int dtype, cudaStream_t st);
mmr_status lx_label_z(const int32_t* label_ids, const float* E, const float* T, const float* P,
                      const float* gamma, const float* beta, const float* wconv, const float* bconv, int rows,
                      void* z16, int dtype, cudaStream_t st, float* z32 = nullptr);
mmr_status lx_box_ln(const float* boxes4, const float* Wb, const float* bb, const float* gamma, const float* beta,
                     float scale, int rows, float* acc32, cudaStream_t st);
// attention for the first query row of every pair only (cls_tail.cu): the last encoder block
mmr_status cls_attention(const void* q, int64_t q_pair_stride, const void* k, const void* v, int64_t ldkv,
                         const int32_t* key_mask, void* out16, int64_t ldo, int B, int Sk, int heads, int dtype,
                         cudaStream_t stream);
// LayerNorm + pooler + 2-way head of the B [CLS] rows in one kernel (cls_tail.cu); head_kind 0 AM-softmax, 1 linear
mmr_status cls_pool_head(const float* y32, const float* gamma, const float* beta, const void* Wp16, const float* bp,
                         int head_kind, const float* hw, const float* hb, const int32_t* labels, int B, float* pooled32,
                         float* probs, float* logits, int dtype, cudaStream_t stream);
// strict precision mode (strict.cu): two-term operand split, precise activations, fp32 attention
mmr_status split3(const float* x, int64_t ldx, int rows, int K, void* out16, int64_t ldo, int act, int weights, int dtype,
                  cudaStream_t stream);
mmr_status act32(float* x, int64_t n, int act, cudaStream_t stream);
mmr_status attention_f32(const float* q, int64_t q_pair, int64_t ldq, const float* k, const float* v, int64_t kv_pair,
                         int64_t ldkv, const int32_t* key_mask, float* out, int64_t o_pair, int64_t ldo, int B, int Sq,
                         int Sk, int heads, cudaStream_t stream);
mmr_status zk_head(const float* pooled, const float* wn, const int32_t* labels, int B, float* probs, float* logits,
                   cudaStream_t st);
mmr_status linear_head(const float* x, int width, const float* ln_gamma, const float* ln_beta, const float* W,
                       const float* bias, int B, float* probs, float* logits, cudaStream_t st);

}  // namespace mmr

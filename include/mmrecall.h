/*
 * mmrecall.h — C ABI of the B200-native (sm_100a) cross-modal match scorer.
 *
 * The reference (zuokai/KDDCUP_2020_MultimodalitiesRecall_2nd_Place) has no FFI layer: its hot path is
 * reached through Python calls into TensorFlow-1 / PyTorch-1 stock ops.  This header is the boundary a
 * maintainer binds instead (ctypes stub in INTEGRATION.md).  Each entry point names the reference
 * call site it replaces.  Plain pointers and sizes only; no torch / C++ types cross this ABI.
 *
 * Conventions
 *   - every pointer marked "dev" is a device pointer valid on the current CUDA device;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - calls are asynchronous w.r.t. the host, ordered on `stream`; nothing here synchronises;
 *   - 16-bit operands are bf16 (MMR_DT_BF16) or fp16 (MMR_DT_FP16); accumulation, LayerNorm statistics,
 *     softmax and the residual stream are fp32;
 *   - weights are [out, in] row-major ("K-major"), i.e. torch nn.Linear.weight; a TF `kernel` [in,out]
 *     is transposed by the packer (mmr_create does it for you);
 *   - return value: MMR_OK or an error code, with text in mmr_last_error(); no exceptions cross the ABI;
 *   - there is NO CPU fallback: on a device that is not sm_100 every entry returns MMR_ERR_ARCH.
 */
#ifndef MMRECALL_H_
#define MMRECALL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  MMR_OK = 0,
  MMR_ERR_INVALID = 1, /* bad argument: shape / alignment / null pointer            */
  MMR_ERR_CUDA = 2,    /* a CUDA runtime / driver call failed                        */
  MMR_ERR_ARCH = 3,    /* device is not sm_100 (B200); no fallback path exists       */
  MMR_ERR_NOMEM = 4,   /* device allocation failed                                   */
  MMR_ERR_WEIGHTS = 5  /* a required weight tensor is missing or has the wrong shape */
} mmr_status;

typedef enum { MMR_DT_FP16 = 0, MMR_DT_BF16 = 1 } mmr_dtype;

/* Arithmetic of the model-level forward (mmr_config.precision).
 *   MMR_PRECISION_FAST    every MMA operand rounded once to `dtype` (11 significand bits for fp16); fp32 accumulate,
 *                         residual stream, LayerNorm, softmax.  Within 1e-3 of the fp32 reference on weights of the
 *                         reference initialisers; up to 3e-3 on weights of trained magnitude.
 *   MMR_PRECISION_STRICT  two-term split operands (x = hi + lo) on EVERY tensor-core GEMM, the three significant partial
 *                         products accumulated in fp32 in one K-concatenated MMA chain; precise GELU / tanh; attention
 *                         in fp32.  ~1e-5 of the fp32 reference on either weight set, at about a third of the throughput. */
typedef enum { MMR_PRECISION_FAST = 0, MMR_PRECISION_STRICT = 1 } mmr_precision;

typedef enum {
  MMR_ACT_NONE = 0,
  MMR_ACT_RELU = 1,      /* slim.conv2d default activation: imagebert_zk/model_triple.py:189,193 */
  MMR_ACT_GELU_TANH = 2, /* imagebert_zk/pixelbert.py:315-328                                   */
  MMR_ACT_GELU_ERF = 3,  /* lxmert/src/lxrt/modeling.py:113-119                                 */
  MMR_ACT_TANH = 4       /* pooler: pixelbert.py:258-266, modeling.py:596-608                   */
} mmr_act;

typedef enum {
  MMR_MODEL_IMAGEBERT_ZK = 0,  /* code/imagebert_zk  (ImageBertB / C) */
  MMR_MODEL_IMAGEBERT_LDS = 1, /* code/imagebert_lds (ImageBertA)     */
  MMR_MODEL_LXMERT = 2         /* code/lxmert                          */
} mmr_model_kind;

const char* mmr_last_error(void);
int mmr_abi_version(void);
/* 1 when the library was built with MMR_EXPERIMENTAL (csrc/build.py): the kernel variants that lost their A/B
 * measurements -- mma.sync attention, first tcgen05 attention, row-owner GEMM+LayerNorm, 4-CTA multicast GEMM -- are
 * compiled in and their knobs below are live.  In the default build (0) those knobs are accepted and ignored:
 * MMR_TUNE_ATTN_TMA, MMR_TUNE_ATTN_TC, MMR_TUNE_LN_ROW_CFG, MMR_TUNE_GEMM_CLUSTER = 2, MMR_TUNE_GEMM_LN = 2 / 3. */
int mmr_experimental_build(void);
/* Kernel-selection knobs, for A/B measurements and tests (defaults are the fastest measured; the environment
 * variable of the same meaning is read once at first use): value 0 disables / selects the older path.
 *   MMR_TUNE_GEMM_PAIR     (env MMR_GEMM_PAIR,    default 1) CTA-pair (cta_group::2) GEMM kernels
 *   MMR_TUNE_GEMM_P16      (env MMR_GEMM_P16,     default 1) 16-bit-output GEMM with the TMA-store epilogue
 *   MMR_TUNE_GEMM_TAIL     (env MMR_GEMM_TAIL,    default 1) split the last partial wave of that GEMM along N
 *   MMR_TUNE_GEMM_CLUSTER  (env MMR_GEMM_CLUSTER, default 1) 1 = lone CTA pairs, 2 = 4-CTA clusters sharing W by
 *                                                            TMA multicast (measured slower: 132 of 148 SMs)
 *   MMR_TUNE_GEMM_LN       (env MMR_GEMM_LN,      default 1) fused projection + residual + LayerNorm kernel: 1 = three
 *                                                            CTA pairs per 256-row block meeting through a global table
 *                                                            (gemm_ln_sm100.cu), 2 = one pair owns the block and all 768
 *                                                            columns (gemm_lnrow_sm100.cu; 3 = that one only for K <= 1024),
 *                                                            0 = GEMM + LayerNorm kernels
 *   MMR_TUNE_PDL           (env MMR_PDL,          default 1) programmatic dependent launch between the kernels
 *   MMR_TUNE_ATTN_TMA      (env MMR_ATTN_TMA,     default 0) persistent TMA-pipelined mma.sync attention kernel
 *                                                            (measured slower than one CTA per (pair, head): 40 vs 32 us)
 *   MMR_TUNE_ATTN_TC       (env MMR_ATTN_TC,      default 2) tcgen05 / TMEM attention: 2 = pipelined four items deep per
 *                                                            SM with P kept in TMEM (attention_tc2.cu: 24-25 us at B=256,
 *                                                            S=68 against 31.7 us for the mma.sync kernel), 1 = first
 *                                                            version, one item in flight per CTA (attention_tc.cu, 40 us),
 *                                                            0 = mma.sync kernels
 *   MMR_TUNE_LN_ROW_CFG    (env MMR_LN_ROW_CFG,   default 0) shared-memory split of the row-owner GEMM+LN kernel as
 *                                                            three digits (operand stages, fp32 chunk slots, 16-bit
 *                                                            stages per epilogue warp): 421, 331, 511, 412, 322; 0 = default 
 *   MMR_TUNE_LABEL_DEDUP   (env MMR_LABEL_DEDUP,  default 1) zk label-text term evaluated once per distinct label phrase
 *                                                            of the batch (bit-identical to the per-box evaluation, 0) 
 *   MMR_TUNE_LX_MERGE      (env MMR_LX_MERGE,     default 1) LXMERT: the projections of the language and the visual
 *                                                            stream (different weights, one activation buffer) as ONE
 *                                                            launch each when batch x query length is a multiple of 256
 *   MMR_TUNE_PRUNE_LAST    (env MMR_PRUNE_LAST,   default 1) last encoder block: the scorers read sequence_output[:, 0]
 *                                                            only (pixelbert.py:258-266, modeling.py:925), so keys / values
 *                                                            are projected for all rows but attention, output projection,
 *                                                            FFN and both LayerNorms run for the B [CLS] rows alone (LXMERT:
 *                                                            the last cross layer's visual half, modeling.py:468-479, feeds
 *                                                            nothing and is skipped); 0 = the full last block.  Forced off
 *                                                            while mmr_set_debug_taps is non-zero (taps want every row).
 *   MMR_TUNE_LX_QUERY_DEDUP (env MMR_LX_QUERY_DEDUP, default 1) LXMERT: honour mmr_inputs.lang_unique / lang_slot (the
 *                                                            language-only blocks once per distinct query of the batch);
 *                                                            0 = ignore them.  Also off while debug taps are on. */
enum { MMR_TUNE_GEMM_PAIR = 0, MMR_TUNE_GEMM_P16 = 1, MMR_TUNE_GEMM_TAIL = 2, MMR_TUNE_GEMM_CLUSTER = 3,
       MMR_TUNE_GEMM_LN = 4, MMR_TUNE_PDL = 5, MMR_TUNE_ATTN_TMA = 6, MMR_TUNE_ATTN_TC = 7, MMR_TUNE_LN_ROW_CFG = 8,
       MMR_TUNE_LABEL_DEDUP = 9, MMR_TUNE_LX_MERGE = 10, MMR_TUNE_PRUNE_LAST = 11, MMR_TUNE_LX_QUERY_DEDUP = 12,
       MMR_TUNE_COUNT = 13 };
mmr_status mmr_set_tuning(int knob, int value);
/* Current value of a knob (-1 for an unknown one). */
int mmr_get_tuning(int knob);
/* Counts mmr_set_tuning calls: a caller that replays captured CUDA graphs of mmr_forward re-captures when it moves. */
unsigned mmr_tuning_generation(void);
/* MMR_OK iff `device` is an sm_100 part. */
mmr_status mmr_device_check(int device);

/* ------------------------------------------------------------------------------------------------
 * Input decode (host code, no GPU): the per-line work of the reference loaders -- tab split, base64 decode of boxes /
 * 2048-d features / class labels, zero padding to the box budget -- imagebert_zk/load_data_v4.py:133-163, 91-102,
 * 380-383; lxmert/src/utils.py:23-36.  Lines are decoded by `n_threads` threads (<= 0: all cores) straight into the
 * caller's batch arrays (use pinned memory to feed the scorer).  Tokenisation of `query` and of the class-label phrases
 * stays with the caller (tokenizer.py); boxes come back RAW, normalisation /[h, w, h, w] (+ area) is mmr_boxes_normalize.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t max_boxes;      /* R: box slots per record (reference MAX_BOX_NUM = 10)                              */
  int32_t feat_dim;       /* 2048                                                                              */
  int64_t* product_id;    /* [n]                                                                               */
  int32_t* image_h;       /* [n]                                                                               */
  int32_t* image_w;       /* [n]                                                                               */
  int32_t* num_boxes;     /* [n] as stored in the file (may exceed R; slots hold the first R)                  */
  float* boxes4;          /* [n, R, 4] raw pixel boxes, zero padded                                            */
  float* feats;           /* [n, R, feat_dim] fp32, zero padded                                                */
  int64_t* class_labels;  /* [n, R] detector class ids, zero padded                                            */
  int64_t* query_id;      /* [n]                                                                               */
  int64_t* query_off;     /* [n, 2] (offset, length) of each query string inside query_text                    */
  char* query_text;       /* concatenated UTF-8 query strings (order of arrival, not of lines)                 */
  size_t query_cap;       /* capacity of query_text in bytes                                                   */
} mmr_decode_out;
mmr_status mmr_decode_tsv(const char* const* lines, const size_t* line_len, int64_t n_lines, const mmr_decode_out* out,
                          int n_threads);
/* The same for batch arrays that are decoded into again and again (a driver's pinned staging arrays): dirty_boxes
 * [n_lines] is caller-kept state, one count per record slot of the arrays = how many leading box slots may hold
 * non-zero data (initialise to -1 or max_boxes for arrays of unknown content, 0 for zero-filled ones).  Only the slots
 * [boxes of this record, dirty) are cleared and the count is updated, instead of zero-padding every record to
 * max_boxes: at 36 slots and ~4 boxes per record the padding is 90 % of the bytes (seq_padding_2's job,
 * load_data_v4.py:91-102).  The arrays end up byte-identical to mmr_decode_tsv's. */
mmr_status mmr_decode_tsv_reuse(const char* const* lines, const size_t* line_len, int64_t n_lines,
                                const mmr_decode_out* out, int32_t* dirty_boxes, int n_threads);
/* CRC-32C (Castagnoli) of a HOST buffer, continuing from `crc` (0 to start): the per-block / per-tensor checksum of
 * the TensorFlow checkpoints the two ImageBert drivers restore (imagebert_zk/evaluate_normal.py:204-212,
 * imagebert_lds/src/run_pretraining_predict_score.py:347-362); used by the pure-Python bundle reader. */
uint32_t mmr_crc32c(const void* data, size_t n, uint32_t crc);
/* boxes5[i, r] = (x1/h, y1/w, x2/h, y2/w, (x2-x1)(y2-y1)/(w h)) exactly as load_data_v4.py:142-145 divides
 * (with_area != 0; zk) or the first four only (with_area == 0; lxmert utils.py:31).  Device pointers; fp32 division. */
mmr_status mmr_boxes_normalize(const float* boxes4, const int32_t* image_h, const int32_t* image_w, int64_t n,
                               int max_boxes, int with_area, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Ensemble on the device: code/main.py:41-104 (weighted merge, product-uniqueness filter, per-query
 * top-k with the < k fall-back) and imagebert_lds/src/evaluation.py:4-38 (nDCG@k).  All pointers are device pointers.
 * Pairs are flattened in the reference's iteration order with the pairs of a query contiguous
 * (query_start [n_queries + 1]); scores missing from a file are back-filled by the caller (main.py:50-58) before
 * upload.  weights4 is a HOST array of 4 doubles.  merged [n_pairs] fp64 (bit-identical to the host ensemble);
 * top [n_queries, topk] pair indices (-1: none); status [n_queries]: 0 = query not written (no survivor),
 * 1 = filtered top-k, 2 = fall-back top-k (written after the status-1 queries, main.py:101-104).
 * workspace: >= 20 * n_products bytes.
 * ---------------------------------------------------------------------------------------------- */
mmr_status mmr_ensemble_topk(const double* s1, const double* s2, const double* s3, const double* s4,
                             const int32_t* product_of, const int32_t* query_start, int64_t n_pairs, int32_t n_queries,
                             int32_t n_products, const double* weights4, double margin, double tie, int32_t topk,
                             double* merged, int32_t* top, int32_t* status, void* workspace, size_t workspace_bytes,
                             void* stream);
/* gt / gt_start: ground-truth product indices per query (CSR); ndcg [n_queries], -1 for queries without a prediction. */
mmr_status mmr_ndcg_at_k(const int32_t* top, const int32_t* product_of, const int32_t* gt, const int32_t* gt_start,
                         int32_t n_queries, int32_t topk, double* ndcg, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Operator level (one kernel each).  Used by the model driver below and by the parity tests.
 * ---------------------------------------------------------------------------------------------- */

/* out = act(A[M,K] * W[N,K]^T + bias) (+ residual).  tcgen05 tensor-core GEMM, TMA-fed, fp32 accumulate
 * in TMEM.  Replaces tf.layers.dense / slim.fully_connected / nn.Linear at pixelbert.py:767-788, 960-985;
 * pixelmodel.py:439-442; modeling.py:325-420, 522-523.
 *   A16   dev, 16-bit, row stride lda (elements; lda*2 bytes must be a multiple of 16)
 *   W16   dev, 16-bit [N,K], row stride ldw
 *   bias  dev fp32 [N] or NULL;  residual dev fp32 [M, ldr] or NULL (added after the activation)
 *   out16 dev 16-bit [M, ldo16] or NULL;  out32 dev fp32 [M, ldo32] or NULL (at least one required)
 * K must be a multiple of 64, N a multiple of 16. */
mmr_status mmr_gemm(const void* A16, int64_t lda, const void* W16, int64_t ldw, int M, int N, int K,
                    const float* bias, const float* residual, int64_t ldr, void* out16, int64_t ldo16,
                    float* out32, int64_t ldo32, int act, int dtype, void* stream);

/* Row LayerNorm over the last dim (biased variance, eps inside the sqrt): tf.contrib.layers.layer_norm
 * (pixelbert.py:414-417) / nn.LayerNorm(eps=1e-12) (modeling.py:266).  x fp32 [M, ldx]; writes the
 * normalised row as 16-bit (GEMM operand) and/or fp32 (residual stream); `scale` multiplies the result
 * (LXMERT's "/3", modeling.py:530); accumulate!=0 adds into out32 instead of overwriting. */
mmr_status mmr_layernorm(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps, int M,
                         int H, void* out16, int64_t ldo16, float* out32, int64_t ldo32, float scale,
                         int accumulate, int dtype, void* stream);

/* out = LayerNorm(A[M,K] * W[768,K]^T + bias + residual) * gamma + beta in ONE kernel (cluster of three CTA
 * pairs, row statistics exchanged through distributed shared memory): the "dense + residual + layer_norm" tails of
 * pixelbert.py:960-966, 977-983 / modeling.py:355-366, 409-420.  N is fixed at 768; residual fp32 [M, ldr] may
 * alias out32 (in-place residual stream).  mmr_gemm_layernorm_supported tells whether (M, K) can take this path on
 * the current device (else call mmr_gemm with a residual, then mmr_layernorm). */
mmr_status mmr_gemm_layernorm(const void* A16, int64_t lda, const void* W16, int64_t ldw, int M, int K,
                              const float* bias, const float* residual, int64_t ldr, const float* gamma,
                              const float* beta, float eps, void* out16, int64_t ldo16, float* out32, int64_t ldo32,
                              int dtype, void* stream);
int mmr_gemm_layernorm_supported(int M, int K, int dtype);

/* Match heads on pooled [B, width] fp32 rows (device pointers), the last step of every scorer.
 * mmr_am_softmax_head: model_triple.amsoftmax_loss, imagebert_zk/model_triple.py:56-86 (inference): x/|x|, cosine with
 *   the COLUMN-NORMALISED kernel wn [2, 768] (l2_normalize(kernel, 0, 1e-10) done by the caller once), clip, margin
 *   0.35 on the fed label's cosine when it exceeds 0.35, x30, softmax.  logits may be NULL.
 * mmr_linear_head: optional LayerNorm(width, eps 1e-12) then W [2, width] + bias, softmax:
 *   get_next_sentence_output, imagebert_lds/src/run_pretraining_predict_score.py:479-501 (ln_gamma = NULL, width 768)
 *   and the tail of logit_fc, lxmert/src/tasks/kdd_model.py:167-172 (width 1536). */
mmr_status mmr_am_softmax_head(const float* pooled, const float* wn, const int32_t* labels, int B, float* probs,
                               float* logits, void* stream);
mmr_status mmr_linear_head(const float* x, int width, const float* ln_gamma, const float* ln_beta, const float* W,
                           const float* bias, int B, float* probs, float* logits, void* stream);

/* Multi-head scaled-dot-product attention, softmax in fp32, additive key mask (1-m)*-10000:
 * pixelbert.py:790-850 / modeling.py:325-352.  q [B*Sq, ldq], k/v [B*Sk, ldk/ldv], head h = columns
 * [64h, 64h+64).  key_mask dev int32 [B,Sk] (1 = attend) or NULL.  Sq, Sk <= 128. */
mmr_status mmr_attention(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                         const int32_t* key_mask, void* out16, int64_t ldo, int B, int Sq, int Sk,
                         int heads, int dtype, void* stream);

/* fp32 -> 16-bit cast with 16-byte vector loads (region features [B*R,2048]). n must be a multiple of 8.  fp16
 * saturates at +-65504 instead of overflowing to inf (detector pool5 features are post-ReLU and far below that; a
 * corrupt record must not poison a whole batch with NaNs). */
mmr_status mmr_cast16(const float* x, void* out16, int64_t n, int dtype, void* stream);

/* Attention for the FIRST query row of every pair only (the [CLS] row the poolers read: pixelbert.py:258-266,
 * modeling.py:596-608): same arithmetic as mmr_attention, one CTA per pair.  q: 16-bit, the query row of pair b at
 * q + b * q_pair_stride (elements); k / v: key row j of pair b at k + (b * Sk + j) * ldkv; out16 [B, ldo]. Sk <= 128. */
mmr_status mmr_cls_attention(const void* q, int64_t q_pair_stride, const void* k, const void* v, int64_t ldkv,
                             const int32_t* key_mask, void* out16, int64_t ldo, int B, int Sk, int heads, int dtype,
                             void* stream);

/* Strict-precision building blocks (MMR_PRECISION_STRICT).
 * mmr_split3: out16[r] = [hi | lo | hi] (weights = 0) or [hi | hi | lo] (weights = 1) of act(x[r, :K]), hi = round16(x),
 *   lo = round16(x - hi); `act` is applied in full precision first (tanhf / erff GELU).  A GEMM over the concatenated 3K
 *   axis of an activation split and a weight split accumulates A_hi W_hi + A_lo W_hi + A_hi W_lo in fp32.
 * mmr_attention_f32: the attention of pixelbert.py:790-850 / modeling.py:325-352 in fp32 on the CUDA cores; pair b has
 *   its Sq query rows at q + b * q_pair + i * ldq, its Sk keys / values at k|v + b * kv_pair + j * ldkv, its output rows
 *   at out + b * o_pair + i * ldo (all in elements); Sk <= 128. */
mmr_status mmr_split3(const float* x, int64_t ldx, int rows, int K, void* out16, int64_t ldo, int act, int weights,
                      int dtype, void* stream);
mmr_status mmr_attention_f32(const float* q, int64_t q_pair, int64_t ldq, const float* k, const float* v, int64_t kv_pair,
                             int64_t ldkv, const int32_t* key_mask, float* out, int64_t o_pair, int64_t ldo, int B,
                             int Sq, int Sk, int heads, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Model level.
 * ---------------------------------------------------------------------------------------------- */

typedef struct mmr_handle mmr_handle;

typedef struct {
  int32_t model_kind;   /* mmr_model_kind                                                         */
  int32_t dtype;        /* mmr_dtype of the MMA operands                                          */
  int32_t hidden;       /* 768  (user_data/bert_config.json)                                      */
  int32_t heads;        /* 12                                                                     */
  int32_t intermediate; /* 3072                                                                   */
  int32_t vocab;        /* 21128                                                                  */
  int32_t max_pos;      /* 512                                                                    */
  int32_t type_vocab;   /* 2                                                                      */
  int32_t feat_dim;     /* 2048                                                                   */
  int32_t label_len;    /* 8 label-text tokens per box                                            */
  int32_t n_layers;     /* single-stream encoder depth (zk / lds); LXMERT: language layers        */
  int32_t n_r_layers;   /* LXMERT relational (visual) layers, else 0                              */
  int32_t n_x_layers;   /* LXMERT cross-modality layers, else 0                                   */
  int32_t lq;           /* query tokens  (reference native: 20 zk/lds, 23 lxmert)                 */
  int32_t nbox;         /* region slots  (reference native: 10)                                   */
  int32_t max_batch;    /* workspace is sized for this many pairs per forward                     */
  int32_t precision;    /* mmr_precision (ABI version 2)                                          */
} mmr_config;

/* One named fp32 host tensor, named exactly as in the reference checkpoint (TF variable name or torch
 * state_dict key). */
typedef struct {
  const char* name;
  const float* data; /* host */
  int32_t ndim;
  int64_t dims[4];
} mmr_tensor;

/* Device inputs of one forward over B pairs.  Shapes follow the reference feeds:
 * evaluate_normal.py:141-152 (zk), run_pretraining_predict_score.py:526-541 (lds), kdd_model.py:74-100
 * (lxmert).  Unused members may be NULL. */
typedef struct {
  const int32_t* query_ids;   /* [B, lq]                                                          */
  const int32_t* segment_ids; /* zk: [B, lq+nbox]; lds: [B, lq]; lxmert: NULL (all 0)              */
  const int32_t* label_ids;   /* [B, nbox, label_len]                                             */
  const float* feats;         /* [B, nbox, feat_dim] fp32                                         */
  const float* boxes;         /* zk: [B, nbox, 5]; lxmert: [B, nbox, 4]; lds: unused              */
  const int32_t* len_query;   /* zk: [B]   (tf.sequence_mask, model_triple.py:198)                */
  const int32_t* num_boxes;   /* zk: [B]   (model_triple.py:199)                                  */
  const int32_t* query_mask;  /* lxmert: [B, lq]  input_mask                                      */
  const int32_t* visn_mask;   /* lxmert: [B, nbox] visual_attention_mask                          */
  const int32_t* labels;      /* zk: [B] AM-softmax margin label (model_triple.py:66-81)          */
  const float* region_sum;    /* zk, optional: [B, nbox, hidden] fp32 = label + box + feat term ALREADY fused
                                 (the `imgfeat` argument of pixelbert.BertModel, pixelbert.py:150-186); when
                                 non-NULL, feats / boxes / label_ids are ignored                     */
  /* lxmert, optional (ABI version 2): the language stream of LXMERT's first n_layers blocks depends on the QUERY only
   * (modeling.py:577-578 runs them before any cross-attention), and a candidate set scores ~30 products per query.
   * lang_unique [n_lang_unique] = ascending pair indices, one representative per distinct (query_ids, query_mask) row of
   * the batch; lang_slot [B] = position in lang_unique of pair b's representative.  When given (and
   * MMR_TUNE_LX_QUERY_DEDUP != 0) those blocks run on n_lang_unique rows instead of B and are expanded to all pairs
   * before the cross-modality blocks: same scores, less work.  NULL / 0 = every pair computes its own.           */
  const int32_t* lang_unique;
  const int32_t* lang_slot;
  int32_t n_lang_unique;
} mmr_inputs;

mmr_status mmr_create(const mmr_config* cfg, const mmr_tensor* weights, int n_weights, int device,
                      mmr_handle** out);
/* Device memory mmr_create will allocate for `cfg`: the packed-weight arena (16-bit matrices, fp32 tables; zk adds the
 * 8 x vocab x 768 label-conv tables) and the per-forward workspace sized for max_batch.  No GPU needed; either output
 * may be NULL.  (sizing query of the boundary.) */
mmr_status mmr_workspace_bytes(const mmr_config* cfg, size_t* weight_bytes, size_t* workspace_bytes);
void mmr_destroy(mmr_handle* h);

/* Concurrency: a handle is not re-entrant, and the fused GEMM+LayerNorm kernel needs its whole grid co-resident, so
 * at most ONE mmr_forward may be in flight per device.  Forwards of different handles issued on different streams of one
 * device are serialised by the library (a cross-stream event wait) -- except while a stream is being captured into a
 * CUDA graph, where the caller must keep replays of different handles' graphs from overlapping on one device.
 *
 * Scores B pairs (B <= max_batch).  probs_out dev fp32 [B,2] (softmax of the 2-way head; the reference
 * score is column 1 — for LXMERT column -1, same thing).  logits_out dev fp32 [B,2] or NULL: the pre-softmax
 * values (zk: 30 * margin-adjusted cosines, model_triple.py:81-85; lds: logits, run_pretraining_predict_score.py:
 * 491-492; lxmert: `logit`, the third return value of KDDModel.forward).  pooled_out dev fp32 [B,hidden] or NULL.
 * Replaces model_triple.model_attention_channel_e (model_triple.py:162-214), bertmodel(...) inference
 * path (run_pretraining_predict_score.py:288-394) and KDDModel.forward (kdd_model.py:183-214). */
mmr_status mmr_forward(mmr_handle* h, const mmr_inputs* in, int B, float* probs_out, float* logits_out,
                       float* pooled_out, void* stream);

/* Debug / parity taps (BertModel.get_embedding_output / get_sequence_output, pixelbert.py:279-309): copies of
 * internal activations after the last forward (dev fp32, rows = pairs x tokens; LXMERT: all language rows, then all
 * visual rows).  which: 0 = embedding output, only kept when mmr_set_debug_taps(h, 1) was called before the forward
 * (one extra device copy per forward); 1 = final encoder layer output, all rows -- also needs taps enabled before the
 * forward, because the default forward computes the last block for the [CLS] rows only (MMR_TUNE_PRUNE_LAST);
 * 2 + i = output of encoder layer i
 * (get_all_encoder_layers; single-stream models only), kept when mmr_set_debug_taps(h, 2) was called (allocates
 * n_layers x rows x hidden floats once, copies after every layer). */
mmr_status mmr_set_debug_taps(mmr_handle* h, int enable);
mmr_status mmr_get_activation(mmr_handle* h, int which, float* dst, int64_t n_floats, void* stream);

/* Per-launch device timing inside a forward, for bench.py's roofline line: with profiling enabled every kernel
 * launch of mmr_forward is followed by a cudaEventRecord on the forward's stream.  mmr_get_profile waits for the
 * last forward and returns the number of launches n (<= cap), filling kinds[i] (0 tcgen05 GEMM, 1 attention,
 * 2 LayerNorm, 3 embedding / head row kernels), ms[i] (device time from the previous event) and flops[i]
 * (algorithmic FLOPs of that launch, multiply-add = 2); -1 on a CUDA error. */
mmr_status mmr_set_profiling(mmr_handle* h, int enable);
int mmr_get_profile(mmr_handle* h, int cap, int32_t* kinds, float* ms, double* flops);

/* Number of kernels the last mmr_forward launched (for bench.py's gpu_launches claim). */
int mmr_launches_per_forward(const mmr_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* MMRECALL_H_ */

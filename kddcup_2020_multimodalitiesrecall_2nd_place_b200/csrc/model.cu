// Model driver behind mmr_create / mmr_forward: packs reference-named fp32 weights into device-resident 16-bit
// [out,in] matrices (+ fp32 biases / LayerNorm / embedding tables) and sequences the kernels of this directory for
// the three scorers.  Replaces, end to end on the device:
//   zk     model_triple.model_attention_channel_e            (imagebert_zk/model_triple.py:162-214)
//   lds    pixelmodel.BertModel + get_next_sentence_output   (imagebert_lds/src/pixelmodel.py:145-270,
//                                                             run_pretraining_predict_score.py:288-394, 479-501)
//   lxmert KDDModel.forward                                  (lxmert/src/tasks/kdd_model.py:183-214,
//                                                             lxrt/modeling.py:872-927)
// Layout in HBM: one arena for weights, one for the per-forward workspace (sized for max_batch at create time, no
// allocation or synchronisation inside mmr_forward).  The residual stream lives in ONE fp32 buffer x32 [M,768]
// that GEMM epilogues (+bias +residual) and the LayerNorm kernel update in place, with its 16-bit mirror x16 as
// the next GEMM's TMA operand.  LXMERT's two streams share these buffers: rows [0, B*Lq) are the language
// stream, rows [B*Lq, B*(Lq+R)) the visual stream, so the weight-shared cross-attention block (modeling.py:462-463)
// runs its QKV and output projections as single GEMMs over all rows.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <algorithm>
#include <string>
#include <vector>

#include "common.cuh"
#include "gemm_common.cuh"
#include "kernels.cuh"

namespace mmr {

// ------------------------------------------------------------------------------------------ pack kernels
// dst16[r, c] = src32[r, c] (transpose = 0) or src32[c, r] (transpose = 1; src is [cols, rows]).
template <class E16>
__global__ void pack16_kernel(const float* __restrict__ src, typename E16::T* __restrict__ dst, int rows, int cols,
                              int transpose) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  if (!transpose) {
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      const int r = r0 + i, c = c0 + threadIdx.x;
      if (r < rows && c < cols) {
        uint32_t pk = E16::pack(src[int64_t(r) * cols + c], 0.f);
        reinterpret_cast<uint16_t*>(dst)[int64_t(r) * cols + c] = uint16_t(pk & 0xffffu);
      }
    }
    return;
  }
  // src is [cols, rows] row-major; read coalesced along its rows-dimension, write coalesced along cols.
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int sc = c0 + i, sr = r0 + threadIdx.x;  // src element (sc, sr)
    tile[i][threadIdx.x] = (sc < cols && sr < rows) ? src[int64_t(sc) * rows + sr] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < rows && c < cols) {
      uint32_t pk = E16::pack(tile[threadIdx.x][i], 0.f);
      reinterpret_cast<uint16_t*>(dst)[int64_t(r) * cols + c] = uint16_t(pk & 0xffffu);
    }
  }
}

// hi = round16(x), lo = round16(x - hi): two-term split used once at create time for the zk label tables.
template <class E16>
__global__ void split16_kernel(const float* __restrict__ x, typename E16::T* __restrict__ hi,
                               typename E16::T* __restrict__ lo, int64_t n) {
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) {
    const float v = x[i];
    const uint32_t ph = E16::pack(v, 0.f);
    const float h = E16::unpack(ph).x;
    const uint32_t pl = E16::pack(v - h, 0.f);
    reinterpret_cast<uint16_t*>(hi)[i] = uint16_t(ph & 0xffffu);
    reinterpret_cast<uint16_t*>(lo)[i] = uint16_t(pl & 0xffffu);
  }
}

static mmr_status pack16(const float* src_dev, void* dst16, int rows, int cols, bool transpose, int dtype,
                         cudaStream_t st) {
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  if (dtype == MMR_DT_BF16)
    pack16_kernel<BF16><<<grid, block, 0, st>>>(src_dev, static_cast<BF16::T*>(dst16), rows, cols, transpose);
  else
    pack16_kernel<FP16><<<grid, block, 0, st>>>(src_dev, static_cast<FP16::T*>(dst16), rows, cols, transpose);
  MMR_CUDA_OK(cudaGetLastError());
  return MMR_OK;
}

static mmr_status split16(const float* x, void* hi, void* lo, int64_t n, int dtype, cudaStream_t st) {
  const int grid = int(std::min<int64_t>((n + 255) / 256, 148 * 8));
  if (dtype == MMR_DT_BF16)
    split16_kernel<BF16><<<grid, 256, 0, st>>>(x, static_cast<BF16::T*>(hi), static_cast<BF16::T*>(lo), n);
  else
    split16_kernel<FP16><<<grid, 256, 0, st>>>(x, static_cast<FP16::T*>(hi), static_cast<FP16::T*>(lo), n);
  MMR_CUDA_OK(cudaGetLastError());
  return MMR_OK;
}

// ------------------------------------------------------------------------------------------ handle
struct Linear {
  void* w16 = nullptr;   // [n, kw] 16-bit: kw = k (fast) or 3 k = [hi | hi | lo] (strict precision, strict.cu)
  float* bias = nullptr; // [n] or null
  int n = 0, k = 0, kw = 0;
};
struct LNp {
  float* gamma = nullptr;
  float* beta = nullptr;
};
struct AttBlock {  // attention + output projection + LayerNorm (pixelbert.py:658-852, 960-966; modeling.py:369-391)
  Linear qkv, out;
  LNp ln;
};
struct FfnBlock {  // pixelbert.py:969-983; modeling.py:394-420
  Linear in, out;
  LNp ln;
};
struct Layer {
  AttBlock att;
  FfnBlock ffn;
};
struct XLayer {  // modeling.py:444-493
  AttBlock cross, lang_self, visn_self;
  FfnBlock lang_ffn, visn_ffn;
};

struct Arena {
  uint8_t* base = nullptr;
  size_t cap = 0, used = 0;
  void* take(size_t bytes) {
    const size_t off = (used + 255) & ~size_t(255);
    if (off + bytes > cap) return nullptr;
    used = off + bytes;
    return base + off;
  }
};

}  // namespace mmr

struct mmr_handle {
  mmr_config cfg{};
  int device = 0;
  int act = MMR_ACT_GELU_TANH;
  mmr::Arena weights, work;
  // embeddings (fp32)
  float *E = nullptr, *T = nullptr, *P = nullptr;
  mmr::LNp emb_ln;
  std::vector<mmr::Layer> layers, r_layers;
  std::vector<mmr::XLayer> x_layers;
  mmr::Linear pooler;
  // zk
  mmr::Linear conv2, featureemb;
  float *tables = nullptr, *bc1 = nullptr, *Wb = nullptr, *bb = nullptr, *am_wn = nullptr;
  // lds
  mmr::Linear lds_feat;
  float *wl = nullptr, *cls_w = nullptr, *cls_b = nullptr;
  // lxmert
  mmr::Linear visn_fc, label_fc, logit0;
  mmr::LNp visn_ln, box_ln, label_ln, logit_ln;
  float *box_w = nullptr, *box_b = nullptr, *wconv = nullptr, *bconv = nullptr, *logit3_w = nullptr,
        *logit3_b = nullptr;
  // workspace
  void *f16 = nullptr, *t16 = nullptr, *x16 = nullptr, *qkv16 = nullptr, *ctx16 = nullptr, *h16 = nullptr,
       *pooled16 = nullptr;
  float *tmp32 = nullptr, *x32 = nullptr, *pooled32 = nullptr, *head32 = nullptr, *emb_tap = nullptr;
  // zk: label term once per distinct phrase (embed.cu): phrase table {epoch, row}, representative of every box, terms
  unsigned long long* lab_tab = nullptr;
  uint32_t lab_tab_mask = 0;
  uint32_t* lab_epoch_dev = nullptr;   // device-resident: moved on by the forward itself (CUDA-graph replay safe)
  int32_t* lab_rep = nullptr;
  float* lab_term32 = nullptr;
  float* layer_tap = nullptr;   // [n_layers, rows_max, hidden], allocated by mmr_set_debug_taps(h, 2)
  int32_t* key_mask = nullptr;
  // lxmert: compact language stream, one row group per distinct query (mmr_inputs.lang_unique), and its key mask
  void* xl16c = nullptr;
  float* xl32c = nullptr;
  int32_t* lang_mask_c = nullptr;
  // [CLS]-row tail of the last block (MMR_TUNE_PRUNE_LAST): compact [B, .] buffers
  void *xc16 = nullptr, *ctxc16 = nullptr, *hc16 = nullptr;
  float* xc32 = nullptr;
  // fused GEMM+LayerNorm row-statistics exchange table: this handle's own (gemm_ln_sm100.cu)
  mmr::LnTable ln_table;
  // strict precision (strict.cu): fp32 activations + ONE split-operand scratch reused by every GEMM
  bool strict = false;
  void* a16s = nullptr;          // [rows, 3 K] split A operand of the GEMM being launched
  float *qkv32 = nullptr, *ctx32 = nullptr, *h32 = nullptr;
  cudaEvent_t done_ev = nullptr; // recorded after every eager forward: cross-stream serialisation on one device
  int64_t rows_max = 0;   // encoder rows at max_batch
  int last_B = 0;
  int pruned_last = 0;   // the last forward computed its final block for the [CLS] rows only
  int launches = 0;
  int keep_taps = 0;
  // per-launch event profiling (mmr_set_profiling / mmr_get_profile)
  int prof_on = 0, prof_n = 0;
  std::vector<cudaEvent_t> prof_ev;
  std::vector<int> prof_kind;
  std::vector<double> prof_flops;
};

namespace mmr {

// Makes `device` current for the lifetime of the object and restores the caller's device afterwards: no entry point of
// this library leaves the calling thread on another device than it came with.
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != device) switched = cudaSetDevice(device) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};

// At most one forward in flight per device (the fused GEMM+LayerNorm kernel needs its whole grid co-resident): the
// last eager forward of every device leaves an event; a forward arriving on ANOTHER stream waits for it on the device.
struct DeviceTail {
  cudaEvent_t ev = nullptr;
  cudaStream_t stream = nullptr;
  const mmr_handle* owner = nullptr;
};
static DeviceTail g_tail[64];
static std::mutex g_tail_mu;   // host threads driving different handles of one device

__global__ void transpose32_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols) {
  // dst [cols, rows] = src [rows, cols]^T
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? src[int64_t(r) * cols + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) dst[int64_t(c) * rows + r] = tile[threadIdx.x][i];
  }
}
static mmr_status transpose32(const float* src, float* dst, int rows, int cols, cudaStream_t st) {
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  transpose32_kernel<<<grid, block, 0, st>>>(src, dst, rows, cols);
  MMR_CUDA_OK(cudaGetLastError());
  return MMR_OK;
}

using TensorMap = std::map<std::string, const mmr_tensor*>;

static int64_t numel(const mmr_tensor* t) {
  int64_t n = 1;
  for (int i = 0; i < t->ndim; ++i) n *= t->dims[i];
  return n;
}

struct Packer {
  mmr_handle* h;
  const TensorMap& tm;
  float* staging;       // device fp32 staging for matrices
  size_t staging_floats;
  cudaStream_t st;
  float* staging_t = nullptr;   // strict precision only: the fp32 [n, k] matrix after the TF-layout transpose

  mmr_status find(const std::string& name, int64_t want_numel, const mmr_tensor** out) const {
    auto it = tm.find(name);
    if (it == tm.end()) return fail(MMR_ERR_WEIGHTS, "weight '%s' is missing", name.c_str());
    if (numel(it->second) != want_numel)
      return fail(MMR_ERR_WEIGHTS, "weight '%s' has %lld elements, expected %lld", name.c_str(),
                  (long long)numel(it->second), (long long)want_numel);
    if (it->second->data == nullptr) return fail(MMR_ERR_WEIGHTS, "weight '%s' has a null data pointer", name.c_str());
    *out = it->second;
    return MMR_OK;
  }
  // fp32 vector / table copied verbatim
  mmr_status f32(const std::string& name, int64_t n, float** out) {
    const mmr_tensor* t;
    MMR_TRY(find(name, n, &t));
    float* d = static_cast<float*>(h->weights.take(size_t(n) * 4));
    if (!d) return fail(MMR_ERR_NOMEM, "weight arena exhausted at '%s'", name.c_str());
    MMR_CUDA_OK(cudaMemcpyAsync(d, t->data, size_t(n) * 4, cudaMemcpyHostToDevice, st));
    *out = d;
    return MMR_OK;
  }
  // 16-bit [n,k] matrix into dst16 (row offset already applied).  tf_layout: source is [k,n] ("kernel").
  mmr_status mat_into(const std::string& name, int n, int k, bool tf_layout, void* dst16) {
    const mmr_tensor* t;
    MMR_TRY(find(name, int64_t(n) * k, &t));
    if (size_t(n) * k > staging_floats) return fail(MMR_ERR_INVALID, "staging too small for '%s'", name.c_str());
    MMR_CUDA_OK(cudaMemcpyAsync(staging, t->data, size_t(n) * k * 4, cudaMemcpyHostToDevice, st));
    if (h->strict) {
      // [n, 3k] = [hi | hi | lo] of the fp32 [n, k] matrix (the transpose of a TF kernel happens in fp32 first)
      const float* src = staging;
      if (tf_layout) {
        MMR_TRY(transpose32(staging, staging_t, k, n, st));
        src = staging_t;
      }
      return split3(src, k, n, k, dst16, 3 * int64_t(k), MMR_ACT_NONE, 1, h->cfg.dtype, st);
    }
    MMR_TRY(pack16(staging, dst16, n, k, tf_layout, h->cfg.dtype, st));
    // staging is reused by the next call: the pageable H2D copy above is synchronous w.r.t. the host buffer but
    // the device-side order on `st` already serialises copy -> pack -> next copy.
    return MMR_OK;
  }
  int kw_of(int k) const { return h->strict ? 3 * k : k; }
  mmr_status linear(const std::string& wname, const std::string& bname, int n, int k, bool tf_layout, Linear* L) {
    L->n = n;
    L->k = k;
    L->kw = kw_of(k);
    L->w16 = h->weights.take(size_t(n) * L->kw * 2);
    if (!L->w16) return fail(MMR_ERR_NOMEM, "weight arena exhausted at '%s'", wname.c_str());
    MMR_TRY(mat_into(wname, n, k, tf_layout, L->w16));
    if (!bname.empty()) MMR_TRY(f32(bname, n, &L->bias));
    return MMR_OK;
  }
  // fused [3H, H] QKV projection from three reference tensors
  mmr_status qkv(const std::string& prefix, const char* wsuffix, const char* bsuffix, bool tf_layout, Linear* L) {
    const int H = h->cfg.hidden;
    L->n = 3 * H;
    L->k = H;
    L->kw = kw_of(H);
    L->w16 = h->weights.take(size_t(3) * H * L->kw * 2);
    L->bias = static_cast<float*>(h->weights.take(size_t(3) * H * 4));
    if (!L->w16 || !L->bias) return fail(MMR_ERR_NOMEM, "weight arena exhausted at '%s'", prefix.c_str());
    const char* names[3] = {"query", "key", "value"};
    for (int i = 0; i < 3; ++i) {
      MMR_TRY(mat_into(prefix + names[i] + wsuffix, H, H, tf_layout,
                       static_cast<uint8_t*>(L->w16) + size_t(i) * H * L->kw * 2));
      const mmr_tensor* t;
      MMR_TRY(find(prefix + names[i] + bsuffix, H, &t));
      MMR_CUDA_OK(cudaMemcpyAsync(L->bias + i * H, t->data, size_t(H) * 4, cudaMemcpyHostToDevice, st));
    }
    return MMR_OK;
  }
  mmr_status ln(const std::string& gname, const std::string& bname, int n, LNp* p) {
    MMR_TRY(f32(gname, n, &p->gamma));
    MMR_TRY(f32(bname, n, &p->beta));
    return MMR_OK;
  }
};

static mmr_status pack_tf_layer(Packer& pk, const std::string& p, Layer* L) {
  const int H = pk.h->cfg.hidden, I = pk.h->cfg.intermediate;
  MMR_TRY(pk.qkv(p + "attention/self/", "/kernel", "/bias", true, &L->att.qkv));
  MMR_TRY(pk.linear(p + "attention/output/dense/kernel", p + "attention/output/dense/bias", H, H, true, &L->att.out));
  MMR_TRY(pk.ln(p + "attention/output/LayerNorm/gamma", p + "attention/output/LayerNorm/beta", H, &L->att.ln));
  MMR_TRY(pk.linear(p + "intermediate/dense/kernel", p + "intermediate/dense/bias", I, H, true, &L->ffn.in));
  MMR_TRY(pk.linear(p + "output/dense/kernel", p + "output/dense/bias", H, I, true, &L->ffn.out));
  MMR_TRY(pk.ln(p + "output/LayerNorm/gamma", p + "output/LayerNorm/beta", H, &L->ffn.ln));
  return MMR_OK;
}

static mmr_status pack_torch_att(Packer& pk, const std::string& att, const std::string& out, AttBlock* A) {
  const int H = pk.h->cfg.hidden;
  MMR_TRY(pk.qkv(att, ".weight", ".bias", false, &A->qkv));
  MMR_TRY(pk.linear(out + "dense.weight", out + "dense.bias", H, H, false, &A->out));
  MMR_TRY(pk.ln(out + "LayerNorm.weight", out + "LayerNorm.bias", H, &A->ln));
  return MMR_OK;
}
static mmr_status pack_torch_ffn(Packer& pk, const std::string& in, const std::string& out, FfnBlock* F) {
  const int H = pk.h->cfg.hidden, I = pk.h->cfg.intermediate;
  MMR_TRY(pk.linear(in + "dense.weight", in + "dense.bias", I, H, false, &F->in));
  MMR_TRY(pk.linear(out + "dense.weight", out + "dense.bias", H, I, false, &F->out));
  MMR_TRY(pk.ln(out + "LayerNorm.weight", out + "LayerNorm.bias", H, &F->ln));
  return MMR_OK;
}

static size_t weight_arena_bytes(const mmr_config& c) {
  const size_t H = c.hidden, I = c.intermediate, V = c.vocab;
  const size_t e16 = c.precision == MMR_PRECISION_STRICT ? 6 : 2;   // bytes per stored matrix element
  const size_t att = 4 * H * H * e16 + 4 * H * 4 + 2 * H * 4 + 8 * 256;
  const size_t ffn = 2 * H * I * e16 + (H + I) * 4 + 2 * H * 4 + 8 * 256;
  size_t n = (V + c.max_pos + c.type_vocab + 2) * H * 4 + 16 * 256;
  n += size_t(c.n_layers + c.n_r_layers) * (att + ffn);
  n += size_t(c.n_x_layers) * (3 * att + 2 * ffn);
  n += H * H * e16 + H * 4;  // pooler
  n += size_t(c.feat_dim) * H * e16 + 4 * H * H * e16 + 64 * H * 4 + (1 << 20);  // projections, heads, small stuff
  if (c.model_kind == MMR_MODEL_IMAGEBERT_ZK) n += size_t(c.label_len) * V * H * 4;  // label tables
  return n;
}

// T_k = E . Wc1[k] in near-fp32 precision through three 16-bit GEMMs on split operands (create time only).
static mmr_status build_zk_tables(Packer& pk) {
  mmr_handle* h = pk.h;
  const int H = h->cfg.hidden, V = h->cfg.vocab, Tn = h->cfg.label_len;
  const int dt = h->cfg.dtype;
  const mmr_tensor* wt;
  MMR_TRY(pk.find("kdd_conv1/weights", int64_t(Tn) * H * H, &wt));
  h->tables = static_cast<float*>(h->weights.take(size_t(Tn) * V * H * 4));
  if (!h->tables) return fail(MMR_ERR_NOMEM, "weight arena exhausted at the zk label tables");
  // scratch: E hi/lo [V,H], W hi/lo [H,H] 16-bit, W^T fp32 [H,H]
  void *Eh, *El, *Wh, *Wl;
  float* WT;
  MMR_CUDA_OK(cudaMalloc(&Eh, size_t(V) * H * 2));
  MMR_CUDA_OK(cudaMalloc(&El, size_t(V) * H * 2));
  MMR_CUDA_OK(cudaMalloc(&Wh, size_t(H) * H * 2));
  MMR_CUDA_OK(cudaMalloc(&Wl, size_t(H) * H * 2));
  MMR_CUDA_OK(cudaMalloc(&WT, size_t(H) * H * 4));
  mmr_status rc = split16(h->E, Eh, El, int64_t(V) * H, dt, pk.st);
  for (int k = 0; k < Tn && rc == MMR_OK; ++k) {
    // Wc1[0,k] is [Hin,Hout]; the GEMM wants [Hout,Hin].  Transpose in fp32 so the split sees unrounded values.
    if (cudaMemcpyAsync(pk.staging, wt->data + size_t(k) * H * H, size_t(H) * H * 4, cudaMemcpyHostToDevice,
                        pk.st) != cudaSuccess) {
      rc = fail(MMR_ERR_CUDA, "upload of kdd_conv1 tap %d failed", k);
      break;
    }
    rc = transpose32(pk.staging, WT, H, H, pk.st);
    if (rc != MMR_OK) break;
    rc = split16(WT, Wh, Wl, int64_t(H) * H, dt, pk.st);
    if (rc != MMR_OK) break;
    float* out = h->tables + size_t(k) * V * H;
    rc = gemm(Eh, H, Wh, H, V, H, H, nullptr, nullptr, 0, nullptr, 0, out, H, MMR_ACT_NONE, dt, pk.st);
    if (rc != MMR_OK) break;
    rc = gemm(Eh, H, Wl, H, V, H, H, nullptr, out, H, nullptr, 0, out, H, MMR_ACT_NONE, dt, pk.st);
    if (rc != MMR_OK) break;
    rc = gemm(El, H, Wh, H, V, H, H, nullptr, out, H, nullptr, 0, out, H, MMR_ACT_NONE, dt, pk.st);
  }
  cudaStreamSynchronize(pk.st);
  cudaFree(Eh); cudaFree(El); cudaFree(Wh); cudaFree(Wl); cudaFree(WT);
  return rc;
}

static mmr_status pack_weights(mmr_handle* h, const TensorMap& tm, cudaStream_t st) {
  const mmr_config& c = h->cfg;
  const int H = c.hidden, F = c.feat_dim;
  const size_t staging_floats = std::max<size_t>(size_t(H) * c.intermediate, size_t(F) * H);
  float* staging = nullptr;
  MMR_CUDA_OK(cudaMalloc(&staging, staging_floats * 4));
  Packer pk{h, tm, staging, staging_floats, st};
  if (h->strict && cudaMalloc(&pk.staging_t, staging_floats * 4) != cudaSuccess) {
    cudaFree(staging);
    return fail(MMR_ERR_NOMEM, "cannot allocate the strict-precision packing scratch");
  }
  mmr_status rc = MMR_OK;
#define PK(expr)                     \
  do {                               \
    rc = (expr);                     \
    if (rc != MMR_OK) goto done;     \
  } while (0)

  if (c.model_kind == MMR_MODEL_IMAGEBERT_ZK || c.model_kind == MMR_MODEL_IMAGEBERT_LDS) {
    h->act = MMR_ACT_GELU_TANH;
    PK(pk.f32("bert/embeddings/word_embeddings", int64_t(c.vocab) * H, &h->E));
    PK(pk.f32("bert/embeddings/token_type_embeddings", int64_t(c.type_vocab) * H, &h->T));
    PK(pk.f32("bert/embeddings/position_embeddings", int64_t(c.max_pos) * H, &h->P));
    PK(pk.ln("bert/embeddings/LayerNorm/gamma", "bert/embeddings/LayerNorm/beta", H, &h->emb_ln));
    h->layers.resize(c.n_layers);
    for (int i = 0; i < c.n_layers; ++i)
      PK(pack_tf_layer(pk, "bert/encoder/layer_" + std::to_string(i) + "/", &h->layers[i]));
    PK(pk.linear("bert/pooler/dense/kernel", "bert/pooler/dense/bias", H, H, true, &h->pooler));
    if (c.model_kind == MMR_MODEL_IMAGEBERT_ZK) {
      PK(pk.linear("kdd_conv2/weights", "kdd_conv2/biases", H, F, true, &h->conv2));
      PK(pk.linear("kdd_featureemb/fully_connected/weights", "kdd_featureemb/fully_connected/biases", H, H, true,
                   &h->featureemb));
      PK(pk.f32("kdd_conv1/biases", H, &h->bc1));
      PK(pk.f32("kdd_dense1/weights", int64_t(5) * H, &h->Wb));
      PK(pk.f32("kdd_dense1/biases", H, &h->bb));
      {
        // column-normalised AM-softmax kernel, stored [2,768] (model_triple.py:64: l2_normalize(kernel, 0, 1e-10))
        const mmr_tensor* t;
        PK(pk.find("cls/seq_relationship/am_kernel", int64_t(H) * 2, &t));
        std::vector<float> wn(size_t(2) * H);
        for (int j = 0; j < 2; ++j) {
          float ss = 0.f;
          for (int i = 0; i < H; ++i) ss += t->data[i * 2 + j] * t->data[i * 2 + j];
          const float inv = 1.0f / std::sqrt(std::max(ss, 1e-10f));
          for (int i = 0; i < H; ++i) wn[size_t(j) * H + i] = t->data[i * 2 + j] * inv;
        }
        h->am_wn = static_cast<float*>(h->weights.take(wn.size() * 4));
        if (!h->am_wn) { rc = fail(MMR_ERR_NOMEM, "weight arena exhausted at am_kernel"); goto done; }
        if (cudaMemcpyAsync(h->am_wn, wn.data(), wn.size() * 4, cudaMemcpyHostToDevice, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess) {
          rc = fail(MMR_ERR_CUDA, "upload of am_kernel failed");
          goto done;
        }
      }
      PK(build_zk_tables(pk));
    } else {
      PK(pk.linear("featureemb/fully_connected/weights", "featureemb/fully_connected/biases", H, F, true,
                   &h->lds_feat));
      PK(pk.f32("bert/embeddings/word_embeddings_labelembedding", c.label_len, &h->wl));
      PK(pk.f32("cls/seq_relationship/output_weights", int64_t(2) * H, &h->cls_w));
      PK(pk.f32("cls/seq_relationship/output_bias", 2, &h->cls_b));
    }
  } else {
    h->act = MMR_ACT_GELU_ERF;
    const std::string b = "lxrt_encoder.model.bert.";
    PK(pk.f32(b + "embeddings.word_embeddings.weight", int64_t(c.vocab) * H, &h->E));
    PK(pk.f32(b + "embeddings.token_type_embeddings.weight", int64_t(c.type_vocab) * H, &h->T));
    PK(pk.f32(b + "embeddings.position_embeddings.weight", int64_t(c.max_pos) * H, &h->P));
    PK(pk.ln(b + "embeddings.LayerNorm.weight", b + "embeddings.LayerNorm.bias", H, &h->emb_ln));
    const std::string v = b + "encoder.visn_fc.";
    PK(pk.linear(v + "visn_fc.weight", v + "visn_fc.bias", H, F, false, &h->visn_fc));
    PK(pk.ln(v + "visn_layer_norm.weight", v + "visn_layer_norm.bias", H, &h->visn_ln));
    PK(pk.f32(v + "box_fc.weight", int64_t(H) * 4, &h->box_w));
    PK(pk.f32(v + "box_fc.bias", H, &h->box_b));
    PK(pk.ln(v + "box_layer_norm.weight", v + "box_layer_norm.bias", H, &h->box_ln));
    PK(pk.f32(v + "label_conv.weight", c.label_len, &h->wconv));
    PK(pk.f32(v + "label_conv.bias", 1, &h->bconv));
    PK(pk.linear(v + "label_fc.weight", v + "label_fc.bias", H, H, false, &h->label_fc));
    PK(pk.ln(v + "label_layer_norm.weight", v + "label_layer_norm.bias", H, &h->label_ln));
    h->layers.resize(c.n_layers);
    h->r_layers.resize(c.n_r_layers);
    h->x_layers.resize(c.n_x_layers);
    for (int i = 0; i < c.n_layers; ++i) {
      const std::string p = b + "encoder.layer." + std::to_string(i) + ".";
      PK(pack_torch_att(pk, p + "attention.self.", p + "attention.output.", &h->layers[i].att));
      PK(pack_torch_ffn(pk, p + "intermediate.", p + "output.", &h->layers[i].ffn));
    }
    for (int i = 0; i < c.n_r_layers; ++i) {
      const std::string p = b + "encoder.r_layers." + std::to_string(i) + ".";
      PK(pack_torch_att(pk, p + "attention.self.", p + "attention.output.", &h->r_layers[i].att));
      PK(pack_torch_ffn(pk, p + "intermediate.", p + "output.", &h->r_layers[i].ffn));
    }
    for (int i = 0; i < c.n_x_layers; ++i) {
      const std::string p = b + "encoder.x_layers." + std::to_string(i) + ".";
      XLayer& X = h->x_layers[i];
      PK(pack_torch_att(pk, p + "visual_attention.att.", p + "visual_attention.output.", &X.cross));
      PK(pack_torch_att(pk, p + "lang_self_att.self.", p + "lang_self_att.output.", &X.lang_self));
      PK(pack_torch_att(pk, p + "visn_self_att.self.", p + "visn_self_att.output.", &X.visn_self));
      PK(pack_torch_ffn(pk, p + "lang_inter.", p + "lang_output.", &X.lang_ffn));
      PK(pack_torch_ffn(pk, p + "visn_inter.", p + "visn_output.", &X.visn_ffn));
    }
    PK(pk.linear(b + "pooler.dense.weight", b + "pooler.dense.bias", H, H, false, &h->pooler));
    PK(pk.linear("logit_fc.0.weight", "logit_fc.0.bias", 2 * H, H, false, &h->logit0));
    PK(pk.ln("logit_fc.2.weight", "logit_fc.2.bias", 2 * H, &h->logit_ln));
    PK(pk.f32("logit_fc.3.weight", int64_t(2) * 2 * H, &h->logit3_w));
    PK(pk.f32("logit_fc.3.bias", 2, &h->logit3_b));
  }
#undef PK
done:
  cudaStreamSynchronize(st);
  cudaFree(staging);
  if (pk.staging_t) cudaFree(pk.staging_t);
  return rc;
}

static mmr_status alloc_workspace(mmr_handle* h, size_t* size_only = nullptr) {
  const mmr_config& c = h->cfg;
  const int64_t H = c.hidden, I = c.intermediate, B = c.max_batch, R = c.nbox;
  int64_t S = c.lq + c.nbox;
  if (c.model_kind == MMR_MODEL_IMAGEBERT_LDS) S = c.lq + 2 * c.nbox;
  const int64_t M = B * S;
  h->rows_max = M;
  const bool strict = h->strict;
  const bool zk = c.model_kind == MMR_MODEL_IMAGEBERT_ZK;
  uint32_t lab_slots = 0;
  if (zk) {
    lab_slots = 1024;
    while (lab_slots < 4 * uint64_t(B * R)) lab_slots <<= 1;   // <= 25 % load
  }
  const size_t ln_bytes = gemm_ln_table_bytes(int(M));
  // split A operand of the widest strict GEMM: FFN-out over all rows, or the 2048-d region projection
  const size_t a16s_elems = size_t(std::max<int64_t>(M * 3 * I, B * R * 3 * int64_t(c.feat_dim)));
  // One pass sizes the arena, a second hands out the pointers: the same list, so they cannot disagree.
  for (int pass = 0; pass < 2; ++pass) {
    size_t bytes = 0;
    auto take = [&](size_t n) -> void* {
      if (pass == 0) {
        bytes += ((n + 255) & ~size_t(255));
        return nullptr;
      }
      return h->work.take(n);
    };
    if (!strict) {
      h->f16 = take(B * R * c.feat_dim * 2);
      h->t16 = take(B * R * H * 2);
      h->qkv16 = take(M * 3 * H * 2);
      h->ctx16 = take(M * H * 2);
      h->h16 = take(M * I * 2);
      h->pooled16 = take(B * H * 2);
      h->ctxc16 = take(B * H * 2);
      h->hc16 = take(B * I * 2);
    } else {
      h->a16s = take(a16s_elems * 2);
      h->qkv32 = static_cast<float*>(take(M * 3 * H * 4));
      h->ctx32 = static_cast<float*>(take(M * H * 4));
      h->h32 = static_cast<float*>(take(M * I * 4));
    }
    h->tmp32 = static_cast<float*>(take(B * R * H * 4));
    h->x16 = take(M * H * 2);
    h->x32 = static_cast<float*>(take(M * H * 4));
    h->key_mask = static_cast<int32_t*>(take(M * 4));
    h->pooled32 = static_cast<float*>(take(B * H * 4));
    h->head32 = static_cast<float*>(take(B * 2 * H * 4));
    h->emb_tap = static_cast<float*>(take(M * H * 4));
    h->xc16 = take(B * H * 2);
    h->xc32 = static_cast<float*>(take(B * H * 4));
    if (c.model_kind == MMR_MODEL_LXMERT && !strict) {
      h->xl16c = take(B * c.lq * H * 2);
      h->xl32c = static_cast<float*>(take(B * c.lq * H * 4));
      h->lang_mask_c = static_cast<int32_t*>(take(B * c.lq * 4));
    }
    void* ln_mem = take(ln_bytes);
    if (zk) {
      h->lab_tab = static_cast<unsigned long long*>(take(size_t(lab_slots) * 8));
      h->lab_rep = static_cast<int32_t*>(take(B * R * 4));
      h->lab_term32 = static_cast<float*>(take(B * R * H * 4));
      h->lab_epoch_dev = static_cast<uint32_t*>(take(256));
    }
    if (pass == 0) {
      bytes += 4096;
      if (size_only != nullptr) {      // mmr_workspace_bytes: the plan without the allocation
        *size_only = bytes;
        return MMR_OK;
      }
      MMR_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&h->work.base), bytes));
      h->work.cap = bytes;
      continue;
    }
    if (!h->xc32 || !ln_mem || (zk && !h->lab_epoch_dev)) return fail(MMR_ERR_NOMEM, "workspace arena mis-sized");
    MMR_TRY(gemm_ln_table_init(ln_mem, int(M), &h->ln_table, nullptr));
  }
  if (zk) {
    h->lab_tab_mask = lab_slots - 1;
    MMR_CUDA_OK(cudaMemset(h->lab_tab, 0, size_t(lab_slots) * 8));   // epoch 0 = never written
    const uint32_t first_epoch = 1;
    MMR_CUDA_OK(cudaMemcpy(h->lab_epoch_dev, &first_epoch, sizeof(first_epoch), cudaMemcpyHostToDevice));
  }
  MMR_CUDA_OK(cudaEventCreateWithFlags(&h->done_ev, cudaEventDisableTiming));
  MMR_CUDA_OK(cudaDeviceSynchronize());
  return MMR_OK;
}

// ------------------------------------------------------------------------------------------ forward pieces
enum LaunchKind { K_GEMM = 0, K_ATTENTION = 1, K_LAYERNORM = 2, K_ROW = 3 };

// The last block only for the rows the pooler reads (include/mmrecall.h, MMR_TUNE_PRUNE_LAST); taps want every row.
static bool prune_last(const mmr_handle* h) {
  if (tuning(MMR_TUNE_PRUNE_LAST) == 0 || h->keep_taps != 0) return false;
  return h->cfg.model_kind == MMR_MODEL_LXMERT ? !h->x_layers.empty() : !h->layers.empty();
}

struct Ctx {
  mmr_handle* h;
  cudaStream_t st;
  int dt;
  int H;
  int launches = 0;
  // the residual stream the block helpers below work on: the handle's x16 / x32, or (LXMERT, language blocks once per
  // distinct query) the compact language stream
  void* x16_base = nullptr;
  float* x32_base = nullptr;
  uint8_t* x16(int64_t row) const { return static_cast<uint8_t*>(x16_base ? x16_base : h->x16) + row * H * 2; }
  float* x32(int64_t row) const { return (x32_base ? x32_base : h->x32) + row * H; }
  uint8_t* qkv(int64_t row, int part) const {
    return static_cast<uint8_t*>(h->qkv16) + (row * 3 * H + int64_t(part) * H) * 2;
  }
  uint8_t* ctx(int64_t row) const { return static_cast<uint8_t*>(h->ctx16) + row * H * 2; }

  // Bookkeeping after every kernel launch: the launch count behind mmr_launches_per_forward and, when profiling
  // is on, one event per launch so bench.py can attribute device time to kernels inside the timed step.
  mmr_status mark(int kind, double flops) {
    ++launches;
    if (h->prof_on && h->prof_n < int(h->prof_kind.size())) {
      MMR_CUDA_OK(cudaEventRecord(h->prof_ev[h->prof_n + 1], st));
      h->prof_kind[h->prof_n] = kind;
      h->prof_flops[h->prof_n] = flops;
      ++h->prof_n;
    }
    return MMR_OK;
  }
  mmr_status G(const void* A, int64_t lda, const Linear& W, int M, const float* residual, void* out16, int64_t ldo16,
               float* out32, int act) {
    MMR_TRY(gemm(A, lda, W.w16, W.k, M, W.n, W.k, W.bias, residual, H, out16, ldo16, out32, H, act, dt, st));
    return mark(K_GEMM, 2.0 * M * W.n * W.k);
  }
  // out = LN(A . W^T + bias + residual), fp32 rows + 16-bit mirror: one fused kernel when the shape allows, else two.
  // residual (row stride ldr) may alias out32 (row stride H) — the residual stream updated in place — or be a strided
  // view of it (the [CLS] rows feeding the compact tail buffers).
  mmr_status G_LN_ex(const void* A, int64_t lda, const Linear& W, const LNp& ln, const float* residual, int64_t ldr,
                     void* out16, float* out32, int rows) {
    if (gemm_ln_eligible(rows, W.n, W.k, dt)) {
      MMR_TRY(gemm_ln(A, lda, W.w16, W.k, rows, W.k, W.bias, residual, ldr, ln.gamma, ln.beta, 1e-12f, out16, H, out32, H,
                      dt, h->ln_table, st));
      return mark(K_GEMM, 2.0 * rows * W.n * W.k);
    }
    MMR_TRY(gemm(A, lda, W.w16, W.k, rows, W.n, W.k, W.bias, residual, ldr, nullptr, 0, out32, H, MMR_ACT_NONE, dt, st));
    MMR_TRY(mark(K_GEMM, 2.0 * rows * W.n * W.k));
    return LN(out32, ln, rows, out16, out32, 1.0f, 0);
  }
  mmr_status G_LN(const void* A, int64_t lda, const Linear& W, const LNp& ln, int64_t row0, int rows) {
    return G_LN_ex(A, lda, W, ln, x32(row0), H, x16(row0), x32(row0), rows);
  }
  // Two streams that share the activation buffers but not the weights (LXMERT: language rows [0, split), visual rows
  // [split, M)) run a projection as ONE launch when the split is a multiple of the 256-row tile: full waves instead
  // of two launches of 32 and 36 row blocks (the fused GEMM+LN kernel loses a third of its SMs at 32 blocks).
  bool mergeable(const Linear& W1, const Linear& W2, int split) const {
    return tuning(MMR_TUNE_LX_MERGE) != 0 && W1.n == W2.n && W1.k == W2.k && split > 0 && split % 256 == 0;
  }
  mmr_status G2(const void* A, int64_t lda, const Linear& W1, const Linear& W2, int M, int split, void* out16,
                int64_t ldo16, int act) {
    if (mergeable(W1, W2, split) && tuning(MMR_TUNE_GEMM_PAIR) != 0 &&
        gemm_pair16_eligible(M, W1.n, W1.k, nullptr, out16, nullptr)) {
      MMR_TRY(gemm_pair16_2w(A, lda, W1.w16, W2.w16, W1.k, M, W1.n, W1.k, W1.bias, W2.bias, split, out16, ldo16, act, dt,
                             st));
      return mark(K_GEMM, 2.0 * M * W1.n * W1.k);
    }
    MMR_TRY(G(A, lda, W1, split, nullptr, out16, ldo16, nullptr, act));
    return G(static_cast<const uint8_t*>(A) + int64_t(split) * lda * 2, lda, W2, M - split, nullptr,
             static_cast<uint8_t*>(out16) + int64_t(split) * ldo16 * 2, ldo16, nullptr, act);
  }
  mmr_status G_LN2(const void* A, int64_t lda, const Linear& W1, const Linear& W2, const LNp& ln1, const LNp& ln2,
                   int rows, int split, int64_t row0 = 0) {
    if (mergeable(W1, W2, split) && gemm_ln_eligible(rows, W1.n, W1.k, dt)) {
      MMR_TRY(gemm_ln_2w(A, lda, W1.w16, W2.w16, W1.k, rows, W1.k, W1.bias, W2.bias, x32(row0), H, ln1.gamma, ln2.gamma,
                         ln1.beta, ln2.beta, split, 1e-12f, x16(row0), H, x32(row0), H, dt, h->ln_table, st));
      return mark(K_GEMM, 2.0 * rows * W1.n * W1.k);
    }
    MMR_TRY(G_LN(A, lda, W1, ln1, row0, split));
    return G_LN(static_cast<const uint8_t*>(A) + int64_t(split) * lda * 2, lda, W2, ln2, row0 + split, rows - split);
  }
  mmr_status LN(float* x, const LNp& p, int rows, void* out16, float* out32, float scale, int accumulate) {
    MMR_TRY(layernorm(x, H, p.gamma, p.beta, 1e-12f, rows, H, out16, H, out32, H, scale, accumulate, dt, st));
    return mark(K_LAYERNORM, 0.0);
  }
};

// x16[rows] -> qkv16[rows]
static mmr_status qkv_proj(Ctx& c, const Linear& W, int64_t row0, int rows) {
  return c.G(c.x16(row0), c.H, W, rows, nullptr, c.qkv(row0, 0), 3 * c.H, nullptr, MMR_ACT_NONE);
}
// ctx16[q rows] = softmax(Q K^T / 8 + mask) V with Q from rows q0.., K/V from rows k0..
static mmr_status attend(Ctx& c, int64_t q0, int Sq, int64_t k0, int Sk, const int32_t* key_mask, int B) {
  MMR_TRY(attention(c.qkv(q0, 0), 3 * c.H, c.qkv(k0, 1), 3 * c.H, c.qkv(k0, 2), 3 * c.H, key_mask, c.ctx(q0), c.H, B,
                    Sq, Sk, c.h->cfg.heads, c.dt, c.st));
  return c.mark(K_ATTENTION, 4.0 * B * Sq * Sk * c.H);
}
// Two attention problems of one batch in ONE launch (LXMERT: both streams' self-attention, or both directions of the
// shared cross-attention): saves a launch, a prologue and a tail of a kernel that is latency-bound at these shapes.
static mmr_status attend2(Ctx& c, int64_t q0, int Sq0, int64_t k0, int Sk0, const int32_t* mask0, int64_t q1, int Sq1,
                          int64_t k1, int Sk1, const int32_t* mask1, int B, int B0 = 0) {
  if (B0 <= 0) B0 = B;   // pairs of the FIRST problem (LXMERT with the language blocks once per query: fewer than B)
  static const bool env_off = [] { const char* e = getenv("MMR_LX_ATTN_PAIR"); return e != nullptr && atoi(e) == 0; }();
  bool separate = tuning(MMR_TUNE_LX_MERGE) == 0 || env_off;   // (environment: A/B measurements only)
#ifdef MMR_EXPERIMENTAL
  separate = separate || tuning(MMR_TUNE_ATTN_TC) != 2;   // the older attention kernels take one problem per launch
#endif
  if (separate) {
    MMR_TRY(attend(c, q0, Sq0, k0, Sk0, mask0, B0));
    return attend(c, q1, Sq1, k1, Sk1, mask1, B);
  }
  const int64_t ld = 3 * int64_t(c.H);
  const AttentionArgs a{c.qkv(q0, 0), c.qkv(k0, 1), c.qkv(k0, 2), ld, ld, ld, mask0, c.ctx(q0), c.H, B0, Sq0, Sk0};
  const AttentionArgs b{c.qkv(q1, 0), c.qkv(k1, 1), c.qkv(k1, 2), ld, ld, ld, mask1, c.ctx(q1), c.H, B, Sq1, Sk1};
  MMR_TRY(attention_pair(a, b, c.h->cfg.heads, c.dt, c.st));
  return c.mark(K_ATTENTION, 4.0 * (double(B0) * Sq0 * Sk0 + double(B) * Sq1 * Sk1) * c.H);
}
// x = LN(ctx . Wo^T + bo + x), in place on the residual stream
static mmr_status out_proj_ln(Ctx& c, const AttBlock& A, int64_t row0, int rows) {
  return c.G_LN(c.ctx(row0), c.H, A.out, A.ln, row0, rows);
}
// x = LN(act(x W1^T + b1) W2^T + b2 + x)
static mmr_status ffn_block(Ctx& c, const FfnBlock& F, int64_t row0, int rows) {
  uint8_t* hbuf = static_cast<uint8_t*>(c.h->h16) + row0 * F.in.n * 2;
  MMR_TRY(c.G(c.x16(row0), c.H, F.in, rows, nullptr, hbuf, F.in.n, nullptr, c.h->act));
  return c.G_LN(hbuf, F.in.n, F.out, F.ln, row0, rows);
}
static mmr_status self_att_block(Ctx& c, const AttBlock& A, int64_t row0, int B, int S, const int32_t* key_mask) {
  MMR_TRY(qkv_proj(c, A.qkv, row0, B * S));
  MMR_TRY(attend(c, row0, S, row0, S, key_mask, B));
  return out_proj_ln(c, A, row0, B * S);
}
static mmr_status bert_layer(Ctx& c, const Layer& L, int64_t row0, int B, int S, const int32_t* key_mask) {
  MMR_TRY(self_att_block(c, L.att, row0, B, S, key_mask));
  return ffn_block(c, L.ffn, row0, B * S);
}

// One self-attention + FFN layer of BOTH LXMERT streams (language rows [0, nl), visual rows [nl, nl + nv)): the four
// projections as merged launches, attention per stream.
static mmr_status two_stream_att(Ctx& c, const AttBlock& A1, const AttBlock& A2, int B, int Lq, int R,
                                 const int32_t* mask1, const int32_t* mask2, int64_t row0 = 0, int B1 = 0) {
  // stream 1 = B1 row groups of Lq rows from row0 (all B pairs, or one group per distinct query), stream 2 = B x R rows
  if (B1 <= 0) B1 = B;
  const int nl = B1 * Lq, nv = B * R;
  MMR_TRY(c.G2(c.x16(row0), c.H, A1.qkv, A2.qkv, nl + nv, nl, c.qkv(row0, 0), 3 * c.H, MMR_ACT_NONE));
  MMR_TRY(attend2(c, row0, Lq, row0, Lq, mask1, row0 + nl, R, row0 + nl, R, mask2, B, B1));
  return c.G_LN2(c.ctx(row0), c.H, A1.out, A2.out, A1.ln, A2.ln, nl + nv, nl, row0);
}
static mmr_status two_stream_ffn(Ctx& c, const FfnBlock& F1, const FfnBlock& F2, int nl, int nv, int64_t row0 = 0) {
  uint8_t* hbuf = static_cast<uint8_t*>(c.h->h16) + row0 * F1.in.n * 2;
  MMR_TRY(c.G2(c.x16(row0), c.H, F1.in, F2.in, nl + nv, nl, hbuf, F1.in.n, c.h->act));
  return c.G_LN2(hbuf, F1.in.n, F1.out, F2.out, F1.ln, F2.ln, nl + nv, nl, row0);
}

// first token of every pair: rows b*S of x16 (row stride S*H), pixelbert.py:258-266 / modeling.py:596-608
static mmr_status pooler(Ctx& c, int B, int S, bool compact) {
  mmr_handle* h = c.h;
  if (!compact) return c.G(h->x16, int64_t(S) * c.H, h->pooler, B, nullptr, h->pooled16, c.H, h->pooled32, MMR_ACT_TANH);
  const Linear& W = h->pooler;
  MMR_TRY(gemm(h->xc16, c.H, W.w16, W.k, B, W.n, W.k, W.bias, nullptr, 0, h->pooled16, c.H, h->pooled32, c.H, MMR_ACT_TANH,
               c.dt, c.st, true));
  return c.mark(K_GEMM, 2.0 * B * W.n * W.k);
}

// The last block for the [CLS] rows only (MMR_TUNE_PRUNE_LAST; cls_tail.cu): the stream rows [row0, row0 + B*S) keep
// their block INPUT, the B first rows leave through the compact xc16 / xc32 buffers.  Keys and values are projected
// for every row (they feed the [CLS] query), everything after the attention runs on B rows.
static mmr_status cls_tail_block(Ctx& c, const AttBlock& A, const FfnBlock& F, int64_t row0, int B, int S,
                                 const int32_t* key_mask, bool final_ln = true) {
  mmr_handle* h = c.h;
  const int H = c.H;
  MMR_TRY(qkv_proj(c, A.qkv, row0, B * S));
  MMR_TRY(cls_attention(c.qkv(row0, 0), int64_t(S) * 3 * H, c.qkv(row0, 1), c.qkv(row0, 2), 3 * H, key_mask, h->ctxc16, H, B,
                        S, h->cfg.heads, c.dt, c.st));
  MMR_TRY(c.mark(K_ATTENTION, 4.0 * B * S * H));
  // B rows only: always the single-CTA GEMM + the row LayerNorm kernel, so that the bits of a pair's score do not depend
  // on how many pairs share its batch (chunking invariance, tests/test_gpu_parity.py)
  auto tail_gemm = [&](const void* A16, int64_t lda, const Linear& W, const float* residual, int64_t ldr, void* out16,
                       int64_t ldo16, float* out32, int act) -> mmr_status {
    MMR_TRY(gemm(A16, lda, W.w16, W.k, B, W.n, W.k, W.bias, residual, ldr, out16, ldo16, out32, H, act, c.dt, c.st, true));
    return c.mark(K_GEMM, 2.0 * B * W.n * W.k);
  };
  MMR_TRY(tail_gemm(h->ctxc16, H, A.out, c.x32(row0), int64_t(S) * H, nullptr, 0, h->xc32, MMR_ACT_NONE));
  MMR_TRY(c.LN(h->xc32, A.ln, B, h->xc16, h->xc32, 1.0f, 0));
  MMR_TRY(tail_gemm(h->xc16, H, F.in, nullptr, 0, h->hc16, F.in.n, nullptr, h->act));
  MMR_TRY(tail_gemm(h->hc16, F.in.n, F.out, h->xc32, H, nullptr, 0, h->xc32, MMR_ACT_NONE));
  if (!final_ln) return MMR_OK;   // the block's last LayerNorm is fused into the pooler + head kernel that follows
  return c.LN(h->xc32, F.ln, B, h->xc16, h->xc32, 1.0f, 0);
}

static mmr_status forward_single_stream(Ctx& c, const mmr_inputs* in, int B, float* probs, float* logits) {
  mmr_handle* h = c.h;
  const mmr_config& cfg = h->cfg;
  const int H = c.H, R = cfg.nbox, Lq = cfg.lq;
  const bool zk = cfg.model_kind == MMR_MODEL_IMAGEBERT_ZK;
  const int S = zk ? Lq + R : Lq + 2 * R;
  const bool fused_in = zk && in->region_sum != nullptr;
  MMR_REQUIRE(in->query_ids && in->segment_ids && (fused_in || (in->label_ids && in->feats)),
              "mmr_forward: missing input pointer");
  if (!fused_in) {
    MMR_TRY(cast16(in->feats, h->f16, int64_t(B) * R * cfg.feat_dim, c.dt, c.st));
    MMR_TRY(c.mark(K_ROW, 0));
  }
  const int32_t* mask = nullptr;
  if (zk) {
    MMR_REQUIRE(in->len_query && in->num_boxes && in->labels && (fused_in || in->boxes),
                "mmr_forward(zk): missing input pointer");
    if (fused_in) {
      // caller already formed label + box + feat (pixelbert.BertModel's imgfeat argument)
      MMR_TRY(cast16(in->region_sum, h->t16, int64_t(B) * R * H, c.dt, c.st));
      MMR_TRY(c.mark(K_ROW, 0));
    } else {
      // feat = ReLU(f . Wc2 + bc2)  (model_triple.py:192-194)
      MMR_TRY(c.G(h->f16, cfg.feat_dim, h->conv2, B * R, nullptr, nullptr, 0, h->tmp32, MMR_ACT_RELU));
      // (the claim kernel reads the 8 ids of a box as two 16-byte words)
      if (tuning(MMR_TUNE_LABEL_DEDUP) != 0 && (reinterpret_cast<uintptr_t>(in->label_ids) & 15) == 0) {
        MMR_TRY(zk_label_terms(in->label_ids, h->tables, cfg.vocab, h->bc1, h->lab_tab, h->lab_tab_mask, h->lab_epoch_dev,
                               h->lab_rep, h->lab_term32, B * R, c.st));
        MMR_TRY(c.mark(K_ROW, 0));
        MMR_TRY(c.mark(K_ROW, 0));
        MMR_TRY(zk_region_sum_rep(h->tmp32, in->boxes, h->lab_rep, h->lab_term32, h->Wb, h->bb, h->t16, B * R,
                                  h->lab_epoch_dev, c.dt, c.st));
      } else {
        MMR_TRY(zk_region_sum(h->tmp32, in->boxes, in->label_ids, h->tables, cfg.vocab, h->bc1, h->Wb, h->bb, h->t16,
                              B * R, c.dt, c.st));
      }
      MMR_TRY(c.mark(K_ROW, 0));
    }
    // region = (label + box + feat) . Wfe + bfe  (pixelbert.py:449-452)
    MMR_TRY(c.G(h->t16, H, h->featureemb, B * R, nullptr, nullptr, 0, h->tmp32, MMR_ACT_NONE));
    MMR_TRY(zk_embed(in->query_ids, in->segment_ids, h->tmp32, in->len_query, in->num_boxes, h->E, h->T, h->P,
                     h->emb_ln.gamma, h->emb_ln.beta, Lq, R, B, h->x16, h->x32, h->key_mask, c.dt, c.st));
    MMR_TRY(c.mark(K_ROW, 0));
    mask = h->key_mask;
  } else {
    // region = f . Wf + bf, no activation (pixelmodel.py:439-442)
    MMR_TRY(c.G(h->f16, cfg.feat_dim, h->lds_feat, B * R, nullptr, nullptr, 0, h->tmp32, MMR_ACT_NONE));
    MMR_TRY(lds_embed(in->query_ids, in->segment_ids, in->label_ids, h->tmp32, h->E, h->T, h->P, h->emb_ln.gamma,
                      h->emb_ln.beta, h->wl, Lq, R, B, h->x16, h->x32, c.dt, c.st));
    MMR_TRY(c.mark(K_ROW, 0));
  }
  if (h->keep_taps)
    MMR_CUDA_OK(cudaMemcpyAsync(h->emb_tap, h->x32, size_t(B) * S * H * 4, cudaMemcpyDeviceToDevice, c.st));
  const bool prune = prune_last(h);
  for (size_t i = 0; i < h->layers.size(); ++i) {
    if (prune && i + 1 == h->layers.size()) {
      // [CLS] rows only, and the block's last LayerNorm + pooler + match head as ONE kernel (cls_tail.cu)
      MMR_TRY(cls_tail_block(c, h->layers[i].att, h->layers[i].ffn, 0, B, S, mask, false));
      const LNp& ln = h->layers[i].ffn.ln;
      MMR_TRY(cls_pool_head(h->xc32, ln.gamma, ln.beta, h->pooler.w16, h->pooler.bias, zk ? 0 : 1, zk ? h->am_wn : h->cls_w,
                            zk ? nullptr : h->cls_b, zk ? in->labels : nullptr, B, h->pooled32, probs, logits, c.dt, c.st));
      return c.mark(K_ROW, 2.0 * B * H * H);
    }
    MMR_TRY(bert_layer(c, h->layers[i], 0, B, S, mask));
    if (h->layer_tap != nullptr && h->keep_taps >= 2)
      MMR_CUDA_OK(cudaMemcpyAsync(h->layer_tap + i * size_t(h->rows_max) * H, h->x32, size_t(B) * S * H * 4,
                                  cudaMemcpyDeviceToDevice, c.st));
  }
  MMR_TRY(pooler(c, B, S, false));
  if (zk)
    MMR_TRY(zk_head(h->pooled32, h->am_wn, in->labels, B, probs, logits, c.st));
  else
    MMR_TRY(linear_head(h->pooled32, H, nullptr, nullptr, h->cls_w, h->cls_b, B, probs, logits, c.st));
  return c.mark(K_ROW, 0);
}

static mmr_status forward_lxmert(Ctx& c, const mmr_inputs* in, int B, float* probs, float* logits) {
  mmr_handle* h = c.h;
  const mmr_config& cfg = h->cfg;
  const int H = c.H, R = cfg.nbox, Lq = cfg.lq;
  MMR_REQUIRE(in->query_ids && in->label_ids && in->feats && in->boxes && in->query_mask && in->visn_mask,
              "mmr_forward(lxmert): missing input pointer");
  const int64_t v0 = int64_t(B) * Lq;  // first visual row
  const int nl = B * Lq, nv = B * R;
  // Language blocks once per distinct query of the batch (mmr_inputs.lang_unique / lang_slot): U row groups instead of B.
  const int U = (tuning(MMR_TUNE_LX_QUERY_DEDUP) != 0 && h->keep_taps == 0 && in->lang_unique != nullptr &&
                 in->lang_slot != nullptr && in->n_lang_unique > 0 && in->n_lang_unique < B && !h->layers.empty())
                    ? in->n_lang_unique : 0;
  // With query grouping, the compact language stream can ride the merged two-stream launches when it is padded to
  // whole 256-row tiles (groups beyond U repeat the last distinct query): it then sits in the main buffers right BEFORE
  // the visual rows, [v0 - nlp, v0) -- rows that belong to the expanded language stream only after the language blocks.
  const int nlp = U > 0 ? ((U * Lq + 255) / 256) * 256 : 0;
  const bool ride = U > 0 && tuning(MMR_TUNE_LX_MERGE) != 0 && nlp % Lq == 0 && nlp <= nl && !h->r_layers.empty();
  const int Up = ride ? nlp / Lq : U;                    // language row groups actually computed
  const int64_t r0 = ride ? v0 - nlp : 0;                // first row of the compact language stream
  // language embedding (modeling.py:913)
  if (U > 0) {
    MMR_TRY(lx_lang_embed(in->query_ids, h->E, h->T, h->P, h->emb_ln.gamma, h->emb_ln.beta, Lq, Up,
                          ride ? static_cast<void*>(c.x16(r0)) : h->xl16c, ride ? c.x32(r0) : h->xl32c, c.dt, c.st,
                          in->lang_unique, U));
    MMR_TRY(c.mark(K_ROW, 0));
    MMR_TRY(lx_gather_mask(in->query_mask, in->lang_unique, U, Lq, Up, h->lang_mask_c, c.st));
  } else {
    MMR_TRY(lx_lang_embed(in->query_ids, h->E, h->T, h->P, h->emb_ln.gamma, h->emb_ln.beta, Lq, B, c.x16(0), c.x32(0),
                          c.dt, c.st));
  }
  MMR_TRY(c.mark(K_ROW, 0));
  // visual embedding (modeling.py:519-533): (LN(fc(f)) + LN(fc(box)) + LN(fc(conv(label_emb)))) / 3
  const float third = 1.0f / 3.0f;
  MMR_TRY(cast16(in->feats, h->f16, int64_t(nv) * cfg.feat_dim, c.dt, c.st));
  MMR_TRY(c.mark(K_ROW, 0));
  MMR_TRY(c.G(h->f16, cfg.feat_dim, h->visn_fc, nv, nullptr, nullptr, 0, h->tmp32, MMR_ACT_NONE));
  MMR_TRY(c.LN(h->tmp32, h->visn_ln, nv, nullptr, c.x32(v0), third, 0));
  MMR_TRY(lx_box_ln(in->boxes, h->box_w, h->box_b, h->box_ln.gamma, h->box_ln.beta, third, nv, c.x32(v0), c.st));
  MMR_TRY(c.mark(K_ROW, 0));
  MMR_TRY(lx_label_z(in->label_ids, h->E, h->T, h->P, h->emb_ln.gamma, h->emb_ln.beta, h->wconv, h->bconv, nv, h->t16,
                     c.dt, c.st));
  MMR_TRY(c.mark(K_ROW, 0));
  MMR_TRY(c.G(h->t16, H, h->label_fc, nv, nullptr, nullptr, 0, h->tmp32, MMR_ACT_NONE));
  MMR_TRY(c.LN(h->tmp32, h->label_ln, nv, c.x16(v0), c.x32(v0), third, 1));
  if (h->keep_taps)
    MMR_CUDA_OK(cudaMemcpyAsync(h->emb_tap, h->x32, size_t(nl + nv) * H * 4, cudaMemcpyDeviceToDevice, c.st));
  // language layers (modeling.py:577-578) and visual layers (:582-583) are independent chains: the first
  // min(9, 5) of each run pairwise through merged launches, the rest alone
  const bool merge = tuning(MMR_TUNE_LX_MERGE) != 0 && nl % 256 == 0;
  if (U > 0 && ride) {
    // language blocks (modeling.py:577-578) and visual blocks (:582-583) pairwise through the merged launches, the
    // rest alone; then the compact stream moves aside and every pair receives its query's rows
    const size_t both = std::min(h->layers.size(), h->r_layers.size());
    for (size_t i = 0; i < both; ++i) {
      MMR_TRY(two_stream_att(c, h->layers[i].att, h->r_layers[i].att, B, Lq, R, h->lang_mask_c, in->visn_mask, r0, Up));
      MMR_TRY(two_stream_ffn(c, h->layers[i].ffn, h->r_layers[i].ffn, nlp, nv, r0));
    }
    for (size_t i = both; i < h->layers.size(); ++i) MMR_TRY(bert_layer(c, h->layers[i], r0, Up, Lq, h->lang_mask_c));
    for (size_t i = both; i < h->r_layers.size(); ++i) MMR_TRY(bert_layer(c, h->r_layers[i], v0, B, R, in->visn_mask));
    MMR_CUDA_OK(cudaMemcpyAsync(h->xl32c, c.x32(r0), size_t(U) * Lq * H * 4, cudaMemcpyDeviceToDevice, c.st));
    MMR_CUDA_OK(cudaMemcpyAsync(h->xl16c, c.x16(r0), size_t(U) * Lq * H * 2, cudaMemcpyDeviceToDevice, c.st));
    MMR_TRY(lx_expand_rows(h->xl32c, h->xl16c, in->lang_slot, Lq, B, c.x32(0), c.x16(0), c.dt, c.st));
    MMR_TRY(c.mark(K_ROW, 0));
  } else if (U > 0) {
    // the language-only blocks on the compact stream in its own buffers, the visual ones on theirs
    Ctx cl = c;
    cl.x16_base = h->xl16c;
    cl.x32_base = h->xl32c;
    for (const Layer& L : h->layers) MMR_TRY(bert_layer(cl, L, 0, U, Lq, h->lang_mask_c));
    c.launches = cl.launches;
    for (const Layer& L : h->r_layers) MMR_TRY(bert_layer(c, L, v0, B, R, in->visn_mask));
    MMR_TRY(lx_expand_rows(h->xl32c, h->xl16c, in->lang_slot, Lq, B, c.x32(0), c.x16(0), c.dt, c.st));
    MMR_TRY(c.mark(K_ROW, 0));
  }
  const size_t n_both = (merge && U == 0) ? std::min(h->layers.size(), h->r_layers.size()) : 0;
  for (size_t i = 0; i < n_both; ++i) {
    MMR_TRY(two_stream_att(c, h->layers[i].att, h->r_layers[i].att, B, Lq, R, in->query_mask, in->visn_mask));
    MMR_TRY(two_stream_ffn(c, h->layers[i].ffn, h->r_layers[i].ffn, nl, nv));
  }
  for (size_t i = n_both; U == 0 && i < h->layers.size(); ++i) MMR_TRY(bert_layer(c, h->layers[i], 0, B, Lq, in->query_mask));
  for (size_t i = n_both; U == 0 && i < h->r_layers.size(); ++i) MMR_TRY(bert_layer(c, h->r_layers[i], v0, B, R, in->visn_mask));
  const bool prune = prune_last(h);
  for (size_t xi = 0; xi < h->x_layers.size(); ++xi) {                                        // :589-591
    const XLayer& X = h->x_layers[xi];
    // cross attention both ways with ONE weight set, both from the pre-update streams (modeling.py:462-463)
    MMR_TRY(qkv_proj(c, X.cross.qkv, 0, nl + nv));
    const bool last_pruned = prune && xi + 1 == h->x_layers.size();
    if (last_pruned) MMR_TRY(attend(c, 0, Lq, v0, R, in->visn_mask, B));
    else MMR_TRY(attend2(c, 0, Lq, v0, R, in->visn_mask, v0, R, 0, Lq, in->query_mask, B));
    if (last_pruned) {
      // last cross layer: the pooler reads lang[:, 0] (modeling.py:925), so its visual half (cross attention into the
      // visual stream, visual self-attention and FFN, modeling.py:468-479) feeds nothing; the language half needs the
      // cross-attended language rows as keys / values of its self-attention, then only the [CLS] rows
      MMR_TRY(out_proj_ln(c, X.cross, 0, nl));
      MMR_TRY(cls_tail_block(c, X.lang_self, X.lang_ffn, 0, B, Lq, in->query_mask));
      break;
    }
    MMR_TRY(out_proj_ln(c, X.cross, 0, nl + nv));
    if (merge) {
      MMR_TRY(two_stream_att(c, X.lang_self, X.visn_self, B, Lq, R, in->query_mask, in->visn_mask));
      MMR_TRY(two_stream_ffn(c, X.lang_ffn, X.visn_ffn, nl, nv));
    } else {
      MMR_TRY(self_att_block(c, X.lang_self, 0, B, Lq, in->query_mask));
      MMR_TRY(self_att_block(c, X.visn_self, v0, B, R, in->visn_mask));
      MMR_TRY(ffn_block(c, X.lang_ffn, 0, nl));
      MMR_TRY(ffn_block(c, X.visn_ffn, v0, nv));
    }
  }
  MMR_TRY(pooler(c, B, Lq, prune));
  // logit_fc: Linear(768,1536) -> erf-GELU -> LayerNorm(1536) -> Linear(1536,2)  (kdd_model.py:167-172)
  MMR_TRY(gemm(h->pooled16, H, h->logit0.w16, H, B, 2 * H, H, h->logit0.bias, nullptr, 0, nullptr, 0, h->head32, 2 * H,
               MMR_ACT_GELU_ERF, c.dt, c.st));
  MMR_TRY(c.mark(K_GEMM, 2.0 * B * 2 * H * H));
  MMR_TRY(linear_head(h->head32, 2 * H, h->logit_ln.gamma, h->logit_ln.beta, h->logit3_w, h->logit3_b, B, probs,
                      logits, c.st));
  return c.mark(K_ROW, 0);
}


// ------------------------------------------------------------------------------------------ strict precision
// Same graphs as above with fp32 activations end to end and every tensor-core GEMM on two-term split operands
// (strict.cu).  Built from plain pieces -- split, GEMM (+bias, +residual), two-pass LayerNorm, fp32 attention -- with
// no fused GEMM+LN and no 16-bit activation buffer: this mode trades throughput for ~1e-5 agreement with the fp32
// reference path and is meant to be easy to audit.
struct SCtx {
  mmr_handle* h;
  cudaStream_t st;
  int dt, H;
  Ctx* book;   // launch bookkeeping (counts, profiling events)
  // out32 = act_out(split(act_in(A32)) . W^T + bias) (+ residual)
  mmr_status GS(const float* A32, int64_t lda, int rows, const Linear& W, int act_in, const float* residual, int64_t ldr,
                float* out32, int64_t ldo, int act_out, bool single_cta = false) {
    MMR_TRY(split3(A32, lda, rows, W.k, h->a16s, 3 * int64_t(W.k), act_in, 0, dt, st));
    MMR_TRY(book->mark(K_ROW, 0));
    MMR_TRY(gemm(h->a16s, 3 * int64_t(W.k), W.w16, W.kw, rows, W.n, W.kw, W.bias, residual, ldr, nullptr, 0, out32, ldo,
                 act_out, dt, st, single_cta));
    return book->mark(K_GEMM, 2.0 * rows * W.n * W.k);   // algorithmic FLOPs (the split triples the MMA work)
  }
  mmr_status LN(float* x, int64_t ldx, const LNp& p, int rows, float* out32, int64_t ldo, float scale, int accumulate) {
    MMR_TRY(layernorm(x, ldx, p.gamma, p.beta, 1e-12f, rows, H, nullptr, 0, out32, ldo, scale, accumulate, dt, st));
    return book->mark(K_LAYERNORM, 0.0);
  }
  // attention + output projection + LayerNorm, in place on x32 rows [row0, row0 + B*Sq); keys from [krow0, + B*Sk)
  mmr_status att_block(const AttBlock& A, int64_t row0, int Sq, int64_t krow0, int Sk, const int32_t* key_mask, int B,
                       bool project) {
    const int64_t H3 = 3 * int64_t(H);
    if (project) MMR_TRY(GS(h->x32 + row0 * H, H, B * Sq, A.qkv, MMR_ACT_NONE, nullptr, 0, h->qkv32 + row0 * H3, H3, MMR_ACT_NONE));
    MMR_TRY(attention_f32(h->qkv32 + row0 * H3, Sq * H3, H3, h->qkv32 + krow0 * H3 + H, h->qkv32 + krow0 * H3 + 2 * H,
                          Sk * H3, H3, key_mask, h->ctx32 + row0 * H, int64_t(Sq) * H, H, B, Sq, Sk, h->cfg.heads, st));
    MMR_TRY(book->mark(K_ATTENTION, 4.0 * B * Sq * Sk * H));
    return MMR_OK;
  }
  mmr_status out_ln(const AttBlock& A, int64_t row0, int rows) {
    float* x = h->x32 + row0 * H;
    MMR_TRY(GS(h->ctx32 + row0 * H, H, rows, A.out, MMR_ACT_NONE, x, H, x, H, MMR_ACT_NONE));
    return LN(x, H, A.ln, rows, x, H, 1.0f, 0);
  }
  mmr_status ffn(const FfnBlock& F, int64_t row0, int rows, int act) {
    float* x = h->x32 + row0 * H;
    float* hb = h->h32 + row0 * F.in.n;
    MMR_TRY(GS(x, H, rows, F.in, MMR_ACT_NONE, nullptr, 0, hb, F.in.n, MMR_ACT_NONE));
    MMR_TRY(GS(hb, F.in.n, rows, F.out, act, x, H, x, H, MMR_ACT_NONE));   // GELU applied in fp32 while splitting
    return LN(x, H, F.ln, rows, x, H, 1.0f, 0);
  }
  mmr_status self_layer(const AttBlock& A, const FfnBlock& F, int64_t row0, int B, int S, const int32_t* key_mask, int act) {
    MMR_TRY(att_block(A, row0, S, row0, S, key_mask, B, true));
    MMR_TRY(out_ln(A, row0, B * S));
    return ffn(F, row0, B * S, act);
  }
  // last block for the [CLS] rows only: results land in xc32 [B, H]
  mmr_status cls_tail(const AttBlock& A, const FfnBlock& F, int64_t row0, int B, int S, const int32_t* key_mask, int act) {
    const int64_t H3 = 3 * int64_t(H);
    MMR_TRY(GS(h->x32 + row0 * H, H, B * S, A.qkv, MMR_ACT_NONE, nullptr, 0, h->qkv32 + row0 * H3, H3, MMR_ACT_NONE));
    // one query row per pair (stride S rows), compact context rows in ctx32 [B, H]
    MMR_TRY(attention_f32(h->qkv32 + row0 * H3, S * H3, H3, h->qkv32 + row0 * H3 + H, h->qkv32 + row0 * H3 + 2 * H, S * H3,
                          H3, key_mask, h->ctx32, H, H, B, 1, S, h->cfg.heads, st));
    MMR_TRY(book->mark(K_ATTENTION, 4.0 * B * S * H));
    MMR_TRY(GS(h->ctx32, H, B, A.out, MMR_ACT_NONE, h->x32 + row0 * H, int64_t(S) * H, h->xc32, H, MMR_ACT_NONE, true));
    MMR_TRY(LN(h->xc32, H, A.ln, B, h->xc32, H, 1.0f, 0));
    MMR_TRY(GS(h->xc32, H, B, F.in, MMR_ACT_NONE, nullptr, 0, h->h32, F.in.n, MMR_ACT_NONE, true));
    MMR_TRY(GS(h->h32, F.in.n, B, F.out, act, h->xc32, H, h->xc32, H, MMR_ACT_NONE, true));
    return LN(h->xc32, H, F.ln, B, h->xc32, H, 1.0f, 0);
  }
  mmr_status pool(int B, int S, bool compact) {
    return GS(compact ? h->xc32 : h->x32, compact ? int64_t(H) : int64_t(S) * H, B, h->pooler, MMR_ACT_NONE, nullptr, 0,
              h->pooled32, H, MMR_ACT_TANH, compact);
  }
};

static mmr_status forward_single_stream_strict(Ctx& c, const mmr_inputs* in, int B, float* probs, float* logits) {
  mmr_handle* h = c.h;
  const mmr_config& cfg = h->cfg;
  const int H = c.H, R = cfg.nbox, Lq = cfg.lq;
  const bool zk = cfg.model_kind == MMR_MODEL_IMAGEBERT_ZK;
  const int S = zk ? Lq + R : Lq + 2 * R;
  SCtx s{h, c.st, c.dt, H, &c};
  const bool fused_in = zk && in->region_sum != nullptr;
  MMR_REQUIRE(in->query_ids && in->segment_ids && (fused_in || (in->label_ids && in->feats)),
              "mmr_forward: missing input pointer");
  const int32_t* mask = nullptr;
  if (zk) {
    MMR_REQUIRE(in->len_query && in->num_boxes && in->labels && (fused_in || in->boxes),
                "mmr_forward(zk): missing input pointer");
    const float* region_sum = in->region_sum;
    if (!fused_in) {
      MMR_TRY(s.GS(in->feats, cfg.feat_dim, B * R, h->conv2, MMR_ACT_NONE, nullptr, 0, h->tmp32, H, MMR_ACT_RELU));
      if ((reinterpret_cast<uintptr_t>(in->label_ids) & 15) == 0) {
        MMR_TRY(zk_label_terms(in->label_ids, h->tables, cfg.vocab, h->bc1, h->lab_tab, h->lab_tab_mask, h->lab_epoch_dev,
                               h->lab_rep, h->lab_term32, B * R, c.st));
        MMR_TRY(c.mark(K_ROW, 0));
        MMR_TRY(c.mark(K_ROW, 0));
        MMR_TRY(zk_region_sum_rep(h->tmp32, in->boxes, h->lab_rep, h->lab_term32, h->Wb, h->bb, nullptr, B * R,
                                  h->lab_epoch_dev, c.dt, c.st, h->ctx32));
      } else {
        MMR_TRY(zk_region_sum(h->tmp32, in->boxes, in->label_ids, h->tables, cfg.vocab, h->bc1, h->Wb, h->bb, nullptr, B * R,
                              c.dt, c.st, h->ctx32));
      }
      MMR_TRY(c.mark(K_ROW, 0));
      region_sum = h->ctx32;
    }
    MMR_TRY(s.GS(region_sum, H, B * R, h->featureemb, MMR_ACT_NONE, nullptr, 0, h->tmp32, H, MMR_ACT_NONE));
    MMR_TRY(zk_embed(in->query_ids, in->segment_ids, h->tmp32, in->len_query, in->num_boxes, h->E, h->T, h->P,
                     h->emb_ln.gamma, h->emb_ln.beta, Lq, R, B, h->x16, h->x32, h->key_mask, c.dt, c.st));
    MMR_TRY(c.mark(K_ROW, 0));
    mask = h->key_mask;
  } else {
    MMR_TRY(s.GS(in->feats, cfg.feat_dim, B * R, h->lds_feat, MMR_ACT_NONE, nullptr, 0, h->tmp32, H, MMR_ACT_NONE));
    MMR_TRY(lds_embed(in->query_ids, in->segment_ids, in->label_ids, h->tmp32, h->E, h->T, h->P, h->emb_ln.gamma,
                      h->emb_ln.beta, h->wl, Lq, R, B, h->x16, h->x32, c.dt, c.st));
    MMR_TRY(c.mark(K_ROW, 0));
  }
  if (h->keep_taps)
    MMR_CUDA_OK(cudaMemcpyAsync(h->emb_tap, h->x32, size_t(B) * S * H * 4, cudaMemcpyDeviceToDevice, c.st));
  const bool prune = prune_last(h);
  for (size_t i = 0; i < h->layers.size(); ++i) {
    const Layer& L = h->layers[i];
    if (prune && i + 1 == h->layers.size()) {
      MMR_TRY(s.cls_tail(L.att, L.ffn, 0, B, S, mask, h->act));
      break;
    }
    MMR_TRY(s.self_layer(L.att, L.ffn, 0, B, S, mask, h->act));
    if (h->layer_tap != nullptr && h->keep_taps >= 2)
      MMR_CUDA_OK(cudaMemcpyAsync(h->layer_tap + i * size_t(h->rows_max) * H, h->x32, size_t(B) * S * H * 4,
                                  cudaMemcpyDeviceToDevice, c.st));
  }
  MMR_TRY(s.pool(B, S, prune));
  if (zk)
    MMR_TRY(zk_head(h->pooled32, h->am_wn, in->labels, B, probs, logits, c.st));
  else
    MMR_TRY(linear_head(h->pooled32, H, nullptr, nullptr, h->cls_w, h->cls_b, B, probs, logits, c.st));
  return c.mark(K_ROW, 0);
}

static mmr_status forward_lxmert_strict(Ctx& c, const mmr_inputs* in, int B, float* probs, float* logits) {
  mmr_handle* h = c.h;
  const mmr_config& cfg = h->cfg;
  const int H = c.H, R = cfg.nbox, Lq = cfg.lq;
  SCtx s{h, c.st, c.dt, H, &c};
  MMR_REQUIRE(in->query_ids && in->label_ids && in->feats && in->boxes && in->query_mask && in->visn_mask,
              "mmr_forward(lxmert): missing input pointer");
  const int64_t v0 = int64_t(B) * Lq;
  const int nl = B * Lq, nv = B * R;
  const int act = h->act;
  MMR_TRY(lx_lang_embed(in->query_ids, h->E, h->T, h->P, h->emb_ln.gamma, h->emb_ln.beta, Lq, B, c.x16(0), c.x32(0),
                        c.dt, c.st));
  MMR_TRY(c.mark(K_ROW, 0));
  const float third = 1.0f / 3.0f;
  float* xv = c.x32(v0);
  MMR_TRY(s.GS(in->feats, cfg.feat_dim, nv, h->visn_fc, MMR_ACT_NONE, nullptr, 0, h->tmp32, H, MMR_ACT_NONE));
  MMR_TRY(s.LN(h->tmp32, H, h->visn_ln, nv, xv, H, third, 0));
  MMR_TRY(lx_box_ln(in->boxes, h->box_w, h->box_b, h->box_ln.gamma, h->box_ln.beta, third, nv, xv, c.st));
  MMR_TRY(c.mark(K_ROW, 0));
  MMR_TRY(lx_label_z(in->label_ids, h->E, h->T, h->P, h->emb_ln.gamma, h->emb_ln.beta, h->wconv, h->bconv, nv, nullptr,
                     c.dt, c.st, h->ctx32));
  MMR_TRY(c.mark(K_ROW, 0));
  MMR_TRY(s.GS(h->ctx32, H, nv, h->label_fc, MMR_ACT_NONE, nullptr, 0, h->tmp32, H, MMR_ACT_NONE));
  MMR_TRY(s.LN(h->tmp32, H, h->label_ln, nv, xv, H, third, 1));
  if (h->keep_taps)
    MMR_CUDA_OK(cudaMemcpyAsync(h->emb_tap, h->x32, size_t(nl + nv) * H * 4, cudaMemcpyDeviceToDevice, c.st));
  for (const Layer& L : h->layers) MMR_TRY(s.self_layer(L.att, L.ffn, 0, B, Lq, in->query_mask, act));
  for (const Layer& L : h->r_layers) MMR_TRY(s.self_layer(L.att, L.ffn, v0, B, R, in->visn_mask, act));
  const bool prune = prune_last(h);
  for (size_t xi = 0; xi < h->x_layers.size(); ++xi) {
    const XLayer& X = h->x_layers[xi];
    const bool last = prune && xi + 1 == h->x_layers.size();
    // one projection of both streams with the shared cross-attention weights, then attention each way from the
    // pre-update streams (modeling.py:462-463)
    MMR_TRY(s.GS(h->x32, H, nl + nv, X.cross.qkv, MMR_ACT_NONE, nullptr, 0, h->qkv32, 3 * int64_t(H), MMR_ACT_NONE));
    MMR_TRY(s.att_block(X.cross, 0, Lq, v0, R, in->visn_mask, B, false));
    if (last) {
      MMR_TRY(s.out_ln(X.cross, 0, nl));
      MMR_TRY(s.cls_tail(X.lang_self, X.lang_ffn, 0, B, Lq, in->query_mask, act));
      break;
    }
    MMR_TRY(s.att_block(X.cross, v0, R, 0, Lq, in->query_mask, B, false));
    MMR_TRY(s.out_ln(X.cross, 0, nl + nv));
    MMR_TRY(s.att_block(X.lang_self, 0, Lq, 0, Lq, in->query_mask, B, true));
    MMR_TRY(s.out_ln(X.lang_self, 0, nl));
    MMR_TRY(s.att_block(X.visn_self, v0, R, v0, R, in->visn_mask, B, true));
    MMR_TRY(s.out_ln(X.visn_self, v0, nv));
    MMR_TRY(s.ffn(X.lang_ffn, 0, nl, act));
    MMR_TRY(s.ffn(X.visn_ffn, v0, nv, act));
  }
  MMR_TRY(s.pool(B, Lq, prune));
  // logit_fc: Linear(768,1536) -> erf-GELU -> LayerNorm(1536) -> Linear(1536,2)  (kdd_model.py:167-172)
  MMR_TRY(s.GS(h->pooled32, H, B, h->logit0, MMR_ACT_NONE, nullptr, 0, h->head32, 2 * int64_t(H), MMR_ACT_NONE));
  MMR_TRY(act32(h->head32, int64_t(B) * 2 * H, MMR_ACT_GELU_ERF, c.st));
  MMR_TRY(c.mark(K_ROW, 0));
  MMR_TRY(linear_head(h->head32, 2 * H, h->logit_ln.gamma, h->logit_ln.beta, h->logit3_w, h->logit3_b, B, probs,
                      logits, c.st));
  return c.mark(K_ROW, 0);
}

}  // namespace mmr

// ------------------------------------------------------------------------------------------ C ABI
extern "C" mmr_status mmr_create(const mmr_config* cfg, const mmr_tensor* weights, int n_weights, int device,
                                 mmr_handle** out) {
  using namespace mmr;
  MMR_REQUIRE(cfg && weights && out && n_weights > 0, "mmr_create: null argument");
  *out = nullptr;
  MMR_TRY(mmr_device_check(device));
  MMR_REQUIRE(device >= 0 && device < 64, "mmr_create: device index %d out of range", device);
  DeviceGuard guard(device);
  MMR_REQUIRE(cfg->precision == MMR_PRECISION_FAST || cfg->precision == MMR_PRECISION_STRICT,
              "mmr_create: bad precision %d", cfg->precision);
  MMR_REQUIRE(cfg->hidden == 768 && cfg->heads == 12, "mmr_create: only hidden=768 / heads=12 (bert_config.json) is built");
  MMR_REQUIRE(cfg->intermediate % 64 == 0 && cfg->intermediate % 16 == 0, "mmr_create: intermediate must be a multiple of 64");
  MMR_REQUIRE(cfg->feat_dim % 64 == 0, "mmr_create: feat_dim must be a multiple of 64");
  MMR_REQUIRE(cfg->label_len == 8, "mmr_create: label_len must be 8 (8-tap label conv / Conv2d(8,1,1))");
  MMR_REQUIRE(cfg->dtype == MMR_DT_BF16 || cfg->dtype == MMR_DT_FP16, "mmr_create: bad dtype");
  MMR_REQUIRE(cfg->model_kind >= 0 && cfg->model_kind <= 2, "mmr_create: bad model_kind %d", cfg->model_kind);
  MMR_REQUIRE(cfg->lq > 0 && cfg->nbox > 0 && cfg->max_batch > 0, "mmr_create: lq / nbox / max_batch must be positive");
  const int S = cfg->model_kind == MMR_MODEL_IMAGEBERT_LDS ? cfg->lq + 2 * cfg->nbox : cfg->lq + cfg->nbox;
  MMR_REQUIRE((cfg->model_kind == MMR_MODEL_LXMERT ? std::max(cfg->lq, cfg->nbox) : S) <= 128,
              "mmr_create: sequence longer than 128 tokens is not supported by the attention kernel");
  MMR_REQUIRE(cfg->lq + 1 <= cfg->max_pos, "mmr_create: lq exceeds max_pos");
  if (cfg->model_kind == MMR_MODEL_LXMERT)
    MMR_REQUIRE(cfg->n_layers >= 0 && cfg->n_r_layers >= 0 && cfg->n_x_layers >= 0, "mmr_create: negative layer count");

  mmr_handle* h = new mmr_handle();
  h->cfg = *cfg;
  h->device = device;
  h->strict = cfg->precision == MMR_PRECISION_STRICT;
  TensorMap tm;
  for (int i = 0; i < n_weights; ++i)
    if (weights[i].name) tm[weights[i].name] = &weights[i];
  mmr_status rc = MMR_OK;
  const size_t wbytes = weight_arena_bytes(*cfg);
  if (cudaMalloc(reinterpret_cast<void**>(&h->weights.base), wbytes) != cudaSuccess) {
    rc = fail(MMR_ERR_NOMEM, "cannot allocate %zu bytes for weights", wbytes);
  } else {
    h->weights.cap = wbytes;
    cudaStream_t st;
    if (cudaStreamCreate(&st) != cudaSuccess) {
      rc = fail(MMR_ERR_CUDA, "cudaStreamCreate failed");
    } else {
      rc = pack_weights(h, tm, st);
      cudaStreamDestroy(st);
    }
  }
  if (rc == MMR_OK) rc = alloc_workspace(h);
  if (rc != MMR_OK) {
    mmr_destroy(h);
    return rc;
  }
  *out = h;
  return MMR_OK;
}

extern "C" mmr_status mmr_workspace_bytes(const mmr_config* cfg, size_t* weight_bytes, size_t* workspace_bytes) {
  using namespace mmr;
  MMR_REQUIRE(cfg && (weight_bytes || workspace_bytes), "mmr_workspace_bytes: null argument");
  MMR_REQUIRE(cfg->model_kind >= 0 && cfg->model_kind <= 2 && cfg->lq > 0 && cfg->nbox > 0 && cfg->max_batch > 0 &&
                  cfg->hidden > 0 && cfg->intermediate > 0 && cfg->feat_dim > 0,
              "mmr_workspace_bytes: bad configuration");
  if (weight_bytes) *weight_bytes = weight_arena_bytes(*cfg);
  if (workspace_bytes) {
    mmr_handle plan;
    plan.cfg = *cfg;
    plan.strict = cfg->precision == MMR_PRECISION_STRICT;
    MMR_TRY(alloc_workspace(&plan, workspace_bytes));
  }
  return MMR_OK;
}

extern "C" void mmr_destroy(mmr_handle* h) {
  if (!h) return;
  mmr::DeviceGuard guard(h->device);
  cudaDeviceSynchronize();   // nothing of this handle may still run when its arenas (and exchange table) go away
  {
    std::lock_guard<std::mutex> lock(mmr::g_tail_mu);
    if (h->device >= 0 && h->device < 64 && mmr::g_tail[h->device].owner == h) mmr::g_tail[h->device] = mmr::DeviceTail();
  }
  if (h->done_ev) cudaEventDestroy(h->done_ev);
  if (h->weights.base) cudaFree(h->weights.base);
  if (h->work.base) cudaFree(h->work.base);
  if (h->layer_tap) cudaFree(h->layer_tap);
  for (cudaEvent_t e : h->prof_ev) cudaEventDestroy(e);
  delete h;
}

extern "C" mmr_status mmr_forward(mmr_handle* h, const mmr_inputs* in, int B, float* probs_out, float* logits_out,
                                  float* pooled_out, void* stream) {
  using namespace mmr;
  MMR_REQUIRE(h && in && probs_out, "mmr_forward: null argument");
  MMR_REQUIRE(B > 0 && B <= h->cfg.max_batch, "mmr_forward: B=%d outside (0, max_batch=%d]", B, h->cfg.max_batch);
  DeviceGuard guard(h->device);
  MMR_TRY(require_sm100());
  Ctx c{h, static_cast<cudaStream_t>(stream), h->cfg.dtype, h->cfg.hidden};
  // one forward in flight per device: wait (on the device) for the previous eager forward if it went to another stream
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  MMR_CUDA_OK(cudaStreamIsCapturing(c.st, &cap));
  const bool eager = cap == cudaStreamCaptureStatusNone;
  DeviceTail& tail = g_tail[h->device];
  if (eager) {
    std::lock_guard<std::mutex> lock(g_tail_mu);
    if (tail.ev != nullptr && tail.stream != c.st) MMR_CUDA_OK(cudaStreamWaitEvent(c.st, tail.ev, 0));
  }
  if (h->prof_on) {
    h->prof_n = 0;
    MMR_CUDA_OK(cudaEventRecord(h->prof_ev[0], c.st));
  }
  mmr_status rc;
  if (h->strict)
    rc = h->cfg.model_kind == MMR_MODEL_LXMERT ? forward_lxmert_strict(c, in, B, probs_out, logits_out)
                                               : forward_single_stream_strict(c, in, B, probs_out, logits_out);
  else
    rc = h->cfg.model_kind == MMR_MODEL_LXMERT ? forward_lxmert(c, in, B, probs_out, logits_out)
                                               : forward_single_stream(c, in, B, probs_out, logits_out);
  if (rc != MMR_OK) return rc;
  if (pooled_out != nullptr)
    MMR_CUDA_OK(cudaMemcpyAsync(pooled_out, h->pooled32, size_t(B) * h->cfg.hidden * 4, cudaMemcpyDeviceToDevice,
                                c.st));
  if (eager) {
    std::lock_guard<std::mutex> lock(g_tail_mu);
    MMR_CUDA_OK(cudaEventRecord(h->done_ev, c.st));
    tail.ev = h->done_ev;
    tail.stream = c.st;
    tail.owner = h;
  }
  h->last_B = B;
  h->pruned_last = prune_last(h) ? 1 : 0;
  h->launches = c.launches;
  return MMR_OK;
}

extern "C" mmr_status mmr_set_debug_taps(mmr_handle* h, int enable) {
  MMR_REQUIRE(h, "mmr_set_debug_taps: null handle");
  MMR_REQUIRE(enable >= 0 && enable <= 2, "mmr_set_debug_taps: enable must be 0, 1 or 2");
  if (enable == 2 && h->layer_tap == nullptr) {
    MMR_REQUIRE(h->cfg.model_kind != MMR_MODEL_LXMERT, "mmr_set_debug_taps: per-layer taps are for single-stream models");
    mmr::DeviceGuard guard(h->device);
    MMR_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&h->layer_tap),
                           h->layers.size() * size_t(h->rows_max) * h->cfg.hidden * 4));
  }
  h->keep_taps = enable;
  return MMR_OK;
}

extern "C" mmr_status mmr_get_activation(mmr_handle* h, int which, float* dst, int64_t n_floats, void* stream) {
  using namespace mmr;
  MMR_REQUIRE(h && dst, "mmr_get_activation: null argument");
  MMR_REQUIRE(h->last_B > 0, "mmr_get_activation: no forward has run yet");
  const mmr_config& c = h->cfg;
  DeviceGuard guard(h->device);
  // bounded by the workspace, not by the last eager batch: a forward replayed from a CUDA graph does not pass through
  // mmr_forward, so the host cannot know its batch
  const int64_t have = h->rows_max * c.hidden;
  MMR_REQUIRE(n_floats > 0 && n_floats <= have, "mmr_get_activation: n_floats=%lld exceeds %lld", (long long)n_floats,
              (long long)have);
  const float* src = nullptr;
  if (which == 0) {
    MMR_REQUIRE(h->keep_taps, "mmr_get_activation: embedding tap needs mmr_set_debug_taps(h, 1) before the forward");
    src = h->emb_tap;
  } else if (which == 1) {
    MMR_REQUIRE(h->keep_taps || !h->pruned_last,
                "mmr_get_activation: the final-layer tap needs mmr_set_debug_taps(h, 1) before the forward (the default "
                "forward computes the last block for the [CLS] rows only)");
    src = h->x32;
  } else if (which >= 2 && which < 2 + int(h->layers.size()) && c.model_kind != MMR_MODEL_LXMERT) {
    MMR_REQUIRE(h->keep_taps >= 2 && h->layer_tap, "mmr_get_activation: layer taps need mmr_set_debug_taps(h, 2)");
    src = h->layer_tap + size_t(which - 2) * size_t(h->rows_max) * c.hidden;
  } else {
    return fail(MMR_ERR_INVALID, "mmr_get_activation: unknown tap %d", which);
  }
  MMR_CUDA_OK(cudaMemcpyAsync(dst, src, size_t(n_floats) * 4, cudaMemcpyDeviceToDevice,
                              static_cast<cudaStream_t>(stream)));
  return MMR_OK;
}

extern "C" mmr_status mmr_set_profiling(mmr_handle* h, int enable) {
  using namespace mmr;
  MMR_REQUIRE(h, "mmr_set_profiling: null handle");
  if (enable && h->prof_ev.empty()) {
    const int cap = 512;
    h->prof_ev.resize(cap + 1);
    for (auto& e : h->prof_ev) MMR_CUDA_OK(cudaEventCreate(&e));
    h->prof_kind.assign(cap, 0);
    h->prof_flops.assign(cap, 0.0);
  }
  h->prof_on = enable ? 1 : 0;
  h->prof_n = 0;
  return MMR_OK;
}

extern "C" int mmr_get_profile(mmr_handle* h, int cap, int32_t* kinds, float* ms, double* flops) {
  if (!h || h->prof_n == 0) return 0;
  if (cudaEventSynchronize(h->prof_ev[h->prof_n]) != cudaSuccess) return -1;
  const int n = h->prof_n < cap ? h->prof_n : cap;
  for (int i = 0; i < n; ++i) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, h->prof_ev[i], h->prof_ev[i + 1]) != cudaSuccess) return -1;
    if (kinds) kinds[i] = h->prof_kind[i];
    if (ms) ms[i] = t;
    if (flops) flops[i] = h->prof_flops[i];
  }
  return n;
}

extern "C" int mmr_launches_per_forward(const mmr_handle* h) { return h ? h->launches : 0; }

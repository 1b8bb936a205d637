// Ensemble merge, product-uniqueness filter, per-query top-k and nDCG@k on the device:
// code/main.py:41-104 and imagebert_lds/src/evaluation.py:4-38 as four small kernels over the flattened pair list, so
// that the cfg5 pipeline (three scorers -> ensemble -> top-5) can stay on the GPU.  fp64 throughout, evaluated in the
// reference's order with contraction off (Python floats: w1*r1 + w2*r2 + w3*r3 + w4*r4 left to right, no FMA), so the
// merged scores are bit-identical to the host implementation (ensemble.py) and to the shipped submission.csv.
//
// Input contract: pairs flattened in the reference's iteration order (queries in first-file order, candidates in
// LXMERT-file order), the pairs of a query contiguous: query_start[q] .. query_start[q+1]; product_of[i] in [0, P).
#include "common.cuh"
#include "ptx.cuh"

namespace mmr {

__device__ __forceinline__ unsigned long long f64_key(double v) {   // order-preserving map double -> uint64
  const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(v));
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_f64(unsigned long long k) {
  const unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double(static_cast<long long>(b));
}

__global__ void ens_merge_kernel(const double* __restrict__ s1, const double* __restrict__ s2,
                                 const double* __restrict__ s3, const double* __restrict__ s4,
                                 const int32_t* __restrict__ product_of, int64_t n, double w1, double w2, double w3,
                                 double w4, double* __restrict__ merged, unsigned long long* __restrict__ best_key) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // main.py:59, left to right, every product and sum rounded separately
  const double m = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(w1, s1[i]), __dmul_rn(w2, s2[i])), __dmul_rn(w3, s3[i])),
                             __dmul_rn(w4, s4[i]));
  merged[i] = m;
  atomicMax(best_key + product_of[i], f64_key(m));                  // main.py:65-68
}

// runner-up per product WITH multiplicity (main.py:69-72, 78-80 sorts all scores of the product): a second copy of the
// best score is a runner-up equal to the best
__global__ void ens_second_kernel(const double* __restrict__ merged, const int32_t* __restrict__ product_of, int64_t n,
                                  const unsigned long long* __restrict__ best_key,
                                  unsigned long long* __restrict__ second_key, int32_t* __restrict__ best_count) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t p = product_of[i];
  const unsigned long long k = f64_key(merged[i]);
  if (k == best_key[p]) atomicAdd(best_count + p, 1);
  else atomicMax(second_key + p, k);
}

// One warp per query: survivors of the uniqueness filter, then the top-k of the survivors (>= k of them), of all
// candidates (1 .. k-1 survivors: main.py:101-104) or nothing (no survivor).  Order: score descending, earlier pair
// first on ties (Python's sorted(..., reverse=True) is stable).
__global__ void ens_topk_kernel(const double* __restrict__ merged, const int32_t* __restrict__ product_of,
                                const int32_t* __restrict__ query_start, int32_t n_queries,
                                const unsigned long long* __restrict__ best_key,
                                const unsigned long long* __restrict__ second_key,
                                const int32_t* __restrict__ best_count, double margin, double tie, int topk,
                                int32_t* __restrict__ top, int32_t* __restrict__ status) {
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= n_queries) return;
  const int lane = threadIdx.x & 31;
  const int a = query_start[q], b = query_start[q + 1];
  auto keep = [&](int i) {
    const int32_t p = product_of[i];
    const double best = key_f64(best_key[p]);
    const bool has_second = best_count[p] >= 2 || second_key[p] != 0ull;
    const double second = best_count[p] >= 2 ? best : key_f64(second_key[p]);
    if (has_second && __dsub_rn(best, second) < margin) return false;           // main.py:80-82
    return fabs(__dsub_rn(merged[i], best)) < tie;                              // main.py:83
  };
  int survivors = 0;
  for (int i = a + lane; i < b; i += 32) survivors += keep(i) ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) survivors += __shfl_xor_sync(0xffffffffu, survivors, o);
  const int mode = survivors == 0 ? 0 : (survivors >= topk ? 1 : 2);
  if (lane == 0) status[q] = mode;
  double prev_v = INFINITY;
  int prev_i = -1;
  for (int k = 0; k < topk; ++k) {
    double bv = -INFINITY;
    int bi = 0x7fffffff;
    if (mode != 0) {
      for (int i = a + lane; i < b; i += 32) {
        if (mode == 1 && !keep(i)) continue;
        const double v = merged[i];
        const bool after_prev = v < prev_v || (v == prev_v && i > prev_i);
        if (after_prev && (v > bv || (v == bv && i < bi))) { bv = v; bi = i; }
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) top[q * topk + k] = bi == 0x7fffffff ? -1 : bi;
    prev_v = bv;
    prev_i = bi;
  }
}

// nDCG@k per query (evaluation.py:4-38): relevance = membership of the predicted product in the query's ground-truth
// list gt[gt_start[q] .. gt_start[q+1]); ideal = min(k, |gt|) ones.  ndcg[q] = -1 for queries without a prediction.
__global__ void ens_ndcg_kernel(const int32_t* __restrict__ top, const int32_t* __restrict__ product_of,
                                const int32_t* __restrict__ gt, const int32_t* __restrict__ gt_start,
                                int32_t n_queries, int topk, double* __restrict__ ndcg) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_queries) return;
  if (top[q * topk] < 0) { ndcg[q] = -1.0; return; }
  const int ga = gt_start[q], gb = gt_start[q + 1];
  double dcg = 0.0, idcg = 0.0;
  for (int k = 0; k < topk; ++k) {
    const double disc = 1.0 / log2(double(k + 2));
    if (k < gb - ga) idcg += disc;
    const int i = top[q * topk + k];
    if (i < 0) continue;
    const int32_t p = product_of[i];
    bool hit = false;
    for (int j = ga; j < gb; ++j) hit = hit || gt[j] == p;
    if (hit) dcg += disc;
  }
  ndcg[q] = idcg > 0.0 ? dcg / idcg : 0.0;
}

}  // namespace mmr

extern "C" mmr_status mmr_ensemble_topk(const double* s1, const double* s2, const double* s3, const double* s4,
                                        const int32_t* product_of, const int32_t* query_start, int64_t n_pairs,
                                        int32_t n_queries, int32_t n_products, const double* weights4, double margin,
                                        double tie, int32_t topk, double* merged, int32_t* top, int32_t* status,
                                        void* workspace, size_t workspace_bytes, void* stream) {
  using namespace mmr;
  MMR_TRY(require_sm100());
  MMR_REQUIRE(s1 && s2 && s3 && s4 && product_of && query_start && weights4 && merged && top && status && workspace,
              "mmr_ensemble_topk: null argument");
  MMR_REQUIRE(n_pairs > 0 && n_queries > 0 && n_products > 0 && topk > 0 && topk <= 32, "mmr_ensemble_topk: bad sizes");
  const size_t need = size_t(n_products) * (8 + 8 + 4);
  MMR_REQUIRE(workspace_bytes >= need, "mmr_ensemble_topk: workspace needs %zu bytes", need);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto* best_key = static_cast<unsigned long long*>(workspace);
  auto* second_key = best_key + n_products;
  auto* best_count = reinterpret_cast<int32_t*>(second_key + n_products);
  MMR_CUDA_OK(cudaMemsetAsync(workspace, 0, need, st));
  const unsigned blocks = unsigned((n_pairs + 255) / 256);
  ens_merge_kernel<<<blocks, 256, 0, st>>>(s1, s2, s3, s4, product_of, n_pairs, weights4[0], weights4[1], weights4[2],
                                           weights4[3], merged, best_key);
  ens_second_kernel<<<blocks, 256, 0, st>>>(merged, product_of, n_pairs, best_key, second_key, best_count);
  ens_topk_kernel<<<unsigned((n_queries + 7) / 8), 256, 0, st>>>(merged, product_of, query_start, n_queries, best_key,
                                                               second_key, best_count, margin, tie, topk, top, status);
  MMR_CUDA_OK(cudaGetLastError());
  return MMR_OK;
}

extern "C" mmr_status mmr_ndcg_at_k(const int32_t* top, const int32_t* product_of, const int32_t* gt,
                                    const int32_t* gt_start, int32_t n_queries, int32_t topk, double* ndcg,
                                    void* stream) {
  using namespace mmr;
  MMR_TRY(require_sm100());
  MMR_REQUIRE(top && product_of && gt && gt_start && ndcg && n_queries > 0 && topk > 0, "mmr_ndcg_at_k: bad argument");
  ens_ndcg_kernel<<<unsigned((n_queries + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      top, product_of, gt, gt_start, n_queries, topk, ndcg);
  MMR_CUDA_OK(cudaGetLastError());
  return MMR_OK;
}

// tcgen05 tensor-core GEMM for sm_100a:  out = act(A[M,K] · W[N,K]^T + bias) (+ residual)
//
// Replaces every dense projection on the hot path of the reference (tf.layers.dense at
// imagebert_zk/pixelbert.py:767-788, 960-985; slim.fully_connected at pixelbert.py:449-452 and
// imagebert_lds/src/pixelmodel.py:439-442; nn.Linear at lxmert/src/lxrt/modeling.py:325-420, 522-523).
//
// Design (one CTA per SM, persistent over output tiles):
//   warp 0      TMA producer: 128x64 A tile + BNx64 W tile per stage, 128B-swizzled, mbarrier complete_tx
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BN x 16, fp32 accum in TMEM)
//   warps 2..9  epilogue: tcgen05.ld (32 lanes x 32 cols per warp), bias / activation / residual in fp32
//               registers, vector stores of 16-bit and/or fp32 rows
//   three pipelines: smem full/empty ring (kStages deep), TMEM accumulator full/empty (2 deep, so the
//   epilogue of tile i overlaps the MMAs of tile i+1), and the static tile schedule.
#include <cuda.h>

#include <cstdlib>

#include "gemm_common.cuh"

namespace mmr {

constexpr int kBM = 128;       // rows per tile (= UMMA M, one TMEM lane per row)
constexpr int kThreads = kGemmThreads;
constexpr uint32_t kABytes = kBM * kBK * 2;
// Two shapes of the same kernel: 256-column tiles with a 4-stage operand ring (general), and 64-column tiles with an
// 8-stage ring for the few-row launches of the last block's [CLS] tail (M = B <= 512): there a tile's K loop is a chain
// of TMA latencies on a handful of CTAs, and narrow tiles put 4x the CTAs on it with twice the loads in flight.
template <int BN_T, int STAGES>
struct SingleCfg {
  static constexpr int kBnT = BN_T, kStages = STAGES;
  static constexpr uint32_t kBBytes = BN_T * kBK * 2;
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;
  static constexpr size_t kSmemBytes = 1024 /*align slack*/ + size_t(STAGES) * kStageBytes + 256 /*barriers*/ + kEpiSmemBytes;
};
using SingleWide = SingleCfg<kBN, 4>;
using SingleNarrow = SingleCfg<64, 8>;

template <int ACT, class E16, class CFG>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                    const GemmParams p) {
  constexpr int kStages = CFG::kStages, kBnT = CFG::kBnT;
  constexpr uint32_t kBBytes = CFG::kBBytes, kStageBytes = CFG::kStageBytes;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                                 // kStages x 16 KB
  uint8_t* smem_b = smem + size_t(kStages) * kABytes;     // kStages x 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + size_t(kStages) * kStageBytes);
  uint64_t* full_bar = bars;                 // [kStages]  TMA -> MMA
  uint64_t* empty_bar = bars + kStages;      // [kStages]  MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * kStages;  // [2]        MMA -> epilogue
  uint64_t* tempty_bar = tfull_bar + 2;      // [2]        epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* epi_stage = reinterpret_cast<float*>(smem + size_t(kStages) * kStageBytes + 256);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (p.M + kBM - 1) / kBM;
  const int n_tiles = (p.N + kBnT - 1) / kBnT;
  const int k_blocks = p.K / kBK;
  const int total_tiles = m_tiles * n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_w);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], kEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                 // the previous kernel's outputs (this one's operands / residual) are complete
  pdl_launch_dependents();    // the next kernel may be scheduled as soon as SMs free up

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
        // TMA always delivers (and counts) the full box: rows past M or N are zero-filled.
        const uint32_t tx = kABytes + p.w_box_rows * kBK * 2;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&full_bar[stage], tx);
          tma_load_2d(smem_a + size_t(stage) * kABytes, &tmap_a, &full_bar[stage], kb * kBK, m_blk * kBM);
          tma_load_2d(smem_b + size_t(stage) * kBBytes, &tmap_w, &full_bar[stage], kb * kBK, n_blk * kBnT);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int n_blk = tile % n_tiles;
        const int bn = min(kBnT, p.N - n_blk * kBnT);
        const uint32_t idesc = umma_idesc_f16(p.idesc_fmt, kBM, uint32_t(bn));
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1u;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + uint32_t(acc) * kBnT;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + size_t(stage) * kABytes);
          const uint32_t b_addr = smem_u32(smem_b + size_t(stage) * kBBytes);
          const uint64_t a_desc = umma_desc_k_sw128(a_addr);
          const uint64_t b_desc = umma_desc_k_sw128(b_addr);
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k) {
            // advance 16 elements (32 B) along K inside the swizzle atom: +2 in the (addr >> 4) field
            umma_f16(tmem_d, a_desc + uint64_t(2 * k), b_desc + uint64_t(2 * k), idesc,
                     (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees this smem stage once the MMAs above retire
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int ew = warp - 2;                // 0..7
    const int quarter = warp & 3;           // TMEM lane quarter this warp may touch
    const int half = ew >> 2;               // which 128-column half of the accumulator
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      const int bn = min(kBnT, p.N - n_blk * kBnT);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int row0 = m_blk * kBM + quarter * 32;
      const uint32_t taddr_row = tmem_base + uint32_t(acc) * kBnT + (uint32_t(quarter * 32) << 16);
      epilogue_warp<ACT, E16>(p, taddr_row, row0, n_blk * kBnT, bn, half, epi_stage + ew * kEpiStageFloats);
      // all of this warp's TMEM reads of accumulator `acc` are complete -> hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  }
  return fn;
}

// 2-D map over a row-major 16-bit matrix [rows, cols] with row stride ld (elements); box = 64 x box_rows,
// 128-byte swizzle (must match umma_desc_k_sw128).
mmr_status make_tmap_2d(void* map_out, const void* base, int64_t rows, int64_t cols, int64_t ld,
                        int box_rows, int dtype) {
  CUtensorMap* map = static_cast<CUtensorMap*>(map_out);
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return fail(MMR_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const cuuint64_t gdim[2] = {cuuint64_t(cols), cuuint64_t(rows)};
  const cuuint64_t gstride[1] = {cuuint64_t(ld) * 2};
  const cuuint32_t box[2] = {cuuint32_t(kBK), cuuint32_t(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt =
      dtype == MMR_DT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = fn(map, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MMR_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", int(r));
  return MMR_OK;
}

// General 2-D map: element kind 0 = fp16, 1 = bf16, 2 = fp32; box = box_cols x box_rows elements; swizzle_bytes is
// 128, 64 or 0 and must equal the box's row size in bytes when non-zero (the epilogue staging layouts rely on it).
mmr_status make_tmap_ex(void* map_out, const void* base, int64_t rows, int64_t cols, int64_t ld, int elem_kind,
                        int box_cols, int box_rows, int swizzle_bytes) {
  CUtensorMap* map = static_cast<CUtensorMap*>(map_out);
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return fail(MMR_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const int esz = elem_kind == 2 ? 4 : 2;
  const cuuint64_t gdim[2] = {cuuint64_t(cols), cuuint64_t(rows)};
  const cuuint64_t gstride[1] = {cuuint64_t(ld) * esz};
  const cuuint32_t box[2] = {cuuint32_t(box_cols), cuuint32_t(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = elem_kind == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : elem_kind == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                  : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  const CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                      : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(map, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MMR_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", int(r));
  return MMR_OK;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

template <int ACT, class E16, class CFG>
static mmr_status launch_gemm(const CUtensorMap& ta, const CUtensorMap& tw, const GemmParams& p, int grid,
                              cudaStream_t stream) {
  auto kern = gemm_tcgen05_kernel<ACT, E16, CFG>;
  static bool configured = false;
  if (!configured) {
    MMR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(CFG::kSmemBytes)));
    configured = true;
  }
  MMR_CUDA_OK(launch_pdl(kern, dim3(grid), dim3(kThreads), CFG::kSmemBytes, stream, ta, tw, p));
  return MMR_OK;
}

template <class E16, class CFG>
static mmr_status dispatch_act(int act, const CUtensorMap& ta, const CUtensorMap& tw, const GemmParams& p,
                               int grid, cudaStream_t s) {
  switch (act) {
    case MMR_ACT_NONE: return launch_gemm<MMR_ACT_NONE, E16, CFG>(ta, tw, p, grid, s);
    case MMR_ACT_RELU: return launch_gemm<MMR_ACT_RELU, E16, CFG>(ta, tw, p, grid, s);
    case MMR_ACT_GELU_TANH: return launch_gemm<MMR_ACT_GELU_TANH, E16, CFG>(ta, tw, p, grid, s);
    case MMR_ACT_GELU_ERF: return launch_gemm<MMR_ACT_GELU_ERF, E16, CFG>(ta, tw, p, grid, s);
    case MMR_ACT_TANH: return launch_gemm<MMR_ACT_TANH, E16, CFG>(ta, tw, p, grid, s);
    default: return fail(MMR_ERR_INVALID, "mmr_gemm: unknown activation %d", act);
  }
}

mmr_status gemm(const void* A16, int64_t lda, const void* W16, int64_t ldw, int M, int N, int K,
                const float* bias, const float* residual, int64_t ldr, void* out16, int64_t ldo16,
                float* out32, int64_t ldo32, int act, int dtype, cudaStream_t stream, bool single_cta_only) {
  MMR_TRY(require_sm100());
  MMR_REQUIRE(A16 && W16, "mmr_gemm: null operand");
  MMR_REQUIRE(out16 || out32, "mmr_gemm: no output given");
  MMR_REQUIRE(M > 0 && N > 0 && K > 0, "mmr_gemm: empty problem M=%d N=%d K=%d", M, N, K);
  MMR_REQUIRE(K % kBK == 0, "mmr_gemm: K=%d must be a multiple of %d", K, kBK);
  MMR_REQUIRE(N % 16 == 0, "mmr_gemm: N=%d must be a multiple of 16", N);
  MMR_REQUIRE(lda % 8 == 0 && ldw % 8 == 0, "mmr_gemm: operand row strides must be multiples of 8 elements");
  MMR_REQUIRE((reinterpret_cast<uintptr_t>(A16) & 15) == 0 && (reinterpret_cast<uintptr_t>(W16) & 15) == 0,
              "mmr_gemm: operands must be 16-byte aligned");
  MMR_REQUIRE(!out16 || (ldo16 % 8 == 0 && (reinterpret_cast<uintptr_t>(out16) & 15) == 0),
              "mmr_gemm: out16 must be 16-byte aligned with ld %% 8 == 0");
  MMR_REQUIRE(!out32 || (ldo32 % 4 == 0 && (reinterpret_cast<uintptr_t>(out32) & 15) == 0),
              "mmr_gemm: out32 must be 16-byte aligned with ld %% 4 == 0");
  MMR_REQUIRE(!residual || (ldr % 4 == 0 && (reinterpret_cast<uintptr_t>(residual) & 15) == 0),
              "mmr_gemm: residual must be 16-byte aligned with ld %% 4 == 0");
  MMR_REQUIRE(!bias || (reinterpret_cast<uintptr_t>(bias) & 15) == 0, "mmr_gemm: bias must be 16-byte aligned");
  MMR_REQUIRE(dtype == MMR_DT_BF16 || dtype == MMR_DT_FP16, "mmr_gemm: bad dtype %d", dtype);

  const bool pair_enabled = tuning(MMR_TUNE_GEMM_PAIR) != 0 && !single_cta_only;
  if (pair_enabled && gemm_pair16_eligible(M, N, K, residual, out16, out32))
    return gemm_pair16(A16, lda, W16, ldw, M, N, K, bias, out16, ldo16, act, dtype, stream);
  if (pair_enabled && gemm_pair_eligible(M, N, K)) {
    GemmParams pp{M, N, K, bias, residual, ldr, out16, ldo16, out32, ldo32, uint32_t(dtype), uint32_t(kBN / 2)};
    return gemm_pair(A16, lda, W16, ldw, pp, act, dtype, stream);
  }
  CUtensorMap ta, tw;
  MMR_TRY(make_tmap_2d(&ta, A16, M, K, lda, kBM, dtype));
  // few rows (the [CLS] tail): 64-column tiles, 8-stage ring; else 256-column tiles
  const bool narrow = single_cta_only && M <= 512 && N >= 64;
  const int bn_t = narrow ? SingleNarrow::kBnT : kBN;
  // Box rows for W: a full tile-wide box when N allows it, else exactly N rows.
  const int w_box = N >= bn_t ? bn_t : N;
  MMR_TRY(make_tmap_2d(&tw, W16, N, K, ldw, w_box, dtype));
  GemmParams p{M, N, K, bias, residual, ldr, out16, ldo16, out32, ldo32, uint32_t(dtype), uint32_t(w_box)};
  const int tiles = ((M + kBM - 1) / kBM) * ((N + bn_t - 1) / bn_t);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  if (narrow) {
    if (dtype == MMR_DT_BF16) return dispatch_act<BF16, SingleNarrow>(act, ta, tw, p, grid, stream);
    return dispatch_act<FP16, SingleNarrow>(act, ta, tw, p, grid, stream);
  }
  if (dtype == MMR_DT_BF16) return dispatch_act<BF16, SingleWide>(act, ta, tw, p, grid, stream);
  return dispatch_act<FP16, SingleWide>(act, ta, tw, p, grid, stream);
}

}  // namespace mmr

extern "C" mmr_status mmr_gemm(const void* A16, int64_t lda, const void* W16, int64_t ldw, int M, int N, int K,
                               const float* bias, const float* residual, int64_t ldr, void* out16,
                               int64_t ldo16, float* out32, int64_t ldo32, int act, int dtype, void* stream) {
  return mmr::gemm(A16, lda, W16, ldw, M, N, K, bias, residual, ldr, out16, ldo16, out32, ldo32, act, dtype,
                   static_cast<cudaStream_t>(stream), false);
}

// CTA-pair (cta_group::2) tcgen05 GEMM for sm_100a:  out = act(A[M,K] · W[N,K]^T + bias) (+ residual)
//
// Same contract and epilogue as gemm_sm100.cu, for the large projections of the encoder (N a multiple of 256):
// two CTAs on the two SMs of a TPC form a cluster and compute one 256 x 256 output tile with UMMA M = 256.
// Each CTA stages only ITS half of both operands per K step (A rows [128r, 128r+128), W rows [128r, 128r+128) of
// the tile: 32 KB instead of the 48 KB a 128 x 256 single-CTA tile needs), the tensor core reads the W halves of
// both SMs, and each CTA's TMEM receives its 128 accumulator rows.  Halving the shared-memory fill per FLOP is what
// lets the MMA pipe run near peak: a single CTA at 128 x 256 x 64 needs 96 B/clk of TMA fill on top of 96 B/clk of
// operand reads, more than one SM's shared memory delivers.
//
//   warps 0..7  epilogue in both CTAs on their own 128 rows; "accumulator drained" arrives on the leader's barrier
//   warp 8      TMA producer (both CTAs): waits the local "empty" barrier, loads its halves, credits the bytes to
//               the LEADER's "full" barrier (cta_group::2 TMA); 6-stage ring
//   warp 9      leader only: single-thread tcgen05.mma.cta_group::2 issuer; commits are multicast to both CTAs
// (the two latency-critical roles sit at the highest warp ids: the SM's arbiter favours them, measured +3 %)
#include <cuda.h>

#include "gemm_common.cuh"

namespace mmr {

constexpr int kPairBM = 256;          // tile rows per CTA pair
constexpr int kHalfM = 128;           // rows per CTA (= TMEM lanes)
constexpr int kHalfN = kBN / 2;       // W rows staged per CTA
constexpr int kPairStages = 6;
constexpr uint32_t kPairABytes = kHalfM * kBK * 2;   // 16 KB
constexpr uint32_t kPairBBytes = kHalfN * kBK * 2;   // 16 KB
constexpr uint32_t kPairStageBytes = kPairABytes + kPairBBytes;
constexpr size_t kPairSmemBytes = 1024 + size_t(kPairStages) * kPairStageBytes + 256 + kEpiSmemBytes;

template <int ACT, class E16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                 const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + size_t(kPairStages) * kPairABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + size_t(kPairStages) * kPairStageBytes);
  uint64_t* full_bar = bars;                      // [stages]  TMA (both CTAs) -> MMA; used in the leader only
  uint64_t* empty_bar = bars + kPairStages;       // [stages]  MMA -> TMA, multicast to both CTAs
  uint64_t* tfull_bar = bars + 2 * kPairStages;   // [2]       MMA -> epilogue, multicast to both CTAs
  uint64_t* tempty_bar = tfull_bar + 2;           // [2]       epilogue (both CTAs) -> MMA; used in the leader only
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* epi_stage = reinterpret_cast<float*>(smem + size_t(kPairStages) * kPairStageBytes + 256);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();        // 0 = leader
  const int pair = blockIdx.x >> 1;
  const int n_pairs = gridDim.x >> 1;

  const int m_tiles = (p.M + kPairBM - 1) / kPairBM;
  const int n_tiles = p.N / kBN;
  const int k_blocks = p.K / kBK;
  const int total_tiles = m_tiles * n_tiles;

  if (warp == kEpiWarps && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_w);
    for (int s = 0; s < kPairStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 2 * kEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == kEpiWarps + 1) {
    tmem_alloc_2sm(tmem_slot, kTmemCols);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer's barriers are initialised before any remote arrive / TMA credit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                 // the previous kernel's outputs (this one's operands / residual) are complete
  pdl_launch_dependents();    // the next kernel may be scheduled as soon as SMs free up

  if (warp == kEpiWarps) {
    // ===================== TMA producer (both CTAs); warps 8 / 9: the arbiter favours higher warp ids =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < total_tiles; tile += n_pairs) {
        const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
        const int a_row = m_blk * kPairBM + int(rank) * kHalfM;
        const int w_row = n_blk * kBN + int(rank) * kHalfN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[stage]), 0);
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * kPairStageBytes);
          tma_load_2d_2sm(smem_a + size_t(stage) * kPairABytes, &tmap_a, full_leader, kb * kBK, a_row);
          tma_load_2d_2sm(smem_b + size_t(stage) * kPairBBytes, &tmap_w, full_leader, kb * kBK, w_row);
          if (++stage == kPairStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == kEpiWarps + 1) {
    // ===================== MMA issuer (leader CTA, one thread) =====================
    if (rank == 0 && lane == 0) {
      const uint32_t idesc = umma_idesc_f16(p.idesc_fmt, kPairBM, kBN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = pair; tile < total_tiles; tile += n_pairs, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1u;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + uint32_t(acc) * kBN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t a_desc = umma_desc_k_sw128(smem_u32(smem_a + size_t(stage) * kPairABytes));
          const uint64_t b_desc = umma_desc_k_sw128(smem_u32(smem_b + size_t(stage) * kPairBBytes));
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k) {
            umma_f16_2sm(tmem_d, a_desc + uint64_t(2 * k), b_desc + uint64_t(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit_2sm_mc(&empty_bar[stage], 0b11);   // frees this stage in BOTH CTAs once the MMAs retire
          if (++stage == kPairStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit_2sm_mc(&tfull_bar[acc], 0b11);       // accumulator complete -> both epilogues
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs, own 128 rows) =====================
    const int ew = warp;
    const int quarter = warp & 3;
    const int half = ew >> 2;
    int it = 0;
    for (int tile = pair; tile < total_tiles; tile += n_pairs, ++it) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int row0 = m_blk * kPairBM + int(rank) * kHalfM + quarter * 32;
      const uint32_t taddr_row = tmem_base + uint32_t(acc) * kBN + (uint32_t(quarter * 32) << 16);
      epilogue_warp<ACT, E16>(p, taddr_row, row0, n_blk * kBN, kBN, half, epi_stage + ew * kEpiStageFloats);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0));
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // no CTA leaves while its peer may still read its smem or signal its barriers
  if (warp == kEpiWarps + 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, kTmemCols);
  }
}

template <int ACT, class E16>
static mmr_status launch_pair(const CUtensorMap& ta, const CUtensorMap& tw, const GemmParams& p, int grid,
                              cudaStream_t stream) {
  auto kern = gemm_pair_kernel<ACT, E16>;
  static bool configured = false;
  if (!configured) {
    MMR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kPairSmemBytes)));
    configured = true;
  }
  MMR_CUDA_OK(launch_pdl(kern, dim3(grid), dim3(kGemmThreads), kPairSmemBytes, stream, ta, tw, p));
  return MMR_OK;
}

template <class E16>
static mmr_status dispatch_pair(int act, const CUtensorMap& ta, const CUtensorMap& tw, const GemmParams& p, int grid,
                                cudaStream_t s) {
  switch (act) {
    case MMR_ACT_NONE: return launch_pair<MMR_ACT_NONE, E16>(ta, tw, p, grid, s);
    case MMR_ACT_RELU: return launch_pair<MMR_ACT_RELU, E16>(ta, tw, p, grid, s);
    case MMR_ACT_GELU_TANH: return launch_pair<MMR_ACT_GELU_TANH, E16>(ta, tw, p, grid, s);
    case MMR_ACT_GELU_ERF: return launch_pair<MMR_ACT_GELU_ERF, E16>(ta, tw, p, grid, s);
    case MMR_ACT_TANH: return launch_pair<MMR_ACT_TANH, E16>(ta, tw, p, grid, s);
    default: return fail(MMR_ERR_INVALID, "mmr_gemm: unknown activation %d", act);
  }
}

bool gemm_pair_eligible(int M, int N, int K) { return N % kBN == 0 && M > kHalfM && K % kBK == 0; }

// Arguments are already validated by mmr::gemm.
mmr_status gemm_pair(const void* A16, int64_t lda, const void* W16, int64_t ldw, const GemmParams& p, int act,
                     int dtype, cudaStream_t stream) {
  CUtensorMap ta, tw;
  MMR_TRY(make_tmap_2d(&ta, A16, p.M, p.K, lda, kHalfM, dtype));
  MMR_TRY(make_tmap_2d(&tw, W16, p.N, p.K, ldw, kHalfN, dtype));
  const int tiles = ((p.M + kPairBM - 1) / kPairBM) * (p.N / kBN);
  const int max_pairs = sm_count() / 2;
  const int grid = 2 * (tiles < max_pairs ? tiles : max_pairs);
  if (dtype == MMR_DT_BF16) return dispatch_pair<BF16>(act, ta, tw, p, grid, stream);
  return dispatch_pair<FP16>(act, ta, tw, p, grid, stream);
}

}  // namespace mmr

// Main loop shared by the CTA-pair (cta_group::2) GEMM kernels of this directory: the TMA producer and the
// single-thread tcgen05.mma issuer of one 256 x BN x K output tile, plus the shared-memory carve-up they agree on.
//
// Two CTAs on the two SMs of a TPC compute one 256-row tile with UMMA M = 256.  Each CTA stages only ITS half of
// both operands per 64-wide K step (A rows [128 r, 128 r + 128), W rows [bn/2 r, bn/2 r + bn/2) of the tile); the
// transaction bytes of both CTAs are credited to the LEADER's "full" barrier, the leader's MMA thread issues for
// the pair and its commits are multicast to both CTAs.
//
// Optionally TWO pairs form a 4-CTA cluster and share the W tile (CP = 2): the pairs work on adjacent 256-row blocks
// of the same column tile, every CTA fetches only a QUARTER of the W tile and TMA-multicasts it to its twin in the
// other pair (same rank within the pair), so the cluster reads 96 KB instead of 128 KB from L2 per K step.  The
// price is lock-step: a stage is refilled only after BOTH pairs' MMAs have consumed it (each "empty" barrier counts
// two commits, multicast to all four CTAs).  The ncu captures in profiles/ show these GEMMs bound by L2 -> SM
// operand traffic (8.3-9.8 TB/s at every shape), which is what this trades SMs for (33 clusters = 132 of 148 SMs).
#pragma once
#include "gemm_common.cuh"

namespace mmr {

constexpr int kPairRows = 256;                       // tile rows per CTA pair
constexpr int kCtaRows = 128;                        // rows per CTA (= TMEM lanes)
constexpr uint32_t kOpABytes = kCtaRows * kBK * 2;   // 16 KB: this CTA's half of the A tile
constexpr uint32_t kOpBBytes = (kBN / 2) * kBK * 2;  // 16 KB: this CTA's half of the W tile (full-width tiles)

template <int STAGES>
struct PairRing {
  uint8_t* a;            // [STAGES][16 KB]
  uint8_t* b;            // [STAGES][16 KB]
  uint64_t* full;        // [STAGES]  TMA (both CTAs) -> MMA; used in the leader only
  uint64_t* empty;       // [STAGES]  MMA -> TMA, multicast to both CTAs
  uint64_t* tfull;       // [2]       MMA -> epilogue, multicast to both CTAs
  uint64_t* tempty;      // [2]       epilogue (both CTAs) -> MMA; used in the leader only
  static constexpr size_t kOperandBytes = size_t(STAGES) * (kOpABytes + kOpBBytes);
  static constexpr int kNumBars = 2 * STAGES + 4;

  __device__ __forceinline__ void carve(uint8_t* smem_1024, uint64_t* bars) {
    a = smem_1024;
    b = smem_1024 + size_t(STAGES) * kOpABytes;
    full = bars;
    empty = bars + STAGES;
    tfull = bars + 2 * STAGES;
    tempty = tfull + 2;
  }
  // one thread, before the cluster-wide sync that precedes any remote arrive; cluster_pairs = MMA threads whose
  // commits free a stage (1, or 2 when two pairs share the W tile)
  __device__ __forceinline__ void init(uint32_t epilogue_arrivals, uint32_t cluster_pairs = 1) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], cluster_pairs);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], epilogue_arrivals);
    }
  }
};

struct RingPos {
  int stage = 0;
  uint32_t phase = 0;
  template <int STAGES>
  __device__ __forceinline__ void advance() {
    if (++stage == STAGES) {
      stage = 0;
      phase ^= 1u;
    }
  }
};

// TMA load with cta_group::2 semantics, multicast to the CTAs of `mask` (same smem offset in each); the transaction
// bytes are credited, in every destination CTA, to the barrier at `mbar_addr`'s offset in the EVEN CTA of that
// destination's pair (mbar_addr = own shared::cta address with the pair-rank bit cleared).
__device__ __forceinline__ void tma_load_2d_2sm_mc(void* smem_dst, const void* tmap, uint32_t mbar_addr, uint16_t mask,
                                                   int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1, {%4, %5}], [%2], %3;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(mbar_addr), "h"(mask), "r"(c0), "r"(c1)
      : "memory");
}
constexpr uint32_t kPairRankBitMask = 0xFEFFFFFFu;   // shared::cluster address bit that selects the odd CTA of a pair

// Producer side of one tile (one thread of each CTA).  a_row / w_row are THIS CTA's first rows; w_rows is the number
// of W rows this CTA's smem holds per K step (bn / 2).  CP = 1: the CTA loads them all (`tmap_w` box = w_rows rows).
// CP = 2: the CTA of pair q loads rows [w_row + q w_rows/2, + w_rows/2) (`tmap_w` box = w_rows / 2 rows) and
// multicasts them to itself and its twin (cluster ranks r and r + 2).
template <int STAGES, int CP>
__device__ __forceinline__ void pair_produce_tile(PairRing<STAGES>& ring, RingPos& pos, const CUtensorMap* tmap_a,
                                                  const CUtensorMap* tmap_w, int a_row, int w_row, int w_rows,
                                                  int k_blocks, uint32_t rank_in_pair, uint32_t leader_rank,
                                                  uint32_t pair_in_cluster) {
  const uint32_t stage_tx = 2u * (kOpABytes + uint32_t(w_rows) * kBK * 2u);   // both CTAs' bytes
  for (int kb = 0; kb < k_blocks; ++kb) {
    mbar_wait(&ring.empty[pos.stage], pos.phase ^ 1u);
    const uint32_t full_leader = mapa_u32(smem_u32(&ring.full[pos.stage]), leader_rank);
    if (rank_in_pair == 0) mbar_arrive_expect_tx(&ring.full[pos.stage], stage_tx);
    tma_load_2d_2sm(ring.a + size_t(pos.stage) * kOpABytes, tmap_a, full_leader, kb * kBK, a_row);
    if constexpr (CP == 1) {
      tma_load_2d_2sm(ring.b + size_t(pos.stage) * kOpBBytes, tmap_w, full_leader, kb * kBK, w_row);
    } else {
      const int part = w_rows / 2;   // rows fetched by this CTA
      tma_load_2d_2sm_mc(ring.b + size_t(pos.stage) * kOpBBytes + size_t(pair_in_cluster) * part * (kBK * 2), tmap_w,
                         smem_u32(&ring.full[pos.stage]) & kPairRankBitMask,
                         uint16_t(0b0101u << rank_in_pair), kb * kBK, w_row + int(pair_in_cluster) * part);
    }
    pos.advance<STAGES>();
  }
}

// MMA side of one tile (one thread of the leader CTA): waits for the accumulator to be drained, issues
// k_blocks x 4 UMMAs of 256 x bn x 16, frees each smem stage and finally publishes the accumulator.
template <int STAGES>
__device__ __forceinline__ void pair_mma_tile(PairRing<STAGES>& ring, RingPos& pos, uint32_t tmem_d, uint32_t idesc,
                                              int k_blocks, int acc, uint32_t acc_phase, uint16_t pair_mask,
                                              uint16_t stage_mask) {
  mbar_wait(&ring.tempty[acc], acc_phase ^ 1u);
  tc_fence_after();
  for (int kb = 0; kb < k_blocks; ++kb) {
    mbar_wait(&ring.full[pos.stage], pos.phase);
    tc_fence_after();
    const uint64_t a_desc = umma_desc_k_sw128(smem_u32(ring.a + size_t(pos.stage) * kOpABytes));
    const uint64_t b_desc = umma_desc_k_sw128(smem_u32(ring.b + size_t(pos.stage) * kOpBBytes));
#pragma unroll
    for (int k = 0; k < kBK / kUmmaK; ++k) {
      // +32 B along K inside the 128-byte swizzle atom = +2 in the descriptor's (address >> 4) field
      umma_f16_2sm(tmem_d, a_desc + uint64_t(2 * k), b_desc + uint64_t(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
    }
    umma_commit_2sm_mc(&ring.empty[pos.stage], stage_mask);  // frees this stage in every CTA that fills it
    pos.advance<STAGES>();
  }
  umma_commit_2sm_mc(&ring.tfull[acc], pair_mask);           // accumulator complete -> both epilogues
}

}  // namespace mmr

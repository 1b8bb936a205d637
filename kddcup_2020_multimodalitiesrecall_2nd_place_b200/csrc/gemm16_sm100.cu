// CTA-pair tcgen05 GEMM with a 16-bit output only:  out16 = act(A[M,K] · W[N,K]^T + bias)
//
// The two big "operand-producing" projections of every encoder block: the fused Q/K/V projection
// (pixelbert.py:767-788, modeling.py:325-337) and the FFN-in projection with its GELU (pixelbert.py:969-974,
// modeling.py:394-406).  Main loop: pair_pipeline.cuh.  What differs from gemm2_sm100.cu is the epilogue, which the
// first ncu capture (profiles/r01d_*) showed to be the bound of those launches (tensor pipe 46-50 % busy, epilogue
// warps stalled on bias loads, on the smem round trip of the transposition and on the membar of a .release arrive):
//
//   * every thread keeps the TMEM-native shape "one thread = one accumulator row" from tcgen05.ld to shared memory:
//     it packs its 32 columns to 16 bit and writes them into a 128-byte-swizzled [32 rows x 64 columns] stage,
//     exactly the layout a SWIZZLE_128B tensor map expects (conflict-free 16-byte stores, no read-back);
//   * one lane hands the stage to the TMA engine (cp.async.bulk.tensor store): no per-thread global address math,
//     no predicates (rows past M are clipped by the tensor map), full 128-byte lines to L2, asynchronous;
//   * two stages per warp, recycled with cp.async.bulk.wait_group.read;
//   * the bias slice of the warp's 128 columns sits in warp-private shared memory, loaded while the MMAs still run;
//   * "accumulator drained" is a relaxed arrive (the tcgen05.wait::ld before it is the ordering that matters).
//
// Tail: tiles are 256 x 256; when the last wave would leave most pairs idle, its tiles are split along N into
// 2 or 4 sub-tiles (UMMA N = 128 / 64) so that the remainder spreads over the pairs.
#include <cuda.h>

#include <cstdlib>

#include "pair_pipeline.cuh"

namespace mmr {

// Warp roles: the SM's warp arbiter favours HIGHER warp ids, so the two single-thread roles whose latency the whole
// pipeline hangs on (TMA producer, MMA issuer) take the highest ids and the eight epilogue warps ids 0..7.
constexpr int kProducerWarp = kEpiWarps, kMmaWarp = kEpiWarps + 1;
constexpr int kP16WarpStage = 2 * 2048;             // two [32 rows x 64 B] store stages per epilogue warp
template <int STAGES>
constexpr size_t p16_smem_bytes() {
  return 1024 + PairRing<STAGES>::kOperandBytes + size_t(kEpiWarps) * kP16WarpStage + 256 /*barriers*/;
}

struct P16Params {
  int M, N, K;
  const float* bias;     // [N] or null
  const float* bias2;    // bias of the SECOND weight matrix (row blocks >= split_blk), or null
  int split_blk;         // first 256-row block that multiplies the second weight matrix (INT_MAX: one matrix)
  uint32_t idesc_fmt;    // 0 fp16 / 1 bf16
  int full_tiles;        // work items [0, full_tiles) are 256 columns wide; the rest are the split tail
  int tail_split_log2;   // 0, 1 or 2: tail items are (256 >> s) columns wide
  int total_items;       // full_tiles + (remaining items << s)
  unsigned long long* trace;   // debug: [cluster][32][4] ns stamps (entry, prologue done, per tile MMA / epilogue), or null
};
__device__ __forceinline__ unsigned long long p16_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// A work item is one column tile of one row GROUP: a group is 256 rows for a lone pair (CP = 1) and 512 rows for a
// 4-CTA cluster (CP = 2), whose two pairs take its upper and lower 256-row block.

struct TileCoord {
  int m_blk, col0, bn;
};
// Work item -> (row group, first column, width).  n_tiles = N / 256.
__device__ __forceinline__ TileCoord p16_tile(const P16Params& p, int tile, int n_tiles) {
  TileCoord t;
  if (tile < p.full_tiles) {
    t.m_blk = tile / n_tiles;
    t.col0 = (tile % n_tiles) * kBN;
    t.bn = kBN;
  } else {
    const int s = p.tail_split_log2;
    const int sub = tile - p.full_tiles;
    const int parent = p.full_tiles + (sub >> s);
    t.bn = kBN >> s;
    t.m_blk = parent / n_tiles;
    t.col0 = (parent % n_tiles) * kBN + (sub & ((1 << s) - 1)) * t.bn;
  }
  return t;
}

template <int ACT, class E16, int STAGES, int CP>
__global__ void __cluster_dims__(2 * CP, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_pair16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                   const __grid_constant__ CUtensorMap tmap_w_tail, const __grid_constant__ CUtensorMap tmap_w2,
                   const __grid_constant__ CUtensorMap tmap_w2_tail, const __grid_constant__ CUtensorMap tmap_o,
                   const P16Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* out_stage = smem + PairRing<STAGES>::kOperandBytes;                     // 1024-aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(out_stage + size_t(kEpiWarps) * kP16WarpStage);
  PairRing<STAGES> ring;
  ring.carve(smem, bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + PairRing<STAGES>::kNumBars);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const unsigned long long t_entry = p16_now();
  const uint32_t crank = cluster_ctarank();
  const uint32_t rank = crank & 1u;               // row half within the pair; 0 = the pair's MMA-issuing CTA
  const uint32_t cpair = crank >> 1;              // which pair of the cluster (always 0 when CP = 1)
  const uint32_t leader = crank & ~1u;            // cluster rank of this pair's leader
  const int pair = blockIdx.x / (2 * CP);         // work-item stride unit: the cluster
  const int n_pairs = gridDim.x / (2 * CP);
  const int n_tiles = p.N / kBN;
  const int k_blocks = p.K / kBK;
  const int total_tiles = p.total_items;

  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_w_tail);
    tma_prefetch_desc(&tmap_w2);
    tma_prefetch_desc(&tmap_w2_tail);
    tma_prefetch_desc(&tmap_o);
    ring.init(2 * kEpiWarps, CP);
    fence_mbar_init();
  }
  if (warp == kMmaWarp) {
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

  if (warp == kProducerWarp) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      RingPos pos;
      for (int tile = pair; tile < total_tiles; tile += n_pairs) {
        const TileCoord t = p16_tile(p, tile, n_tiles);
        const int w_rows = t.bn / 2;
        // two weight matrices over one row range (LXMERT: language rows, then visual rows): chosen by the row block
        const bool second = t.m_blk >= p.split_blk;
        pair_produce_tile<STAGES, CP>(ring, pos, &tmap_a,
                                      t.bn == kBN ? (second ? &tmap_w2 : &tmap_w) : (second ? &tmap_w2_tail : &tmap_w_tail),
                                      (t.m_blk * CP + int(cpair)) * kPairRows + int(rank) * kCtaRows,
                                      t.col0 + int(rank) * w_rows, w_rows, k_blocks, rank, leader, cpair);
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer (leader CTA, one thread) =====================
    if (rank == 0 && lane == 0) {
      RingPos pos;
      int it = 0;
      if (p.trace != nullptr && cpair == 0) {
        p.trace[size_t(pair) * 128 + 0] = t_entry;
        p.trace[size_t(pair) * 128 + 1] = p16_now();
      }
      for (int tile = pair; tile < total_tiles; tile += n_pairs, ++it) {
        const TileCoord t = p16_tile(p, tile, n_tiles);
        const int acc = it & 1;
        if (p.trace != nullptr && cpair == 0 && it < 30) p.trace[size_t(pair) * 128 + 4 + 4 * it] = p16_now();
        pair_mma_tile<STAGES>(ring, pos, tmem_base + uint32_t(acc) * kBN,
                                  umma_idesc_f16(p.idesc_fmt, kPairRows, uint32_t(t.bn)), k_blocks, acc,
                                  (it >> 1) & 1u, uint16_t(0b11u << leader), uint16_t((1u << (2 * CP)) - 1u));
        if (p.trace != nullptr && cpair == 0 && it < 30) p.trace[size_t(pair) * 128 + 4 + 4 * it + 1] = p16_now();
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs, own 128 rows) =====================
    const int ew = warp;                   // epilogue warps are warps 0..7
    const int quarter = warp & 3;          // TMEM lane quarter this warp may touch
    const int half = ew >> 2;              // which half of the tile's columns
    uint8_t* stage0 = out_stage + size_t(ew) * kP16WarpStage;
    const uint32_t tempty_leader0 = mapa_u32(smem_u32(&ring.tempty[0]), leader);
    const uint32_t tempty_leader1 = mapa_u32(smem_u32(&ring.tempty[1]), leader);
    const uint32_t sw64 = uint32_t((lane >> 1) & 3);
    uint32_t n_store = 0;
    int it = 0;
    for (int tile = pair; tile < total_tiles; tile += n_pairs, ++it) {
      const TileCoord t = p16_tile(p, tile, n_tiles);
      const int acc = it & 1;
      const float* bias = t.m_blk >= p.split_blk ? p.bias2 : p.bias;
      const int wcols = t.bn / 2;                       // columns owned by this warp: 128, 64 or 32
      const int wcol0 = t.col0 + half * wcols;
      const int n_chunks = wcols / 32;                  // 4, 2 or 1
      const int row0 = (t.m_blk * CP + int(cpair)) * kPairRows + int(rank) * kCtaRows + quarter * 32;
      const uint32_t taddr = tmem_base + uint32_t(acc) * kBN + (uint32_t(quarter * 32) << 16) + uint32_t(half * wcols);
      mbar_wait(&ring.tfull[acc], (it >> 1) & 1u);
      tc_fence_after();
      if (p.trace != nullptr && crank == 0 && ew == 0 && lane == 0 && it < 30) p.trace[size_t(pair) * 128 + 4 + 4 * it + 2] = p16_now();
#pragma unroll 1
      for (int c = 0; c < n_chunks; ++c) {
        uint8_t* buf = stage0 + ((n_store & 1u) ? 2048 : 0);
        // bias of this chunk: the same 128 bytes for every lane (L1 broadcast), in flight across the TMEM load
        float4 bb[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          bb[j] = bias != nullptr ? __ldg(reinterpret_cast<const float4*>(bias + wcol0 + c * 32 + 4 * j))
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t r[32];
        tmem_ld_32x32(taddr + uint32_t(c * 32), r);
        // this stage was last used two stores ago: its TMA store must have finished READING it
        if (lane == 0) bulk_wait_read<1>();
        tmem_ld_wait();
        if (c == n_chunks - 1) {
          // all of this warp's TMEM reads of accumulator `acc` are complete -> hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster_relaxed(acc ? tempty_leader1 : tempty_leader0);
        }
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float v0 = apply_act<ACT>(__uint_as_float(r[4 * j]) + bb[j].x);
          const float v1 = apply_act<ACT>(__uint_as_float(r[4 * j + 1]) + bb[j].y);
          const float v2 = apply_act<ACT>(__uint_as_float(r[4 * j + 2]) + bb[j].z);
          const float v3 = apply_act<ACT>(__uint_as_float(r[4 * j + 3]) + bb[j].w);
          pk[2 * j] = E16::pack(v0, v1);
          pk[2 * j + 1] = E16::pack(v2, v3);
        }
        __syncwarp();   // lane 0's wait on the stage precedes every lane's writes to it
        // 32 columns x 2 B = the four 16-byte units of this thread's 64-byte stage row (64-byte swizzle)
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(buf + lane * 64 + ((uint32_t(q) ^ sw64) << 4)) =
              make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(buf, &tmap_o, wcol0 + c * 32, row0);
          bulk_commit();
        }
        ++n_store;
      }
      if (p.trace != nullptr && crank == 0 && ew == 0 && lane == 0 && it < 30) p.trace[size_t(pair) * 128 + 4 + 4 * it + 3] = p16_now();
    }
    if (lane == 0) bulk_wait<0>();   // stores complete (and their smem reads with them) before the CTA retires
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // no CTA leaves while its peer may still read its smem or signal its barriers
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, kTmemCols);
  }
  if (p.trace != nullptr && crank == 0 && threadIdx.x == 0) p.trace[size_t(pair) * 128 + 2] = p16_now();
}

template <int ACT, class E16, int STAGES, int CP>
static mmr_status launch_p16s(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& twt,
                              const CUtensorMap& tw2, const CUtensorMap& twt2, const CUtensorMap& to,
                              const P16Params& p, int grid, cudaStream_t stream) {
  auto kern = gemm_pair16_kernel<ACT, E16, STAGES, CP>;
  static bool configured = false;
  if (!configured) {
    MMR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(p16_smem_bytes<STAGES>())));
    configured = true;
  }
  MMR_CUDA_OK(launch_pdl(kern, dim3(grid), dim3(kGemmThreads), p16_smem_bytes<STAGES>(), stream, ta, tw, twt, tw2, twt2, to, p));
  return MMR_OK;
}
constexpr int kP16Stages = 6;
template <int ACT, class E16>
static mmr_status launch_p16(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& twt, const CUtensorMap& tw2,
                             const CUtensorMap& twt2, const CUtensorMap& to, const P16Params& p, int grid,
                             int cluster_pairs, cudaStream_t stream) {
#ifdef MMR_EXPERIMENTAL
  if (cluster_pairs == 2) return launch_p16s<ACT, E16, kP16Stages, 2>(ta, tw, twt, tw2, twt2, to, p, grid, stream);
#endif
  (void)cluster_pairs;
  return launch_p16s<ACT, E16, kP16Stages, 1>(ta, tw, twt, tw2, twt2, to, p, grid, stream);
}

// Co-resident 4-CTA clusters of this kernel on the current device (0: none).  All instantiations share the launch
// shape, so one query serves them all.
static int p16_max_quads() {
#ifndef MMR_EXPERIMENTAL
  return 0;   // the 4-CTA multicast variant (measured slower: 33 clusters = 132 of 148 SMs) is a lab note
#else
  static int cached = -1;
  if (cached >= 0) return cached;
  auto kern = gemm_pair16_kernel<MMR_ACT_NONE, FP16, kP16Stages, 2>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(p16_smem_bytes<kP16Stages>())) !=
      cudaSuccess) {
    cudaGetLastError();
    return cached = 0;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(4 * 37, 1, 1);
  cfg.blockDim = dim3(kGemmThreads, 1, 1);
  cfg.dynamicSmemBytes = p16_smem_bytes<kP16Stages>();
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = 4;
  attr.val.clusterDim.y = 1;
  attr.val.clusterDim.z = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  return cached = n;
#endif
}

template <class E16>
static mmr_status dispatch_p16(int act, const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& twt,
                               const CUtensorMap& tw2, const CUtensorMap& twt2, const CUtensorMap& to,
                               const P16Params& p, int grid, int cp, cudaStream_t s) {
  switch (act) {
    case MMR_ACT_NONE: return launch_p16<MMR_ACT_NONE, E16>(ta, tw, twt, tw2, twt2, to, p, grid, cp, s);
    case MMR_ACT_RELU: return launch_p16<MMR_ACT_RELU, E16>(ta, tw, twt, tw2, twt2, to, p, grid, cp, s);
    case MMR_ACT_GELU_TANH: return launch_p16<MMR_ACT_GELU_TANH, E16>(ta, tw, twt, tw2, twt2, to, p, grid, cp, s);
    case MMR_ACT_GELU_ERF: return launch_p16<MMR_ACT_GELU_ERF, E16>(ta, tw, twt, tw2, twt2, to, p, grid, cp, s);
    case MMR_ACT_TANH: return launch_p16<MMR_ACT_TANH, E16>(ta, tw, twt, tw2, twt2, to, p, grid, cp, s);
    default: return fail(MMR_ERR_INVALID, "mmr_gemm: unknown activation %d", act);
  }
}

static unsigned long long* g_p16_trace = nullptr;
bool gemm_pair16_eligible(int M, int N, int K, const float* residual, const void* out16, const float* out32) {
  return tuning(MMR_TUNE_GEMM_P16) != 0 && N % kBN == 0 && M > kCtaRows && K % kBK == 0 && residual == nullptr && out32 == nullptr &&
         out16 != nullptr;
}

// Arguments are already validated by mmr::gemm.  W16b / biasb / split_row: optional SECOND weight matrix (same N, K,
// ldw) for the rows from split_row on (a multiple of 256) — LXMERT's language and visual streams, which share the
// activation buffers but not the weights, then run their projection as ONE launch with full waves.
mmr_status gemm_pair16_2w(const void* A16, int64_t lda, const void* W16, const void* W16b, int64_t ldw, int M, int N,
                          int K, const float* bias, const float* biasb, int split_row, void* out16, int64_t ldo16,
                          int act, int dtype, cudaStream_t stream) {
  const bool two = W16b != nullptr;
  const bool split_tail = tuning(MMR_TUNE_GEMM_TAIL) != 0;
  const int want_cp = two ? 1 : tuning(MMR_TUNE_GEMM_CLUSTER);   // a 4-CTA cluster spans two row blocks
  const int m_tiles = (M + kPairRows - 1) / kPairRows, n_tiles = N / kBN;
  // 4-CTA clusters (two pairs sharing the W tile) when the device places enough of them and there are at least two
  // row blocks; else lone pairs
  const int quads = want_cp == 2 ? p16_max_quads() : 0;
  const int cp = (quads >= 8 && m_tiles >= 2) ? 2 : 1;
  const int units = cp == 2 ? quads : sm_count() / 2;            // clusters that run concurrently
  const int items = ((m_tiles + cp - 1) / cp) * n_tiles;
  P16Params p{M, N, K, bias, two ? biasb : bias, two ? split_row / kPairRows : 0x7fffffff, uint32_t(dtype), items, 0,
              items, g_p16_trace};
  const int rem = items % units;
  if (split_tail && items > units && rem != 0) {
    // split the last partial wave so that it fills (at most) all clusters once
    int s = 0;
    while (s < 2 && (rem << (s + 1)) <= units) ++s;
    if (s > 0) {
      p.full_tiles = items - rem;
      p.tail_split_log2 = s;
      p.total_items = p.full_tiles + (rem << s);
    }
  }
  CUtensorMap ta, tw, twt, tw2, twt2, to;
  MMR_TRY(make_tmap_2d(&ta, A16, M, K, lda, kCtaRows, dtype));
  MMR_TRY(make_tmap_2d(&tw, W16, N, K, ldw, kBN / 2 / cp, dtype));
  MMR_TRY(make_tmap_2d(&twt, W16, N, K, ldw, (kBN >> p.tail_split_log2) / 2 / cp, dtype));
  MMR_TRY(make_tmap_2d(&tw2, two ? W16b : W16, N, K, ldw, kBN / 2 / cp, dtype));
  MMR_TRY(make_tmap_2d(&twt2, two ? W16b : W16, N, K, ldw, (kBN >> p.tail_split_log2) / 2 / cp, dtype));
  MMR_TRY(make_tmap_ex(&to, out16, M, N, ldo16, dtype == MMR_DT_BF16 ? 1 : 0, 32, 32, 64));
  const int grid = 2 * cp * (p.total_items < units ? p.total_items : units);
  if (dtype == MMR_DT_BF16) return dispatch_p16<BF16>(act, ta, tw, twt, tw2, twt2, to, p, grid, cp, stream);
  return dispatch_p16<FP16>(act, ta, tw, twt, tw2, twt2, to, p, grid, cp, stream);
}
mmr_status gemm_pair16(const void* A16, int64_t lda, const void* W16, int64_t ldw, int M, int N, int K,
                       const float* bias, void* out16, int64_t ldo16, int act, int dtype, cudaStream_t stream) {
  return gemm_pair16_2w(A16, lda, W16, nullptr, ldw, M, N, K, bias, nullptr, 0, out16, ldo16, act, dtype, stream);
}

}  // namespace mmr

/* Debug only (not in the public header): device buffer of [pairs][128] uint64 ns stamps, or null. */
extern "C" void mmr_debug_set_p16_trace(unsigned long long* dev_buf) { mmr::g_p16_trace = dev_buf; }

// Output projection + bias + residual + LayerNorm in ONE kernel:  x = LN(A[M,K] · W[768,K]^T + bias + x) * gamma + beta
//
// Replaces the "dense -> dropout(identity) -> layer_norm(out + input)" tails of the reference blocks
// (imagebert_zk/pixelbert.py:960-966 and 977-983; lxmert/src/lxrt/modeling.py:355-366 and 409-420), which the first
// version of this repo ran as a GEMM that wrote y (fp32) and a LayerNorm kernel that re-read it: 132 MB of avoidable
// traffic and one launch per LayerNorm, 24 times per 12-layer forward.
//
// A LayerNorm row spans N = 768 = 3 tiles of 256 fp32 accumulator columns, more than one SM's TMEM holds twice over,
// so three CTA pairs (a GROUP: pairs 3g, 3g+1, 3g+2 of the grid) share each 256-row block: pair n computes columns
// [256 n, 256 n + 256) with the pair_pipeline.cuh main loop, and the epilogues meet once per block:
//   pass 1    y = acc + bias + residual, parked back into the TMEM accumulator (tcgen05.st); shifted sums per thread
//             (one thread = one row of its warp's 128 columns); residual chunks arrive by TMA into a 2-slot ring
//   exchange  every warp publishes (mean_i, M2_i) of its 32 rows x 128 columns to a small global table and bumps the
//             counter of its (block, 32-row quarter); it then waits for the 6 partials of its rows (3 column tiles x
//             2 halves).  The wait is on warps that run the SAME step on neighbouring SMs at the same time; the MMAs
//             of the next tile proceed meanwhile in the other accumulator.
//   combine   Chan's parallel formula over the 6 partials -> mean, rstd
//   pass 2    normalise from TMEM, affine, write the fp32 rows and their 16-bit mirror into swizzled stages, TMA-store
// The groups are plain CTA pairs, not a 6-CTA cluster (the first version of this file): a 6-CTA cluster fits only 22
// times on the 148 SMs (GPC granularity) which turns 68 row blocks into 4 waves; 24 groups of pairs make it 3.
// Requirement: every CTA of the grid is resident at once (the grid never exceeds the co-residency the occupancy API
// reports, and the spin is bounded: a violation traps instead of hanging).
#include <cuda.h>

#include <cstdlib>

#include "kernels.cuh"
#include "pair_pipeline.cuh"

namespace mmr {

constexpr int kLnN = 768;
constexpr int kLnTiles = kLnN / kBN;            // 3 column tiles = 3 pairs per group
// the warp arbiter favours higher warp ids: the two latency-critical single-thread roles get ids 8 and 9
constexpr int kLnProducerWarp = kEpiWarps, kLnMmaWarp = kEpiWarps + 1;
// Shared-memory split between the operand ring and the epilogue, per K (chosen by A/B runs of the whole 12-layer
// forward, three alternations on one box):
//   K <= 1024 (out-projection): 4 operand stages, a 2-slot residual / fp32-staging ring + 1 16-bit stage per warp.
//     The launch is bound by its epilogue chain (tools/ln_trace.py), but with 3 stages its main loop is
//     load-latency-bound as well (7 us per tile instead of 4.4, measured in gemm_lnrow_sm100.cu), which delays the
//     first tile of every pair; 3 stages + 3 slots measured 0.6 % slower over the forward.
//   larger K (FFN-out): MMA-bound (23 us of MMAs per tile against a 12 us epilogue that hides behind them), so the
//     operand ring gets everything: 5 stages, ONE slot + 1 16-bit stage per warp (the residual chunks then arrive
//     one at a time, still inside the MMA time).  4 stages + 2 slots measured 1.6 % slower over the forward.
template <int STAGES, int SLOTS, int O16>
struct LnCfg {
  static constexpr int kStages = STAGES, kSlots = SLOTS, kO16 = O16;
  static constexpr int kWarpBytes = SLOTS * 4096 + O16 * 2048;
  // bulk store groups are committed per chunk as {16-bit}, {fp32}: this many of the most recent may still be reading
  // their buffers when the next chunk starts writing (slot reuse distance SLOTS, 16-bit stage reuse distance O16)
  static constexpr int kPending = (2 * SLOTS - 2) < (2 * O16 - 1) ? (2 * SLOTS - 2) : (2 * O16 - 1);
  static constexpr size_t kSmemBytes =
      1024 + PairRing<STAGES>::kOperandBytes + size_t(kEpiWarps) * kWarpBytes + 2 * 3 * kBN * 4 + 512;
};
using LnCfgShortK = LnCfg<4, 2, 1>;
using LnCfgLongK = LnCfg<5, 1, 1>;
constexpr int kLnSlots = 2 * kLnTiles;          // partial statistics per row: 3 column tiles x 2 halves

struct GemmLnParams {
  int M, K;
  const float* bias;      // [768]
  const float* gamma;     // [768]
  const float* beta;      // [768]
  // second parameter set, for the row blocks >= split_blk (LXMERT: language rows, then visual rows, one launch);
  // equal to the first when the launch has one weight matrix
  const float* bias2;
  const float* gamma2;
  const float* beta2;
  int split_blk;
  float eps;
  uint4* stats;           // [m_tiles][6][256]  {mean_i, tag, M2_i, tag}: each half is one 8-byte store carrying its flag
  uint32_t* epoch;        // [0] tag of this launch (read at kernel start), [1] CTAs finished; the last CTA bumps [0]
  uint32_t idesc_fmt;
  unsigned long long* trace;   // debug: per (CTA, epilogue warp, tile) phase timestamps in ns, or null
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define MMR_LN_STAMP(k)                                                                              \
  do {                                                                                               \
    if (p.trace != nullptr && lane == 0 && it < 4)                                                   \
      p.trace[((size_t(blockIdx.x) * kEpiWarps + ew) * 4 + it) * 8 + (k)] = globaltimer_ns();        \
  } while (0)

template <class E16, class CFG>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_ln_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
               const __grid_constant__ CUtensorMap tmap_w2, const __grid_constant__ CUtensorMap tmap_r,
               const __grid_constant__ CUtensorMap tmap_o32, const __grid_constant__ CUtensorMap tmap_o16,
               const GemmLnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kLnStages = CFG::kStages, kSlots = CFG::kSlots, kO16 = CFG::kO16;
  uint8_t* epi = smem + PairRing<kLnStages>::kOperandBytes;                       // 1024-aligned
  float* vec_s = reinterpret_cast<float*>(epi + size_t(kEpiWarps) * CFG::kWarpBytes);   // [2 sets][3][256]: bias, gamma, beta
  uint64_t* bars = reinterpret_cast<uint64_t*>(vec_s + 2 * 3 * kBN);
  PairRing<kLnStages> ring;
  ring.carve(smem, bars);
  uint64_t* res_bar = bars + PairRing<kLnStages>::kNumBars;    // [8 warps][kSlots] residual chunk landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + kSlots * kEpiWarps);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();        // which 128-row half of the block this CTA owns
  const int pair = blockIdx.x >> 1;
  const int n_groups = (gridDim.x >> 1) / kLnTiles;
  const int group = pair / kLnTiles;
  const int n_tile = pair % kLnTiles;             // which 256-column tile this pair owns
  const int m_tiles = (p.M + kPairRows - 1) / kPairRows;
  const int k_blocks = p.K / kBK;

  if (warp == kLnProducerWarp && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_w2);
    tma_prefetch_desc(&tmap_r);
    tma_prefetch_desc(&tmap_o32);
    tma_prefetch_desc(&tmap_o16);
    ring.init(2 * kEpiWarps);
    for (int i = 0; i < kSlots * kEpiWarps; ++i) mbar_init(&res_bar[i], 1);
    fence_mbar_init();
  }
  if (warp == kLnMmaWarp) {
    tmem_alloc_2sm(tmem_slot, kTmemCols);
    tmem_relinquish_2sm();
  }
  // this pair's column tile never changes: its bias / gamma / beta slices live in shared memory for the whole kernel
  for (int i = threadIdx.x; i < 2 * 3 * kBN; i += kGemmThreads) {
    const int set = i / (3 * kBN), j = i % (3 * kBN);
    const float* src = set == 0 ? (j < kBN ? p.bias : (j < 2 * kBN ? p.gamma : p.beta))
                                : (j < kBN ? p.bias2 : (j < 2 * kBN ? p.gamma2 : p.beta2));
    vec_s[i] = __ldg(src + n_tile * kBN + (j & (kBN - 1)));
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  const uint32_t tag_of_launch = *reinterpret_cast<const volatile uint32_t*>(p.epoch);                 // the previous kernel's outputs (this one's operands / residual) are complete
  pdl_launch_dependents();    // the next kernel may be scheduled as soon as SMs free up

  if (warp == kLnProducerWarp) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      RingPos pos;
      const int w_row = n_tile * kBN + int(rank) * (kBN / 2);
      for (int m_blk = group; m_blk < m_tiles; m_blk += n_groups)
        pair_produce_tile<kLnStages, 1>(ring, pos, &tmap_a, m_blk >= p.split_blk ? &tmap_w2 : &tmap_w,
                                        m_blk * kPairRows + int(rank) * kCtaRows, w_row, kBN / 2, k_blocks, rank, 0, 0);
    }
  } else if (warp == kLnMmaWarp) {
    // ===================== MMA issuer (pair leader, one thread) =====================
    if (rank == 0 && lane == 0) {
      RingPos pos;
      const uint32_t idesc = umma_idesc_f16(p.idesc_fmt, kPairRows, kBN);
      int it = 0;
      for (int m_blk = group; m_blk < m_tiles; m_blk += n_groups, ++it) {
        const int acc = it & 1;
        pair_mma_tile<kLnStages>(ring, pos, tmem_base + uint32_t(acc) * kBN, idesc, k_blocks, acc, (it >> 1) & 1u,
                                 0b11, 0b11);
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int ew = warp;                   // epilogue warps are warps 0..7
    const int quarter = warp & 3;
    const int half = ew >> 2;
    uint8_t* wbuf = epi + size_t(ew) * CFG::kWarpBytes;
    // wbuf + 4096 s               : fp32 slot s    [32 rows x 32 cols], 128-byte swizzle
    // wbuf + 4096 kSlots + 2048 t : 16-bit stage t [32 rows x 32 cols], 64-byte swizzle
    const float* vec_w = vec_s + half * 128;               // this warp's 128 columns of the three vectors, set 0
    uint64_t* rfull = res_bar + kSlots * ew;
    const uint32_t tag = tag_of_launch;   // this launch's flag value (never 0)
    const uint32_t tempty_leader0 = mapa_u32(smem_u32(&ring.tempty[0]), 0);
    const uint32_t tempty_leader1 = mapa_u32(smem_u32(&ring.tempty[1]), 0);
    const int col_w = n_tile * kBN + half * 128;            // first of this warp's 128 columns
    const int row_in_blk = int(rank) * kCtaRows + quarter * 32;
    const uint32_t sw128 = uint32_t(lane & 7), sw64 = uint32_t((lane >> 1) & 3);
    uint32_t rph = 0;                                       // parity bits of the residual barriers

    // ---- pass 1 of this warp's `it`-th tile: y = acc + bias + residual -> back into TMEM; shifted sums (shift = this
    // thread's first y); publish (mean_i, M2_i) of the thread's 128 columns
    auto pass1 = [&](int it) {
      const int m_blk = group + it * n_groups;
      const int acc = it & 1;
      const float* bias_w = vec_w + (m_blk >= p.split_blk ? 3 * kBN : 0);
      const int row0 = m_blk * kPairRows + row_in_blk;
      const uint32_t taddr = tmem_base + uint32_t(acc) * kBN + (uint32_t(quarter * 32) << 16) + uint32_t(half * 128);
      MMR_LN_STAMP(0);
      // the first residual chunks -> slots (the previous block's stores must have finished reading them)
      if (lane == 0) {
        bulk_wait_read<0>();
#pragma unroll
        for (int s = 0; s < kSlots; ++s) {
          mbar_arrive_expect_tx(&rfull[s], 4096);
          tma_load_2d(wbuf + 4096 * s, &tmap_r, &rfull[s], col_w + 32 * s, row0);
        }
      }
      mbar_wait(&ring.tfull[acc], (it >> 1) & 1u);
      tc_fence_after();
      MMR_LN_STAMP(1);
      float shift = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        const int s = c % kSlots;
        uint8_t* slot_s = wbuf + 4096 * s;
        mbar_wait(&rfull[s], (rph >> s) & 1u);
        rph ^= 1u << s;
        uint32_t v[32];
        tmem_ld_32x32(taddr + uint32_t(c * 32), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 x = *reinterpret_cast<const float4*>(slot_s + lane * 128 + ((uint32_t(j) ^ sw128) << 4));
          const float4 bb = *reinterpret_cast<const float4*>(bias_w + c * 32 + 4 * j);
          const float y0 = __uint_as_float(v[4 * j]) + bb.x + x.x, y1 = __uint_as_float(v[4 * j + 1]) + bb.y + x.y;
          const float y2 = __uint_as_float(v[4 * j + 2]) + bb.z + x.z, y3 = __uint_as_float(v[4 * j + 3]) + bb.w + x.w;
          if (c == 0 && j == 0) shift = y0;
          const float d0 = y0 - shift, d1 = y1 - shift, d2 = y2 - shift, d3 = y3 - shift;
          s1 += (d0 + d1) + (d2 + d3);
          s2 = fmaf(d0, d0, s2); s2 = fmaf(d1, d1, s2); s2 = fmaf(d2, d2, s2); s2 = fmaf(d3, d3, s2);
          v[4 * j] = __float_as_uint(y0); v[4 * j + 1] = __float_as_uint(y1);
          v[4 * j + 2] = __float_as_uint(y2); v[4 * j + 3] = __float_as_uint(y3);
        }
        tmem_st_32x32(taddr + uint32_t(c * 32), v);
        __syncwarp();   // every lane has read its slot row
        if (lane == 0 && c + kSlots < 4) {
          mbar_arrive_expect_tx(&rfull[s], 4096);
          tma_load_2d(slot_s, &tmap_r, &rfull[s], col_w + 32 * (c + kSlots), row0);
        }
      }
      tmem_st_wait();
      MMR_LN_STAMP(2);
      // Each half of an entry is ONE 8-byte store {value, tag}: a reader that sees this launch's tag sees the value
      // (no fence, no atomic, no counter).
      const float mean_i = shift + s1 * (1.0f / 128.0f);
      const float m2_i = fmaxf(s2 - s1 * s1 * (1.0f / 128.0f), 0.f);
      uint4* tab = p.stats + size_t(m_blk) * kLnSlots * kPairRows + row_in_blk + lane;
      uint8_t* mine = reinterpret_cast<uint8_t*>(tab + size_t(n_tile * 2 + half) * kPairRows);
      st_volatile_u32x2(mine, __float_as_uint(mean_i), tag);
      st_volatile_u32x2(mine + 8, __float_as_uint(m2_i), tag);
    };

    // ---- exchange + pass 2 of the `it`-th tile: wait for the 6 partials of this thread's row, Chan's formula,
    // normalise from TMEM, affine, swizzled stages, TMA stores
    auto pass2 = [&](int it) {
      const int m_blk = group + it * n_groups;
      const int acc = it & 1;
      const float* bias_w = vec_w + (m_blk >= p.split_blk ? 3 * kBN : 0);
      const float* gamma_w = bias_w + kBN;
      const float* beta_w = bias_w + 2 * kBN;
      const int row0 = m_blk * kPairRows + row_in_blk;
      const uint32_t taddr = tmem_base + uint32_t(acc) * kBN + (uint32_t(quarter * 32) << 16) + uint32_t(half * 128);
      float mean, rstd;
      {
        const uint4* tab = p.stats + size_t(m_blk) * kLnSlots * kPairRows + row_in_blk + lane;
        // all 12 words are requested together on every probe: the wait costs one L2 round trip after the last
        // partner's publish, not one per entry
        float means[kLnSlots], m2_tot = 0.f;
        mean = 0.f;
        uint2 a[kLnSlots], b[kLnSlots];
        uint32_t spins = 0;
        for (;;) {
          bool all_in = true;
#pragma unroll
          for (int s = 0; s < kLnSlots; ++s) {
            const uint8_t* e = reinterpret_cast<const uint8_t*>(tab + size_t(s) * kPairRows);
            a[s] = ld_volatile_u32x2(e);
            b[s] = ld_volatile_u32x2(e + 8);
          }
#pragma unroll
          for (int s = 0; s < kLnSlots; ++s) all_in = all_in && a[s].y == tag && b[s].y == tag;
          if (all_in) break;
          if (++spins > MMR_SPIN_LIMIT) __trap();
        }
#pragma unroll
        for (int s = 0; s < kLnSlots; ++s) {
          means[s] = __uint_as_float(a[s].x);
          mean += means[s];
          m2_tot += __uint_as_float(b[s].x);
        }
        mean *= (1.0f / kLnSlots);
#pragma unroll
        for (int s = 0; s < kLnSlots; ++s) {
          const float d = means[s] - mean;
          m2_tot = fmaf(128.0f * d, d, m2_tot);
        }
        rstd = rsqrtf(m2_tot * (1.0f / kLnN) + p.eps);
      }
      MMR_LN_STAMP(3);
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint8_t* slot_s = wbuf + 4096 * (c % kSlots);
        uint8_t* o16_s = wbuf + 4096 * kSlots + 2048 * (c % kO16);
        const int col0 = col_w + c * 32;
        if (lane == 0) bulk_wait_read<CFG::kPending>();   // this chunk's slot and 16-bit stage have been read
        __syncwarp();
        uint32_t v[32];
        tmem_ld_32x32(taddr + uint32_t(c * 32), v);
        tmem_ld_wait();
        if (c == 3) {
          // accumulator drained -> back to the pair leader's MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster_relaxed(acc ? tempty_leader1 : tempty_leader0);
        }
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 gg = *reinterpret_cast<const float4*>(gamma_w + c * 32 + 4 * j);
          const float4 be = *reinterpret_cast<const float4*>(beta_w + c * 32 + 4 * j);
          float4 y;
          y.x = (__uint_as_float(v[4 * j]) - mean) * rstd * gg.x + be.x;
          y.y = (__uint_as_float(v[4 * j + 1]) - mean) * rstd * gg.y + be.y;
          y.z = (__uint_as_float(v[4 * j + 2]) - mean) * rstd * gg.z + be.z;
          y.w = (__uint_as_float(v[4 * j + 3]) - mean) * rstd * gg.w + be.w;
          *reinterpret_cast<float4*>(slot_s + lane * 128 + ((uint32_t(j) ^ sw128) << 4)) = y;
          pk[2 * j] = E16::pack(y.x, y.y);
          pk[2 * j + 1] = E16::pack(y.z, y.w);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(o16_s + lane * 64 + ((uint32_t(q) ^ sw64) << 4)) =
              make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(o16_s, &tmap_o16, col0, row0);
          bulk_commit();
          tma_store_2d(slot_s, &tmap_o32, col0, row0);
          bulk_commit();
        }
      }
      MMR_LN_STAMP(4);
    };

    // (Taking the tiles two at a time -- pass 1 of both, then pass 2 of both, so that a tile's statistics travel while
    // the warp works on the other tile -- was built and measured in round 2: 45.8 vs 46.1 us at K = 768, nothing.  The
    // "exchange wait" of the timeline absorbs the skew of warps that share one bound, L2 -> SM bandwidth: this launch
    // moves ~290 MB through L2 at the ~9 TB/s every GEMM of the forward runs at.)
    const int my_tiles = group < m_tiles ? (m_tiles - group + n_groups - 1) / n_groups : 0;
    for (int it = 0; it < my_tiles; ++it) {
      pass1(it);
      pass2(it);
    }
    if (lane == 0) bulk_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == kLnMmaWarp) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, kTmemCols);
  }
  // The last CTA to finish moves the tag on for the next launch (nobody can still be polling with the old one).
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(p.epoch + 1, 1u) == gridDim.x - 1) {
      p.epoch[1] = 0;
      const uint32_t next = tag_of_launch + 1u;
      p.epoch[0] = next == 0u ? 1u : next;
      __threadfence();
    }
  }
}

// ------------------------------------------------------------------------------------------------ host side
// Exchange table + launch tag.  The table belongs to whoever launches: every mmr_handle carves its own out of its
// workspace arena at create time (model.cu), so two handles -- or graphs captured for them -- never share or outlive
// one; the standalone operator (mmr_gemm_layernorm) takes a stream-ordered scratch allocation per call.  The kernel
// still needs its whole grid co-resident: ONE fused GEMM+LN launch may run on a device at a time (mmr_forward
// serialises forwards that arrive on different streams of one device, see model.cu).
static unsigned long long* g_ln_trace = nullptr;

size_t gemm_ln_table_bytes(int M) {
  const size_t m_tiles = size_t((M + kPairRows - 1) / kPairRows);
  return 256 + m_tiles * kLnSlots * kPairRows * sizeof(uint4);
}
mmr_status gemm_ln_table_init(void* mem, int M, LnTable* out, cudaStream_t stream) {
  MMR_REQUIRE(mem != nullptr && out != nullptr && (reinterpret_cast<uintptr_t>(mem) & 255) == 0,
              "gemm_ln: exchange table memory must be 256-byte aligned");
  const size_t bytes = gemm_ln_table_bytes(M);
  MMR_CUDA_OK(cudaMemsetAsync(mem, 0, bytes, stream));            // tag 0 = "never written"
  MMR_CUDA_OK(cudaMemsetAsync(mem, 1, 1, stream));                // epoch[0] = 1 (little endian), epoch[1] = 0
  out->epoch = static_cast<uint32_t*>(mem);
  out->stats = static_cast<uint8_t*>(mem) + 256;
  out->m_tiles = (M + kPairRows - 1) / kPairRows;
  return MMR_OK;
}

// Largest number of co-resident CTA pairs of this kernel (0 when the device cannot place one): queried once.
template <class E16, class CFG>
static int ln_max_pairs() {
  static int cached = -1;
  if (cached >= 0) return cached;
  auto kern = gemm_ln_kernel<E16, CFG>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(CFG::kSmemBytes)) != cudaSuccess) {
    cudaGetLastError();
    return cached = 0;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * 74, 1, 1);
  cfg.blockDim = dim3(kGemmThreads, 1, 1);
  cfg.dynamicSmemBytes = CFG::kSmemBytes;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = 2;
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
}
template <class E16>
static int ln_max_pairs_for(int K) {
  return K <= 1024 ? ln_max_pairs<E16, LnCfgShortK>() : ln_max_pairs<E16, LnCfgLongK>();
}

bool gemm_ln_eligible(int M, int N, int K, int dtype) {
  if (tuning(MMR_TUNE_GEMM_LN) == 0 || N != kLnN || M <= kCtaRows || K % kBK != 0) return false;
  return (dtype == MMR_DT_BF16 ? ln_max_pairs_for<BF16>(K) : ln_max_pairs_for<FP16>(K)) >= kLnTiles;
}

template <class E16, class CFG>
static mmr_status launch_ln_cfg(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& tw2, const CUtensorMap& tr,
                                const CUtensorMap& to32, const CUtensorMap& to16, const GemmLnParams& p,
                                cudaStream_t stream) {
  const int m_tiles = (p.M + kPairRows - 1) / kPairRows;
  const int max_groups = ln_max_pairs<E16, CFG>() / kLnTiles;
  const int groups = m_tiles < max_groups ? m_tiles : max_groups;
  MMR_CUDA_OK(launch_pdl(gemm_ln_kernel<E16, CFG>, dim3(2 * kLnTiles * groups), dim3(kGemmThreads), CFG::kSmemBytes,
                         stream, ta, tw, tw2, tr, to32, to16, p));
  return MMR_OK;
}
template <class E16>
static mmr_status launch_ln(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& tw2, const CUtensorMap& tr,
                            const CUtensorMap& to32, const CUtensorMap& to16, const GemmLnParams& p, cudaStream_t stream) {
  if (p.K <= 1024) return launch_ln_cfg<E16, LnCfgShortK>(ta, tw, tw2, tr, to32, to16, p, stream);
  return launch_ln_cfg<E16, LnCfgLongK>(ta, tw, tw2, tr, to32, to16, p, stream);
}

// W16b / biasb / gammab / betab / split_row: optional SECOND parameter set for the rows from split_row on (a multiple
// of 256): LXMERT's two streams share the activation buffers, so their projection + LayerNorm tails run as one launch.
mmr_status gemm_ln_2w(const void* A16, int64_t lda, const void* W16, const void* W16b, int64_t ldw, int M, int K,
                      const float* bias, const float* biasb, const float* residual, int64_t ldr, const float* gamma,
                      const float* gammab, const float* beta, const float* betab, int split_row, float eps, void* out16,
                      int64_t ldo16, float* out32, int64_t ldo32, int dtype, const LnTable& table, cudaStream_t stream) {
  MMR_TRY(require_sm100());
  const bool two = W16b != nullptr;
  MMR_REQUIRE(A16 && W16 && bias && residual && gamma && beta && out16 && out32, "gemm_ln: null argument");
  MMR_REQUIRE(!two || (biasb && gammab && betab && split_row > 0 && split_row % kPairRows == 0),
              "gemm_ln: second parameter set incomplete or split row %d not a multiple of %d", split_row, kPairRows);
  MMR_REQUIRE(gemm_ln_eligible(M, kLnN, K, dtype), "gemm_ln: shape M=%d K=%d not eligible", M, K);
  MMR_REQUIRE(lda % 8 == 0 && ldw % 8 == 0 && ldr % 4 == 0 && ldo32 % 4 == 0 && ldo16 % 8 == 0,
              "gemm_ln: row strides break 16-byte alignment");
  MMR_REQUIRE(((reinterpret_cast<uintptr_t>(A16) | reinterpret_cast<uintptr_t>(W16) | reinterpret_cast<uintptr_t>(W16b) |
                reinterpret_cast<uintptr_t>(residual) | reinterpret_cast<uintptr_t>(out16) |
                reinterpret_cast<uintptr_t>(out32) | reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(gamma) |
                reinterpret_cast<uintptr_t>(beta) | reinterpret_cast<uintptr_t>(biasb) | reinterpret_cast<uintptr_t>(gammab) |
                reinterpret_cast<uintptr_t>(betab)) & 15) == 0,
              "gemm_ln: pointers must be 16-byte aligned");
#ifdef MMR_EXPERIMENTAL
  // row-owner decomposition (gemm_lnrow_sm100.cu): 2 = always, 3 = only for K <= 1024 (one weight matrix only)
  if (!two && (tuning(MMR_TUNE_GEMM_LN) == 2 || (tuning(MMR_TUNE_GEMM_LN) == 3 && K <= 1024)))
    return gemm_lnrow(A16, lda, W16, ldw, M, K, bias, residual, ldr, gamma, beta, eps, out16, ldo16, out32, ldo32, dtype,
                      stream);
#endif
  MMR_REQUIRE(table.stats != nullptr && table.epoch != nullptr && (M + kPairRows - 1) / kPairRows <= table.m_tiles,
              "gemm_ln: exchange table missing or too small for M=%d", M);
  const int ek = dtype == MMR_DT_BF16 ? 1 : 0;
  CUtensorMap ta, tw, tw2, tr, to32, to16;
  MMR_TRY(make_tmap_2d(&ta, A16, M, K, lda, kCtaRows, dtype));
  MMR_TRY(make_tmap_2d(&tw, W16, kLnN, K, ldw, kBN / 2, dtype));
  MMR_TRY(make_tmap_2d(&tw2, two ? W16b : W16, kLnN, K, ldw, kBN / 2, dtype));
  MMR_TRY(make_tmap_ex(&tr, residual, M, kLnN, ldr, 2, 32, 32, 128));
  MMR_TRY(make_tmap_ex(&to32, out32, M, kLnN, ldo32, 2, 32, 32, 128));
  MMR_TRY(make_tmap_ex(&to16, out16, M, kLnN, ldo16, ek, 32, 32, 64));
  GemmLnParams p{M, K, bias, gamma, beta, two ? biasb : bias, two ? gammab : gamma, two ? betab : beta,
                 two ? split_row / kPairRows : 0x7fffffff, eps, static_cast<uint4*>(table.stats), table.epoch,
                 uint32_t(dtype), g_ln_trace};
  if (dtype == MMR_DT_BF16) return launch_ln<BF16>(ta, tw, tw2, tr, to32, to16, p, stream);
  return launch_ln<FP16>(ta, tw, tw2, tr, to32, to16, p, stream);
}
mmr_status gemm_ln(const void* A16, int64_t lda, const void* W16, int64_t ldw, int M, int K, const float* bias,
                   const float* residual, int64_t ldr, const float* gamma, const float* beta, float eps, void* out16,
                   int64_t ldo16, float* out32, int64_t ldo32, int dtype, const LnTable& table, cudaStream_t stream) {
  return gemm_ln_2w(A16, lda, W16, nullptr, ldw, M, K, bias, nullptr, residual, ldr, gamma, nullptr, beta, nullptr, 0, eps,
                    out16, ldo16, out32, ldo32, dtype, table, stream);
}

}  // namespace mmr

extern "C" mmr_status mmr_gemm_layernorm(const void* A16, int64_t lda, const void* W16, int64_t ldw, int M, int K,
                                         const float* bias, const float* residual, int64_t ldr, const float* gamma,
                                         const float* beta, float eps, void* out16, int64_t ldo16, float* out32,
                                         int64_t ldo32, int dtype, void* stream) {
  // per-call scratch table, stream-ordered: nothing is shared with any handle or any other call
  using namespace mmr;
  MMR_TRY(require_sm100());
  MMR_REQUIRE(M > 0, "mmr_gemm_layernorm: M=%d", M);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  void* mem = nullptr;
  MMR_CUDA_OK(cudaMallocAsync(&mem, gemm_ln_table_bytes(M), st));
  LnTable table;
  mmr_status rc = gemm_ln_table_init(mem, M, &table, st);
  if (rc == MMR_OK)
    rc = gemm_ln(A16, lda, W16, ldw, M, K, bias, residual, ldr, gamma, beta, eps, out16, ldo16, out32, ldo32, dtype,
                 table, st);
  cudaFreeAsync(mem, st);
  return rc;
}
/* Debug only (not in the public header): device buffer of [grid][8 warps][4 tiles][8] uint64 phase stamps, or null. */
extern "C" void mmr_debug_set_ln_trace(unsigned long long* dev_buf) { mmr::g_ln_trace = dev_buf; }
extern "C" int mmr_gemm_layernorm_supported(int M, int K, int dtype) {
  return mmr::gemm_ln_eligible(M, 768, K, dtype) ? 1 : 0;
}

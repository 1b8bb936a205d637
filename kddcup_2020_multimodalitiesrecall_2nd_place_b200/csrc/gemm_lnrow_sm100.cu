// Output projection + bias + residual + LayerNorm in ONE kernel, row-owner form:
//     x = LN(A[M,K] · W[768,K]^T + bias + x) * gamma + beta
//
// Same contract as gemm_ln_sm100.cu (the reference tails pixelbert.py:960-966 / 977-983, modeling.py:355-366 /
// 409-420), different decomposition.  There, three CTA pairs shared a 256-row block (one 256-column tile each) and
// met through a table in global memory; every block cost pass 1 + a cross-SM exchange + pass 2 per pair, and the
// launch was bound by that 11 us epilogue chain three times over (tools/ln_trace.py).  Here ONE CTA pair owns a
// 256-row block and ALL 768 columns, computed as three consecutive 256-column tiles with the pair_pipeline.cuh
// main loop, so that
//   * a LayerNorm row never leaves its thread: the statistics of a row are two register accumulators per thread
//     (one thread = one row of 128 of each tile's 256 columns) and one shared-memory hand-off between the two
//     column halves — no global table, no tags, no spin, no co-residency requirement;
//   * the first two tiles' epilogues (pass 1: y = acc + bias + residual) hide entirely behind the next tile's MMAs;
//     only the last tile's pass 1 and the block's pass 2 are exposed;
//   * 68 row blocks (cfg2) are one wave on the 74 CTA pairs of the chip instead of three waves of 24 groups.
// Where y waits for the row statistics: tiles 1 and 2 are parked in their own TMEM accumulators (tcgen05.st, in
// place — the two accumulators are exactly the 512 TMEM columns); tile 0 cannot stay (its accumulator is needed for
// tile 2), so its y goes out to its final location in the fp32 output by TMA store and is re-read by TMA in pass 2
// (the same warp stores and re-loads its own chunks, ordered by cp.async.bulk.wait_group; 256 KB per block through
// L2 each way, +5 % of the launch's L2 traffic).
#include <cuda.h>

#include "kernels.cuh"
#include "pair_pipeline.cuh"

namespace mmr {

constexpr int kRowN = 768;
constexpr int kRowTiles = kRowN / kBN;           // 3 column tiles, all owned by the same pair
constexpr int kRowChunks = 4 * kRowTiles;        // 32-column chunks per warp and block
constexpr int kRowProducerWarp = kEpiWarps, kRowMmaWarp = kEpiWarps + 1;   // highest warp ids: favoured by the arbiter
constexpr int kRowCols = kRowN / 2;              // columns per thread: 128 of each of the 3 tiles

// Shared memory: STAGES operand stages (32 KB each) + per epilogue warp SLOTS fp32 chunk slots (4 KB: residual in,
// y / normalised rows out) and one 16-bit stage (2 KB) + the three 768-vectors + the statistics hand-off.
template <int STAGES, int SLOTS, int O16>
struct RowCfg {
  static_assert(SLOTS >= 1 && SLOTS <= 3 && O16 >= 1 && O16 <= 2 && !(SLOTS == 3 && O16 == 2), "unsupported ring");
  static constexpr int kStages = STAGES, kSlots = SLOTS, kO16 = O16;
  static constexpr int kWarpBytes = SLOTS * 4096 + O16 * 2048;
  static constexpr int kStatBytes = 2 * 2 * kCtaRows * 8;   // [block parity][column half][row] {mean, M2}
  static constexpr size_t kSmemBytes =
      1024 + PairRing<STAGES>::kOperandBytes + size_t(kEpiWarps) * kWarpBytes + 3 * kRowN * 4 + kStatBytes + 512;
  static_assert(kSmemBytes <= 232448, "exceeds the 227 KB of dynamic shared memory per CTA");
  // pass 2a (TMEM tiles) alternates between min(SLOTS, 2) slots and O16 16-bit stages; bulk store groups are
  // committed per chunk as {16-bit}, {fp32}: this many of the most recent may still be reading their buffers when the
  // next chunk starts writing
  static constexpr int kSlots2a = SLOTS < 2 ? SLOTS : 2;
  static constexpr int kPend2a = (2 * kSlots2a - 2) < (2 * O16 - 1) ? (2 * kSlots2a - 2) : (2 * O16 - 1);
  // pass 2b re-loads tile 0's chunk j into this slot; with 3 slots the re-load of chunk j + 1 may be issued while the
  // previous chunk's fp32 store still reads its slot (kPend2b = 1), with fewer slots everything older must be done
  __device__ static constexpr int reload_slot(int j) { return SLOTS == 3 ? (2 + j) % 3 : (SLOTS == 2 ? (j & 1) : 0); }
  static constexpr int kPend2b = SLOTS == 3 ? 1 : 0;
};
using RowCfg331 = RowCfg<3, 3, 1>;
using RowCfg421 = RowCfg<4, 2, 1>;
using RowCfg511 = RowCfg<5, 1, 1>;
using RowCfg412 = RowCfg<4, 1, 2>;
using RowCfg322 = RowCfg<3, 2, 2>;

struct GemmLnRowParams {
  int M, K;
  const float* bias;      // [768]
  const float* gamma;     // [768]
  const float* beta;      // [768]
  float eps;
  uint32_t idesc_fmt;
  unsigned long long* trace;   // debug: per (CTA, epilogue warp, block) phase timestamps in ns, or null
};

__device__ __forceinline__ unsigned long long row_globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define MMR_ROW_STAMP(k)                                                                               \
  do {                                                                                                 \
    if (p.trace != nullptr && lane == 0 && blk_iter < 2)                                               \
      p.trace[((size_t(blockIdx.x) * kEpiWarps + ew) * 2 + blk_iter) * 16 + (k)] = row_globaltimer_ns(); \
  } while (0)

__device__ __forceinline__ void epi_bar_sync() {   // the 8 epilogue warps only (named barrier 1)
  asm volatile("bar.sync 1, 256;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() {
  asm volatile("fence.proxy.async;" ::: "memory");
}

template <class E16, class CFG>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_lnrow_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                  const __grid_constant__ CUtensorMap tmap_r, const __grid_constant__ CUtensorMap tmap_o32,
                  const __grid_constant__ CUtensorMap tmap_o16, const GemmLnRowParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kStages = CFG::kStages, kSlots = CFG::kSlots, kO16 = CFG::kO16;
  uint8_t* epi = smem + PairRing<kStages>::kOperandBytes;                              // 1024-aligned
  float* vec_s = reinterpret_cast<float*>(epi + size_t(kEpiWarps) * CFG::kWarpBytes);  // [3][768]: bias, gamma, beta
  float2* stat_s = reinterpret_cast<float2*>(vec_s + 3 * kRowN);                       // [2][2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(stat_s) + CFG::kStatBytes);
  PairRing<kStages> ring;
  ring.carve(smem, bars);
  uint64_t* res_bar = bars + PairRing<kStages>::kNumBars;    // [8 warps][kSlots] fp32 chunk landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + kSlots * kEpiWarps);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();        // which 128-row half of the block this CTA owns
  const int pair = blockIdx.x >> 1;
  const int n_pairs = gridDim.x >> 1;
  const int m_tiles = (p.M + kPairRows - 1) / kPairRows;
  const int k_blocks = p.K / kBK;
  // Column tiles are visited in a per-pair rotation (pair p starts with tile p % 3), so that the 74 pairs, which run
  // in step, do not all pull the same 256 rows of W through the same L2 lines at the same moment.
  const int n_rot = pair % kRowTiles;

  if (warp == kRowProducerWarp && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_r);
    tma_prefetch_desc(&tmap_o32);
    tma_prefetch_desc(&tmap_o16);
    ring.init(2 * kEpiWarps);
    for (int i = 0; i < kSlots * kEpiWarps; ++i) mbar_init(&res_bar[i], 1);
    fence_mbar_init();
  }
  if (warp == kRowMmaWarp) {
    tmem_alloc_2sm(tmem_slot, kTmemCols);
    tmem_relinquish_2sm();
  }
  // weights only (not produced by the previous kernel of the stream): before the dependency wait
  for (int i = threadIdx.x; i < 3 * kRowN; i += kGemmThreads) {
    const float* src = i < kRowN ? p.bias : (i < 2 * kRowN ? p.gamma : p.beta);
    vec_s[i] = __ldg(src + (i % kRowN));
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch_dependents();

  if (warp == kRowProducerWarp) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      RingPos pos;
      for (int m_blk = pair; m_blk < m_tiles; m_blk += n_pairs)
        for (int n = 0; n < kRowTiles; ++n)
          pair_produce_tile<kStages, 1>(ring, pos, &tmap_a, &tmap_w, m_blk * kPairRows + int(rank) * kCtaRows,
                                        ((n + n_rot) % kRowTiles) * kBN + int(rank) * (kBN / 2), kBN / 2, k_blocks,
                                        rank, 0, 0);
    }
  } else if (warp == kRowMmaWarp) {
    // ===================== MMA issuer (pair leader, one thread) =====================
    if (rank == 0 && lane == 0) {
      RingPos pos;
      const uint32_t idesc = umma_idesc_f16(p.idesc_fmt, kPairRows, kBN);
      int it = 0;
      for (int m_blk = pair; m_blk < m_tiles; m_blk += n_pairs)
        for (int n = 0; n < kRowTiles; ++n, ++it) {
          const int acc = it & 1;
          pair_mma_tile<kStages>(ring, pos, tmem_base + uint32_t(acc) * kBN, idesc, k_blocks, acc, (it >> 1) & 1u,
                                 0b11, 0b11);
        }
    }
  } else {
    // ===================== epilogue warps =====================
    const int ew = warp;                   // epilogue warps are warps 0..7
    const int quarter = warp & 3;          // TMEM lane quarter
    const int half = ew >> 2;              // which 128 columns of every tile
    uint8_t* wbuf = epi + size_t(ew) * CFG::kWarpBytes;
    // wbuf + 4096 s        : fp32 slot s    [32 rows x 32 cols], 128-byte swizzle
    // wbuf + 4096 kSlots + 2048 t : 16-bit stage t [32 rows x 32 cols], 64-byte swizzle
    uint8_t* o16_base = wbuf + 4096 * kSlots;
    uint64_t* rfull = res_bar + kSlots * ew;
    const uint32_t tempty_leader0 = mapa_u32(smem_u32(&ring.tempty[0]), 0);
    const uint32_t tempty_leader1 = mapa_u32(smem_u32(&ring.tempty[1]), 0);
    const int row_in_cta = quarter * 32 + lane;
    const uint32_t sw128 = uint32_t(lane & 7), sw64 = uint32_t((lane >> 1) & 3);
    const uint32_t lane_taddr = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(half * 128);
    uint32_t rph = 0;                      // parity bits of the chunk barriers
    int it = 0;                            // tiles processed by this pair so far (same count as the MMA thread's)
    int blk_iter = 0;
    for (int m_blk = pair; m_blk < m_tiles; m_blk += n_pairs, ++blk_iter) {
      const int row0 = m_blk * kPairRows + int(rank) * kCtaRows + quarter * 32;   // global row of lane 0

      MMR_ROW_STAMP(0);
      // chunk g of the block = the (g / 4)-th tile visited, columns [256 tile + 128 half + 32 (g % 4), +32)
      auto chunk_col = [&](int g) { return (((g >> 2) + n_rot) % kRowTiles) * kBN + half * 128 + (g & 3) * 32; };
      if (lane == 0) {
        bulk_wait_read<0>();               // the previous block's stores have finished reading the slots
#pragma unroll
        for (int s = 0; s < kSlots; ++s) {
          mbar_arrive_expect_tx(&rfull[s], 4096);
          tma_load_2d(wbuf + 4096 * s, &tmap_r, &rfull[s], chunk_col(s), row0);
        }
      }

      // ---- pass 1 over the three tiles: y = acc + bias + residual; shifted sums (shift = this thread's first y)
      float shift = 0.f;
      float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f, s2c = 0.f, s2d = 0.f;
#pragma unroll 1
      for (int n = 0; n < kRowTiles; ++n, ++it) {
        const int acc = it & 1;
        const uint32_t taddr = lane_taddr + uint32_t(acc) * kBN;
        mbar_wait(&ring.tfull[acc], (it >> 1) & 1u);
        tc_fence_after();
        MMR_ROW_STAMP(1 + 2 * n);
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          const int g = n * 4 + c;
          const int s = g % kSlots;
          uint8_t* slot_s = wbuf + 4096 * s;
          uint32_t v[32];
          tmem_ld_32x32(taddr + uint32_t(c * 32), v);
          mbar_wait(&rfull[s], (rph >> s) & 1u);
          rph ^= 1u << s;
          tmem_ld_wait();
          const float* bias_c = vec_s + chunk_col(g);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4* xp = reinterpret_cast<float4*>(slot_s + lane * 128 + ((uint32_t(j) ^ sw128) << 4));
            const float4 x = *xp;
            const float4 bb = *reinterpret_cast<const float4*>(bias_c + 4 * j);
            const float y0 = __uint_as_float(v[4 * j]) + bb.x + x.x, y1 = __uint_as_float(v[4 * j + 1]) + bb.y + x.y;
            const float y2 = __uint_as_float(v[4 * j + 2]) + bb.z + x.z, y3 = __uint_as_float(v[4 * j + 3]) + bb.w + x.w;
            if (g == 0 && j == 0) shift = y0;
            const float d0 = y0 - shift, d1 = y1 - shift, d2 = y2 - shift, d3 = y3 - shift;
            s1a += d0 + d1;
            s1b += d2 + d3;
            s2a = fmaf(d0, d0, s2a); s2b = fmaf(d1, d1, s2b); s2c = fmaf(d2, d2, s2c); s2d = fmaf(d3, d3, s2d);
            if (n == 0) {
              *xp = make_float4(y0, y1, y2, y3);      // y replaces the residual in the slot: it leaves by TMA below
            } else {
              v[4 * j] = __float_as_uint(y0); v[4 * j + 1] = __float_as_uint(y1);
              v[4 * j + 2] = __float_as_uint(y2); v[4 * j + 3] = __float_as_uint(y3);
            }
          }
          if (n == 0) {
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(slot_s, &tmap_o32, chunk_col(g), row0);
              bulk_commit();
              if (g + kSlots < kRowChunks) {
                bulk_wait_read<0>();        // off the critical path: tile 1's MMAs take longer than this tile's pass 1
                mbar_arrive_expect_tx(&rfull[s], 4096);
                tma_load_2d(slot_s, &tmap_r, &rfull[s], chunk_col(g + kSlots), row0);
              }
            }
          } else {
            tmem_st_32x32(taddr + uint32_t(c * 32), v);   // parked in place until the row statistics are known
            __syncwarp();                                  // every lane has read its slot row
            if (lane == 0 && g + kSlots < kRowChunks) {
              mbar_arrive_expect_tx(&rfull[s], 4096);
              tma_load_2d(slot_s, &tmap_r, &rfull[s], chunk_col(g + kSlots), row0);
            }
          }
        }
        if (n == 0) {
          // tile 0's accumulator is drained -> back to the pair leader's MMA warp (it becomes tile 2's)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster_relaxed(acc ? tempty_leader1 : tempty_leader0);
        }
        MMR_ROW_STAMP(2 + 2 * n);
      }
      tmem_st_wait();

      // ---- row statistics: this thread's 384 columns + the other half's, through shared memory
      float mean, rstd;
      {
        const float s1 = s1a + s1b, s2 = (s2a + s2b) + (s2c + s2d);
        const float mean_i = shift + s1 * (1.0f / kRowCols);
        const float m2_i = fmaxf(s2 - s1 * s1 * (1.0f / kRowCols), 0.f);
        float2* st = stat_s + (blk_iter & 1) * (2 * kCtaRows);
        st[half * kCtaRows + row_in_cta] = make_float2(mean_i, m2_i);
        // tile 0's y must be complete in global memory before it is re-read (this lane issued those stores itself)
        if (lane == 0) {
          bulk_wait<0>();
          fence_proxy_async_all();
          constexpr int rs0 = CFG::reload_slot(0);
          if (kSlots == 3) {
            mbar_arrive_expect_tx(&rfull[rs0], 4096);
            tma_load_2d(wbuf + 4096 * rs0, &tmap_o32, &rfull[rs0], chunk_col(0), row0);
          }
        }
        epi_bar_sync();
        const float2 o = st[(half ^ 1) * kCtaRows + row_in_cta];
        mean = 0.5f * (mean_i + o.x);
        const float d = 0.5f * (mean_i - o.x);
        const float m2 = (m2_i + o.y) + 2.0f * kRowCols * d * d;   // Chan: two partials of 384 columns each
        rstd = rsqrtf(m2 * (1.0f / kRowN) + p.eps);
      }
      MMR_ROW_STAMP(7);

      // ---- pass 2a: the two parked tiles, from TMEM (tile 1 first: its accumulator is the next block's first)
#pragma unroll 1
      for (int i = 0; i < 8; ++i) {
        const int n = 1 + (i >> 2), c = i & 3;
        const int it_n = it - kRowTiles + n;
        const int acc = it_n & 1;
        const uint32_t taddr = lane_taddr + uint32_t(acc) * kBN;
        uint8_t* slot_s = wbuf + 4096 * (i % CFG::kSlots2a);
        uint8_t* o16_s = o16_base + 2048 * (i % kO16);
        const int col0 = chunk_col(n * 4 + c);
        if (lane == 0) {
          if (kSlots == 2 && i == 7) {
            // the 2-slot ring has no spare slot: tile 0's first chunk comes back into slot 0 once chunk 6 has left it
            bulk_wait_read<0>();
            mbar_arrive_expect_tx(&rfull[0], 4096);
            tma_load_2d(wbuf, &tmap_o32, &rfull[0], chunk_col(0), row0);
          } else {
            bulk_wait_read<CFG::kPend2a>();   // this chunk's slot and 16-bit stage have been read by their last stores
          }
        }
        __syncwarp();
        uint32_t v[32];
        tmem_ld_32x32(taddr + uint32_t(c * 32), v);
        tmem_ld_wait();
        if (c == 3) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster_relaxed(acc ? tempty_leader1 : tempty_leader0);
        }
        const float* gamma_c = vec_s + kRowN + col0;
        const float* beta_c = vec_s + 2 * kRowN + col0;
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 gg = *reinterpret_cast<const float4*>(gamma_c + 4 * j);
          const float4 be = *reinterpret_cast<const float4*>(beta_c + 4 * j);
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
      MMR_ROW_STAMP(8);

      // ---- pass 2b: tile 0, re-read from the fp32 output; normalised in place in the slot it landed in
#pragma unroll 1
      for (int j4 = 0; j4 < 4; ++j4) {
        const int rs = CFG::reload_slot(j4);
        uint8_t* slot_s = wbuf + 4096 * rs;
        uint8_t* o16_s = o16_base + 2048 * (j4 % kO16);
        const int col0 = chunk_col(j4);
        if (lane == 0) {
          bulk_wait_read<CFG::kPend2b>();
          if (kSlots == 1) {               // a single slot: this chunk itself comes back now
            mbar_arrive_expect_tx(&rfull[0], 4096);
            tma_load_2d(wbuf, &tmap_o32, &rfull[0], col0, row0);
          } else if (j4 + 1 < 4) {
            const int rn = CFG::reload_slot(j4 + 1);
            mbar_arrive_expect_tx(&rfull[rn], 4096);
            tma_load_2d(wbuf + 4096 * rn, &tmap_o32, &rfull[rn], chunk_col(j4 + 1), row0);
          }
        }
        __syncwarp();
        mbar_wait(&rfull[rs], (rph >> rs) & 1u);
        rph ^= 1u << rs;
        const float* gamma_c = vec_s + kRowN + col0;
        const float* beta_c = vec_s + 2 * kRowN + col0;
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4* yp = reinterpret_cast<float4*>(slot_s + lane * 128 + ((uint32_t(j) ^ sw128) << 4));
          const float4 t = *yp;
          const float4 gg = *reinterpret_cast<const float4*>(gamma_c + 4 * j);
          const float4 be = *reinterpret_cast<const float4*>(beta_c + 4 * j);
          float4 y;
          y.x = (t.x - mean) * rstd * gg.x + be.x;
          y.y = (t.y - mean) * rstd * gg.y + be.y;
          y.z = (t.z - mean) * rstd * gg.z + be.z;
          y.w = (t.w - mean) * rstd * gg.w + be.w;
          *yp = y;
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
      MMR_ROW_STAMP(9);
    }
    if (lane == 0) bulk_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == kRowMmaWarp) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------ host side
static unsigned long long* g_row_trace = nullptr;

template <class E16, class CFG>
static mmr_status launch_row_cfg(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& tr,
                                 const CUtensorMap& to32, const CUtensorMap& to16, const GemmLnRowParams& p,
                                 cudaStream_t stream) {
  auto kern = gemm_lnrow_kernel<E16, CFG>;
  static bool configured = false;
  if (!configured) {
    MMR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(CFG::kSmemBytes)));
    configured = true;
  }
  const int m_tiles = (p.M + kPairRows - 1) / kPairRows;
  const int max_pairs = sm_count() / 2;
  const int pairs = m_tiles < max_pairs ? m_tiles : max_pairs;
  MMR_CUDA_OK(launch_pdl(kern, dim3(2 * pairs), dim3(kGemmThreads), CFG::kSmemBytes, stream, ta, tw, tr, to32, to16, p));
  return MMR_OK;
}
template <class E16>
static mmr_status launch_row(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& tr, const CUtensorMap& to32,
                             const CUtensorMap& to16, const GemmLnRowParams& p, cudaStream_t stream) {
  // MMR_TUNE_LN_ROW_CFG: (operand stages, fp32 slots, 16-bit stages) as three digits; 0 = default for this K
  int cfg = tuning(MMR_TUNE_LN_ROW_CFG);
  if (cfg == 0) cfg = 421;
  switch (cfg) {
    case 331: return launch_row_cfg<E16, RowCfg331>(ta, tw, tr, to32, to16, p, stream);
    case 511: return launch_row_cfg<E16, RowCfg511>(ta, tw, tr, to32, to16, p, stream);
    case 412: return launch_row_cfg<E16, RowCfg412>(ta, tw, tr, to32, to16, p, stream);
    case 322: return launch_row_cfg<E16, RowCfg322>(ta, tw, tr, to32, to16, p, stream);
    default: return launch_row_cfg<E16, RowCfg421>(ta, tw, tr, to32, to16, p, stream);
  }
}

// Arguments were validated by gemm_ln() (gemm_ln_sm100.cu), which dispatches here when MMR_TUNE_GEMM_LN == 2.
mmr_status gemm_lnrow(const void* A16, int64_t lda, const void* W16, int64_t ldw, int M, int K, const float* bias,
                      const float* residual, int64_t ldr, const float* gamma, const float* beta, float eps, void* out16,
                      int64_t ldo16, float* out32, int64_t ldo32, int dtype, cudaStream_t stream) {
  const int ek = dtype == MMR_DT_BF16 ? 1 : 0;
  CUtensorMap ta, tw, tr, to32, to16;
  MMR_TRY(make_tmap_2d(&ta, A16, M, K, lda, kCtaRows, dtype));
  MMR_TRY(make_tmap_2d(&tw, W16, kRowN, K, ldw, kBN / 2, dtype));
  MMR_TRY(make_tmap_ex(&tr, residual, M, kRowN, ldr, 2, 32, 32, 128));
  MMR_TRY(make_tmap_ex(&to32, out32, M, kRowN, ldo32, 2, 32, 32, 128));
  MMR_TRY(make_tmap_ex(&to16, out16, M, kRowN, ldo16, ek, 32, 32, 64));
  GemmLnRowParams p{M, K, bias, gamma, beta, eps, uint32_t(dtype), g_row_trace};
  if (dtype == MMR_DT_BF16) return launch_row<BF16>(ta, tw, tr, to32, to16, p, stream);
  return launch_row<FP16>(ta, tw, tr, to32, to16, p, stream);
}

}  // namespace mmr

/* Debug only (not in the public header): device buffer of [grid][8 warps][2 blocks][16] uint64 phase stamps, or null. */
extern "C" void mmr_debug_set_lnrow_trace(unsigned long long* dev_buf) { mmr::g_row_trace = dev_buf; }

// Multi-head scaled-dot-product attention on tcgen05 / TMEM, pipelined four (pair, head) items deep per SM.
//
// Same contract as attention.cu (pixelbert.py:790-850, modeling.py:325-352): additive key mask (1 - m) * -10000,
// softmax over keys in fp32, merged-head 16-bit context rows.  attention_tc.cu (first tcgen05 version) ran ONE item
// per CTA through a 3.9 us chain (TMA -> QK^T -> softmax -> P to shared memory -> PV -> output) and lost to the
// mma.sync kernel (40.9 vs 32.5 us at B = 256, S = 68).  An item is 1.2 MFLOP: nothing here is throughput, everything
// is latency, so this version keeps FOUR items in flight per SM and gives every step of the chain its own warp(s):
//   mask warp     the additive mask row of the next items, in the log2 domain (raw int32 rows fetched by cp.async
//                 three items ahead: an L2 round trip must not be the period of anything)
//   TMA producer  Q, K, V boxes ([rows x 64] 16-bit, 128-byte swizzle) of the next items into a ring of up to 6 stages
//   QK^T issuer   one thread: S = Q K^T (UMMA 128 x SkP x 16, operands from shared memory) into TMEM slot n % 4 as soon
//                 as the item's stage is full and the slot's previous O has been read
//   PV issuer     one thread: O = P V as soon as that slot's P is ready (A operand from TMEM, B = V used MN-major
//                 straight from its [keys x 64] tile)
//   softmax       four warpgroups; group g owns TMEM slot g and the items n = g (mod 4) of this CTA.  One thread = one
//                 query row = one TMEM lane.  For <= 80 keys the S row is read ONCE into registers (max, then
//                 exp2 / sum); P goes back INTO TMEM as packed 16-bit pairs over the columns S came from
//                 (tcgen05.st), so it never touches shared memory and needs no proxy fence.  O lands in the slot's
//                 upper 64 columns, is normalised by 1 / sum and leaves as one 128-byte row segment per thread.
// TMEM: 4 slots x 128 columns: S in [0, SkP), P (16-bit pairs) in [0, SkP / 2), O in [64, 128) — S is dead when the
// PV MMA is issued.  Shared memory holds nothing but the operand ring and the mask rows.  Rows >= Sq of the 128-row
// MMAs read whatever follows the item's Q rows in shared memory and are never stored (every output row depends on
// its own query row only); key rows >= Sk are zero (zeroed once, TMA never writes them) and carry a -inf mask.
// What the timeline (tools/attn_trace.py, profiles/) showed on the way: a key_mask load inside the producer loop,
// then the producer loop itself (~1 us of small dependent latencies per item), then ONE thread issuing all nine MMAs
// of an item (~50 ns each) were, in turn, the period of the whole CTA.
#include <cuda.h>

#include <algorithm>

#include "gemm_common.cuh"
#include "kernels.cuh"

namespace mmr {

constexpr int kT2HeadDim = 64;
constexpr int kT2Slots = 4;                        // TMEM slots = softmax warpgroups = items in flight
constexpr int kT2SlotCols = 128;
constexpr int kT2MaxStages = 6;
// Warp roles.  A softmax warp may only touch the TMEM lanes [32 (id % 4), +32), so group g = id / 4 and lane quarter
// w = id % 4.  WPG = 4 (up to 128 query rows): 16 softmax warps + warps 16 (producer) and 17 (MMA); with 18 warps one
// SM sub-partition hosts 5, which caps the kernel at 96 registers per thread.  WPG = 3 (up to 96 query rows — every
// shape of the three scorers but lds): quarter 3 of every group has no rows, so those warp ids take the two special
// roles instead (warp 3 = mask rows, 7 = TMA producer, 11 = Q K^T issuer, 15 = P V issuer): 16 warps, 128 registers,
// the S row of a thread fits in registers, and sub-partition 3 runs nothing but the four single-purpose warps.
template <int WPG>
struct T2Roles {
  static constexpr int kWarps = WPG == 3 ? 16 : 20;
  static constexpr int kThreads = 32 * kWarps;
  static constexpr int kMaskWarp = WPG == 3 ? 3 : 16;
  static constexpr int kProducerWarp = WPG == 3 ? 7 : 17;
  static constexpr int kMmaWarp = WPG == 3 ? 11 : 18;     // issues S = Q K^T; owns the TMEM allocation
  static constexpr int kPvWarp = WPG == 3 ? 15 : 19;      // issues O = P V
  // O rows leave through a per-warp 4 KB shared-memory stage as 128-byte coalesced rows (4 rows per store
  // instruction instead of 32 different lines); the 4-warps-per-group kernel (lds, S = 104) needs the shared
  // memory for its operand ring and stores straight from registers
  static constexpr uint32_t kOutStageBytes = WPG == 3 ? 12 * 4096 : 0;
  __device__ static bool is_softmax(int warp) { return WPG == 3 ? (warp & 3) != 3 : warp < 16; }
};
constexpr uint32_t kT2QReadBytes = 128 * 128;      // what the 128-row UMMA reads from an item's Q base

// D[tmem] (+)= A[tmem] * B[smem]: A is K-major in TMEM (lane = row, one 32-bit column = two consecutive K elements)
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// registers -> TMEM: this warp's 32 lanes x 8 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// Non-blocking probe of an mbarrier phase (try_wait may suspend the thread for a system-dependent time, which would
// starve the other queue of the polling MMA thread).
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ unsigned long long t2_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// debug timeline: [CTA][item n < 24][16] = {tma issued, qk issued, pv issued, s ready seen, p arrived, o ready seen,
// row stored, -} in ns
#define MMR_T2_STAMP(n, k)                                                                          \
  do {                                                                                              \
    if (trace != nullptr && (n) < 24) trace[(size_t(blockIdx.x) * 24 + (n)) * 16 + (k)] = t2_now();  \
  } while (0)

__device__ __forceinline__ void cp_async_4(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}

struct T2Layout {
  uint32_t q_bytes, kv_bytes, stage_bytes, pad_bytes;
  int n_stages;
  size_t smem_bytes;
};
// out_stage_bytes: per-CTA staging for the coalesced output path (12 warps x 4 KB in the 3-warps-per-group kernels)
__host__ __device__ inline T2Layout t2_layout(int Sq, int Sk, uint32_t out_stage_bytes) {
  T2Layout L;
  const int SkP = (Sk + 15) & ~15, Sq8 = (Sq + 7) & ~7;
  L.q_bytes = uint32_t(Sq8) * 128u;
  L.kv_bytes = uint32_t(SkP) * 128u;
  L.stage_bytes = L.q_bytes + 2u * L.kv_bytes;
  // the last stage's Q read (128 rows) must stay inside the allocation
  L.pad_bytes = L.stage_bytes >= kT2QReadBytes ? 0u : kT2QReadBytes - L.stage_bytes;
  const size_t fixed = 1024 + L.pad_bytes + size_t(kT2MaxStages) * 512 + 4 * 512 + 512 + out_stage_bytes;
  int n = int((size_t(227) * 1024 - fixed) / L.stage_bytes);
  L.n_stages = n > kT2MaxStages ? kT2MaxStages : n;
  L.smem_bytes = fixed + size_t(L.n_stages) * L.stage_bytes;
  return L;
}

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kScaleLog2e = 0.125f * kLog2e;     // 1 / sqrt(64), in the log2 domain
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// One attention problem of a launch: B pairs x heads items of Sq query rows x Sk keys (device pointers).
struct T2Seg {
  const int32_t* key_mask;   // [B, Sk] or null
  void* out;                 // [B * Sq, ldo] 16-bit
  int64_t ldo;
  int Sq, Sk;
};
struct T2Segs {
  T2Seg s[2];
  int n_items0;              // items [0, n_items0) are segment 0, [n_items0, n_items) segment 1
};

// NCH = number of 16-key chunks of an S row, kept in registers (2, 3 or 5: the key counts of the three scorers at
// the bench and native shapes); 0 = any row of up to 128 keys, two passes over TMEM.
// TWO = the launch carries two segments (see below); single-problem launches compile the segment selects away.
template <class E16, int NCH, int WPG, bool TWO>
__global__ void __launch_bounds__(T2Roles<WPG>::kThreads, 1)
attention_tc2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                     const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_q1,
                     const __grid_constant__ CUtensorMap tmap_k1, const __grid_constant__ CUtensorMap tmap_v1,
                     const T2Segs segs, int heads, int n_items, uint32_t idesc_fmt, unsigned long long* trace,
                     int ablate) {
  using T = typename E16::T;
  extern __shared__ __align__(1024) uint8_t smem_t2[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_t2) + 1023) & ~uintptr_t(1023));
  // Two SEGMENTS may share one launch (LXMERT: the self-attention of the language and of the visual stream, or the two
  // directions of a cross-attention block): items [0, n_items0) belong to segment 0, the rest to segment 1, each with
  // its own tensor maps, lengths, mask and output.  Stages, TMEM slots and the S-row width are laid out for the LARGER
  // shapes; an item with fewer keys leaves the key rows [Sk_i, SkP) of its stage as the previous item left them --
  // finite numbers behind a -inf mask, i.e. probabilities that are exactly 0.
  const int Sq = max(segs.s[0].Sq, segs.s[1].Sq), Sk = max(segs.s[0].Sk, segs.s[1].Sk);
  const int SkMin = min(segs.s[0].Sk, segs.s[1].Sk);
  const T2Layout L = t2_layout(Sq, Sk, T2Roles<WPG>::kOutStageBytes);
  const int SkP = (Sk + 15) & ~15;
  const int n_items0 = segs.n_items0;
  // work item n of this CTA -> (segment, pair, head)
  auto item_of = [&](int n, int& b, int& h) -> int {
    const int item = int(blockIdx.x) + n * int(gridDim.x);
    const int seg = (TWO && item >= n_items0) ? 1 : 0;
    const int li = item - (seg ? n_items0 : 0);
    b = li / heads;
    h = li - b * heads;
    return seg;
  };
  const int n_stages = L.n_stages;
  float* mask_s = reinterpret_cast<float*>(ring + size_t(n_stages) * L.stage_bytes + L.pad_bytes);   // [stages][128]
  int32_t* raw_s = reinterpret_cast<int32_t*>(mask_s + kT2MaxStages * 128);                          // [4][128]
  uint8_t* out_stage = reinterpret_cast<uint8_t*>(raw_s + 4 * 128);                                  // [12][4 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(out_stage + T2Roles<WPG>::kOutStageBytes);
  uint64_t* full_bar = bars;                          // [6] TMA -> MMA / softmax (mask row)
  uint64_t* empty_bar = bars + kT2MaxStages;          // [6] PV retired -> producer
  uint64_t* s_ready = bars + 2 * kT2MaxStages;        // [4] S complete in TMEM
  uint64_t* p_ready = s_ready + kT2Slots;             // [4] P written back to TMEM (4 warps)
  uint64_t* o_ready = p_ready + kT2Slots;             // [4] O complete in TMEM
  uint64_t* slot_free = o_ready + kT2Slots;           // [4] O read out (4 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(slot_free + kT2Slots);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kT2ProducerWarp = T2Roles<WPG>::kProducerWarp, kT2MmaWarp = T2Roles<WPG>::kMmaWarp;
  constexpr int kT2MaskWarp = T2Roles<WPG>::kMaskWarp, kT2PvWarp = T2Roles<WPG>::kPvWarp;

  // zero the padding key rows [Sk, SkP) of K and V in every stage once: TMA never writes them, and they must read as
  // finite numbers (their probabilities are exactly 0; 0 x NaN is not).  Q rows >= Sq may hold anything.
  {
    const uint32_t pad16 = uint32_t(SkP - SkMin) * 8u;   // 16-byte units per K (or V) tile
    for (uint32_t i = threadIdx.x; i < uint32_t(n_stages) * 2u * pad16; i += blockDim.x) {
      const uint32_t st = i / (2u * pad16), r = i - st * 2u * pad16;
      const uint32_t tile = r / pad16, o = r - tile * pad16;
      uint8_t* base = ring + size_t(st) * L.stage_bytes + L.q_bytes + tile * L.kv_bytes + uint32_t(SkMin) * 128u;
      reinterpret_cast<uint4*>(base)[o] = make_uint4(0, 0, 0, 0);
    }
  }
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    if (TWO) {
      tma_prefetch_desc(&tmap_q1);
      tma_prefetch_desc(&tmap_k1);
      tma_prefetch_desc(&tmap_v1);
    }
    for (int s = 0; s < kT2MaxStages; ++s) {
      mbar_init(&full_bar[s], 2);    // the producer's expect_tx arrive + the mask warp's
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < kT2Slots; ++s) {
      mbar_init(&s_ready[s], 1);
      mbar_init(&p_ready[s], WPG);
      mbar_init(&o_ready[s], 1);
      mbar_init(&slot_free[s], WPG);
    }
    fence_mbar_init();
  }
  if (warp == kT2MmaWarp) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async();   // the zero fill (generic proxy) is ordered before TMA writes / UMMA reads (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch_dependents();
  const int my_items = (n_items - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);

  if (warp == kT2ProducerWarp) {
    // ===================== TMA producer =====================
    // Nothing but "stage free -> three boxes": every dependent latency in this loop is paid once per item by the CTA.
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int n = 0; n < my_items; ++n) {
        int b, h;
        const int seg = item_of(n, b, h);
        const int sq = seg ? segs.s[1].Sq : segs.s[0].Sq, sk = seg ? segs.s[1].Sk : segs.s[0].Sk;
        const CUtensorMap* mq = seg ? &tmap_q1 : &tmap_q;
        const CUtensorMap* mk = seg ? &tmap_k1 : &tmap_k;
        const CUtensorMap* mv = seg ? &tmap_v1 : &tmap_v;
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        uint8_t* st = ring + size_t(stage) * L.stage_bytes;
        if (ablate & 8) {
          mbar_arrive_expect_tx(&full_bar[stage], uint32_t(sq) * 128u);
          tma_load_2d(st, mq, &full_bar[stage], h * kT2HeadDim, b * sq);
        } else {
        mbar_arrive_expect_tx(&full_bar[stage], uint32_t(sq + 2 * sk) * 128u);
        tma_load_2d(st, mq, &full_bar[stage], h * kT2HeadDim, b * sq);
        tma_load_2d(st + L.q_bytes, mk, &full_bar[stage], h * kT2HeadDim, b * sk);
        tma_load_2d(st + L.q_bytes + L.kv_bytes, mv, &full_bar[stage], h * kT2HeadDim, b * sk);
        }
        MMR_T2_STAMP(n, 0);
        if (++stage == n_stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == kT2MaskWarp) {
    // ===================== mask rows =====================
    // The additive mask row of an item, in the log2 domain (x log2 e), second arrival on the stage's "full" barrier.
    // A key_mask load is an L2 round trip (~1 us): the raw int32 rows travel global -> shared with cp.async THREE
    // ITEMS AHEAD into a 4-entry staging ring (each lane reads back only what it fetched itself).
    auto issue_mask = [&](int n) {
      if (n < my_items && !(ablate & 2)) {
        int b, h;
        const int seg = item_of(n, b, h);
        const int32_t* km = seg ? segs.s[1].key_mask : segs.s[0].key_mask;   // (selects, not indexing: a dynamically
        const int sk = seg ? segs.s[1].Sk : segs.s[0].Sk;                    //  indexed parameter struct moves to local memory)
        int32_t* dst = raw_s + (n & 3) * 128;
        if (km != nullptr)
          for (int i = lane; i < sk; i += 32) cp_async_4(dst + i, km + int64_t(b) * sk + i);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");   // (possibly empty) group: one per item, in order
    };
    int stage = 0;
    uint32_t phase = 0;
    issue_mask(0);
    issue_mask(1);
    issue_mask(2);
    for (int n = 0; n < my_items; ++n) {
      issue_mask(n + 3);
      asm volatile("cp.async.wait_group 3;" ::: "memory");    // item n's row has landed
      mbar_wait(&empty_bar[stage], phase ^ 1u);
      float* sMask = mask_s + stage * 128;
      const int32_t* raw = raw_s + (n & 3) * 128;
      int b_, h_;
      const int seg = item_of(n, b_, h_);
      const bool has_mask = (seg ? segs.s[1].key_mask : segs.s[0].key_mask) != nullptr;
      const int sk = seg ? segs.s[1].Sk : segs.s[0].Sk;
      for (int i = lane; i < SkP; i += 32) {
        float m = -INFINITY;   // padding keys (>= this item's Sk) do not exist for the softmax
        if (i < sk) m = (!has_mask || raw[i] != 0) ? 0.0f : -10000.0f * kLog2e;
        sMask[i] = m;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[stage]);
      if (++stage == n_stages) {
        stage = 0;
        phase ^= 1u;
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else if (warp == kT2MmaWarp) {
    // ===================== S = Q K^T issuer (one thread) =====================
    // 4 steps of 16 along d into TMEM slot n % 4, as soon as the item's stage is full and the slot's O has been read.
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_f16(idesc_fmt, 128, uint32_t(SkP));
      int stage = 0;
      uint32_t phase = 0;
      for (int n = 0; n < my_items; ++n) {
        const int slot = n & (kT2Slots - 1);
        mbar_wait(&full_bar[stage], phase);
        mbar_wait(&slot_free[slot], (uint32_t(n >> 2) & 1u) ^ 1u);
        tc_fence_after();
        MMR_T2_STAMP(n, 11);
        const uint32_t st = smem_u32(ring + size_t(stage) * L.stage_bytes);
        const uint64_t q_desc = umma_desc_k_sw128(st), k_desc = umma_desc_k_sw128(st + L.q_bytes);
        const uint32_t tmem_s = tmem_base + uint32_t(slot * kT2SlotCols);
#pragma unroll
        for (int k = 0; k < kT2HeadDim / kUmmaK; ++k)
          if (k == 0 || !(ablate & 32))
            umma_f16(tmem_s, q_desc + uint64_t(2 * k), k_desc + uint64_t(2 * k), idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&s_ready[slot]);
        MMR_T2_STAMP(n, 1);
        if (++stage == n_stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == kT2PvWarp) {
    // ===================== O = P V issuer (one thread) =====================
    // Its own thread: issuing one tcgen05.mma costs this code ~50 ns (descriptor arithmetic and the uniform-register
    // hand-over around every instruction), so the 9 MMAs of an item issued by ONE thread were the period of the
    // whole CTA (tools/attn_trace.py).  SkP / 16 steps of 16 keys: 8 packed P columns from TMEM, +2048 B of V.
    if (lane == 0) {
      const uint32_t idesc_o = umma_idesc_f16(idesc_fmt, 128, kT2HeadDim) | (1u << 16);   // B (= V) is MN-major
      int stage = 0;
      for (int n = 0; n < my_items; ++n) {
        const int slot = n & (kT2Slots - 1);
        mbar_wait(&p_ready[slot], uint32_t(n >> 2) & 1u);
        tc_fence_after();
        MMR_T2_STAMP(n, 9);
        const uint32_t st = smem_u32(ring + size_t(stage) * L.stage_bytes);
        const uint64_t v_desc = umma_desc_k_sw128(st + L.q_bytes + L.kv_bytes);
        const uint32_t tmem_p = tmem_base + uint32_t(slot * kT2SlotCols), tmem_o = tmem_p + 64u;
        for (int ks = 0; ks < ((ablate & 32) ? 1 : SkP / kUmmaK); ++ks)
          umma_f16_ts(tmem_o, tmem_p + uint32_t(8 * ks), v_desc + uint64_t(128 * ks), idesc_o, ks != 0 ? 1u : 0u);
        MMR_T2_STAMP(n, 10);
        umma_commit(&o_ready[slot]);
        umma_commit(&empty_bar[stage]);   // Q, K, V and the mask row of this stage are no longer read
        MMR_T2_STAMP(n, 2);
        if (++stage == n_stages) stage = 0;
      }
    }
  } else if (T2Roles<WPG>::is_softmax(warp)) {
    // ===================== softmax / output warpgroups: one thread = one query row =====================
    const int g = warp >> 2, w = warp & 3;
    const int row = w * 32 + lane;                         // TMEM lane of this thread = query row
    const uint32_t tmem_s = tmem_base + uint32_t(g * kT2SlotCols) + (uint32_t(w * 32) << 16);
    const uint32_t tmem_o = tmem_s + 64u;
    const int n_chunks = SkP >> 4;
    for (int n = g; n < my_items; n += kT2Slots) {
      const uint32_t use = uint32_t(n >> 2);
      const int stage = n % n_stages;
      const uint32_t ring_phase = uint32_t(n / n_stages) & 1u;
      int b, h;
      const int seg = item_of(n, b, h);
      const int sq_i = seg ? segs.s[1].Sq : segs.s[0].Sq;
      T* __restrict__ out = static_cast<T*>(seg ? segs.s[1].out : segs.s[0].out);
      const int64_t ldo = seg ? segs.s[1].ldo : segs.s[0].ldo;
      const bool live = w * 32 < sq_i;                     // this warp owns at least one real query row of the item
      const float* sMask = mask_s + stage * 128;
      mbar_wait(&full_bar[stage], ring_phase);             // the mask row (generic writes of the producer warp)
      mbar_wait(&s_ready[g], use & 1u);
      tc_fence_after();
      if (w == 0 && lane == 0) MMR_T2_STAMP(n, 3);
      float inv = 0.f;
      if (live && !(ablate & 4)) {
        // t = S / 8 + mask in the log2 domain; p = 2^(t - max); P as packed 16-bit pairs over the first SkP / 2
        // columns of S (columns [8c, 8c + 8) were read, as S columns, by this very thread before it overwrites them)
        const float4* m4 = reinterpret_cast<const float4*>(sMask);
        float sum = 0.f;
        if constexpr (NCH > 0) {
          // the whole S row lives in registers: ONE TMEM round trip, one pass for the maximum, one for exp / sum
          uint32_t v[NCH][16];
#pragma unroll
          for (int c = 0; c < NCH; ++c) tmem_ld_32x16(tmem_s + uint32_t(c * 16), v[c]);
          tmem_ld_wait();
          float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 mm = m4[c * 4 + q];
                const float t0 = fmaf(__uint_as_float(v[c][4 * q]), kScaleLog2e, mm.x);
                const float t1 = fmaf(__uint_as_float(v[c][4 * q + 1]), kScaleLog2e, mm.y);
                const float t2 = fmaf(__uint_as_float(v[c][4 * q + 2]), kScaleLog2e, mm.z);
                const float t3 = fmaf(__uint_as_float(v[c][4 * q + 3]), kScaleLog2e, mm.w);
                mx0 = fmaxf(mx0, t0); mx1 = fmaxf(mx1, t1); mx2 = fmaxf(mx2, t2); mx3 = fmaxf(mx3, t3);
                v[c][4 * q] = __float_as_uint(t0); v[c][4 * q + 1] = __float_as_uint(t1);
                v[c][4 * q + 2] = __float_as_uint(t2); v[c][4 * q + 3] = __float_as_uint(t3);
              }
            }
          }
          const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
          float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            {
              uint32_t pk[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float e0 = ex2_approx(__uint_as_float(v[c][2 * j]) - mx);
                const float e1 = ex2_approx(__uint_as_float(v[c][2 * j + 1]) - mx);
                sum0 += e0;
                sum1 += e1;
                pk[j] = E16::pack(e0, e1);
              }
              tmem_st_32x8(tmem_s + uint32_t(c * 8), pk);
            }
          }
          sum = sum0 + sum1;
        } else {
          // long rows (more than 80 keys): two passes over the S row in TMEM
          float mx = -INFINITY, sum0 = 0.f, sum1 = 0.f;
          for (int c = 0; c < n_chunks; ++c) {
            uint32_t v[16];
            tmem_ld_32x16(tmem_s + uint32_t(c * 16), v);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 mm = m4[c * 4 + q];
              mx = fmaxf(mx, fmaxf(fmaxf(fmaf(__uint_as_float(v[4 * q]), kScaleLog2e, mm.x),
                                         fmaf(__uint_as_float(v[4 * q + 1]), kScaleLog2e, mm.y)),
                                   fmaxf(fmaf(__uint_as_float(v[4 * q + 2]), kScaleLog2e, mm.z),
                                         fmaf(__uint_as_float(v[4 * q + 3]), kScaleLog2e, mm.w))));
            }
          }
          for (int c = 0; c < n_chunks; ++c) {
            uint32_t v[16];
            tmem_ld_32x16(tmem_s + uint32_t(c * 16), v);
            tmem_ld_wait();
            uint32_t pk[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 mm = m4[c * 4 + q];
              const float e0 = ex2_approx(fmaf(__uint_as_float(v[4 * q]), kScaleLog2e, mm.x) - mx);
              const float e1 = ex2_approx(fmaf(__uint_as_float(v[4 * q + 1]), kScaleLog2e, mm.y) - mx);
              const float e2 = ex2_approx(fmaf(__uint_as_float(v[4 * q + 2]), kScaleLog2e, mm.z) - mx);
              const float e3 = ex2_approx(fmaf(__uint_as_float(v[4 * q + 3]), kScaleLog2e, mm.w) - mx);
              sum0 += e0;   // the same two running sums, in the same order, as the row-in-registers path: a row's bits do
              sum1 += e1;   // not depend on which specialisation of the kernel a launch (or a paired launch) picks
              sum0 += e2;
              sum1 += e3;
              pk[2 * q] = E16::pack(e0, e1);
              pk[2 * q + 1] = E16::pack(e2, e3);
            }
            tmem_st_32x8(tmem_s + uint32_t(c * 8), pk);
          }
          sum = sum0 + sum1;
        }
        tmem_st_wait();
        inv = 1.0f / sum;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[g]);
      if (w == 0 && lane == 0) MMR_T2_STAMP(n, 4);
      if (w == 1 && lane == 0) MMR_T2_STAMP(n, 7);
      if (w == 2 && lane == 0) MMR_T2_STAMP(n, 8);

      // ---- O row: normalise, 16 bit, one 128-byte segment per thread
      mbar_wait(&o_ready[g], use & 1u);
      tc_fence_after();
      if (w == 0 && lane == 0) MMR_T2_STAMP(n, 5);
      if (live && !(ablate & 16)) {
        uint32_t o0[32], o1[32];
        tmem_ld_32x32(tmem_o, o0);
        tmem_ld_32x32(tmem_o + 32u, o1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&slot_free[g]);
        if constexpr (WPG == 3) {
          // this thread's row -> the warp's stage (16-byte units XOR-swizzled by the row: conflict-free both ways),
          // then row-major read-back: every store instruction covers four whole 128-byte row segments
          uint8_t* stg = out_stage + size_t(g * 3 + w) * 4096;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            *reinterpret_cast<uint4*>(stg + lane * 128 + ((uint32_t(u) ^ uint32_t(lane & 7)) << 4)) = make_uint4(
                E16::pack(__uint_as_float(o0[8 * u]) * inv, __uint_as_float(o0[8 * u + 1]) * inv),
                E16::pack(__uint_as_float(o0[8 * u + 2]) * inv, __uint_as_float(o0[8 * u + 3]) * inv),
                E16::pack(__uint_as_float(o0[8 * u + 4]) * inv, __uint_as_float(o0[8 * u + 5]) * inv),
                E16::pack(__uint_as_float(o0[8 * u + 6]) * inv, __uint_as_float(o0[8 * u + 7]) * inv));
            *reinterpret_cast<uint4*>(stg + lane * 128 + ((uint32_t(4 + u) ^ uint32_t(lane & 7)) << 4)) = make_uint4(
                E16::pack(__uint_as_float(o1[8 * u]) * inv, __uint_as_float(o1[8 * u + 1]) * inv),
                E16::pack(__uint_as_float(o1[8 * u + 2]) * inv, __uint_as_float(o1[8 * u + 3]) * inv),
                E16::pack(__uint_as_float(o1[8 * u + 4]) * inv, __uint_as_float(o1[8 * u + 5]) * inv),
                E16::pack(__uint_as_float(o1[8 * u + 6]) * inv, __uint_as_float(o1[8 * u + 7]) * inv));
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rl = i * 4 + (lane >> 3), u = lane & 7;
            const uint4 val = *reinterpret_cast<const uint4*>(stg + rl * 128 + ((uint32_t(u) ^ uint32_t(rl & 7)) << 4));
            const int r = w * 32 + rl;
            if (r < sq_i && !(ablate & 1)) *reinterpret_cast<uint4*>(out + (int64_t(b) * sq_i + r) * ldo + h * kT2HeadDim + u * 8) = val;
          }
          __syncwarp();   // the stage is rewritten by this warp's next item
        } else if (row < sq_i) {
          uint4* orow = reinterpret_cast<uint4*>(out + (int64_t(b) * sq_i + row) * ldo + h * kT2HeadDim);
#pragma unroll
          for (int u = 0; u < 4; ++u)
            orow[u] = make_uint4(E16::pack(__uint_as_float(o0[8 * u]) * inv, __uint_as_float(o0[8 * u + 1]) * inv),
                                 E16::pack(__uint_as_float(o0[8 * u + 2]) * inv, __uint_as_float(o0[8 * u + 3]) * inv),
                                 E16::pack(__uint_as_float(o0[8 * u + 4]) * inv, __uint_as_float(o0[8 * u + 5]) * inv),
                                 E16::pack(__uint_as_float(o0[8 * u + 6]) * inv, __uint_as_float(o0[8 * u + 7]) * inv));
#pragma unroll
          for (int u = 0; u < 4; ++u)
            orow[4 + u] = make_uint4(E16::pack(__uint_as_float(o1[8 * u]) * inv, __uint_as_float(o1[8 * u + 1]) * inv),
                                     E16::pack(__uint_as_float(o1[8 * u + 2]) * inv, __uint_as_float(o1[8 * u + 3]) * inv),
                                     E16::pack(__uint_as_float(o1[8 * u + 4]) * inv, __uint_as_float(o1[8 * u + 5]) * inv),
                                     E16::pack(__uint_as_float(o1[8 * u + 6]) * inv, __uint_as_float(o1[8 * u + 7]) * inv));
        }
      } else {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&slot_free[g]);
      }
      if (w == 0 && lane == 0) MMR_T2_STAMP(n, 6);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kT2MmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static unsigned long long* g_t2_trace = nullptr;
static int g_t2_ablate = 0;   // timing experiments only (tools/attn_ablate.py): results are wrong when non-zero

// One attention problem as the host sees it.
struct T2Problem {
  const void *q, *k, *v;
  int64_t ldq, ldk, ldv;
  const int32_t* key_mask;
  void* out;
  int64_t ldo;
  int B, Sq, Sk;
};

template <class E16, int NCH, int WPG, bool TWO>
static mmr_status launch_attention_tc2(const T2Problem& p0, const T2Problem* p1, int heads, int dtype, cudaStream_t stream) {
  auto kern = attention_tc2_kernel<E16, NCH, WPG, TWO>;
  static bool configured = false;
  if (!configured) {
    MMR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  const int ek = dtype == MMR_DT_BF16 ? 1 : 0;
  const T2Problem& pb = p1 != nullptr ? *p1 : p0;
  CUtensorMap tq, tk, tv, tq1, tk1, tv1;
  const int64_t cols = int64_t(heads) * kT2HeadDim;
  MMR_TRY(make_tmap_ex(&tq, p0.q, int64_t(p0.B) * p0.Sq, cols, p0.ldq, ek, kT2HeadDim, p0.Sq, 128));
  MMR_TRY(make_tmap_ex(&tk, p0.k, int64_t(p0.B) * p0.Sk, cols, p0.ldk, ek, kT2HeadDim, p0.Sk, 128));
  MMR_TRY(make_tmap_ex(&tv, p0.v, int64_t(p0.B) * p0.Sk, cols, p0.ldv, ek, kT2HeadDim, p0.Sk, 128));
  MMR_TRY(make_tmap_ex(&tq1, pb.q, int64_t(pb.B) * pb.Sq, cols, pb.ldq, ek, kT2HeadDim, pb.Sq, 128));
  MMR_TRY(make_tmap_ex(&tk1, pb.k, int64_t(pb.B) * pb.Sk, cols, pb.ldk, ek, kT2HeadDim, pb.Sk, 128));
  MMR_TRY(make_tmap_ex(&tv1, pb.v, int64_t(pb.B) * pb.Sk, cols, pb.ldv, ek, kT2HeadDim, pb.Sk, 128));
  const T2Layout L = t2_layout(std::max(p0.Sq, pb.Sq), std::max(p0.Sk, pb.Sk), T2Roles<WPG>::kOutStageBytes);
  MMR_REQUIRE(L.n_stages >= 2, "attention_tc2: operand ring does not fit (Sq=%d Sk=%d)", std::max(p0.Sq, pb.Sq),
              std::max(p0.Sk, pb.Sk));
  T2Segs segs;
  segs.s[0] = T2Seg{p0.key_mask, p0.out, p0.ldo, p0.Sq, p0.Sk};
  segs.s[1] = T2Seg{pb.key_mask, pb.out, pb.ldo, pb.Sq, pb.Sk};
  segs.n_items0 = p0.B * heads;
  const int n_items = segs.n_items0 + (p1 != nullptr ? p1->B * heads : 0);
  const int grid = std::min(n_items, sm_count());   // one CTA per SM: it allocates all 512 TMEM columns
  MMR_CUDA_OK(launch_pdl(kern, dim3(grid), dim3(T2Roles<WPG>::kThreads), L.smem_bytes, stream, tq, tk, tv, tq1, tk1, tv1,
                         segs, heads, n_items, uint32_t(dtype), g_t2_trace, g_t2_ablate));
  return MMR_OK;
}
template <class E16>
static mmr_status launch_attention_tc2_e(const T2Problem& p0, const T2Problem* p1, int heads, int dtype,
                                         cudaStream_t stream) {
  const int Sq = std::max(p0.Sq, p1 ? p1->Sq : 0), Sk = std::max(p0.Sk, p1 ? p1->Sk : 0);
  const int chunks = (Sk + 15) / 16;
  if (p1 != nullptr) {
    // two-problem launches are built for the row-in-registers kernels of the LXMERT shapes and the generic ones
    if (Sq <= 96 && chunks == 2) return launch_attention_tc2<E16, 2, 3, true>(p0, p1, heads, dtype, stream);
    if (Sq <= 96 && chunks == 3) return launch_attention_tc2<E16, 3, 3, true>(p0, p1, heads, dtype, stream);
    if (Sq <= 96) return launch_attention_tc2<E16, 0, 3, true>(p0, p1, heads, dtype, stream);
    return launch_attention_tc2<E16, 0, 4, true>(p0, p1, heads, dtype, stream);
  }
  if (Sq <= 96 && chunks == 2) return launch_attention_tc2<E16, 2, 3, false>(p0, p1, heads, dtype, stream);
  if (Sq <= 96 && chunks == 3) return launch_attention_tc2<E16, 3, 3, false>(p0, p1, heads, dtype, stream);
  if (Sq <= 96 && chunks == 5) return launch_attention_tc2<E16, 5, 3, false>(p0, p1, heads, dtype, stream);
  if (Sq <= 96) return launch_attention_tc2<E16, 0, 3, false>(p0, p1, heads, dtype, stream);
  return launch_attention_tc2<E16, 0, 4, false>(p0, p1, heads, dtype, stream);
}

// Arguments are validated by mmr::attention (attention.cu); on top of those this path needs 16-byte aligned output rows.
bool attention_tc2_eligible(const void* out16, int64_t ldo) {
#ifdef MMR_EXPERIMENTAL
  if (tuning(MMR_TUNE_ATTN_TC) != 2) return false;   // the older kernels exist only in experimental builds
#endif
  return ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(out16) & 15) == 0;
}
mmr_status attention_tc2(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                         const int32_t* key_mask, void* out16, int64_t ldo, int B, int Sq, int Sk, int heads, int dtype,
                         cudaStream_t stream) {
  const T2Problem p{q, k, v, ldq, ldk, ldv, key_mask, out16, ldo, B, Sq, Sk};
  if (dtype == MMR_DT_BF16) return launch_attention_tc2_e<BF16>(p, nullptr, heads, dtype, stream);
  return launch_attention_tc2_e<FP16>(p, nullptr, heads, dtype, stream);
}

// Two attention problems in ONE launch (see the kernel's comment on segments): what LXMERT issues back to back for its
// two streams (modeling.py:381-391 per stream, :462-463 for the two directions of the shared cross-attention block).
mmr_status attention_pair(const AttentionArgs& a, const AttentionArgs& b, int heads, int dtype, cudaStream_t stream) {
  MMR_TRY(require_sm100());
  for (const AttentionArgs* x : {&a, &b}) {
    MMR_REQUIRE(x->q && x->k && x->v && x->out16 && x->B > 0, "attention_pair: null pointer or empty batch");
    MMR_REQUIRE(x->Sq > 0 && x->Sq <= 128 && x->Sk > 0 && x->Sk <= 128, "attention_pair: Sq=%d Sk=%d must be in [1,128]",
                x->Sq, x->Sk);
    MMR_REQUIRE(x->ldq % 8 == 0 && x->ldk % 8 == 0 && x->ldv % 8 == 0 && x->ldo % 8 == 0 &&
                    ((reinterpret_cast<uintptr_t>(x->q) | reinterpret_cast<uintptr_t>(x->k) |
                      reinterpret_cast<uintptr_t>(x->v) | reinterpret_cast<uintptr_t>(x->out16)) & 15) == 0,
                "attention_pair: operands and output rows must be 16-byte aligned");
  }
  const T2Problem p0{a.q, a.k, a.v, a.ldq, a.ldk, a.ldv, a.key_mask, a.out16, a.ldo, a.B, a.Sq, a.Sk};
  const T2Problem p1{b.q, b.k, b.v, b.ldq, b.ldk, b.ldv, b.key_mask, b.out16, b.ldo, b.B, b.Sq, b.Sk};
  if (dtype == MMR_DT_BF16) return launch_attention_tc2_e<BF16>(p0, &p1, heads, dtype, stream);
  return launch_attention_tc2_e<FP16>(p0, &p1, heads, dtype, stream);
}

}  // namespace mmr

/* Debug only (not in the public header): device buffer of [grid][24 items][16] uint64 stamps, or null. */
extern "C" void mmr_debug_set_attn_trace(unsigned long long* dev_buf) { mmr::g_t2_trace = dev_buf; }
extern "C" void mmr_debug_set_attn_ablate(int bits) { mmr::g_t2_ablate = bits; }

// Multi-head scaled-dot-product attention on the 5th-generation tensor cores (tcgen05 + TMEM), for the short sequences
// of this workload (Sq, Sk <= 128, d = 64).  Same contract as attention.cu (pixelbert.py:790-850, modeling.py:325-352):
// additive key mask (1 - m) * -10000, softmax over keys in fp32, merged-head 16-bit context rows.
//
// One work item = one (pair, head).  Persistent CTAs (two per SM) walk the items; inside a CTA
//   warp 4      TMA producer: Q, K, V boxes ([rows x 64], 128-byte swizzle) of the next item into a 2-stage ring, plus
//               the additive mask row
//   warp 5      one thread issues the MMAs:  S[128 x SkP] = Q K^T  (UMMA 128 x SkP x 16, both operands K-major from
//               shared memory, fp32 accumulator in TMEM), and after the softmax  O[128 x 64] = P V  (A = P from shared
//               memory, B = V used MN-major straight from its [keys x 64] tile)
//   warps 0-3   warp-specialised softmax: one thread = one query row (its TMEM lane).  Two passes over the S row in
//               TMEM (max, then exp / sum), P packed to 16 bit into a K-major swizzled tile, no shuffles at all; then
//               the O row is normalised and leaves through the (dead) Q rows as 128-byte coalesced stores.
// Hand-offs are mbarriers: full/empty (ring), s_ready, p_ready, o_ready, o_free.  Rows >= Sq of the 128-row MMA are
// computed on whatever the tile holds and never stored (every output row depends on its own query row only); key
// rows >= Sk are zero (ring zeroed once, TMA never writes them) and carry a -inf mask.
#include <cuda.h>

#include <algorithm>

#include "gemm_common.cuh"
#include "kernels.cuh"

namespace mmr {

constexpr int kTcHeadDim = 64;
constexpr int kTcStages = 2;
constexpr int kTcQBytes = 128 * 128;          // Q tile: 128 rows (UMMA M) x 128 B
constexpr int kTcPBytes = 2 * 128 * 128;      // P tile: two 64-key atoms of 128 rows x 128 B
constexpr int kTcTmemCols = 256;              // S: columns [0, SkP), O: columns [128, 192)
constexpr int kTcThreads = 192;

__device__ __forceinline__ uint32_t tc_sw128(int row, int unit) {   // byte offset inside a swizzled [rows x 128 B] tile
  return uint32_t(row) * 128u + (uint32_t(unit ^ (row & 7)) << 4);
}

template <class E16>
__global__ void __launch_bounds__(kTcThreads, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                    const __grid_constant__ CUtensorMap tmap_v, const int32_t* __restrict__ key_mask,
                    typename E16::T* __restrict__ out, int64_t ldo, int Sq, int Sk, int heads, int n_items,
                    uint32_t idesc_fmt) {
  using T = typename E16::T;
  extern __shared__ __align__(1024) uint8_t smem_tc[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_tc) + 1023) & ~uintptr_t(1023));
  const int SkP = (Sk + 15) & ~15;
  const uint32_t kv_bytes = uint32_t(SkP) * 128u;
  const uint32_t stage_bytes = ((uint32_t(kTcQBytes) + 2 * kv_bytes + uint32_t(SkP) * 4u) + 1023u) & ~1023u;
  uint8_t* p_tile = smem + size_t(kTcStages) * stage_bytes;                        // 1024-aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(p_tile + kTcPBytes);
  uint64_t* full_bar = bars;                  // [2] TMA -> MMA / softmax
  uint64_t* empty_bar = bars + kTcStages;     // [2] PV retired (1) + the four softmax warps done with the Q rows (4)
  uint64_t* s_ready = bars + 2 * kTcStages;   // S complete in TMEM
  uint64_t* p_ready = s_ready + 1;            // P written (4 warps)
  uint64_t* o_ready = s_ready + 2;            // O complete in TMEM
  uint64_t* o_free = s_ready + 3;             // O read out by the 4 warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_ready + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // zero the ring and the P tile once: TMA only ever writes rows < Sq / Sk
  for (uint32_t i = threadIdx.x; i < (kTcStages * stage_bytes + kTcPBytes) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    for (int s = 0; s < kTcStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 5);
    }
    mbar_init(s_ready, 1);
    mbar_init(p_ready, 4);
    mbar_init(o_ready, 1);
    mbar_init(o_free, 4);
    fence_mbar_init();
  }
  if (warp == 5) {
    tmem_alloc(tmem_slot, kTcTmemCols);
    tmem_relinquish();
  }
  fence_proxy_async();   // the zero fill (generic proxy) is ordered before TMA writes / UMMA reads (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + 128u;
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 4) {
    // ===================== TMA producer =====================
    int n = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
      const int stage = n & 1;
      const uint32_t phase = (n >> 1) & 1u;
      const int b = item / heads, h = item - b * heads;
      mbar_wait(&empty_bar[stage], phase ^ 1u);
      uint8_t* st = smem + size_t(stage) * stage_bytes;
      float* sMask = reinterpret_cast<float*>(st + kTcQBytes + 2 * kv_bytes);
      for (int i = lane; i < SkP; i += 32) {
        float m = -INFINITY;   // padding keys (>= Sk) do not exist for the softmax
        if (i < Sk) m = (key_mask == nullptr || key_mask[int64_t(b) * Sk + i] != 0) ? 0.0f : -10000.0f;
        sMask[i] = m;
      }
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_expect_tx(&full_bar[stage], uint32_t(Sq + 2 * Sk) * 128u);
        tma_load_2d(st, &tmap_q, &full_bar[stage], h * kTcHeadDim, b * Sq);
        tma_load_2d(st + kTcQBytes, &tmap_k, &full_bar[stage], h * kTcHeadDim, b * Sk);
        tma_load_2d(st + kTcQBytes + kv_bytes, &tmap_v, &full_bar[stage], h * kTcHeadDim, b * Sk);
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_f16(idesc_fmt, 128, uint32_t(SkP));
      const uint32_t idesc_o = umma_idesc_f16(idesc_fmt, 128, kTcHeadDim) | (1u << 16);   // B (= V) is MN-major
      const uint64_t p_desc0 = umma_desc_k_sw128(smem_u32(p_tile));
      const uint64_t p_desc1 = umma_desc_k_sw128(smem_u32(p_tile + 128 * 128));
      int n = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
        const int stage = n & 1;
        const uint32_t ring_phase = (n >> 1) & 1u, item_phase = uint32_t(n) & 1u;
        const uint32_t st = smem_u32(smem + size_t(stage) * stage_bytes);
        mbar_wait(&full_bar[stage], ring_phase);
        tc_fence_after();
        // S = Q K^T: 4 steps of 16 along d
        const uint64_t q_desc = umma_desc_k_sw128(st), k_desc = umma_desc_k_sw128(st + kTcQBytes);
#pragma unroll
        for (int k = 0; k < kTcHeadDim / kUmmaK; ++k)
          umma_f16(tmem_s, q_desc + uint64_t(2 * k), k_desc + uint64_t(2 * k), idesc_s, k != 0 ? 1u : 0u);
        umma_commit(s_ready);
        // O = P V: SkP / 16 steps along the keys; the previous item's O must have been read out
        mbar_wait(p_ready, item_phase);
        mbar_wait(o_free, item_phase ^ 1u);
        tc_fence_after();
        const uint64_t v_desc = umma_desc_k_sw128(st + kTcQBytes + kv_bytes);
        for (int ks = 0; ks < SkP / kUmmaK; ++ks) {
          const uint64_t pa = (ks < 4 ? p_desc0 : p_desc1) + uint64_t(2 * (ks & 3));
          umma_f16(tmem_o, pa, v_desc + uint64_t(128 * ks), idesc_o, ks != 0 ? 1u : 0u);   // +16 keys = +2048 B
        }
        umma_commit(o_ready);
        umma_commit(&empty_bar[stage]);   // K, V (and Q) of this stage are no longer read by the tensor core
      }
    }
  } else {
    // ===================== softmax / output warps: one thread = one query row =====================
    const int row = warp * 32 + lane;                      // TMEM lane of this thread
    const bool live = warp * 32 < Sq;                      // this warp owns at least one real query row
    const uint32_t lane_off = uint32_t(warp * 32) << 16;
    const int n_chunks = SkP >> 4;
    int n = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
      const int stage = n & 1;
      const uint32_t ring_phase = (n >> 1) & 1u, item_phase = uint32_t(n) & 1u;
      const int b = item / heads, h = item - b * heads;
      uint8_t* st = smem + size_t(stage) * stage_bytes;
      const float* sMask = reinterpret_cast<const float*>(st + kTcQBytes + 2 * (SkP * 128));
      mbar_wait(&full_bar[stage], ring_phase);             // the mask row (generic writes of the producer warp)
      mbar_wait(s_ready, item_phase);
      tc_fence_after();
      float inv = 0.f;
      if (live) {
        // ---- pass A: row maximum of s = S / 8 + mask
        float mx = -INFINITY;
        for (int c = 0; c < n_chunks; ++c) {
          uint32_t v[16];
          tmem_ld_32x16(tmem_s + lane_off + uint32_t(c * 16), v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) mx = fmaxf(mx, fmaf(__uint_as_float(v[j]), 0.125f, sMask[c * 16 + j]));
        }
        // ---- pass B: p = exp(s - max), row sum, 16-bit P into the K-major swizzled tile
        float sum = 0.f;
        for (int c = 0; c < n_chunks; ++c) {
          uint32_t v[16];
          tmem_ld_32x16(tmem_s + lane_off + uint32_t(c * 16), v);
          tmem_ld_wait();
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float e0 = __expf(fmaf(__uint_as_float(v[2 * j]), 0.125f, sMask[c * 16 + 2 * j]) - mx);
            const float e1 = __expf(fmaf(__uint_as_float(v[2 * j + 1]), 0.125f, sMask[c * 16 + 2 * j + 1]) - mx);
            sum += e0 + e1;
            pk[j] = E16::pack(e0, e1);
          }
          // keys [16c, 16c+16) = 32 B = units 2(c&3), 2(c&3)+1 of this row in atom c>>2
          uint8_t* prow = p_tile + (c >> 2) * (128 * 128);
          *reinterpret_cast<uint4*>(prow + tc_sw128(row, 2 * (c & 3))) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(prow + tc_sw128(row, 2 * (c & 3) + 1)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
        inv = 1.0f / sum;
      }
      fence_proxy_async();     // P (generic stores) -> UMMA operand reads (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);

      // ---- O row: normalise, 16 bit, out through this thread's Q row
      mbar_wait(o_ready, item_phase);
      tc_fence_after();
      if (live) {
        uint32_t o0[32], o1[32];
        tmem_ld_32x32(tmem_o + lane_off, o0);
        tmem_ld_32x32(tmem_o + lane_off + 32u, o1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_free);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          *reinterpret_cast<uint4*>(st + tc_sw128(row, u)) = make_uint4(
              E16::pack(__uint_as_float(o0[8 * u]) * inv, __uint_as_float(o0[8 * u + 1]) * inv),
              E16::pack(__uint_as_float(o0[8 * u + 2]) * inv, __uint_as_float(o0[8 * u + 3]) * inv),
              E16::pack(__uint_as_float(o0[8 * u + 4]) * inv, __uint_as_float(o0[8 * u + 5]) * inv),
              E16::pack(__uint_as_float(o0[8 * u + 6]) * inv, __uint_as_float(o0[8 * u + 7]) * inv));
          *reinterpret_cast<uint4*>(st + tc_sw128(row, 4 + u)) = make_uint4(
              E16::pack(__uint_as_float(o1[8 * u]) * inv, __uint_as_float(o1[8 * u + 1]) * inv),
              E16::pack(__uint_as_float(o1[8 * u + 2]) * inv, __uint_as_float(o1[8 * u + 3]) * inv),
              E16::pack(__uint_as_float(o1[8 * u + 4]) * inv, __uint_as_float(o1[8 * u + 5]) * inv),
              E16::pack(__uint_as_float(o1[8 * u + 6]) * inv, __uint_as_float(o1[8 * u + 7]) * inv));
        }
        __syncwarp();
        // coalesced read-back: 8 lanes x 16 B per row, 4 rows per instruction
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = warp * 32 + i * 4 + (lane >> 3), u = lane & 7;
          const uint4 v = *reinterpret_cast<const uint4*>(st + tc_sw128(r, u));
          if (r < Sq) *reinterpret_cast<uint4*>(out + (int64_t(b) * Sq + r) * ldo + h * kTcHeadDim + u * 8) = v;
        }
        fence_proxy_async();   // these generic writes of the Q rows precede the next TMA fill of the stage
      } else {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_free);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[stage]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTcTmemCols);
  }
}

static size_t attention_tc_smem_bytes(int Sk) {
  const int SkP = ((Sk + 15) / 16) * 16;
  const size_t stage = ((size_t(kTcQBytes) + 2 * size_t(SkP) * 128 + size_t(SkP) * 4) + 1023) & ~size_t(1023);
  return 1024 + kTcStages * stage + kTcPBytes + 128;
}

template <class E16>
static mmr_status launch_attention_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                      const int32_t* key_mask, void* out, int64_t ldo, int B, int Sq, int Sk, int heads,
                                      int dtype, cudaStream_t stream) {
  using T = typename E16::T;
  auto kern = attention_tc_kernel<E16>;
  static bool configured = false;
  if (!configured) {
    MMR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(attention_tc_smem_bytes(128))));
    configured = true;
  }
  const int ek = dtype == MMR_DT_BF16 ? 1 : 0;
  CUtensorMap tq, tk, tv;
  MMR_TRY(make_tmap_ex(&tq, q, int64_t(B) * Sq, int64_t(heads) * kTcHeadDim, ldq, ek, kTcHeadDim, Sq, 128));
  MMR_TRY(make_tmap_ex(&tk, k, int64_t(B) * Sk, int64_t(heads) * kTcHeadDim, ldk, ek, kTcHeadDim, Sk, 128));
  MMR_TRY(make_tmap_ex(&tv, v, int64_t(B) * Sk, int64_t(heads) * kTcHeadDim, ldv, ek, kTcHeadDim, Sk, 128));
  const size_t smem = attention_tc_smem_bytes(Sk);
  const int per_sm = smem <= 113 * 1024 ? 2 : 1;   // two CTAs share an SM's 227 KB (and its 512 TMEM columns)
  const int n_items = B * heads;
  const int grid = std::min(n_items, sm_count() * per_sm);
  MMR_CUDA_OK(launch_pdl(kern, dim3(grid), dim3(kTcThreads), smem, stream, tq, tk, tv, key_mask, static_cast<T*>(out),
                         ldo, Sq, Sk, heads, n_items, uint32_t(dtype)));
  return MMR_OK;
}

// Arguments are validated by mmr::attention (attention.cu); on top of those this path needs 16-byte aligned output rows.
bool attention_tc_eligible(const void* out16, int64_t ldo) {
  return tuning(MMR_TUNE_ATTN_TC) == 1 && ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(out16) & 15) == 0;
}
mmr_status attention_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                        const int32_t* key_mask, void* out16, int64_t ldo, int B, int Sq, int Sk, int heads, int dtype,
                        cudaStream_t stream) {
  if (dtype == MMR_DT_BF16)
    return launch_attention_tc<BF16>(q, ldq, k, ldk, v, ldv, key_mask, out16, ldo, B, Sq, Sk, heads, dtype, stream);
  return launch_attention_tc<FP16>(q, ldq, k, ldk, v, ldv, key_mask, out16, ldo, B, Sq, Sk, heads, dtype, stream);
}

}  // namespace mmr

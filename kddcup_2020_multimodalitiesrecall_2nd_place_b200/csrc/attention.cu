// Multi-head scaled-dot-product attention for the short sequences of this workload (Sq, Sk <= 128, d = 64).
//
// Replaces pixelbert.py:790-850 (transpose_for_scores, QK^T / sqrt(d), additive mask, softmax, PV, merge heads)
// and lxmert modeling.py:325-352 (BertAttention, self- and cross-), i.e. everything between the QKV projection
// and the output projection.  ~1.5 % of the layer FLOPs: one CTA per (pair, head) keeps Q, K, V of that head in
// shared memory, each warp owns 16 query rows, scores / softmax stay in fp32 registers, the two small matmuls
// run on the warp-level tensor path (mma.sync m16n8k16, fp32 accumulate).  Semantics kept from the reference:
// additive key mask (1 - m) * -10000 (NOT -inf), queries never masked, softmax over keys in fp32.
#include <cuda.h>

#include <algorithm>

#include "gemm_common.cuh"
#include "kernels.cuh"

namespace mmr {

constexpr int kHeadDim = 64;
constexpr int kPitch = 72;       // smem row pitch in elements (144 B): conflict-free ldmatrix
constexpr int kMaxSeq = 128;

// The mma.sync kernels below (one CTA per (pair, head), and its persistent TMA-pipelined sibling) lost to the tcgen05
// kernel of attention_tc2.cu (31.7 / 40 us against 24-25 us at B = 256, S = 68) and are lab notes: they are compiled
// only into MMR_EXPERIMENTAL builds (csrc/build.py, environment MMR_EXPERIMENTAL=1).
#ifdef MMR_EXPERIMENTAL

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                                  uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
// 16-byte global -> shared copy without a register stage; src_bytes = 0 zero-fills the destination.
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src),
               "r"(src_bytes)
               : "memory");
}
template <class E16>
__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                          uint32_t b0, uint32_t b1) {
  if constexpr (E16::kFmt == MMR_DT_BF16) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  } else {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
}

// kMaxKT = compile-time bound on 8-key score tiles (keys padded to 16): the score / probability fragments live in
// registers, so instantiating for the actual key count (40 / 72 / 104 / 128) instead of the maximum is what keeps
// enough CTAs resident per SM to hide the global-load latency of these short, synchronous CTAs.
template <class E16, int kMaxKT>
__global__ void __launch_bounds__(256)
attention_kernel(const typename E16::T* __restrict__ q, int64_t ldq, const typename E16::T* __restrict__ k,
                 int64_t ldk, const typename E16::T* __restrict__ v, int64_t ldv,
                 const int32_t* __restrict__ key_mask, typename E16::T* __restrict__ out, int64_t ldo, int Sq,
                 int Sk) {
  using T = typename E16::T;
  extern __shared__ __align__(16) uint8_t smem_att[];
  pdl_wait();
  pdl_launch_dependents();
  const int h = blockIdx.x, b = blockIdx.y;
  const int nwarps = blockDim.x >> 5;
  const int SqP = nwarps * 16;
  const int SkP = (Sk + 15) & ~15;
  T* sQ = reinterpret_cast<T*>(smem_att);
  T* sK = sQ + SqP * kPitch;
  T* sV = sK + SkP * kPitch;
  float* sMask = reinterpret_cast<float*>(sV + SkP * kPitch);  // [SkP] additive mask

  // ---- cooperative load of this (pair, head): 8 x 16-byte chunks per 64-wide row.  cp.async (LDGSTS) puts every
  // chunk in flight at once, without a register round trip per loop iteration; padding rows are zero-filled
  // (src-size 0).
  const int tid = threadIdx.x;
  for (int i = tid; i < SqP * 8; i += blockDim.x) {
    const int r = i >> 3, c = (i & 7) * 8;
    const int rr = r < Sq ? r : 0;
    cp_async_16(sQ + r * kPitch + c, q + (int64_t(b) * Sq + rr) * ldq + h * kHeadDim + c, r < Sq ? 16 : 0);
  }
  for (int i = tid; i < SkP * 8; i += blockDim.x) {
    const int r = i >> 3, c = (i & 7) * 8;
    const int rr = r < Sk ? r : 0;
    const int nbytes = r < Sk ? 16 : 0;
    cp_async_16(sK + r * kPitch + c, k + (int64_t(b) * Sk + rr) * ldk + h * kHeadDim + c, nbytes);
    cp_async_16(sV + r * kPitch + c, v + (int64_t(b) * Sk + rr) * ldv + h * kHeadDim + c, nbytes);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int i = tid; i < SkP; i += blockDim.x) {
    float m = -INFINITY;  // padding keys (>= Sk) do not exist for the softmax
    if (i < Sk) m = (key_mask == nullptr || key_mask[int64_t(b) * Sk + i] != 0) ? 0.0f : -10000.0f;
    sMask[i] = m;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int q0 = warp * 16;
  if (q0 >= Sq) return;
  const int nkt = SkP >> 3;   // 8-key score tiles
  const int nks = SkP >> 4;   // 16-key steps for P.V

  // ---- S = Q K^T (fp32 accumulate) ----
  uint32_t qa[4][4];
  {
    // A fragment rows: lanes 0-15 -> rows q0 + lane (cols +0), lanes 16-31 -> rows q0 + lane-16 (cols +8)
    const int r = q0 + (lane & 15), cofs = (lane >> 4) * 8;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      ldmatrix_x4(smem_u32(sQ + r * kPitch + ks * 16 + cofs), qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);
    }
  }
  float s[kMaxKT][4];
#pragma unroll
  for (int nt = 0; nt < kMaxKT; ++nt) { s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f; }
#pragma unroll
  for (int np = 0; np < kMaxKT / 2; ++np) {
    if (np * 2 < nkt) {
      // B fragments for keys [16np, 16np+16): matrices (keys 0-7, d lo), (keys 0-7, d hi), (keys 8-15, d lo), (.., d hi)
      const int kr = np * 16 + (lane & 7) + ((lane >> 4) << 3);
      const int dofs = ((lane >> 3) & 1) * 8;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4(smem_u32(sK + kr * kPitch + ks * 16 + dofs), b0, b1, b2, b3);
        mma_16816<E16>(s[2 * np], qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3], b0, b1);
        mma_16816<E16>(s[2 * np + 1], qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3], b2, b3);
      }
    }
  }

  // ---- scale, additive mask, softmax over keys (rows g and g+8 of this warp's 16) ----
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < kMaxKT; ++nt) {
    if (nt < nkt) {
      const float m0 = sMask[nt * 8 + 2 * t4], m1 = sMask[nt * 8 + 2 * t4 + 1];
      s[nt][0] = fmaf(s[nt][0], 0.125f, m0);
      s[nt][1] = fmaf(s[nt][1], 0.125f, m1);
      s[nt][2] = fmaf(s[nt][2], 0.125f, m0);
      s[nt][3] = fmaf(s[nt][3], 0.125f, m1);
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  float sum0 = 0.f, sum1 = 0.f;
  uint32_t pa[kMaxKT][2];  // P as 16-bit A fragments
#pragma unroll
  for (int nt = 0; nt < kMaxKT; ++nt) {
    if (nt < nkt) {
      const float e0 = __expf(s[nt][0] - mx0), e1 = __expf(s[nt][1] - mx0);
      const float e2 = __expf(s[nt][2] - mx1), e3 = __expf(s[nt][3] - mx1);
      sum0 += e0 + e1;
      sum1 += e2 + e3;
      pa[nt][0] = E16::pack(e0, e1);
      pa[nt][1] = E16::pack(e2, e3);
    }
  }
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);

  // ---- O = P V ----
  float o[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) { o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f; }
#pragma unroll
  for (int ks = 0; ks < kMaxKT / 2; ++ks) {
    if (ks < nks) {
      // V^T fragments via ldmatrix.trans: matrices (keys lo, d n0), (keys hi, d n0), (keys lo, d n0+8), (keys hi, d n0+8)
      const int vr = ks * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
      const int dofs = (lane >> 4) * 8;
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4_trans(smem_u32(sV + vr * kPitch + dp * 16 + dofs), b0, b1, b2, b3);
        mma_16816<E16>(o[2 * dp], pa[2 * ks][0], pa[2 * ks][1], pa[2 * ks + 1][0], pa[2 * ks + 1][1], b0, b1);
        mma_16816<E16>(o[2 * dp + 1], pa[2 * ks][0], pa[2 * ks][1], pa[2 * ks + 1][0], pa[2 * ks + 1][1], b2, b3);
      }
    }
  }

  // ---- normalise and store the merged-head context rows ----
  const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
  const int r0 = q0 + g, r1 = q0 + g + 8;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int col = h * kHeadDim + nt * 8 + 2 * t4;
    if (r0 < Sq) {
      *reinterpret_cast<uint32_t*>(out + (int64_t(b) * Sq + r0) * ldo + col) = E16::pack(o[nt][0] * inv0, o[nt][1] * inv0);
    }
    if (r1 < Sq) {
      *reinterpret_cast<uint32_t*>(out + (int64_t(b) * Sq + r1) * ldo + col) = E16::pack(o[nt][2] * inv1, o[nt][3] * inv1);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Persistent, TMA-pipelined version (default).  The kernel above spends half of every CTA's life waiting for its own
// loads (3 CTAs of 5 warps per SM by registers, each doing load -> sync -> compute once).  Here a CTA stays resident,
// a producer warp streams the (pair, head) work items through a ring of shared-memory stages with TMA
// (128-byte-swizzled [rows x 64] boxes of Q, K and V straight out of the fused QKV matrix) and the compute warps
// never wait for memory once the ring is primed.  Same arithmetic, fragment for fragment, as the kernel above; the
// context rows leave through the warp's own (dead) Q rows as 128-byte coalesced stores.
constexpr int kAttStages = 3;

__device__ __forceinline__ uint32_t sw128_off(int row, int chunk) {   // byte offset inside a swizzled [rows x 128 B] tile
  return uint32_t(row) * 128u + (uint32_t(chunk ^ (row & 7)) << 4);
}

template <class E16, int kMaxKT>
__global__ void __launch_bounds__(288)
attention_tma_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                     const __grid_constant__ CUtensorMap tmap_v, const int32_t* __restrict__ key_mask,
                     typename E16::T* __restrict__ out, int64_t ldo, int Sq, int Sk, int heads, int n_items) {
  using T = typename E16::T;
  extern __shared__ __align__(1024) uint8_t smem_att2[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_att2) + 1023) & ~uintptr_t(1023));
  const int nwarps = (blockDim.x >> 5) - 1;         // compute warps; the last warp is the producer
  const int SqP = nwarps * 16;
  const int SkP = (Sk + 15) & ~15;
  const uint32_t q_bytes = uint32_t(SqP) * 128u, kv_bytes = uint32_t(SkP) * 128u;
  const uint32_t stage_bytes = ((q_bytes + 2 * kv_bytes + uint32_t(SkP) * 4u) + 1023u) & ~1023u;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + size_t(kAttStages) * stage_bytes);
  uint64_t* empty_bar = full_bar + kAttStages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // zero the ring once: TMA only ever writes rows < Sq / Sk, so the padding rows of K and V stay zero for good
  for (uint32_t i = threadIdx.x; i < kAttStages * stage_bytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    for (int s = 0; s < kAttStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], uint32_t(nwarps));
    }
    fence_mbar_init();
  }
  fence_proxy_async();   // the zero fill (generic proxy) is ordered before the TMA writes (async proxy)
  __syncthreads();
  pdl_wait();
  pdl_launch_dependents();

  if (warp == nwarps) {
    // ===================== producer warp =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int b = item / heads, h = item - b * heads;
      mbar_wait(&empty_bar[stage], phase ^ 1u);
      uint8_t* st = smem + size_t(stage) * stage_bytes;
      float* sMask = reinterpret_cast<float*>(st + q_bytes + 2 * kv_bytes);
      for (int i = lane; i < SkP; i += 32) {
        float m = -INFINITY;   // padding keys (>= Sk) do not exist for the softmax
        if (i < Sk) m = (key_mask == nullptr || key_mask[int64_t(b) * Sk + i] != 0) ? 0.0f : -10000.0f;
        sMask[i] = m;
      }
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_expect_tx(&full_bar[stage], uint32_t(Sq + 2 * Sk) * 128u);
        tma_load_2d(st, &tmap_q, &full_bar[stage], h * kHeadDim, b * Sq);
        tma_load_2d(st + q_bytes, &tmap_k, &full_bar[stage], h * kHeadDim, b * Sk);
        tma_load_2d(st + q_bytes + kv_bytes, &tmap_v, &full_bar[stage], h * kHeadDim, b * Sk);
      }
      if (++stage == kAttStages) { stage = 0; phase ^= 1u; }
    }
    return;
  }

  // ===================== compute warps: 16 query rows each =====================
  const int g = lane >> 2, t4 = lane & 3;
  const int q0 = warp * 16;
  const int nkt = SkP >> 3;   // 8-key score tiles
  const int nks = SkP >> 4;   // 16-key steps for P.V
  int stage = 0;
  uint32_t phase = 0;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int b = item / heads, h = item - b * heads;
    const uint32_t sQ = smem_u32(smem + size_t(stage) * stage_bytes);
    const uint32_t sK = sQ + q_bytes, sV = sK + kv_bytes;
    const float* sMask = reinterpret_cast<const float*>(smem + size_t(stage) * stage_bytes + q_bytes + 2 * kv_bytes);
    mbar_wait(&full_bar[stage], phase);

    // ---- S = Q K^T (fp32 accumulate) ----
    uint32_t qa[4][4];
    {
      const int r = q0 + (lane & 15), c = lane >> 4;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        ldmatrix_x4(sQ + sw128_off(r, ks * 2 + c), qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);
    }
    float s[kMaxKT][4];
#pragma unroll
    for (int nt = 0; nt < kMaxKT; ++nt) { s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f; }
#pragma unroll
    for (int np = 0; np < kMaxKT / 2; ++np) {
      if (np * 2 < nkt) {
        const int kr = np * 16 + (lane & 7) + ((lane >> 4) << 3);
        const int c = (lane >> 3) & 1;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t b0, b1, b2, b3;
          ldmatrix_x4(sK + sw128_off(kr, ks * 2 + c), b0, b1, b2, b3);
          mma_16816<E16>(s[2 * np], qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3], b0, b1);
          mma_16816<E16>(s[2 * np + 1], qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3], b2, b3);
        }
      }
    }

    // ---- scale, additive mask, softmax over keys (rows g and g+8 of this warp's 16) ----
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < kMaxKT; ++nt) {
      if (nt < nkt) {
        const float2 m = *reinterpret_cast<const float2*>(sMask + nt * 8 + 2 * t4);
        s[nt][0] = fmaf(s[nt][0], 0.125f, m.x);
        s[nt][1] = fmaf(s[nt][1], 0.125f, m.y);
        s[nt][2] = fmaf(s[nt][2], 0.125f, m.x);
        s[nt][3] = fmaf(s[nt][3], 0.125f, m.y);
        mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float sum0 = 0.f, sum1 = 0.f;
    uint32_t pa[kMaxKT][2];  // P as 16-bit A fragments
#pragma unroll
    for (int nt = 0; nt < kMaxKT; ++nt) {
      if (nt < nkt) {
        const float e0 = __expf(s[nt][0] - mx0), e1 = __expf(s[nt][1] - mx0);
        const float e2 = __expf(s[nt][2] - mx1), e3 = __expf(s[nt][3] - mx1);
        sum0 += e0 + e1;
        sum1 += e2 + e3;
        pa[nt][0] = E16::pack(e0, e1);
        pa[nt][1] = E16::pack(e2, e3);
      }
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);

    // ---- O = P V ----
    float o[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < kMaxKT / 2; ++ks) {
      if (ks < nks) {
        const int vr = ks * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
        const int c = lane >> 4;
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {
          uint32_t b0, b1, b2, b3;
          ldmatrix_x4_trans(sV + sw128_off(vr, dp * 2 + c), b0, b1, b2, b3);
          mma_16816<E16>(o[2 * dp], pa[2 * ks][0], pa[2 * ks][1], pa[2 * ks + 1][0], pa[2 * ks + 1][1], b0, b1);
          mma_16816<E16>(o[2 * dp + 1], pa[2 * ks][0], pa[2 * ks][1], pa[2 * ks + 1][0], pa[2 * ks + 1][1], b2, b3);
        }
      }
    }

    // ---- normalise; context rows leave through this warp's own Q rows (dead since the qa loads) ----
    const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
    __syncwarp();   // every lane's ldmatrix of the Q rows has completed
    uint8_t* qrows = smem + size_t(stage) * stage_bytes;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      *reinterpret_cast<uint32_t*>(qrows + sw128_off(q0 + g, nt) + 4 * t4) = E16::pack(o[nt][0] * inv0, o[nt][1] * inv0);
      *reinterpret_cast<uint32_t*>(qrows + sw128_off(q0 + g + 8, nt) + 4 * t4) = E16::pack(o[nt][2] * inv1, o[nt][3] * inv1);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = q0 + i * 4 + (lane >> 3), c = lane & 7;
      const uint4 v = *reinterpret_cast<const uint4*>(qrows + sw128_off(r, c));
      if (r < Sq) *reinterpret_cast<uint4*>(out + (int64_t(b) * Sq + r) * ldo + h * kHeadDim + c * 8) = v;
    }
    fence_proxy_async();   // this warp's generic-proxy writes to its Q rows precede the next TMA fill of the stage
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[stage]);   // this warp is done with the stage (K, V, mask and its Q rows)
    if (++stage == kAttStages) { stage = 0; phase ^= 1u; }
  }
}

static size_t attention2_smem_bytes(int Sq, int Sk) {
  const int SqP = ((Sq + 15) / 16) * 16, SkP = ((Sk + 15) / 16) * 16;
  const size_t stage = ((size_t(SqP) * 128 + 2 * size_t(SkP) * 128 + size_t(SkP) * 4) + 1023) & ~size_t(1023);
  return 1024 + kAttStages * stage + 2 * kAttStages * 8;
}

template <class E16, int kMaxKT>
static mmr_status launch_attention2(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                    const int32_t* key_mask, void* out, int64_t ldo, int B, int Sq, int Sk, int heads,
                                    int dtype, cudaStream_t stream) {
  using T = typename E16::T;
  auto kern = attention_tma_kernel<E16, kMaxKT>;
  static bool configured = false;
  if (!configured) {
    MMR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     int(attention2_smem_bytes(kMaxSeq, kMaxSeq))));
    configured = true;
  }
  const int ek = dtype == MMR_DT_BF16 ? 1 : 0;
  CUtensorMap tq, tk, tv;
  MMR_TRY(make_tmap_ex(&tq, q, int64_t(B) * Sq, int64_t(heads) * kHeadDim, ldq, ek, kHeadDim, Sq, 128));
  MMR_TRY(make_tmap_ex(&tk, k, int64_t(B) * Sk, int64_t(heads) * kHeadDim, ldk, ek, kHeadDim, Sk, 128));
  MMR_TRY(make_tmap_ex(&tv, v, int64_t(B) * Sk, int64_t(heads) * kHeadDim, ldv, ek, kHeadDim, Sk, 128));
  const int nwarps = (Sq + 15) / 16;
  const size_t smem = attention2_smem_bytes(Sq, Sk);
  const int per_sm = int(std::min<size_t>(4, (227 * 1024) / smem));   // resident CTAs per SM (shared memory bound)
  const int n_items = B * heads;
  const int grid = std::min(n_items, sm_count() * std::max(per_sm, 1));
  MMR_CUDA_OK(launch_pdl(kern, dim3(grid), dim3((nwarps + 1) * 32), smem, stream, tq, tk, tv, key_mask,
                         static_cast<T*>(out), ldo, Sq, Sk, heads, n_items));
  return MMR_OK;
}

static size_t attention_smem_bytes(int Sq, int Sk) {
  const int SqP = ((Sq + 15) / 16) * 16, SkP = ((Sk + 15) / 16) * 16;
  return size_t(SqP + 2 * SkP) * kPitch * 2 + size_t(SkP) * 4;
}

template <class E16, int kMaxKT>
static mmr_status launch_attention(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                                   int64_t ldv, const int32_t* key_mask, void* out, int64_t ldo, int B, int Sq,
                                   int Sk, int heads, cudaStream_t stream) {
  using T = typename E16::T;
  auto kern = attention_kernel<E16, kMaxKT>;
  static bool configured = false;
  if (!configured) {
    MMR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     int(attention_smem_bytes(kMaxSeq, kMaxSeq))));
    configured = true;
  }
  const int nwarps = (Sq + 15) / 16;
  dim3 grid(heads, B);
  MMR_CUDA_OK(launch_pdl(kern, grid, dim3(nwarps * 32), attention_smem_bytes(Sq, Sk), stream, static_cast<const T*>(q),
                         ldq, static_cast<const T*>(k), ldk, static_cast<const T*>(v), ldv, key_mask,
                         static_cast<T*>(out), ldo, Sq, Sk));
  return MMR_OK;
}

#endif  // MMR_EXPERIMENTAL

mmr_status attention(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                     const int32_t* key_mask, void* out16, int64_t ldo, int B, int Sq, int Sk, int heads,
                     int dtype, cudaStream_t stream) {
  MMR_TRY(require_sm100());
  MMR_REQUIRE(q && k && v && out16, "mmr_attention: null pointer");
  MMR_REQUIRE(B > 0 && heads > 0, "mmr_attention: B=%d heads=%d", B, heads);
  MMR_REQUIRE(Sq > 0 && Sq <= kMaxSeq && Sk > 0 && Sk <= kMaxSeq, "mmr_attention: Sq=%d Sk=%d must be in [1,%d]",
              Sq, Sk, kMaxSeq);
  MMR_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 2 == 0,
              "mmr_attention: row strides must keep 16-byte alignment of 64-wide head slices");
  MMR_REQUIRE(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v)) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(out16) & 3) == 0,
              "mmr_attention: q/k/v must be 16-byte aligned");
  MMR_REQUIRE(dtype == MMR_DT_BF16 || dtype == MMR_DT_FP16, "mmr_attention: bad dtype %d", dtype);
#ifndef MMR_EXPERIMENTAL
  MMR_REQUIRE(ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(out16) & 15) == 0,
              "mmr_attention: output rows must be 16-byte aligned (the kernels that take any alignment are built only "
              "with MMR_EXPERIMENTAL)");
  return attention_tc2(q, ldq, k, ldk, v, ldv, key_mask, out16, ldo, B, Sq, Sk, heads, dtype, stream);
}
#else
  if (attention_tc2_eligible(out16, ldo))
    return attention_tc2(q, ldq, k, ldk, v, ldv, key_mask, out16, ldo, B, Sq, Sk, heads, dtype, stream);
  if (attention_tc_eligible(out16, ldo))
    return attention_tc(q, ldq, k, ldk, v, ldv, key_mask, out16, ldo, B, Sq, Sk, heads, dtype, stream);
  // the TMA kernel needs out16 rows 16-byte aligned (vector stores) on top of the operand alignment checked above
  const bool tma_path = tuning(MMR_TUNE_ATTN_TMA) != 0 && ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(out16) & 15) == 0;
#define MMR_ATT(E, KT)                                                                                               \
  return tma_path ? launch_attention2<E, KT>(q, ldq, k, ldk, v, ldv, key_mask, out16, ldo, B, Sq, Sk, heads, dtype,  \
                                             stream)                                                                 \
                  : launch_attention<E, KT>(q, ldq, k, ldk, v, ldv, key_mask, out16, ldo, B, Sq, Sk, heads, stream)
  const int kt = ((Sk + 15) / 16) * 2;   // 8-key tiles after padding the keys to a multiple of 16
  if (dtype == MMR_DT_BF16) {
    if (kt <= 6) MMR_ATT(BF16, 6);
    if (kt <= 10) MMR_ATT(BF16, 10);
    if (kt <= 14) MMR_ATT(BF16, 14);
    MMR_ATT(BF16, 16);
  }
  if (kt <= 6) MMR_ATT(FP16, 6);
  if (kt <= 10) MMR_ATT(FP16, 10);
  if (kt <= 14) MMR_ATT(FP16, 14);
  MMR_ATT(FP16, 16);
#undef MMR_ATT
}
#endif

}  // namespace mmr

extern "C" mmr_status mmr_attention(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                                    int64_t ldv, const int32_t* key_mask, void* out16, int64_t ldo, int B, int Sq,
                                    int Sk, int heads, int dtype, void* stream) {
  return mmr::attention(q, ldq, k, ldk, v, ldv, key_mask, out16, ldo, B, Sq, Sk, heads, dtype,
                        static_cast<cudaStream_t>(stream));
}

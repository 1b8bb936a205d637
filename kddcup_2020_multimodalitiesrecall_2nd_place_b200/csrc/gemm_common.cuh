// Pieces shared by the single-CTA and the CTA-pair tcgen05 GEMM kernels: parameters, fused activations and the
// per-warp epilogue (TMEM -> registers -> bias / activation / residual -> vector stores).
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace mmr {

constexpr int kBN = 256;       // max columns per tile (= UMMA N); narrower N-tail tiles use UMMA N = 16k
constexpr int kBK = 64;        // K per stage: 64 x 2 B = one 128-byte swizzle atom row
constexpr int kUmmaK = 16;     // K per tcgen05.mma for 16-bit operands
constexpr int kEpiWarps = 8;
constexpr int kGemmThreads = 32 * (2 + kEpiWarps);
constexpr int kTmemCols = 512;  // 2 accumulators x 256 fp32 columns

struct GemmParams {
  int M, N, K;
  const float* bias;      // [N] or null
  const float* residual;  // [M, ldr] or null
  int64_t ldr;
  void* out16;            // [M, ldo16] or null
  int64_t ldo16;
  float* out32;           // [M, ldo32] or null
  int64_t ldo32;
  uint32_t idesc_fmt;     // 0 fp16 / 1 bf16
  uint32_t w_box_rows;    // rows of the W tensor-map box (256, or N when N < 256)
};

template <int ACT>
__device__ __forceinline__ float apply_act(float x) {
  if constexpr (ACT == MMR_ACT_RELU) return fmaxf(x, 0.0f);
  if constexpr (ACT == MMR_ACT_GELU_TANH) return gelu_tanh_f(x);
  if constexpr (ACT == MMR_ACT_GELU_ERF) return gelu_erf_f(x);
  if constexpr (ACT == MMR_ACT_TANH) return tanh_precise_f(x);
  return x;
}

constexpr int kEpiStageFloats = 32 * 32;                       // one 32-row x 32-column fp32 chunk per warp
constexpr int kEpiSmemBytes = kEpiWarps * kEpiStageFloats * 4;  // 32 KB

// One epilogue warp's share of a 128-row x bn-column accumulator: TMEM lanes [32*quarter, +32), columns
// [128*half, 128*half + 128) in chunks of 32.  `taddr_row` carries the lane and accumulator offsets, `row0` is the
// global row of lane 0.
//
// TMEM hands every thread one ROW of the chunk, which is the worst possible shape for global memory (each warp
// instruction would touch 32 different 128-byte lines).  So: phase 1 applies bias + activation per row and parks the
// chunk in this warp's shared-memory stage (XOR-swizzled float4 slots, conflict-free both ways); phase 2 re-reads
// it row-major, so that every global instruction of the residual read and of the fp32 / 16-bit stores covers
// whole contiguous row segments (4 rows x 128 B, or 8 rows x 64 B on the 16-bit-only path).
template <int ACT, class E16>
__device__ __forceinline__ void epilogue_warp(const GemmParams& p, uint32_t taddr_row, int row0, int col_tile0, int bn,
                                              int half, float* stage) {
  const int lane = threadIdx.x & 31;
  const bool only16 = p.out32 == nullptr && p.residual == nullptr;
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    const int col_in_tile = half * 128 + c * 32;
    if (col_in_tile >= bn) break;  // warp-uniform
    const int col0 = col_tile0 + col_in_tile;
    const int ncols = min(32, bn - col_in_tile);  // multiple of 16 (N % 16 == 0)
    // Residual rows of phase 2, fetched first so that their L2 latency hides behind the TMEM load and phase 1.
    // (out32 may alias residual — the model driver updates the residual stream in place — so these loads must be
    // issued explicitly before any store of this chunk; each element is read and written by the same thread.)
    float4 res[8];
    if (p.residual != nullptr) {
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int grow = row0 + it * 4 + (lane >> 3), c4 = lane & 7;
        res[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (grow < p.M && 4 * c4 < ncols)
          res[it] = *reinterpret_cast<const float4*>(p.residual + int64_t(grow) * p.ldr + col0 + 4 * c4);
      }
    }
    uint32_t r[32];
    tmem_ld_32x32(taddr_row + uint32_t(col_in_tile), r);
    tmem_ld_wait();
    // ---- phase 1: this thread's row -> bias, activation -> swizzled stage
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 v = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                             __uint_as_float(r[4 * j + 3]));
      if (p.bias != nullptr && 4 * j < ncols) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 4 * j));
        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
      }
      v.x = apply_act<ACT>(v.x); v.y = apply_act<ACT>(v.y); v.z = apply_act<ACT>(v.z); v.w = apply_act<ACT>(v.w);
      *reinterpret_cast<float4*>(stage + lane * 32 + ((j ^ (lane & 7)) << 2)) = v;
    }
    __syncwarp();
    // ---- phase 2: row-major read-back, coalesced global traffic
    if (only16) {
      typename E16::T* o16 = reinterpret_cast<typename E16::T*>(p.out16);
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int rl = it * 8 + (lane >> 2), c8 = lane & 3;
        const float4 x0 = *reinterpret_cast<const float4*>(stage + rl * 32 + (((2 * c8) ^ (rl & 7)) << 2));
        const float4 x1 = *reinterpret_cast<const float4*>(stage + rl * 32 + (((2 * c8 + 1) ^ (rl & 7)) << 2));
        const int grow = row0 + rl;
        if (grow < p.M && 8 * c8 < ncols) {
          uint4 q;
          q.x = E16::pack(x0.x, x0.y); q.y = E16::pack(x0.z, x0.w);
          q.z = E16::pack(x1.x, x1.y); q.w = E16::pack(x1.z, x1.w);
          *reinterpret_cast<uint4*>(o16 + int64_t(grow) * p.ldo16 + col0 + 8 * c8) = q;
        }
      }
    } else {
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int rl = it * 4 + (lane >> 3), c4 = lane & 7;
        float4 x = *reinterpret_cast<const float4*>(stage + rl * 32 + ((c4 ^ (rl & 7)) << 2));
        const int grow = row0 + rl;
        if (grow < p.M && 4 * c4 < ncols) {
          const int gcol = col0 + 4 * c4;
          if (p.residual != nullptr) {
            x.x += res[it].x; x.y += res[it].y; x.z += res[it].z; x.w += res[it].w;
          }
          if (p.out32 != nullptr) *reinterpret_cast<float4*>(p.out32 + int64_t(grow) * p.ldo32 + gcol) = x;
          if (p.out16 != nullptr) {
            uint2 q;
            q.x = E16::pack(x.x, x.y); q.y = E16::pack(x.z, x.w);
            *reinterpret_cast<uint2*>(reinterpret_cast<typename E16::T*>(p.out16) + int64_t(grow) * p.ldo16 + gcol) = q;
          }
        }
      }
    }
    __syncwarp();
  }
}

// 2-D tensor map over a row-major 16-bit matrix [rows, cols] with row stride ld (elements); box = 64 x box_rows,
// 128-byte swizzle (must match umma_desc_k_sw128).
mmr_status make_tmap_2d(void* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int dtype);
mmr_status make_tmap_ex(void* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int elem_kind,
                        int box_cols, int box_rows, int swizzle_bytes);
int sm_count();

// CTA-pair kernel (gemm2_sm100.cu): used when N is a multiple of 256 and M spans more than one 128-row block.
bool gemm_pair_eligible(int M, int N, int K);
mmr_status gemm_pair(const void* A16, int64_t lda, const void* W16, int64_t ldw, const GemmParams& p, int act,
                     int dtype, cudaStream_t stream);

// 16-bit-output variant with a TMA-store epilogue and a split tail wave (gemm16_sm100.cu).
bool gemm_pair16_eligible(int M, int N, int K, const float* residual, const void* out16, const float* out32);
mmr_status gemm_pair16(const void* A16, int64_t lda, const void* W16, int64_t ldw, int M, int N, int K,
                       const float* bias, void* out16, int64_t ldo16, int act, int dtype, cudaStream_t stream);

// the same with a second weight matrix for the rows from split_row on (see gemm16_sm100.cu)
mmr_status gemm_pair16_2w(const void* A16, int64_t lda, const void* W16, const void* W16b, int64_t ldw, int M, int N,
                          int K, const float* bias, const float* biasb, int split_row, void* out16, int64_t ldo16,
                          int act, int dtype, cudaStream_t stream);

}  // namespace mmr

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

// One epilogue warp's share of a 128-row x bn-column accumulator: lanes [32*quarter, +32) (one row per thread),
// columns [128*half, 128*half + 128) in chunks of 32.  `taddr_row` already carries the lane and accumulator offsets.
template <int ACT, class E16>
__device__ __forceinline__ void epilogue_warp(const GemmParams& p, uint32_t taddr_row, int row, int col_tile0, int bn,
                                              int half) {
  const bool row_ok = row < p.M;
#pragma unroll 1
for (int c = 0; c < 4; ++c) {
  const int col_in_tile = half * 128 + c * 32;
  if (col_in_tile >= bn) break;  // warp-uniform
  uint32_t r[32];
  tmem_ld_32x32(taddr_row + uint32_t(col_in_tile), r);
  tmem_ld_wait();
  const int col0 = col_tile0 + col_in_tile;
  const int ncols = min(32, bn - col_in_tile);  // multiple of 16 (N % 16 == 0)
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  if (p.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      if (j < ncols) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = apply_act<ACT>(v[j]);
  if (row_ok) {
    if (p.residual != nullptr) {
      const float* rp = p.residual + int64_t(row) * p.ldr + col0;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        if (j < ncols) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(rp + j));
          v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
        }
      }
    }
    if (p.out32 != nullptr) {
      float* op = p.out32 + int64_t(row) * p.ldo32 + col0;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        if (j < ncols) *reinterpret_cast<float4*>(op + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    }
    if (p.out16 != nullptr) {
      typename E16::T* op = reinterpret_cast<typename E16::T*>(p.out16) + int64_t(row) * p.ldo16 + col0;
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        if (j < ncols) {
          uint4 q;
          q.x = E16::pack(v[j], v[j + 1]);
          q.y = E16::pack(v[j + 2], v[j + 3]);
          q.z = E16::pack(v[j + 4], v[j + 5]);
          q.w = E16::pack(v[j + 6], v[j + 7]);
          *reinterpret_cast<uint4*>(op + j) = q;
        }
      }
    }
  }
}
}

// 2-D tensor map over a row-major 16-bit matrix [rows, cols] with row stride ld (elements); box = 64 x box_rows,
// 128-byte swizzle (must match umma_desc_k_sw128).
mmr_status make_tmap_2d(void* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int dtype);
int sm_count();

// CTA-pair kernel (gemm2_sm100.cu): used when N is a multiple of 256 and M spans more than one 128-row block.
bool gemm_pair_eligible(int M, int N, int K);
mmr_status gemm_pair(const void* A16, int64_t lda, const void* W16, int64_t ldw, const GemmParams& p, int act,
                     int dtype, cudaStream_t stream);

}  // namespace mmr

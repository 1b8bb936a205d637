// Microbenchmark: what limits tcgen05.mma.cta_group::2 (256 x 256 x 16, SS operands) on this part?
//   mode 0  MMA only: one thread per pair issues UMMAs back to back on one resident smem stage (no TMA, no waits)
//   mode 1  mode 0 + a free-running TMA producer refilling OTHER stages (no dependency): smem write interference
//   mode 2  the real producer/consumer ring (full/empty barriers), no epilogue
//   mode 3  mode 0 + 8 warps hammering shared memory with 16-byte stores (epilogue-like traffic)
// Prints MMA instructions per second as a fraction of 1 per 128 clk per SM pair at the measured SM clock.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

#include "../../kddcup_2020_multimodalitiesrecall_2nd_place_b200/csrc/pair_pipeline.cuh"

using namespace mmr;
constexpr int ST = 6;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1)
probe(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
      const __grid_constant__ CUtensorMap tmap_o, int mode, int iters,
      long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* scratch = smem + PairRing<ST>::kOperandBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(scratch + 32768);
  PairRing<ST> ring;
  ring.carve(smem, bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + PairRing<ST>::kNumBars + 2);
  uint64_t* done_bar = bars + PairRing<ST>::kNumBars;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  if (warp == 0 && lane == 0) {
    ring.init(1);
    mbar_init(done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc_2sm(tmem_slot, 512); tmem_relinquish_2sm(); }
  tc_fence_before(); __syncthreads(); cluster_sync_all(); tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const long long t0 = clock64();
  if (warp == 0 && lane == 0) {
    if (mode == 1) {
      // free-running loads into stages 1..ST-1, credited to barriers nobody waits on
      for (int i = 0; i < iters; ++i) {
        const int s = 1 + (i % (ST - 1));
        const uint32_t full_leader = mapa_u32(smem_u32(&ring.full[s]), 0);
        if (rank == 0) mbar_arrive_expect_tx(&ring.full[s], 65536);
        tma_load_2d_2sm(ring.a + size_t(s) * kOpABytes, &tmap_a, full_leader, (i % 12) * 64, ((blockIdx.x >> 1) % 68) * 256 + rank * 128);
        tma_load_2d_2sm(ring.b + size_t(s) * kOpBBytes, &tmap_w, full_leader, (i % 12) * 64, rank * 128);
        // pace: roughly one stage per 512 clk
        const long long t = clock64();
        while (clock64() - t < 400) {}
      }
    } else if (mode == 2) {
      RingPos pos;
      pair_produce_tile<ST, 1>(ring, pos, &tmap_a, &tmap_w, ((blockIdx.x >> 1) % 68) * 256 + rank * 128, rank * 128, 128, iters, rank, 0, 0);
    } else if (mode >= 4) {
      // realistic operand traffic (modes 7-9 use the mode-4 pattern): tiles of 12 in-bounds K blocks, A rows walk the whole matrix, W tiles cycle
      RingPos pos;
      const int pair = blockIdx.x >> 1, tiles = iters / 12;
      for (int t = 0; t < tiles; ++t) {
        int m_blk = (mode == 4 || mode >= 7) ? (pair + t * 74) % 68 : pair % 68;     // mode 5: A tile stays the same (L2-hot)
        int n_blk = mode == 6 ? 0 : (pair + t) % 12;                 // mode 6: W tile stays the same
        if (mode >= 10) {   // the real kernel's raster: neighbouring pairs share the A block (9 column tiles per row block)
          const int tile = pair + t * 74;
          m_blk = (tile / 9) % 68;
          n_blk = tile % 9;
        }
        pair_produce_tile<ST, 1>(ring, pos, &tmap_a, &tmap_w, m_blk * 256 + rank * 128, n_blk * 256 + rank * 128, 128, 12, rank, 0, 0);
      }
    }
  } else if (warp == 1 && lane == 0 && rank == 0) {
    const uint32_t idesc = umma_idesc_f16(0, 256, 256);
    if (mode == 2 || mode >= 4) {
      RingPos pos;
      // tempty barrier was initialised with count 1 and never completes a phase: wait parity 1 passes immediately
      pair_mma_tile<ST>(ring, pos, tmem_base, idesc, iters, 0, 0, 0b11, 0b11);
    } else {
      const uint64_t a_desc = umma_desc_k_sw128(smem_u32(ring.a));
      const uint64_t b_desc = umma_desc_k_sw128(smem_u32(ring.b));
      for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_2sm(tmem_base + (i & 1) * 256, a_desc + 2 * k, b_desc + 2 * k, idesc, 1u);
      }
    }
    umma_commit_2sm_mc(done_bar, 0b11);
  } else if (warp >= 2 && mode >= 7 && mode != 10) {
    // epilogue emulation next to the realistic ring (mode 4 traffic): 7 = TMEM loads only, 8 = + pack/STS + TMA
    // stores of [32 x 32] 16-bit blocks, 9 = TMA stores only; one "tile" (4 chunks) per 12 K blocks
    const int ew = warp - 2, quarter = warp & 3, half = ew >> 2;
    uint8_t* buf = scratch + ew * 4096;
    const int tiles = iters / 12;
    uint32_t n_store = 0;
    for (int t = 0; t < tiles; ++t) {
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        if (mode != 9) {  // (mode 11 = real raster + full epilogue emulation)
          tmem_ld_32x32(tmem_base + (t & 1) * 256 + (uint32_t(quarter * 32) << 16) + half * 128 + c * 32, r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = t + j;
        }
        if (mode >= 8) {
          uint8_t* b = buf + (n_store & 1) * 2048;
          if (lane == 0) bulk_wait_read<1>();
          __syncwarp();
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(b + lane * 64 + ((uint32_t(q) ^ ((lane >> 1) & 3)) << 4)) =
                make_uint4(r[8 * q] ^ r[8 * q + 1], r[8 * q + 2] ^ r[8 * q + 3], r[8 * q + 4] ^ r[8 * q + 5], r[8 * q + 6] ^ r[8 * q + 7]);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(b, &tmap_o, ((t % 12) * 256 + half * 128 + c * 32), (((blockIdx.x >> 1) + t * 74) % 68) * 256 + rank * 128 + quarter * 32);
            bulk_commit();
          }
          ++n_store;
        } else {
          uint32_t acc = 0;
#pragma unroll
          for (int j = 0; j < 32; ++j) acc ^= r[j];
          if (acc == 0x12345678u) scratch[ew] = 1;
        }
      }
      // pace one tile per ~12 K blocks of MMA time
      const long long t_ = clock64();
      while (clock64() - t_ < 3000) {}
    }
    if (lane == 0) bulk_wait<0>();
  } else if (warp >= 2 && mode == 3) {
    uint4 v = make_uint4(1, 2, 3, 4);
    for (int i = 0; i < iters * 4; ++i) {
      *reinterpret_cast<uint4*>(scratch + (warp - 2) * 4096 + lane * 64 + (((i & 3) ^ ((lane >> 1) & 3)) << 4)) = v;
      v.x += i;
    }
  }
  mbar_wait(done_bar, 0);
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  tc_fence_before(); __syncthreads(); cluster_sync_all();
  if (warp == 1) { tc_fence_after(); tmem_dealloc_2sm(tmem_base, 512); }
}

int main() {
  const int M = 68 * 256, K = 768, N = 3072;
  void *A, *W; long long* cyc;
  cudaMalloc(&A, size_t(M) * K * 2); cudaMalloc(&W, size_t(N) * K * 2); cudaMalloc(&cyc, 148 * 8);
  cudaMemset(A, 0, size_t(M) * K * 2); cudaMemset(W, 0, size_t(N) * K * 2);
  void* O; cudaMalloc(&O, size_t(M) * N * 2);
  CUtensorMap ta, tw, to;
  if (make_tmap_ex(&to, O, M, N, N, 0, 32, 32, 64) != MMR_OK) { printf("tmap o failed\n"); return 1; }
  if (make_tmap_2d(&ta, A, M, K, K, 128, 0) != MMR_OK || make_tmap_2d(&tw, W, N, K, K, 128, 0) != MMR_OK) { printf("tmap failed\n"); return 1; }
  const size_t smem = 1024 + PairRing<ST>::kOperandBytes + 32768 + 512;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  const int iters = 12 * 40;   // K blocks (4 UMMAs each)
  for (int mode = 4; mode < 12; ++mode) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<<<148, 320, smem>>>(ta, tw, to, mode, iters, cyc);
    cudaEventRecord(e0);
    probe<<<148, 320, smem>>>(ta, tw, to, mode, iters, cyc);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    const double clk_per_mma = double(mx) / (iters * 4.0);
    const double tflops = 74.0 * iters * 4 * 2.0 * 256 * 256 * 16 / (ms * 1e-3) / 1e12;
    printf("mode %d: %s  %.3f ms  max cycles %lld  -> %.1f clk per UMMA (floor 128), %.0f TFLOP/s, SM clock ~%.2f GHz\n", mode,
           cudaGetErrorString(err), ms, mx, clk_per_mma, tflops, mx / (ms * 1e6));
  }
  return 0;
}

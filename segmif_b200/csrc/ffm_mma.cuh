// mma.sync (m16n8k16 bf16) building blocks shared by the FFM forward (ffm.cu) and backward (ffm_bwd.cu) kernels:
// 64-pixel tiles, 4 warps x 16 rows, 128-byte swizzled shared-memory rows.
#pragma once
#include "common.cuh"

namespace segmif {

constexpr int kTilePx = 64;          // pixels per tile (4 warps x 16 rows)
constexpr int kFfmThreads = 128;

__device__ __forceinline__ int swz128(int row, int chunk) { return chunk ^ (row & 7); }

// copy a [rows x K] bf16 tile (K multiple of 8, rows of K*2 bytes) from pixel-major global memory
__device__ __forceinline__ void load_rows_async(bf16* s, const bf16* g, int64_t first_row, int64_t nrows_total, int ld,
                                                int K, int rows, int tid) {
  const int cpr = K >> 3;
  for (int i = tid; i < rows * cpr; i += kFfmThreads) {
    const int row = i / cpr, chunk = i - row * cpr;
    const bool ok = (first_row + row) < nrows_total;
    const bf16* src = ok ? g + (first_row + row) * ld + chunk * 8 : g;
    cp_async16_cg(smem_u32(s + row * K + swz128(row, chunk) * 8), src, ok ? 16 : 0);
  }
}

// acc[8][4] (16 px x 64 out) = X[16 x K] * W[64 x K]^T for this warp's 16 rows
__device__ __forceinline__ void proj16x64(float (&acc)[8][4], const bf16* sX, int K, int row0, const bf16* sW, int lane) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int ks = 0; ks < (K >> 4); ++ks) {
    uint32_t af[4];
    {
      const int row = row0 + (lane & 15), chunk = ks * 2 + (lane >> 4);
      ldmatrix_x4(af, smem_u32(sX + row * K + swz128(row, chunk) * 8));
    }
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t bfr[4];
      const int row = np * 16 + (lane & 7) + ((lane >> 4) << 3), chunk = ks * 2 + ((lane >> 3) & 1);
      ldmatrix_x4(bfr, smem_u32(sW + row * K + swz128(row, chunk) * 8));
      mma_bf16_16816(acc[np * 2], af, bfr[0], bfr[1]);
      mma_bf16_16816(acc[np * 2 + 1], af, bfr[2], bfr[3]);
    }
  }
}

__device__ __forceinline__ void relu_bias_to_afrag(uint32_t (&af)[4][4], const float (&acc)[8][4], const float* bias, int tq) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
    for (int sub = 0; sub < 2; ++sub) {
      const int nt = kk * 2 + sub;
      const float b0 = bias[nt * 8 + tq * 2], b1 = bias[nt * 8 + tq * 2 + 1];
      af[kk][sub * 2 + 0] = pack_bf16x2(fmaxf(acc[nt][0] + b0, 0.f), fmaxf(acc[nt][1] + b1, 0.f));
      af[kk][sub * 2 + 1] = pack_bf16x2(fmaxf(acc[nt][2] + b0, 0.f), fmaxf(acc[nt][3] + b1, 0.f));
    }
  }
}

// acc += A(16x64, register fragments) * M^T, M stored [64 out][64 in] bf16 swizzled
__device__ __forceinline__ void apply64(float (&acc)[8][4], const uint32_t (&af)[4][4], const bf16* sM, int lane) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t bfr[4];
      const int row = np * 16 + (lane & 7) + ((lane >> 4) << 3), chunk = kk * 2 + ((lane >> 3) & 1);
      ldmatrix_x4(bfr, smem_u32(sM + row * 64 + swz128(row, chunk) * 8));
      mma_bf16_16816(acc[np * 2], af[kk], bfr[0], bfr[1]);
      mma_bf16_16816(acc[np * 2 + 1], af[kk], bfr[2], bfr[3]);
    }
  }
}

// acc += X[16 x K] * W[64 x K]^T (same operands as proj16x64, accumulating)
__device__ __forceinline__ void mm16x64_acc(float (&acc)[8][4], const bf16* sX, int K, int row0, const bf16* sW, int lane) {
  for (int ks = 0; ks < (K >> 4); ++ks) {
    uint32_t af[4];
    {
      const int row = row0 + (lane & 15), chunk = ks * 2 + (lane >> 4);
      ldmatrix_x4(af, smem_u32(sX + row * K + swz128(row, chunk) * 8));
    }
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t bfr[4];
      const int row = np * 16 + (lane & 7) + ((lane >> 4) << 3), chunk = ks * 2 + ((lane >> 3) & 1);
      ldmatrix_x4(bfr, smem_u32(sW + row * K + swz128(row, chunk) * 8));
      mma_bf16_16816(acc[np * 2], af, bfr[0], bfr[1]);
      mma_bf16_16816(acc[np * 2 + 1], af, bfr[2], bfr[3]);
    }
  }
}

// gacc[16 rows of this warp][64] += A^T B over the 64 pixels of a tile; sA, sB: [64 px][64] swizzled bf16
__device__ __forceinline__ void gram16x64_acc(float (&gacc)[8][4], const bf16* sA, const bf16* sB, int warp, int lane) {
#pragma unroll
  for (int ks = 0; ks < kTilePx / 16; ++ks) {
    uint32_t af[4];
    {
      const int row = ks * 16 + (lane & 7) + ((lane >> 4) << 3), chunk = warp * 2 + ((lane >> 3) & 1);
      ldmatrix_x4_trans(af, smem_u32(sA + row * 64 + swz128(row, chunk) * 8));
    }
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t bfr[4];
      const int row = ks * 16 + (lane & 7) + (((lane >> 3) & 1) << 3), chunk = np * 2 + (lane >> 4);
      ldmatrix_x4_trans(bfr, smem_u32(sB + row * 64 + swz128(row, chunk) * 8));
      mma_bf16_16816(gacc[np * 2], af, bfr[0], bfr[1]);
      mma_bf16_16816(gacc[np * 2 + 1], af, bfr[2], bfr[3]);
    }
  }
}

}  // namespace segmif

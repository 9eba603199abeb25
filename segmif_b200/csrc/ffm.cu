// Hierarchical interactive attention (FeatureFusionModule -> CrossPath -> MoAM + SoAM) at full image
// resolution (N = H*W tokens, 64 channels, 8 heads of 8).  See include/segmif_b200.h for the algebra:
//   k^T v = Wk (P^T P) Wv^T   because the kv Linears carry no bias, so pass 1 only accumulates three
//   64x64 Gram matrices per image; pass 2 (one tiny CTA per image) turns them into the per-head 8x8
//   column-softmaxed contexts and folds them with end_proj into four 64x64 matrices; pass 3 recomputes
//   the three needed 64-channel projections per pixel tile, applies the folded matrices, adds the
//   residual and LayerNorms -- 384 B/px read in pass 1, 384 B/px read + 256 B/px written in pass 3.
// conv3 / conv4 (1x1 on the segmentation features) are folded into channel_proj3 by the host.
#include <algorithm>

#include "ffm_mma.cuh"

namespace segmif {

// ------------------------------------------------------------------------------------------------ pass 1
__global__ void __launch_bounds__(kFfmThreads) ffm_gram_kernel(const bf16* __restrict__ x1, int ld1,
                                                               const bf16* __restrict__ x2, int ld2,
                                                               const bf16* __restrict__ x3, int ld3, int C3,
                                                               const bf16* __restrict__ wproj,
                                                               const float* __restrict__ bproj,
                                                               float* __restrict__ partials, int64_t HW) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  bf16* sW = reinterpret_cast<bf16*>(smem_raw);            // [w1y 64x64][w2y 64x64][w3u 64xC3]
  bf16* sX = sW + 64 * (128 + C3);                         // [64 px][<=128]
  bf16* sP = sX + kTilePx * 128;                           // [64 px][64]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y, chunk_id = blockIdx.x, nchunk = gridDim.x;

  // weights (global layout is dense row-major; smem rows are swizzled)
  {
    const int koff[3] = {0, 64 * 64, 2 * 64 * 64};
    const int kk[3] = {64, 64, C3};
    for (int s = 0; s < 3; ++s) {
      const int cpr = kk[s] >> 3;
      for (int i = tid; i < 64 * cpr; i += kFfmThreads) {
        const int row = i / cpr, chunk = i - row * cpr;
        cp_async16_cg(smem_u32(sW + koff[s] + row * kk[s] + swz128(row, chunk) * 8), wproj + koff[s] + row * kk[s] + chunk * 8, 16);
      }
    }
    cp_async_commit();
  }
  float gacc[3][8][4];
#pragma unroll
  for (int s = 0; s < 3; ++s)
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) gacc[s][i][j] = 0.f;

  const int64_t ntiles = (HW + kTilePx - 1) / kTilePx;
  const int g = lane >> 2, tq = lane & 3;
  for (int64_t t = chunk_id; t < ntiles; t += nchunk) {
    const int64_t p0 = t * kTilePx;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const bf16* xs = (s == 0 ? x1 : s == 1 ? x2 : x3) + (int64_t)b * HW * (s == 0 ? ld1 : s == 1 ? ld2 : ld3);
      const int ld = s == 0 ? ld1 : s == 1 ? ld2 : ld3;
      const int K = s == 2 ? C3 : 64;
      const bf16* w = sW + (s == 0 ? 0 : s == 1 ? 64 * 64 : 2 * 64 * 64);
      __syncthreads();                                  // previous users of sX / sP are done
      load_rows_async(sX, xs, p0, HW, ld, K, kTilePx, tid);
      cp_async_commit();
      cp_async_wait<0>();
      __syncthreads();
      float acc[8][4];
      proj16x64(acc, sX, K, warp * 16, w, lane);
      // bias + ReLU, zero rows past the image, write P (bf16) to smem
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int row = warp * 16 + g + half * 8;
        const bool live = (p0 + row) < HW;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const int ch = nt * 8 + tq * 2;
          float v0 = fmaxf(acc[nt][half * 2] + bproj[s * 64 + ch], 0.f);
          float v1 = fmaxf(acc[nt][half * 2 + 1] + bproj[s * 64 + ch + 1], 0.f);
          if (!live) { v0 = 0.f; v1 = 0.f; }
          *reinterpret_cast<uint32_t*>(sP + row * 64 + swz128(row, nt) * 8 + tq * 2) = pack_bf16x2(v0, v1);
        }
      }
      __syncthreads();
      // G_s[16 rows of this warp][64] += P^T P over the 64 pixels of the tile
#pragma unroll
      for (int ks = 0; ks < kTilePx / 16; ++ks) {
        uint32_t af[4];
        {
          const int row = ks * 16 + (lane & 7) + ((lane >> 4) << 3), chunk = warp * 2 + ((lane >> 3) & 1);
          ldmatrix_x4_trans(af, smem_u32(sP + row * 64 + swz128(row, chunk) * 8));
        }
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t bfr[4];
          const int row = ks * 16 + (lane & 7) + (((lane >> 3) & 1) << 3), chunk = np * 2 + (lane >> 4);
          ldmatrix_x4_trans(bfr, smem_u32(sP + row * 64 + swz128(row, chunk) * 8));
          mma_bf16_16816(gacc[s][np * 2], af, bfr[0], bfr[1]);
          mma_bf16_16816(gacc[s][np * 2 + 1], af, bfr[2], bfr[3]);
        }
      }
    }
  }
  cp_async_wait<0>();
  float* out = partials + ((int64_t)b * nchunk + chunk_id) * 3 * 4096;
#pragma unroll
  for (int s = 0; s < 3; ++s)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int i = warp * 16 + g + half * 8, j = nt * 8 + tq * 2;
        *reinterpret_cast<float2*>(out + s * 4096 + i * 64 + j) = make_float2(gacc[s][nt][half * 2], gacc[s][nt][half * 2 + 1]);
      }
}

// ------------------------------------------------------------------------------------------------ pass 2
// (a) one CTA per (stream, image): deterministic reduction of the Gram partials, T = Wk G, logits = T Wv^T per
//     head, column softmax (dim=-2)  ->  ctx[b][s][8][8][8] fp32
__global__ void __launch_bounds__(256) ffm_ctx_kernel(const float* __restrict__ partials, int nchunk,
                                                      const float* __restrict__ wkv, float* __restrict__ ctx_out) {
  // Round 2: Wk / Wv are staged in shared memory (the T = Wk G loop read one global weight per FMA: 70 us for this 24-block
  // launch, two launches per step); summation orders unchanged, results bit-identical.
  extern __shared__ __align__(16) float ctx_smem[];
  float* G = ctx_smem;                 // [64][64]
  float* T = G + 4096;                 // [64][64]
  float* Wsh = T + 4096;               // [64][65]: Wk, then Wv (row pitch 65: the logits loop reads eight rows per warp)
  float* lg = Wsh + 64 * 65;           // [512]
  const int s = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const float* pb = partials + ((int64_t)b * nchunk * 3 + s) * 4096;
  {
    // Only 24 blocks run this kernel, so the reduction is latency bound: all 16 outputs of a thread advance together (64 loads
    // in flight per thread instead of 4).  Per output the summation order is the one of round 1: four interleaved partial sums
    // over the chunks, then (a0 + a1) + (a2 + a3).
    float acc[16][4];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.f;
    int c = 0;
    for (; c + 4 <= nchunk; c += 4) {
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const float* q = pb + (int64_t)c * 3 * 4096 + tid + 256 * k;
        acc[k][0] += q[0];
        acc[k][1] += q[(int64_t)1 * 3 * 4096];
        acc[k][2] += q[(int64_t)2 * 3 * 4096];
        acc[k][3] += q[(int64_t)3 * 3 * 4096];
      }
    }
    for (; c < nchunk; ++c) {
#pragma unroll
      for (int k = 0; k < 16; ++k) acc[k][0] += pb[(int64_t)c * 3 * 4096 + tid + 256 * k];
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) G[tid + 256 * k] = (acc[k][0] + acc[k][1]) + (acc[k][2] + acc[k][3]);
  }
  __syncthreads();
  const float* Wk = wkv + (int64_t)s * 128 * 64;
  const float* Wv = Wk + 64 * 64;
  for (int idx = tid; idx < 4096; idx += 256) Wsh[(idx >> 6) * 65 + (idx & 63)] = Wk[idx];
  __syncthreads();
  for (int idx = tid; idx < 4096; idx += 256) {       // T = Wk G_s
    const int r = idx >> 6, c = idx & 63;
    float a = 0.f;
#pragma unroll 8
    for (int k = 0; k < 64; ++k) a = fmaf(Wsh[r * 65 + k], G[k * 64 + c], a);
    T[idx] = a;
  }
  __syncthreads();
  for (int idx = tid; idx < 4096; idx += 256) Wsh[(idx >> 6) * 65 + (idx & 63)] = Wv[idx];
  __syncthreads();
  const float scale = 0.35355339059327379f;           // 8^-1/2 (head_dim 8)
  for (int idx = tid; idx < 512; idx += 256) {        // logits[h][i][j] = scale * T[h8+i,:] . Wv[h8+j,:]
    const int h = idx >> 6, i = (idx >> 3) & 7, j = idx & 7;
    float a = 0.f;
#pragma unroll 8
    for (int c = 0; c < 64; ++c) a = fmaf(T[(h * 8 + i) * 64 + c], Wsh[(h * 8 + j) * 65 + c], a);
    lg[idx] = a * scale;
  }
  __syncthreads();
  if (tid < 64) {                                     // softmax over i (dim=-2) for every (h, j)
    const int h = tid >> 3, j = tid & 7;
    float* c = lg + h * 64 + j;
    float m = -INFINITY;
    for (int i = 0; i < 8; ++i) m = fmaxf(m, c[i * 8]);
    float e[8], sum = 0.f;
    for (int i = 0; i < 8; ++i) { e[i] = expf(c[i * 8] - m); sum += e[i]; }
    for (int i = 0; i < 8; ++i) c[i * 8] = e[i] / sum;
  }
  __syncthreads();
  for (int idx = tid; idx < 512; idx += 256) ctx_out[((int64_t)b * 3 + s) * 512 + idx] = lg[idx];
}

// (b) folded[b][m][o][c]: m=0 Mz1 (ctx1, We1[:, :64]); 1 Mv1 (ctx3, We1[:, 64:]); 2 Mz2 (ctx2, We2[:, :64]); 3 Mv2 (ctx3, We2[:, 64:])
__global__ void __launch_bounds__(256) ffm_fold_kernel(const float* __restrict__ ctx, const float* __restrict__ wend,
                                                       bf16* __restrict__ folded) {
  const int m = blockIdx.x, b = blockIdx.y;
  const int stream = m >> 1, is_v = m & 1;
  const float* cb = ctx + ((int64_t)b * 3 + (is_v ? 2 : stream)) * 512;
  for (int idx = threadIdx.x; idx < 4096; idx += 256) {
    const int o = idx >> 6, c = idx & 63;
    const int h = c >> 3, i = c & 7;
    const float* cx = cb + h * 64 + i * 8;
    const float* we = wend + (int64_t)stream * 64 * 128 + o * 128 + (is_v ? 64 : 0) + h * 8;
    float a = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) a = fmaf(cx[j], we[j], a);
    folded[((int64_t)b * 4 + m) * 4096 + idx] = __float2bfloat16_rn(a);
  }
}

// ------------------------------------------------------------------------------------------------ pass 3
__global__ void __launch_bounds__(kFfmThreads) ffm_apply_kernel(
    const bf16* __restrict__ x1, int ld1, const bf16* __restrict__ x2, int ld2, const bf16* __restrict__ x3, int ld3,
    int C3, const bf16* __restrict__ wproj, const float* __restrict__ bproj, const bf16* __restrict__ folded,
    const float* __restrict__ bend, const float* __restrict__ ln_g, const float* __restrict__ ln_b, float eps,
    bf16* __restrict__ out1, int ldo1, bf16* __restrict__ out2, int ldo2, int64_t HW, bf16* __restrict__ pre1,
    bf16* __restrict__ pre2) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  bf16* sW3 = reinterpret_cast<bf16*>(smem_raw);   // [64][C3]   channel_proj3 (y half) folded with conv3|conv4
  bf16* sW1 = sW3 + 64 * C3;                       // [64][64]   channel_proj1 (u half)
  bf16* sW2 = sW1 + 64 * 64;                       // [64][64]   channel_proj2 (u half)
  bf16* sM = sW2 + 64 * 64;                        // [4][64][64] folded Mz1, Mv1, Mz2, Mv2 of this image
  bf16* sX1 = sM + 4 * 64 * 64;                    // [64 px][64]
  bf16* sX2 = sX1 + kTilePx * 64;
  bf16* sX3 = sX2 + kTilePx * 64;                  // [64 px][C3]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y;
  {
    // weights: global [w3y 64xC3][w1u 64x64][w2u 64x64] dense; folded [4][64][64] dense
    const int cpr3 = C3 >> 3;
    for (int i = tid; i < 64 * cpr3; i += kFfmThreads) {
      const int row = i / cpr3, chunk = i - row * cpr3;
      cp_async16_cg(smem_u32(sW3 + row * C3 + swz128(row, chunk) * 8), wproj + row * C3 + chunk * 8, 16);
    }
    for (int i = tid; i < 2 * 64 * 8; i += kFfmThreads) {
      const int m = i >> 9, row = (i >> 3) & 63, chunk = i & 7;
      cp_async16_cg(smem_u32(sW1 + m * 4096 + row * 64 + swz128(row, chunk) * 8), wproj + 64 * C3 + m * 4096 + row * 64 + chunk * 8, 16);
    }
    const bf16* fb = folded + (int64_t)b * 4 * 4096;
    for (int i = tid; i < 4 * 64 * 8; i += kFfmThreads) {
      const int m = i >> 9, row = (i >> 3) & 63, chunk = i & 7;
      cp_async16_cg(smem_u32(sM + m * 4096 + row * 64 + swz128(row, chunk) * 8), fb + m * 4096 + row * 64 + chunk * 8, 16);
    }
    cp_async_commit();
  }
  const int g = lane >> 2, tq = lane & 3;
  const int64_t ntiles = (HW + kTilePx - 1) / kTilePx;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t p0 = t * kTilePx;
    __syncthreads();
    load_rows_async(sX1, x1 + (int64_t)b * HW * ld1, p0, HW, ld1, 64, kTilePx, tid);
    load_rows_async(sX2, x2 + (int64_t)b * HW * ld2, p0, HW, ld2, 64, kTilePx, tid);
    load_rows_async(sX3, x3 + (int64_t)b * HW * ld3, p0, HW, ld3, C3, kTilePx, tid);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    const int row0 = warp * 16;
    uint32_t ay[4][4];
    {
      float acc[8][4];
      proj16x64(acc, sX3, C3, row0, sW3, lane);
      relu_bias_to_afrag(ay, acc, bproj, tq);
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const bf16* sX = s == 0 ? sX1 : sX2;
      uint32_t au[4][4];
      {
        float acc[8][4];
        proj16x64(acc, sX, 64, row0, s == 0 ? sW1 : sW2, lane);
        relu_bias_to_afrag(au, acc, bproj + 64 * (1 + s), tq);
      }
      float acc[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      apply64(acc, ay, sM + (2 * s) * 4096, lane);
      apply64(acc, au, sM + (2 * s + 1) * 4096, lane);
      // + end_proj bias + residual x_s, LayerNorm over the 64 channels of each row
      float sum[2] = {0.f, 0.f};
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int ch = nt * 8 + tq * 2;
        const float b0 = bend[s * 64 + ch], b1 = bend[s * 64 + ch + 1];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int row = row0 + g + half * 8;
          const float2 xr = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(sX + row * 64 + swz128(row, nt) * 8 + tq * 2));
          acc[nt][half * 2] += b0 + xr.x;
          acc[nt][half * 2 + 1] += b1 + xr.y;
          sum[half] += acc[nt][half * 2] + acc[nt][half * 2 + 1];
        }
      }
      float mean[2], rstd[2];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        sum[half] += __shfl_xor_sync(0xffffffffu, sum[half], 1);
        sum[half] += __shfl_xor_sync(0xffffffffu, sum[half], 2);
        mean[half] = sum[half] * (1.f / 64.f);
      }
      float sq[2] = {0.f, 0.f};
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float d = acc[nt][j] - mean[j >> 1]; sq[j >> 1] += d * d; }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        sq[half] += __shfl_xor_sync(0xffffffffu, sq[half], 1);
        sq[half] += __shfl_xor_sync(0xffffffffu, sq[half], 2);
        rstd[half] = rsqrtf(sq[half] * (1.f / 64.f) + eps);
      }
      bf16* op = (s == 0 ? out1 : out2);
      const int ldo = s == 0 ? ldo1 : ldo2;
      bf16* pre = s == 0 ? pre1 : pre2;             // training: the LayerNorm input, dense [B*HW, 64]
      if (pre != nullptr) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int64_t px = p0 + row0 + g + half * 8;
          if (px >= HW) continue;
          bf16* o = pre + ((int64_t)b * HW + px) * 64 + tq * 2;
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) *reinterpret_cast<uint32_t*>(o + nt * 8) = pack_bf16x2(acc[nt][half * 2], acc[nt][half * 2 + 1]);
        }
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int64_t px = p0 + row0 + g + half * 8;
        if (px >= HW) continue;
        bf16* o = op + ((int64_t)b * HW + px) * ldo + tq * 2;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const int ch = nt * 8 + tq * 2;
          const float v0 = (acc[nt][half * 2] - mean[half]) * rstd[half] * ln_g[s * 64 + ch] + ln_b[s * 64 + ch];
          const float v1 = (acc[nt][half * 2 + 1] - mean[half]) * rstd[half] * ln_g[s * 64 + ch + 1] + ln_b[s * 64 + ch + 1];
          *reinterpret_cast<uint32_t*>(o + nt * 8) = pack_bf16x2(v0, v1);
        }
      }
    }
  }
  cp_async_wait<0>();
}

// ================================================================================================ low-resolution seg stream
// channel_proj3 (already folded with conv3 | conv4) is linear and bilinear resizing is linear with weights that sum to
// one, so relu(channel_proj3(conv(upsample(f)))) == relu(upsample(Q)) with Q = W_eff f + b_eff computed ONCE at the
// encoder's resolution ([B, h, w, 128] bf16, 39 MB / 10 MB at cfg 2: L2 resident).  The *_lr kernels interpolate Q on the
// fly instead of projecting a materialised full-resolution feature map: the third projection GEMM of both passes, the
// 0.9 GB of upsampled features and their upsample kernel all disappear.  Q channels: [0,64) = y3 half, [64,128) = u3.
// Both *_lr passes run on tcgen05 tensor cores: ffm_tc.cu.

// one-time opt-in to the largest configuration of each kernel (never called again, e.g. during graph capture)
static int set_smem(const void* fn, size_t bytes, const char* what, bool* done) {
  if (*done) return SEGMIF_OK;
  *done = true;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) { set_error("%s: cudaFuncSetAttribute(%zu) failed: %s", what, bytes, cudaGetErrorString(e)); return SEGMIF_ERR_CUDA; }
  return SEGMIF_OK;
}

}  // namespace segmif

using namespace segmif;

extern "C" int segmif_ffm_gram_fwd(const void* x1, int ld1, int coff1, const void* x2, int ld2, int coff2,
                                   const void* x3, int ld3, int C3, const void* wproj, const float* bproj,
                                   float* partials, int nchunk, int B, int64_t HW, segmif_stream_t stream) {
  SEGMIF_REQUIRE(x1 && x2 && x3 && wproj && bproj && partials, "ffm_gram: null pointer");
  SEGMIF_REQUIRE(C3 == 64 || C3 == 128, "ffm_gram: C3=%d must be 64 or 128", C3);
  SEGMIF_REQUIRE(ld1 % 8 == 0 && ld2 % 8 == 0 && ld3 % 8 == 0 && coff1 % 8 == 0 && coff2 % 8 == 0, "ffm_gram: pitches/offsets must be multiples of 8");
  SEGMIF_REQUIRE(nchunk > 0 && B > 0 && HW > 0, "ffm_gram: bad sizes");
  const size_t smem = (size_t)(64 * (128 + C3) + kTilePx * 128 + kTilePx * 64) * sizeof(bf16);
  static bool cfg = false;
  int rc = set_smem((const void*)ffm_gram_kernel, (size_t)(64 * 256 + kTilePx * 128 + kTilePx * 64) * sizeof(bf16), "ffm_gram", &cfg);
  if (rc) return rc;
  dim3 grid(nchunk, B);
  ffm_gram_kernel<<<grid, kFfmThreads, smem, as_stream(stream)>>>((const bf16*)x1 + coff1, ld1, (const bf16*)x2 + coff2, ld2,
                                                                   (const bf16*)x3, ld3, C3, (const bf16*)wproj, bproj, partials, HW);
  return check_launch("segmif_ffm_gram_fwd");
}

extern "C" int segmif_ffm_ctx_fwd(const float* partials, int nchunk, const float* wkv, const float* wend, void* folded,
                                  float* ctx_out, int B, segmif_stream_t stream) {
  SEGMIF_REQUIRE(partials && wkv && wend && folded && ctx_out, "ffm_ctx: null pointer (ctx_out [B,3,8,8,8] is required)");
  SEGMIF_REQUIRE(nchunk > 0 && B > 0, "ffm_ctx: bad sizes");
  cudaStream_t st = as_stream(stream);
  constexpr int ctx_smem = (2 * 4096 + 64 * 65 + 512) * (int)sizeof(float);
  static bool ctx_cfg = false;
  if (!ctx_cfg) {
    cudaError_t e = cudaFuncSetAttribute(ffm_ctx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx_smem);
    if (e != cudaSuccess) { set_error("ffm_ctx: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return SEGMIF_ERR_CUDA; }
    ctx_cfg = true;
  }
  ffm_ctx_kernel<<<dim3(3, B), 256, ctx_smem, st>>>(partials, nchunk, wkv, ctx_out);
  int rc = check_launch("segmif_ffm_ctx_fwd");
  if (rc) return rc;
  ffm_fold_kernel<<<dim3(4, B), 256, 0, st>>>(ctx_out, wend, (bf16*)folded);
  return check_launch("segmif_ffm_ctx_fwd(fold)");
}

static int ffm_apply_impl(const void* x1, int ld1, int coff1, const void* x2, int ld2, int coff2,
                                    const void* x3, int ld3, int C3, const void* wproj, const float* bproj,
                                    const void* folded, const float* bend, const float* ln_gamma, const float* ln_beta,
                                    float eps, void* out1, int ldo1, int coffo1, void* out2, int ldo2, int coffo2, int B,
                                    int64_t HW, void* pre1, void* pre2, segmif_stream_t stream) {
  SEGMIF_REQUIRE(x1 && x2 && x3 && wproj && bproj && folded && bend && ln_gamma && ln_beta && out1 && out2, "ffm_apply: null pointer");
  SEGMIF_REQUIRE(C3 == 64 || C3 == 128, "ffm_apply: C3=%d must be 64 or 128", C3);
  SEGMIF_REQUIRE(ld1 % 8 == 0 && ld2 % 8 == 0 && ld3 % 8 == 0 && coff1 % 8 == 0 && coff2 % 8 == 0, "ffm_apply: input pitches/offsets must be multiples of 8");
  SEGMIF_REQUIRE(ldo1 % 2 == 0 && ldo2 % 2 == 0 && coffo1 % 2 == 0 && coffo2 % 2 == 0, "ffm_apply: output pitches/offsets must be even");
  const size_t smem = (size_t)(64 * C3 + 2 * 4096 + 4 * 4096 + 2 * kTilePx * 64 + kTilePx * C3) * sizeof(bf16);
  static bool cfg = false;
  int rc = set_smem((const void*)ffm_apply_kernel, (size_t)(64 * 128 + 2 * 4096 + 4 * 4096 + 2 * kTilePx * 64 + kTilePx * 128) * sizeof(bf16), "ffm_apply", &cfg);
  if (rc) return rc;
  const int64_t ntiles = (HW + kTilePx - 1) / kTilePx;
  const int per_image = (int)std::min<int64_t>(ntiles, std::max<int64_t>(1, (148 * 2 * 4) / B));
  dim3 grid(per_image, B);
  ffm_apply_kernel<<<grid, kFfmThreads, smem, as_stream(stream)>>>(
      (const bf16*)x1 + coff1, ld1, (const bf16*)x2 + coff2, ld2, (const bf16*)x3, ld3, C3, (const bf16*)wproj, bproj,
      (const bf16*)folded, bend, ln_gamma, ln_beta, eps, (bf16*)out1 + coffo1, ldo1, (bf16*)out2 + coffo2, ldo2, HW, (bf16*)pre1, (bf16*)pre2);
  return check_launch("segmif_ffm_apply_fwd");
}

extern "C" int segmif_ffm_apply_fwd(const void* x1, int ld1, int coff1, const void* x2, int ld2, int coff2,
                                    const void* x3, int ld3, int C3, const void* wproj, const float* bproj,
                                    const void* folded, const float* bend, const float* ln_gamma, const float* ln_beta,
                                    float eps, void* out1, int ldo1, int coffo1, void* out2, int ldo2, int coffo2, int B,
                                    int64_t HW, segmif_stream_t stream) {
  return ffm_apply_impl(x1, ld1, coff1, x2, ld2, coff2, x3, ld3, C3, wproj, bproj, folded, bend, ln_gamma, ln_beta, eps,
                        out1, ldo1, coffo1, out2, ldo2, coffo2, B, HW, nullptr, nullptr, stream);
}

extern "C" int segmif_ffm_apply_train_fwd(const void* x1, int ld1, int coff1, const void* x2, int ld2, int coff2,
                                          const void* x3, int ld3, int C3, const void* wproj, const float* bproj,
                                          const void* folded, const float* bend, const float* ln_gamma,
                                          const float* ln_beta, float eps, void* out1, int ldo1, int coffo1, void* out2,
                                          int ldo2, int coffo2, int B, int64_t HW, void* pre1, void* pre2,
                                          segmif_stream_t stream) {
  SEGMIF_REQUIRE(pre1 && pre2, "ffm_apply_train: the pre-LayerNorm outputs are required");
  return ffm_apply_impl(x1, ld1, coff1, x2, ld2, coff2, x3, ld3, C3, wproj, bproj, folded, bend, ln_gamma, ln_beta, eps,
                        out1, ldo1, coffo1, out2, ldo2, coffo2, B, HW, pre1, pre2, stream);
}

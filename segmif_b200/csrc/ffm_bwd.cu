// Backward of the hierarchical interactive attention (core/model_fusion.py:350-361, :263-288, :303-328) in the same
// three-pass shape as the forward (ffm.cu).  With r_i = x_i + y3 Mz_i^T + u_i Mv_i^T + b_i the LayerNorm input and
// dr_i its gradient (segmif_layernorm_bwd), everything that couples pixels goes through 64x64 matrices per image:
//   (1) gram:  R_iy = dr_i^T y3,  R_iu = dr_i^T u_i     (gradients of the folded matrices Mz_i, Mv_i), fp32 partials;
//   (2) ctx:   one CTA per image: dW_end, d ctx -> column-softmax backward -> dA_s (block diagonal) -> dW_kv and
//              dG_s = Wk^T dA_s Wv; because G_s = P_s^T P_s the pixel-side gradient is dP_s = P_s (dG_s + dG_s^T);
//              emits seven bf16 matrices per image: S_1, S_2, S_3 (= dG + dG^T) and Mz_1^T, Mv_1^T, Mz_2^T, Mv_2^T;
//   (3) apply: per 64-pixel tile recompute the six 64-channel projections, then
//                 dP_i = [ (y_i S_i) * 1[y_i>0] , (dr_i Mv_i) * 1[u_i>0] ]            i = 1, 2
//                 dP_3 = [ (dr_1 Mz_1 + dr_2 Mz_2) * 1[y3>0] , (u3 S_3) * 1[u3>0] ]
//              written as three [pixels, 128] bf16 tensors = gradients of the channel_proj pre-activations.  dx_i, dW and
//              db of the projections then come from the generic GEMM / wgrad / colsum kernels.
#include <algorithm>

#include "ffm_mma.cuh"

namespace segmif {

// relu(acc + bias) > 0 as a bit per accumulator element (bit nt*4 + j)
__device__ __forceinline__ uint32_t relu_mask(const float (&acc)[8][4], const float* bias, int tq) {
  uint32_t m = 0;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (acc[nt][j] + bias[nt * 8 + tq * 2 + (j & 1)] > 0.f) m |= 1u << (nt * 4 + j);
  return m;
}

__device__ __forceinline__ void zero_acc(float (&acc)[8][4]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
}

// relu(acc + bias) as bf16 into a swizzled [64 px][64] smem tile (rows past the image zeroed)
__device__ __forceinline__ void relu_to_smem(bf16* sP, const float (&acc)[8][4], const float* bias, int row0, int g, int tq,
                                             int64_t p0, int64_t HW) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int row = row0 + g + half * 8;
    const bool live = (p0 + row) < HW;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int ch = nt * 8 + tq * 2;
      float v0 = fmaxf(acc[nt][half * 2] + bias[ch], 0.f);
      float v1 = fmaxf(acc[nt][half * 2 + 1] + bias[ch + 1], 0.f);
      if (!live) { v0 = 0.f; v1 = 0.f; }
      *reinterpret_cast<uint32_t*>(sP + row * 64 + swz128(row, nt) * 8 + tq * 2) = pack_bf16x2(v0, v1);
    }
  }
}

__device__ __forceinline__ void load_w64(bf16* s, const bf16* g, int tid) {      // dense [64][64] -> swizzled
  for (int i = tid; i < 64 * 8; i += kFfmThreads) {
    const int row = i >> 3, chunk = i & 7;
    cp_async16_cg(smem_u32(s + row * 64 + swz128(row, chunk) * 8), g + row * 64 + chunk * 8, 16);
  }
}

// ------------------------------------------------------------------------------------------------ (1) cross-Grams
// wfull: bf16 [3][128][64] = channel_proj1|2|3 weights (rows 0..63 = y half, 64..127 = u half); bfull fp32 [3][128]
__global__ void __launch_bounds__(kFfmThreads) ffm_bwd_gram_kernel(const bf16* __restrict__ x1, int ld1,
                                                                   const bf16* __restrict__ x2, int ld2,
                                                                   const bf16* __restrict__ x3, int ld3,
                                                                   const bf16* __restrict__ dr1, const bf16* __restrict__ dr2,
                                                                   const bf16* __restrict__ wfull, const float* __restrict__ bfull,
                                                                   float* __restrict__ partials, int64_t HW) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  bf16* sW3y = reinterpret_cast<bf16*>(smem_raw);
  bf16* sW1u = sW3y + 4096;
  bf16* sW2u = sW1u + 4096;
  bf16* sX1 = sW2u + 4096;
  bf16* sX2 = sX1 + 4096;
  bf16* sX3 = sX2 + 4096;
  bf16* sD1 = sX3 + 4096;
  bf16* sD2 = sD1 + 4096;
  bf16* sY3 = sD2 + 4096;
  bf16* sU1 = sY3 + 4096;
  bf16* sU2 = sU1 + 4096;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y, chunk_id = blockIdx.x, nchunk = gridDim.x;
  load_w64(sW3y, wfull + 2 * 8192, tid);
  load_w64(sW1u, wfull + 4096, tid);
  load_w64(sW2u, wfull + 8192 + 4096, tid);
  cp_async_commit();
  float R[4][8][4];
#pragma unroll
  for (int m = 0; m < 4; ++m) zero_acc(R[m]);
  const int g = lane >> 2, tq = lane & 3;
  const int64_t ntiles = (HW + kTilePx - 1) / kTilePx;
  for (int64_t t = chunk_id; t < ntiles; t += nchunk) {
    const int64_t p0 = t * kTilePx;
    __syncthreads();
    load_rows_async(sX1, x1 + (int64_t)b * HW * ld1, p0, HW, ld1, 64, kTilePx, tid);
    load_rows_async(sX2, x2 + (int64_t)b * HW * ld2, p0, HW, ld2, 64, kTilePx, tid);
    load_rows_async(sX3, x3 + (int64_t)b * HW * ld3, p0, HW, ld3, 64, kTilePx, tid);
    load_rows_async(sD1, dr1 + (int64_t)b * HW * 64, p0, HW, 64, 64, kTilePx, tid);
    load_rows_async(sD2, dr2 + (int64_t)b * HW * 64, p0, HW, 64, 64, kTilePx, tid);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    const int row0 = warp * 16;
    {
      float acc[8][4];
      proj16x64(acc, sX3, 64, row0, sW3y, lane);
      relu_to_smem(sY3, acc, bfull + 2 * 128, row0, g, tq, p0, HW);
      proj16x64(acc, sX1, 64, row0, sW1u, lane);
      relu_to_smem(sU1, acc, bfull + 64, row0, g, tq, p0, HW);
      proj16x64(acc, sX2, 64, row0, sW2u, lane);
      relu_to_smem(sU2, acc, bfull + 128 + 64, row0, g, tq, p0, HW);
    }
    __syncthreads();
    gram16x64_acc(R[0], sD1, sY3, warp, lane);
    gram16x64_acc(R[1], sD1, sU1, warp, lane);
    gram16x64_acc(R[2], sD2, sY3, warp, lane);
    gram16x64_acc(R[3], sD2, sU2, warp, lane);
  }
  cp_async_wait<0>();
  float* out = partials + ((int64_t)b * nchunk + chunk_id) * 4 * 4096;
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int i = warp * 16 + g + half * 8, j = nt * 8 + tq * 2;
        *reinterpret_cast<float2*>(out + m * 4096 + i * 64 + j) = make_float2(R[m][nt][half * 2], R[m][nt][half * 2 + 1]);
      }
}

// ------------------------------------------------------------------------------------------------ (2) per-image algebra
__device__ __forceinline__ void reduce_partials(float* dst, const float* src, int nchunk, int64_t stride, int tid, int nthreads) {
  for (int idx = tid; idx < 4096; idx += nthreads) {
    float a0 = 0.f, a1 = 0.f;
    int c = 0;
    for (; c + 2 <= nchunk; c += 2) { a0 += src[(int64_t)c * stride + idx]; a1 += src[(int64_t)(c + 1) * stride + idx]; }
    if (c < nchunk) a0 += src[(int64_t)c * stride + idx];
    dst[idx] = a0 + a1;
  }
}

// grid (3, B): CTA (s, b) owns context stream s of image b (s = 0: kv1 / y1, 1: kv2 / y2, 2: kv3 / u3) and the folded
// matrices that use it (m = 0 | 2 | {1, 3}), so the twelve CTAs of a batch of four run concurrently; Wk / Wv are staged
// in shared memory (the inner products below walk them with stride 64).
constexpr int kCtxThreads = 512;

__global__ void __launch_bounds__(kCtxThreads) ffm_bwd_ctx_kernel(const float* __restrict__ Rpart, int nchunkR,
                                                                  const float* __restrict__ Gpart, int nchunkG,
                                                                  const float* __restrict__ ctx, const float* __restrict__ wkv,
                                                                  const float* __restrict__ wend, const bf16* __restrict__ folded,
                                                                  bf16* __restrict__ mats, float* __restrict__ dwkv,
                                                                  float* __restrict__ dwend) {
  extern __shared__ float smf[];
  float* A0 = smf;             // reduced partial (R_m, then G_s)
  float* T = smf + 4096;       // Wk G
  float* U = smf + 8192;       // G Wv^T, then V
  float* DG = smf + 12288;
  float* sWk = smf + 16384;    // [64][65] padded
  float* sWv = sWk + 64 * 65;
  float* dctx = sWv + 64 * 65; // [512]
  float* dA = dctx + 512;      // [512]  [h][i][j], scale folded in
  float* sWe = dA + 512;       // [64][8*8 + 1]: the We columns of the current m (64 rows x 64 cols), padded
  const int s = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const float* cx = ctx + ((int64_t)b * 3 + s) * 512;
  for (int i = tid; i < 512; i += kCtxThreads) dctx[i] = 0.f;
  {
    const float* Wk = wkv + (int64_t)s * 128 * 64;
    for (int i = tid; i < 4096; i += kCtxThreads) { sWk[(i >> 6) * 65 + (i & 63)] = Wk[i]; sWv[(i >> 6) * 65 + (i & 63)] = Wk[4096 + i]; }
  }
  __syncthreads();
  const int nm = s == 2 ? 2 : 1;
  for (int q = 0; q < nm; ++q) {
    const int m = s == 2 ? 1 + 2 * q : 2 * s;             // folded matrix index: Mz1, Mv1, Mz2, Mv2
    const int stream = m >> 1, is_v = m & 1;
    reduce_partials(A0, Rpart + ((int64_t)b * nchunkR * 4 + m) * 4096, nchunkR, 4 * 4096, tid, kCtxThreads);
    const float* We = wend + (int64_t)stream * 64 * 128 + (is_v ? 64 : 0);
    for (int i = tid; i < 4096; i += kCtxThreads) sWe[(i >> 6) * 65 + (i & 63)] = We[(i >> 6) * 128 + (i & 63)];
    __syncthreads();
    for (int idx = tid; idx < 4096; idx += kCtxThreads) {    // dW_end[o][h8+j] += sum_i dM[o][h8+i] ctx[h][i][j]
      const int o = idx >> 6, c = idx & 63, h = c >> 3, j = c & 7;
      float a = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) a = fmaf(A0[o * 64 + h * 8 + i], cx[h * 64 + i * 8 + j], a);
      atomicAdd(dwend + (int64_t)stream * 64 * 128 + o * 128 + (is_v ? 64 : 0) + c, a);
    }
    {                                                        // d ctx[h][i][j] += sum_o dM[o][h8+i] We[o][h8+j]
      const int idx = tid;                                   // kCtxThreads == 512 == number of context entries
      const int h = idx >> 6, i = (idx >> 3) & 7, j = idx & 7;
      float a = 0.f;
#pragma unroll 8
      for (int o = 0; o < 64; ++o) a = fmaf(A0[o * 64 + h * 8 + i], sWe[o * 65 + h * 8 + j], a);
      dctx[idx] += a;
    }
    __syncthreads();
  }
  const float scale = 0.35355339059327379f;
  if (tid < 64) {                                           // softmax (over i) backward for every (h, j)
    const int h = tid >> 3, j = tid & 7;
    float dot = 0.f;
    for (int i = 0; i < 8; ++i) dot = fmaf(dctx[h * 64 + i * 8 + j], cx[h * 64 + i * 8 + j], dot);
    for (int i = 0; i < 8; ++i) dA[h * 64 + i * 8 + j] = scale * cx[h * 64 + i * 8 + j] * (dctx[h * 64 + i * 8 + j] - dot);
  }
  reduce_partials(A0, Gpart + ((int64_t)b * nchunkG * 3 + s) * 4096, nchunkG, 3 * 4096, tid, kCtxThreads);
  __syncthreads();
  for (int idx = tid; idx < 4096; idx += kCtxThreads) {
    const int r = idx >> 6, c = idx & 63;
    float t = 0.f, u = 0.f;
#pragma unroll 8
    for (int k = 0; k < 64; ++k) {
      t = fmaf(sWk[r * 65 + k], A0[k * 64 + c], t);         // T[r][c] = sum_k Wk[r][k] G[k][c]
      u = fmaf(A0[r * 64 + k], sWv[c * 65 + k], u);         // U[r][c] = sum_k G[r][k] Wv[c][k]
    }
    T[idx] = t;
    U[idx] = u;
  }
  __syncthreads();
  for (int idx = tid; idx < 4096; idx += kCtxThreads) {
    const int a = idx >> 6, c = idx & 63, h = a >> 3, ij = a & 7;
    float dk = 0.f, dv = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      dk = fmaf(dA[h * 64 + ij * 8 + q], U[c * 64 + h * 8 + q], dk);     // dWk[8h+i][c] = sum_j dA[h][i][j] U[c][8h+j]
      dv = fmaf(dA[h * 64 + q * 8 + ij], T[(h * 8 + q) * 64 + c], dv);   // dWv[8h+j][c] = sum_i dA[h][i][j] T[8h+i][c]
    }
    atomicAdd(dwkv + (int64_t)s * 8192 + a * 64 + c, dk);
    atomicAdd(dwkv + (int64_t)s * 8192 + 4096 + a * 64 + c, dv);
  }
  __syncthreads();
  for (int idx = tid; idx < 4096; idx += kCtxThreads) {     // V[8h+i][c] = sum_j dA[h][i][j] Wv[8h+j][c]
    const int a = idx >> 6, c = idx & 63, h = a >> 3, i = a & 7;
    float v = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) v = fmaf(dA[h * 64 + i * 8 + j], sWv[(h * 8 + j) * 65 + c], v);
    U[idx] = v;
  }
  __syncthreads();
  for (int idx = tid; idx < 4096; idx += kCtxThreads) {     // dG[c][c'] = sum_a Wk[a][c] V[a][c']
    const int c = idx >> 6, c2 = idx & 63;
    float v = 0.f;
#pragma unroll 8
    for (int a = 0; a < 64; ++a) v = fmaf(sWk[a * 65 + c], U[a * 64 + c2], v);
    DG[idx] = v;
  }
  __syncthreads();
  bf16* ms = mats + ((int64_t)b * 7 + s) * 4096;
  for (int idx = tid; idx < 4096; idx += kCtxThreads) {
    const int c = idx >> 6, c2 = idx & 63;
    ms[idx] = __float2bfloat16_rn(DG[idx] + DG[c2 * 64 + c]);
  }
  for (int q = 0; q < nm; ++q) {                            // transposes of the folded forward matrices this CTA owns
    const int m = s == 2 ? 1 + 2 * q : 2 * s;
    const bf16* f = folded + ((int64_t)b * 4 + m) * 4096;
    bf16* mt = mats + ((int64_t)b * 7 + 3 + m) * 4096;
    for (int idx = tid; idx < 4096; idx += kCtxThreads) mt[idx] = f[(idx & 63) * 64 + (idx >> 6)];
  }
}

// ------------------------------------------------------------------------------------------------ (3) pixel side
__device__ __forceinline__ void store_masked(bf16* dst, int64_t row_base, int coff, const float (&acc)[8][4], uint32_t mask,
                                             int row0, int g, int tq, int64_t p0, int64_t HW) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int64_t px = p0 + row0 + g + half * 8;
    if (px >= HW) continue;
    bf16* o = dst + (row_base + px) * 128 + coff + tq * 2;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float v0 = (mask >> (nt * 4 + half * 2)) & 1u ? acc[nt][half * 2] : 0.f;
      const float v1 = (mask >> (nt * 4 + half * 2 + 1)) & 1u ? acc[nt][half * 2 + 1] : 0.f;
      *reinterpret_cast<uint32_t*>(o + nt * 8) = pack_bf16x2(v0, v1);
    }
  }
}

__global__ void __launch_bounds__(kFfmThreads) ffm_bwd_apply_kernel(const bf16* __restrict__ x1, int ld1,
                                                                    const bf16* __restrict__ x2, int ld2,
                                                                    const bf16* __restrict__ x3, int ld3,
                                                                    const bf16* __restrict__ dr1, const bf16* __restrict__ dr2,
                                                                    const bf16* __restrict__ wfull, const float* __restrict__ bfull,
                                                                    const bf16* __restrict__ mats, bf16* __restrict__ dP1,
                                                                    bf16* __restrict__ dP2, bf16* __restrict__ dP3, int64_t HW) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  bf16* sW = reinterpret_cast<bf16*>(smem_raw);      // [6][64][64]: W1y W1u W2y W2u W3y W3u
  bf16* sM = sW + 6 * 4096;                          // [7][64][64]: S1 S2 S3 Mz1T Mv1T Mz2T Mv2T
  bf16* sX1 = sM + 7 * 4096;
  bf16* sX2 = sX1 + 4096;
  bf16* sX3 = sX2 + 4096;
  bf16* sD1 = sX3 + 4096;
  bf16* sD2 = sD1 + 4096;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y;
  for (int m = 0; m < 6; ++m) load_w64(sW + m * 4096, wfull + m * 4096, tid);
  for (int m = 0; m < 7; ++m) load_w64(sM + m * 4096, mats + ((int64_t)b * 7 + m) * 4096, tid);
  cp_async_commit();
  const int g = lane >> 2, tq = lane & 3;
  const int64_t ntiles = (HW + kTilePx - 1) / kTilePx;
  const int64_t rb = (int64_t)b * HW;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t p0 = t * kTilePx;
    __syncthreads();
    load_rows_async(sX1, x1 + rb * ld1, p0, HW, ld1, 64, kTilePx, tid);
    load_rows_async(sX2, x2 + rb * ld2, p0, HW, ld2, 64, kTilePx, tid);
    load_rows_async(sX3, x3 + rb * ld3, p0, HW, ld3, 64, kTilePx, tid);
    load_rows_async(sD1, dr1 + rb * 64, p0, HW, 64, 64, kTilePx, tid);
    load_rows_async(sD2, dr2 + rb * 64, p0, HW, 64, 64, kTilePx, tid);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    const int row0 = warp * 16;
    float acc[8][4], d[8][4];
    // stream 3, u half: dP3u = (u3 S3) * 1[u3 > 0]
    {
      uint32_t au[4][4];
      proj16x64(acc, sX3, 64, row0, sW + 5 * 4096, lane);
      const uint32_t mk = relu_mask(acc, bfull + 2 * 128 + 64, tq);
      relu_bias_to_afrag(au, acc, bfull + 2 * 128 + 64, tq);
      zero_acc(d);
      apply64(d, au, sM + 2 * 4096, lane);
      store_masked(dP3, rb, 64, d, mk, row0, g, tq, p0, HW);
    }
    // stream 3, y half: dP3y = (dr1 Mz1 + dr2 Mz2) * 1[y3 > 0]
    {
      proj16x64(acc, sX3, 64, row0, sW + 4 * 4096, lane);
      const uint32_t mk = relu_mask(acc, bfull + 2 * 128, tq);
      zero_acc(d);
      mm16x64_acc(d, sD1, 64, row0, sM + 3 * 4096, lane);
      mm16x64_acc(d, sD2, 64, row0, sM + 5 * 4096, lane);
      store_masked(dP3, rb, 0, d, mk, row0, g, tq, p0, HW);
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const bf16* sX = s == 0 ? sX1 : sX2;
      const bf16* sD = s == 0 ? sD1 : sD2;
      bf16* dP = s == 0 ? dP1 : dP2;
      {   // y half: dPy = (y S_s) * 1[y > 0]
        uint32_t ay[4][4];
        proj16x64(acc, sX, 64, row0, sW + (2 * s) * 4096, lane);
        const uint32_t mk = relu_mask(acc, bfull + s * 128, tq);
        relu_bias_to_afrag(ay, acc, bfull + s * 128, tq);
        zero_acc(d);
        apply64(d, ay, sM + s * 4096, lane);
        store_masked(dP, rb, 0, d, mk, row0, g, tq, p0, HW);
      }
      {   // u half: dPu = (dr Mv_s) * 1[u > 0]
        proj16x64(acc, sX, 64, row0, sW + (2 * s + 1) * 4096, lane);
        const uint32_t mk = relu_mask(acc, bfull + s * 128 + 64, tq);
        zero_acc(d);
        mm16x64_acc(d, sD, 64, row0, sM + (4 + 2 * s) * 4096, lane);
        store_masked(dP, rb, 64, d, mk, row0, g, tq, p0, HW);
      }
    }
  }
  cp_async_wait<0>();
}

}  // namespace segmif

using namespace segmif;

static int opt_in_smem(const void* fn, size_t bytes, const char* what, bool* done) {
  if (*done) return SEGMIF_OK;
  *done = true;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) { set_error("%s: cudaFuncSetAttribute(%zu) failed: %s", what, bytes, cudaGetErrorString(e)); return SEGMIF_ERR_CUDA; }
  return SEGMIF_OK;
}

extern "C" int segmif_ffm_bwd_gram(const void* x1, int ld1, int coff1, const void* x2, int ld2, int coff2, const void* x3,
                                   int ld3, int coff3, const void* dr1, const void* dr2, const void* wfull,
                                   const float* bfull, float* partials, int nchunk, int B, int64_t HW,
                                   segmif_stream_t stream) {
  SEGMIF_REQUIRE(x1 && x2 && x3 && dr1 && dr2 && wfull && bfull && partials, "ffm_bwd_gram: null pointer");
  SEGMIF_REQUIRE(ld1 % 8 == 0 && ld2 % 8 == 0 && ld3 % 8 == 0 && coff1 % 8 == 0 && coff2 % 8 == 0 && coff3 % 8 == 0, "ffm_bwd_gram: pitches/offsets must be multiples of 8");
  SEGMIF_REQUIRE(nchunk > 0 && B > 0 && HW > 0, "ffm_bwd_gram: bad sizes");
  const size_t smem = 11 * 4096 * sizeof(bf16);
  static bool cfg = false;
  int rc = opt_in_smem((const void*)ffm_bwd_gram_kernel, smem, "ffm_bwd_gram", &cfg);
  if (rc) return rc;
  ffm_bwd_gram_kernel<<<dim3(nchunk, B), kFfmThreads, smem, as_stream(stream)>>>(
      (const bf16*)x1 + coff1, ld1, (const bf16*)x2 + coff2, ld2, (const bf16*)x3 + coff3, ld3, (const bf16*)dr1, (const bf16*)dr2,
      (const bf16*)wfull, bfull, partials, HW);
  return check_launch("segmif_ffm_bwd_gram");
}

extern "C" int segmif_ffm_bwd_ctx(const float* r_partials, int nchunk_r, const float* g_partials, int nchunk_g,
                                  const float* ctx, const float* wkv, const float* wend, const void* folded, void* mats,
                                  float* dwkv, float* dwend, int B, segmif_stream_t stream) {
  SEGMIF_REQUIRE(r_partials && g_partials && ctx && wkv && wend && folded && mats && dwkv && dwend, "ffm_bwd_ctx: null pointer");
  SEGMIF_REQUIRE(nchunk_r > 0 && nchunk_g > 0 && B > 0, "ffm_bwd_ctx: bad sizes");
  const size_t smem = (size_t)(4 * 4096 + 3 * 64 * 65 + 2 * 512) * sizeof(float);
  static bool cfg = false;
  int rc = opt_in_smem((const void*)ffm_bwd_ctx_kernel, smem, "ffm_bwd_ctx", &cfg);
  if (rc) return rc;
  ffm_bwd_ctx_kernel<<<dim3(3, B), kCtxThreads, smem, as_stream(stream)>>>(r_partials, nchunk_r, g_partials, nchunk_g, ctx, wkv, wend,
                                                          (const bf16*)folded, (bf16*)mats, dwkv, dwend);
  return check_launch("segmif_ffm_bwd_ctx");
}

extern "C" int segmif_ffm_bwd_apply(const void* x1, int ld1, int coff1, const void* x2, int ld2, int coff2, const void* x3,
                                    int ld3, int coff3, const void* dr1, const void* dr2, const void* wfull,
                                    const float* bfull, const void* mats, void* dP1, void* dP2, void* dP3, int B, int64_t HW,
                                    segmif_stream_t stream) {
  SEGMIF_REQUIRE(x1 && x2 && x3 && dr1 && dr2 && wfull && bfull && mats && dP1 && dP2 && dP3, "ffm_bwd_apply: null pointer");
  SEGMIF_REQUIRE(ld1 % 8 == 0 && ld2 % 8 == 0 && ld3 % 8 == 0 && coff1 % 8 == 0 && coff2 % 8 == 0 && coff3 % 8 == 0, "ffm_bwd_apply: pitches/offsets must be multiples of 8");
  SEGMIF_REQUIRE(B > 0 && HW > 0, "ffm_bwd_apply: bad sizes");
  const size_t smem = (size_t)(6 + 7 + 5) * 4096 * sizeof(bf16);
  static bool cfg = false;
  int rc = opt_in_smem((const void*)ffm_bwd_apply_kernel, smem, "ffm_bwd_apply", &cfg);
  if (rc) return rc;
  const int64_t ntiles = (HW + kTilePx - 1) / kTilePx;
  const int per_image = (int)std::min<int64_t>(ntiles, std::max<int64_t>(1, (148 * 4) / B));
  ffm_bwd_apply_kernel<<<dim3(per_image, B), kFfmThreads, smem, as_stream(stream)>>>(
      (const bf16*)x1 + coff1, ld1, (const bf16*)x2 + coff2, ld2, (const bf16*)x3 + coff3, ld3, (const bf16*)dr1, (const bf16*)dr2,
      (const bf16*)wfull, bfull, (const bf16*)mats, (bf16*)dP1, (bf16*)dP2, (bf16*)dP3, HW);
  return check_launch("segmif_ffm_bwd_apply");
}

// Bandwidth-bound row / pixel kernels of the MiT encoder and the decode head:
// LayerNorm, depthwise 3x3 + GELU, 7x7 patch embedding + LayerNorm, bilinear resize,
// upsample+argmax, NCHW<->NHWC converters.  All fp32 math, 128-bit accesses where alignment allows.
#include <algorithm>

#include "common.cuh"

namespace segmif {

// ------------------------------------------------------------------------------------ LayerNorm
// One warp per row; each lane owns VEC consecutive channels per 32*VEC-wide slab (<= 16 values).
template <typename TI, typename TO, int VEC>
__global__ void __launch_bounds__(256) layernorm_kernel(const TI* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, TO* __restrict__ y,
                                                        int64_t rows, int C, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  constexpr int MAXIT = 16 / VEC;
  const int nit = C / (32 * VEC);
  const TI* xr = x + row * C;
  float v[MAXIT][VEC];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXIT; ++i) {
    if (i < nit) {
      const int c = (i * 32 + lane) * VEC;
#pragma unroll
      for (int j = 0; j < VEC; ++j) { v[i][j] = ld_as_float(xr + c + j); s += v[i][j]; }
    }
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXIT; ++i)
    if (i < nit) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) { const float d = v[i][j] - mean; q += d * d; }
    }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
  TO* yr = y + row * C;
#pragma unroll
  for (int i = 0; i < MAXIT; ++i)
    if (i < nit) {
      const int c = (i * 32 + lane) * VEC;
#pragma unroll
      for (int j = 0; j < VEC; ++j) st_from_float(yr + c + j, (v[i][j] - mean) * rstd * gamma[c + j] + beta[c + j]);
    }
}

// Narrow rows (C <= 128): LPR = C/4 lanes own one row (one 128-bit load each), a warp covers 32/LPR rows per pass and
// RPW passes are unrolled so that several independent loads are in flight per lane -- a warp-per-row layout leaves a
// single 8-byte load per lane outstanding and runs at ~20 % of the HBM roofline on the [153600, 64] stage-1 tensors.
template <typename TO, int C, int PASSES>
__global__ void __launch_bounds__(256) layernorm_narrow_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, TO* __restrict__ y,
                                                               int64_t rows, float eps) {
  constexpr int LPR = C / 4, RPP = 32 / LPR;                 // lanes per row, rows per pass
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, li = lane % LPR;
  const int64_t warp_global = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t row0 = warp_global * (RPP * PASSES) + sub;
  const float4 g = *reinterpret_cast<const float4*>(gamma + li * 4);
  const float4 b = *reinterpret_cast<const float4*>(beta + li * 4);
  float4 v[PASSES];
#pragma unroll
  for (int p = 0; p < PASSES; ++p) {
    const int64_t row = row0 + (int64_t)p * RPP;
    v[p] = row < rows ? *reinterpret_cast<const float4*>(x + row * C + li * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int p = 0; p < PASSES; ++p) {
    const int64_t row = row0 + (int64_t)p * RPP;
    float s = (v[p].x + v[p].y) + (v[p].z + v[p].w);
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / C);
    const float dx = v[p].x - mean, dy = v[p].y - mean, dz = v[p].z - mean, dw = v[p].w - mean;
    float q = (dx * dx + dy * dy) + (dz * dz + dw * dw);
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.0f / C) + eps);
    if (row < rows) {
      const float o0 = dx * rstd * g.x + b.x, o1 = dy * rstd * g.y + b.y, o2 = dz * rstd * g.z + b.z, o3 = dw * rstd * g.w + b.w;
      if (sizeof(TO) == 2) {
        *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(y) + row * C + li * 4) = make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
      } else {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + row * C + li * 4) = make_float4(o0, o1, o2, o3);
      }
    }
  }
}

template <typename TO, int C>
static int launch_ln_narrow(const void* x, const float* g, const float* b, void* y, int64_t rows, float eps, cudaStream_t st) {
  constexpr int PASSES = 4, RPW = (32 / (C / 4)) * PASSES;   // rows per warp
  const int64_t warps = ceil_div(rows, RPW);
  layernorm_narrow_kernel<TO, C, PASSES><<<(unsigned)ceil_div(warps, 8), 256, 0, st>>>((const float*)x, g, b, (TO*)y, rows, eps);
  return check_launch("segmif_layernorm_fwd");
}

template <typename TI, typename TO>
static int launch_ln(const void* x, const float* g, const float* b, void* y, int64_t rows, int C, float eps,
                     cudaStream_t st) {
  const int wpb = 8;
  dim3 grid((unsigned)ceil_div(rows, wpb));
  if (C % 128 == 0 && C / 128 <= 4)
    layernorm_kernel<TI, TO, 4><<<grid, wpb * 32, 0, st>>>((const TI*)x, g, b, (TO*)y, rows, C, eps);
  else if (C % 64 == 0 && C / 64 <= 8)
    layernorm_kernel<TI, TO, 2><<<grid, wpb * 32, 0, st>>>((const TI*)x, g, b, (TO*)y, rows, C, eps);
  else
    layernorm_kernel<TI, TO, 1><<<grid, wpb * 32, 0, st>>>((const TI*)x, g, b, (TO*)y, rows, C, eps);
  return check_launch("segmif_layernorm_fwd");
}

// ------------------------------------------------------------------------------------ DWConv + GELU
// erf through Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, far below the bf16 output resolution): two MUFU ops and
// seven FMAs instead of libdevice's ~28-instruction erff -- the kernel is instruction-issue bound, not HBM bound.
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __frcp_rn(fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = exp2f(-1.4426950408889634f * z * z);
  const float erf_abs = fmaf(-p * t, e, 1.0f);
  return 0.5f * x * (1.0f + copysignf(erf_abs, x));
}

// thread = (pixel, 8-channel group); neighbours come through L1/L2 (each input line is reused 9x).  grid = (chunks of
// one image row, H, B) so that every index is 32-bit: the first version decomposed a flat 64-bit index with four 64-bit
// divisions per thread (12 CALLs, 1144 SASS instructions, 78% issue-bound at 151 us for the stage-1 map).  (A two-pixel
// sliding window with the weights in registers was tried: 118 registers, lower occupancy, 1.5x slower.)
__global__ void __launch_bounds__(256) dwconv3x3_gelu_kernel(const bf16* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ bias, bf16* __restrict__ y,
                                                             int H, int W, int C) {
  const unsigned cg = (unsigned)C >> 3;
  const unsigned t = blockIdx.x * 256u + threadIdx.x;
  if (t >= (unsigned)W * cg) return;
  const unsigned xw = t / cg;
  const int c = (int)(t - xw * cg) * 8;
  const int yh = blockIdx.y;
  const unsigned ctr = ((blockIdx.z * (unsigned)H + yh) * (unsigned)W + xw) * (unsigned)C + c;   // element index of the centre tap
  const int rs = W * C;                                            // row stride in elements
  float acc[8];
  load8(bias + c, acc);
  if (yh >= 1 && yh + 1 < H && xw >= 1 && xw + 1 < (unsigned)W) {       // interior pixel: nine unconditional taps
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        float v[8], wv[8];
        load8(x + (ctr + (unsigned)((ky - 1) * rs + (kx - 1) * C)), v);
        load8(w + (unsigned)((ky * 3 + kx) * C + c), wv);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(v[j], wv[j], acc[j]);
      }
  } else {
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      if ((unsigned)(yh + ky - 1) >= (unsigned)H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        if (xw + kx - 1 >= (unsigned)W) continue;                    // unsigned wrap covers xw + kx - 1 < 0
        float v[8], wv[8];
        load8(x + (ctr + (unsigned)((ky - 1) * rs + (kx - 1) * C)), v);
        load8(w + (unsigned)((ky * 3 + kx) * C + c), wv);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(v[j], wv[j], acc[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = gelu_fast(acc[j]);
  store8(y + ctr, acc);
}

// ------------------------------------------------------------------------------------ patch embed 7x7 s4 + LN
// One warp produces PX=4 horizontally adjacent output pixels; lane owns CPL = C0/32 channels.
// Weights [147][C0] live in smem; the 7 x 19 x 3 input window is staged per warp.
template <int CPL>
__global__ void __launch_bounds__(256) patch_embed7_ln_kernel(const float* __restrict__ img, const float* __restrict__ w,
                                                              const float* __restrict__ bias,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, float eps,
                                                              const float* __restrict__ in_scale,
                                                              const float* __restrict__ in_shift,
                                                              float* __restrict__ tokens, int B, int H, int W, int Ho,
                                                              int Wo) {
  constexpr int C0 = CPL * 32, PX = 4, WIN = 7 + 4 * (PX - 1);   // 19 input columns
  extern __shared__ float smem[];
  float* sw = smem;                                   // [147][C0]
  float* swin = smem + 147 * C0;                      // [warps][3][7][WIN]
  for (int i = threadIdx.x; i < 147 * C0; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  float* win = swin + warp * (3 * 7 * WIN);
  const int groups_x = (Wo + PX - 1) / PX;
  const int64_t ngroups = (int64_t)B * Ho * groups_x;
  float sc[3], sh[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) { sc[c] = in_scale ? in_scale[c] : 1.f; sh[c] = in_shift ? in_shift[c] : 0.f; }
  for (int64_t gidx = (int64_t)blockIdx.x * nwarp + warp; gidx < ngroups; gidx += (int64_t)gridDim.x * nwarp) {
    const int gx = (int)(gidx % groups_x);
    const int oy = (int)((gidx / groups_x) % Ho);
    const int b = (int)(gidx / ((int64_t)groups_x * Ho));
    const int ox0 = gx * PX;
    const int iy0 = oy * 4 - 3, ix0 = ox0 * 4 - 3;
    __syncwarp();
    for (int i = lane; i < 3 * 7 * WIN; i += 32) {
      const int c = i / (7 * WIN), r = (i / WIN) % 7, col = i % WIN;
      const int iy = iy0 + r, ix = ix0 + col;
      float v = 0.f;
      if ((unsigned)iy < (unsigned)H && (unsigned)ix < (unsigned)W)
        v = img[(((int64_t)b * 3 + c) * H + iy) * W + ix] * sc[c] + sh[c];
      win[i] = v;
    }
    __syncwarp();
    float acc[PX][CPL];
#pragma unroll
    for (int p = 0; p < PX; ++p)
#pragma unroll
      for (int j = 0; j < CPL; ++j) acc[p][j] = bias[lane * CPL + j];
    for (int c = 0; c < 3; ++c)
      for (int ky = 0; ky < 7; ++ky) {
        const float* wr = win + (c * 7 + ky) * WIN;
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
          const float* wp = sw + ((c * 7 + ky) * 7 + kx) * C0 + lane * CPL;
          float wv[CPL];
#pragma unroll
          for (int j = 0; j < CPL; ++j) wv[j] = wp[j];
#pragma unroll
          for (int p = 0; p < PX; ++p) {
            const float xv = wr[kx + 4 * p];
#pragma unroll
            for (int j = 0; j < CPL; ++j) acc[p][j] = fmaf(xv, wv[j], acc[p][j]);
          }
        }
      }
#pragma unroll
    for (int p = 0; p < PX; ++p) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < CPL; ++j) s += acc[p][j];
      const float mean = warp_sum(s) / (float)C0;
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < CPL; ++j) { const float d = acc[p][j] - mean; q += d * d; }
      const float rstd = rsqrtf(warp_sum(q) / (float)C0 + eps);
      const int ox = ox0 + p;
      if (ox < Wo) {
        float* o = tokens + (((int64_t)b * Ho + oy) * Wo + ox) * C0 + lane * CPL;
#pragma unroll
        for (int j = 0; j < CPL; ++j) o[j] = (acc[p][j] - mean) * rstd * gamma[lane * CPL + j] + beta[lane * CPL + j];
      }
    }
  }
}

// ------------------------------------------------------------------------------------ bilinear (align_corners=False)
// Index arithmetic mirrors ATen's upsample_bilinear2d: src = scale*(dst+0.5)-0.5 clamped at 0, fp32.
__device__ __forceinline__ void bilinear_src(int dst, float scale, int in_size, int& i0, int& i1, float& l0, float& l1) {
  float s = scale * ((float)dst + 0.5f) - 0.5f;
  s = s < 0.f ? 0.f : s;
  i0 = (int)s;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  l1 = s - (float)i0;
  l0 = 1.f - l1;
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) bilinear_nhwc_kernel(const TI* __restrict__ src, int B, int h, int w, int C,
                                                            int ld_src, TO* __restrict__ dst, int H, int W, int ld_dst,
                                                            int dst_coff, float sy, float sx) {
  const int cg = C >> 3;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)B * H * W * cg) return;
  const int c = (int)(idx % cg) * 8;
  const int64_t pix = idx / cg;
  const int X = (int)(pix % W), Y = (int)((pix / W) % H);
  const int64_t b = pix / ((int64_t)W * H);
  int y0, y1, x0, x1;
  float hy0, hy1, wx0, wx1;
  bilinear_src(Y, sy, h, y0, y1, hy0, hy1);
  bilinear_src(X, sx, w, x0, x1, wx0, wx1);
  const TI* base = src + b * h * w * ld_src + c;
  float a[8], bq[8], cc[8], d[8], o[8];
  load8(base + ((int64_t)y0 * w + x0) * ld_src, a);
  load8(base + ((int64_t)y0 * w + x1) * ld_src, bq);
  load8(base + ((int64_t)y1 * w + x0) * ld_src, cc);
  load8(base + ((int64_t)y1 * w + x1) * ld_src, d);
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = hy0 * (wx0 * a[j] + wx1 * bq[j]) + hy1 * (wx0 * cc[j] + wx1 * d[j]);
  store8(dst + pix * ld_dst + dst_coff + c, o);
}

// logits fp32 [B,h,w,nc] -> labels; lowest index wins ties (strict >), like torch.argmax.
__global__ void __launch_bounds__(256) upsample_argmax_kernel(const float* __restrict__ logits, int B, int h, int w,
                                                              int nc, int64_t* __restrict__ labels, int H, int W,
                                                              float sy, float sx) {
  const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= (int64_t)B * H * W) return;
  const int X = (int)(pix % W), Y = (int)((pix / W) % H);
  const int64_t b = pix / ((int64_t)W * H);
  int y0, y1, x0, x1;
  float hy0, hy1, wx0, wx1;
  bilinear_src(Y, sy, h, y0, y1, hy0, hy1);
  bilinear_src(X, sx, w, x0, x1, wx0, wx1);
  const float* base = logits + b * h * w * nc;
  const float* p00 = base + ((int64_t)y0 * w + x0) * nc;
  const float* p01 = base + ((int64_t)y0 * w + x1) * nc;
  const float* p10 = base + ((int64_t)y1 * w + x0) * nc;
  const float* p11 = base + ((int64_t)y1 * w + x1) * nc;
  float best = -INFINITY;
  int arg = 0;
  for (int c = 0; c < nc; ++c) {
    const float v = hy0 * (wx0 * p00[c] + wx1 * p01[c]) + hy1 * (wx0 * p10[c] + wx1 * p11[c]);
    if (v > best || c == 0) { best = v; arg = c; }
  }
  labels[pix] = arg;
}

// ------------------------------------------------------------------------------------ layout converters
// 32-pixel x 32-channel smem tiles so that both sides are coalesced.
template <typename TI>
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const TI* __restrict__ src, int ld_src, int src_coff,
                                                           float* __restrict__ dst, int64_t HW, int C) {
  __shared__ float tile[32][33];
  const int64_t b = blockIdx.z;
  const int64_t p0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 8 rows per pass
  for (int r = ty; r < 32; r += 8) {
    const int64_t p = p0 + r;
    const int c = c0 + tx;
    tile[r][tx] = (p < HW && c < C) ? ld_as_float(src + (b * HW + p) * ld_src + src_coff + c) : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r;
    const int64_t p = p0 + tx;
    if (p < HW && c < C) dst[(b * C + c) * HW + p] = tile[tx][r];
  }
}

template <typename TO>
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ src, TO* __restrict__ dst,
                                                           int ld_dst, int dst_coff, int64_t HW, int C) {
  __shared__ float tile[32][33];
  const int64_t b = blockIdx.z;
  const int64_t p0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r;
    const int64_t p = p0 + tx;
    tile[r][tx] = (p < HW && c < C) ? src[(b * C + c) * HW + p] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int64_t p = p0 + r;
    const int c = c0 + tx;
    if (p < HW && c < C) st_from_float(dst + (b * HW + p) * ld_dst + dst_coff + c, tile[tx][r]);
  }
}

}  // namespace segmif

using namespace segmif;

extern "C" int segmif_layernorm_fwd(const void* x, int x_dtype, const float* gamma, const float* beta, void* y,
                                    int y_dtype, int64_t rows, int C, float eps, segmif_stream_t stream) {
  SEGMIF_REQUIRE(x && y && gamma && beta, "layernorm: null pointer");
  SEGMIF_REQUIRE(C > 0 && C % 32 == 0 && C <= 512, "layernorm: C=%d must be a multiple of 32 and <= 512", C);
  if (rows == 0) return SEGMIF_OK;
  cudaStream_t st = as_stream(stream);
  if (x_dtype == SEGMIF_F32 && (C == 32 || C == 64 || C == 128) && ((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0 &&
      ((uintptr_t)gamma & 15) == 0 && ((uintptr_t)beta & 15) == 0) {
    const bool ob = y_dtype == SEGMIF_BF16;
    if (C == 32) return ob ? launch_ln_narrow<bf16, 32>(x, gamma, beta, y, rows, eps, st) : launch_ln_narrow<float, 32>(x, gamma, beta, y, rows, eps, st);
    if (C == 64) return ob ? launch_ln_narrow<bf16, 64>(x, gamma, beta, y, rows, eps, st) : launch_ln_narrow<float, 64>(x, gamma, beta, y, rows, eps, st);
    return ob ? launch_ln_narrow<bf16, 128>(x, gamma, beta, y, rows, eps, st) : launch_ln_narrow<float, 128>(x, gamma, beta, y, rows, eps, st);
  }
  if (x_dtype == SEGMIF_F32 && y_dtype == SEGMIF_BF16) return launch_ln<float, bf16>(x, gamma, beta, y, rows, C, eps, st);
  if (x_dtype == SEGMIF_F32 && y_dtype == SEGMIF_F32) return launch_ln<float, float>(x, gamma, beta, y, rows, C, eps, st);
  if (x_dtype == SEGMIF_BF16 && y_dtype == SEGMIF_BF16) return launch_ln<bf16, bf16>(x, gamma, beta, y, rows, C, eps, st);
  if (x_dtype == SEGMIF_BF16 && y_dtype == SEGMIF_F32) return launch_ln<bf16, float>(x, gamma, beta, y, rows, C, eps, st);
  set_error("layernorm: unsupported dtypes %d -> %d", x_dtype, y_dtype);
  return SEGMIF_ERR_INVALID;
}

extern "C" int segmif_dwconv3x3_gelu_fwd(const void* x, const float* w9c, const float* bias, void* y, int B, int H,
                                         int W, int C, segmif_stream_t stream) {
  SEGMIF_REQUIRE(x && w9c && bias && y, "dwconv: null pointer");
  SEGMIF_REQUIRE(C % 8 == 0, "dwconv: C=%d must be a multiple of 8", C);
  if ((int64_t)B * H * W == 0) return SEGMIF_OK;
  if (dwconv_tma_ok(B, H, W, C)) return dwconv_tma_fwd(x, w9c, bias, y, B, H, W, C, 0, 1, as_stream(stream));
  SEGMIF_REQUIRE((int64_t)B * H * W * C < (1ll << 32) && H <= 65535 && B <= 65535, "dwconv: tensor of %d x %d x %d x %d elements exceeds the 32-bit element index", B, H, W, C);
  dim3 grid((unsigned)ceil_div((int64_t)W * (C / 8), 256), (unsigned)H, (unsigned)B);
  dwconv3x3_gelu_kernel<<<grid, 256, 0, as_stream(stream)>>>((const bf16*)x, w9c, bias, (bf16*)y, H, W, C);
  return check_launch("segmif_dwconv3x3_gelu_fwd");
}

extern "C" int segmif_patch_embed7_ln_fwd(const float* img, const float* w, const float* bias, const float* gamma,
                                          const float* beta, float eps, const float* in_scale3,
                                          const float* in_shift3, float* tokens, int B, int H, int W, int C0,
                                          segmif_stream_t stream) {
  SEGMIF_REQUIRE(img && w && bias && gamma && beta && tokens, "patch_embed: null pointer");
  SEGMIF_REQUIRE(C0 == 32 || C0 == 64, "patch_embed: C0=%d must be 32 or 64", C0);
  const int Ho = (H + 6 - 7) / 4 + 1, Wo = (W + 6 - 7) / 4 + 1;
  const int warps = 8;
  const size_t smem = (size_t)(147 * C0 + warps * 3 * 7 * 19) * sizeof(float);
  const int64_t groups = (int64_t)B * Ho * ((Wo + 3) / 4);
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(groups, warps), 148 * 8);
  cudaStream_t st = as_stream(stream);
  if (C0 == 64) {
    static bool cfg = false;
    if (!cfg) { cudaFuncSetAttribute(patch_embed7_ln_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); cfg = true; }
    patch_embed7_ln_kernel<2><<<grid, warps * 32, smem, st>>>(img, w, bias, gamma, beta, eps, in_scale3, in_shift3, tokens, B, H, W, Ho, Wo);
  } else {
    patch_embed7_ln_kernel<1><<<grid, warps * 32, smem, st>>>(img, w, bias, gamma, beta, eps, in_scale3, in_shift3, tokens, B, H, W, Ho, Wo);
  }
  return check_launch("segmif_patch_embed7_ln_fwd");
}

extern "C" int segmif_bilinear_nhwc_fwd(const void* src, int src_dtype, int B, int h, int w, int C, int ld_src,
                                        void* dst, int dst_dtype, int H, int W, int ld_dst, int dst_coff,
                                        segmif_stream_t stream) {
  SEGMIF_REQUIRE(src && dst, "bilinear: null pointer");
  SEGMIF_REQUIRE(C % 8 == 0 && ld_src % 8 == 0 && ld_dst % 8 == 0 && dst_coff % 8 == 0, "bilinear: channels must be multiples of 8");
  const int64_t total = (int64_t)B * H * W * (C / 8);
  if (total == 0) return SEGMIF_OK;
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  const unsigned grid = (unsigned)ceil_div(total, 256);
  cudaStream_t st = as_stream(stream);
#define SEGMIF_BL(TI, TO) bilinear_nhwc_kernel<TI, TO><<<grid, 256, 0, st>>>((const TI*)src, B, h, w, C, ld_src, (TO*)dst, H, W, ld_dst, dst_coff, sy, sx)
  if (src_dtype == SEGMIF_F32 && dst_dtype == SEGMIF_F32) SEGMIF_BL(float, float);
  else if (src_dtype == SEGMIF_F32 && dst_dtype == SEGMIF_BF16) SEGMIF_BL(float, bf16);
  else if (src_dtype == SEGMIF_BF16 && dst_dtype == SEGMIF_BF16) SEGMIF_BL(bf16, bf16);
  else if (src_dtype == SEGMIF_BF16 && dst_dtype == SEGMIF_F32) SEGMIF_BL(bf16, float);
  else { set_error("bilinear: bad dtypes"); return SEGMIF_ERR_INVALID; }
#undef SEGMIF_BL
  return check_launch("segmif_bilinear_nhwc_fwd");
}

extern "C" int segmif_upsample_argmax_fwd(const float* logits, int B, int h, int w, int nc, int64_t* labels, int H,
                                          int W, segmif_stream_t stream) {
  SEGMIF_REQUIRE(logits && labels && nc > 0, "upsample_argmax: bad arguments");
  const int64_t total = (int64_t)B * H * W;
  if (total == 0) return SEGMIF_OK;
  upsample_argmax_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, as_stream(stream)>>>(
      logits, B, h, w, nc, labels, H, W, (float)h / (float)H, (float)w / (float)W);
  return check_launch("segmif_upsample_argmax_fwd");
}

extern "C" int segmif_nhwc_to_nchw(const void* src, int src_dtype, int ld_src, int src_coff, float* dst, int B,
                                   int HW, int C, segmif_stream_t stream) {
  SEGMIF_REQUIRE(src && dst, "nhwc_to_nchw: null pointer");
  if ((int64_t)B * HW * C == 0) return SEGMIF_OK;
  dim3 grid((unsigned)ceil_div(HW, 32), (unsigned)ceil_div(C, 32), B);
  if (src_dtype == SEGMIF_F32) nhwc_to_nchw_kernel<float><<<grid, 256, 0, as_stream(stream)>>>((const float*)src, ld_src, src_coff, dst, HW, C);
  else nhwc_to_nchw_kernel<bf16><<<grid, 256, 0, as_stream(stream)>>>((const bf16*)src, ld_src, src_coff, dst, HW, C);
  return check_launch("segmif_nhwc_to_nchw");
}

extern "C" int segmif_nchw_to_nhwc(const float* src, void* dst, int dst_dtype, int ld_dst, int dst_coff, int B, int HW,
                                   int C, segmif_stream_t stream) {
  SEGMIF_REQUIRE(src && dst, "nchw_to_nhwc: null pointer");
  if ((int64_t)B * HW * C == 0) return SEGMIF_OK;
  dim3 grid((unsigned)ceil_div(HW, 32), (unsigned)ceil_div(C, 32), B);
  if (dst_dtype == SEGMIF_F32) nchw_to_nhwc_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(src, (float*)dst, ld_dst, dst_coff, HW, C);
  else nchw_to_nhwc_kernel<bf16><<<grid, 256, 0, as_stream(stream)>>>(src, (bf16*)dst, ld_dst, dst_coff, HW, C);
  return check_launch("segmif_nchw_to_nhwc");
}

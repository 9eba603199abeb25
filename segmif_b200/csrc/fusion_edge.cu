// Edge layers of Fusion_Network3_ac (1-channel in / 1-channel out 3x3 convs with the shared PReLU)
// and the colour transforms.  All HBM-bound; fp32 math.
#include "common.cuh"

namespace segmif {

// conv1_ir / conv1_vis: fp32 plane [B,H,W] -> bf16 pixel-major, Cout channels (multiple of 8).
// A thread owns one 8-channel group (its 72 weights stay in registers) and 4 horizontally consecutive pixels; a warp covers
// 16 consecutive pixels x 8 groups, so every store instruction writes whole 128-byte pixel records.  The warp walks kIn1Rows
// rows of its 16-pixel column strip with a three-row register window (6 new input loads per row instead of 18) -- the first
// version reloaded the 72 weights + bias for every 4 pixels and reached 1.7 TB/s of stores; this one amortises them over
// 4 * kIn1Rows pixels.  Requires Cout == 64 (8 groups) for full warps; other widths use more warps per pixel block.
constexpr int kIn1Rows = 16;
// packed fp32 pairs (FFMA2): two output channels per issue slot; same roundings as scalar fmaf
__device__ __forceinline__ float2 in1_fma2(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                     rc = *reinterpret_cast<unsigned long long*>(&c), rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
// Round 2: 146 registers gave ONE 256-thread block per SM (12 % occupancy, issue-active 48 %: profiles/r2_ncu_misc_summary.csv);
// capped at 128 for two blocks per SM, and the 72 FMAs per pixel and channel group issued as 36 FFMA2.
__global__ void __launch_bounds__(256, 2) conv3x3_in1_kernel(const float* __restrict__ plane, int64_t bstride,
                                                          const float* __restrict__ w, const float* __restrict__ bias,
                                                          const float* __restrict__ alpha_p, bf16* __restrict__ dst,
                                                          int ld_dst, int dst_coff, int B, int H, int W, int Cout) {
  const int lane = threadIdx.x & 31, warp_in_blk = threadIdx.x >> 5;
  const int ngroups = Cout >> 3;                       // channel groups of 8
  const int gpw = ngroups < 8 ? ngroups : 8;           // groups handled by one warp (8 lanes-groups)
  const int cg_sets = (ngroups + 7) / 8;               // warps needed to cover all channels of one pixel block
  const int xblocks = (W + 15) / 16;
  const int ystrips = (H + kIn1Rows - 1) / kIn1Rows;
  const int64_t units = (int64_t)B * ystrips * xblocks * cg_sets;
  const int64_t unit = (int64_t)blockIdx.x * 8 + warp_in_blk;
  if (unit >= units) return;
  const unsigned uu = (unsigned)unit;                  // units < 2^31 (checked by the host): 32-bit divisions only
  const int cset = (int)(uu % (unsigned)cg_sets);
  const unsigned u2 = uu / (unsigned)cg_sets;
  const int xb = (int)(u2 % (unsigned)xblocks);
  const unsigned u3 = u2 / (unsigned)xblocks;
  const int ys = (int)(u3 % (unsigned)ystrips);
  const int64_t b = u3 / (unsigned)ystrips;
  const int grp = cset * 8 + (lane & 7), q = lane >> 3;
  if ((lane & 7) >= gpw || grp >= ngroups) return;
  const int c = grp * 8;
  const float alpha = *alpha_p;
  float2 wr[9][4], bs[4];
  {
    float t8[8];
    load8(bias + c, t8);
#pragma unroll
    for (int j = 0; j < 4; ++j) bs[j] = make_float2(t8[2 * j], t8[2 * j + 1]);
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      load8(w + t * Cout + c, t8);
#pragma unroll
      for (int j = 0; j < 4; ++j) wr[t][j] = make_float2(t8[2 * j], t8[2 * j + 1]);
    }
  }
  const int x0 = xb * 16 + q * 4;
  const float* src = plane + b * bstride;
  const int y_begin = ys * kIn1Rows, y_end = min(H, y_begin + kIn1Rows);
  auto load_row = [&](int iy, float (&row)[6]) {
#pragma unroll
    for (int cc = 0; cc < 6; ++cc) {
      const int ix = x0 + cc - 1;
      row[cc] = ((unsigned)iy < (unsigned)H && (unsigned)ix < (unsigned)W) ? __ldg(src + (int64_t)iy * W + ix) : 0.f;
    }
  };
  float in[3][6];
  load_row(y_begin - 1, in[0]);
  load_row(y_begin, in[1]);
  for (int y = y_begin; y < y_end; ++y) {
    load_row(y + 1, in[2]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int x = x0 + i;
      if (x >= W) break;
      float2 a2[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) a2[j] = bs[j];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float2 v = make_float2(in[r][i + kx], in[r][i + kx]);
#pragma unroll
          for (int j = 0; j < 4; ++j) a2[j] = in1_fma2(v, wr[r * 3 + kx][j], a2[j]);
        }
      float acc[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[2 * j] = a2[j].x >= 0.f ? a2[j].x : alpha * a2[j].x;
        acc[2 * j + 1] = a2[j].y >= 0.f ? a2[j].y : alpha * a2[j].y;
      }
      store8(dst + ((b * H + y) * W + x) * ld_dst + dst_coff + c, acc);
    }
#pragma unroll
    for (int cc = 0; cc < 6; ++cc) { in[0][cc] = in[1][cc]; in[1][cc] = in[2][cc]; }
  }
}

// conv22: bf16 pixel-major Cin channels -> fp32 plane [B,1,H,W]; a quad of 4 lanes shares one pixel (each lane owns
// Cin/4 = 8 channels, its 72 weights in registers), reduced with two shuffles.  A quad walks kOut1Run consecutive pixels of
// a row with a three-COLUMN window of packed bf16 (3 new 16-byte loads per pixel instead of 9; the first version also
// re-read the 9 x 8 weights from L1 for every pixel and reached 0.6 TB/s).  Cin == 32.
constexpr int kOut1Run = 8;
__global__ void __launch_bounds__(256) conv3x3_out1_c32_kernel(const bf16* __restrict__ src, int ld_src,
                                                               const float* __restrict__ w, const float* __restrict__ bias,
                                                               const float* __restrict__ alpha_p, float* __restrict__ dst,
                                                               int B, int H, int W) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t run = gid >> 2;
  const int q = (int)(gid & 3);
  const int runs_x = (W + kOut1Run - 1) / kOut1Run;
  const int64_t nruns = (int64_t)B * H * runs_x;
  const bool live = run < nruns;
  const int64_t rr = live ? run : 0;
  const int rx = (int)(rr % runs_x);
  const int Y = (int)((rr / runs_x) % H);
  const int64_t b = rr / ((int64_t)runs_x * H);
  float wr[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t) load8(w + t * 32 + q * 8, wr[t]);
  const float alpha = *alpha_p, b0 = bias[0];
  const int x_begin = rx * kOut1Run;
  auto load_col = [&](int ix, uint4 (&col)[3]) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int iy = Y + r - 1;
      col[r] = ((unsigned)iy < (unsigned)H && (unsigned)ix < (unsigned)W)
                   ? __ldg(reinterpret_cast<const uint4*>(src + ((b * H + iy) * W + ix) * ld_src + q * 8)) : make_uint4(0, 0, 0, 0);
    }
  };
  uint4 win[3][3];                        // [column][row]
  load_col(x_begin - 1, win[0]);
  load_col(x_begin, win[1]);
#pragma unroll 1
  for (int i = 0; i < kOut1Run; ++i) {
    const int X = x_begin + i;
    load_col(X + 1, win[2]);
    float acc = 0.f;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const uint4 u = win[kx][ky];
        const float2 p0 = unpack_bf16x2(u.x), p1 = unpack_bf16x2(u.y), p2 = unpack_bf16x2(u.z), p3 = unpack_bf16x2(u.w);
        const float* ww = wr[ky * 3 + kx];
        acc = fmaf(p0.x, ww[0], acc); acc = fmaf(p0.y, ww[1], acc); acc = fmaf(p1.x, ww[2], acc); acc = fmaf(p1.y, ww[3], acc);
        acc = fmaf(p2.x, ww[4], acc); acc = fmaf(p2.y, ww[5], acc); acc = fmaf(p3.x, ww[6], acc); acc = fmaf(p3.y, ww[7], acc);
      }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (live && q == 0 && X < W) {
      const float v = acc + b0;
      dst[(b * H + Y) * W + X] = v >= 0.f ? v : alpha * v;
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) { win[0][r] = win[1][r]; win[1][r] = win[2][r]; }
  }
}

// general Cin (multiple of 32): one pixel per quad, taps re-read (kept for shapes other than conv22's 32 channels)
__global__ void __launch_bounds__(256) conv3x3_out1_kernel(const bf16* __restrict__ src, int ld_src,
                                                           const float* __restrict__ w, const float* __restrict__ bias,
                                                           const float* __restrict__ alpha_p, float* __restrict__ dst,
                                                           int B, int H, int W, int Cin) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t pix = gid >> 2;
  const int q = (int)(gid & 3);
  const int64_t npix = (int64_t)B * H * W;
  const bool live = pix < npix;
  const int64_t pp = live ? pix : 0;
  const unsigned pu = (unsigned)pp;                        // B*H*W < 2^31 (checked by the host)
  const unsigned prow = pu / (unsigned)W;
  const int X = (int)(pu - prow * (unsigned)W), Y = (int)(prow % (unsigned)H);
  const int64_t b = prow / (unsigned)H;
  const int cpl = Cin >> 2;                 // channels per lane (multiple of 8)
  float acc = 0.f;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = Y + ky - 1;
    if ((unsigned)iy >= (unsigned)H) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = X + kx - 1;
      if ((unsigned)ix >= (unsigned)W) continue;
      const bf16* p = src + ((b * H + iy) * W + ix) * ld_src + q * cpl;
      const float* wp = w + (ky * 3 + kx) * Cin + q * cpl;
      for (int c = 0; c < cpl; c += 8) {
        float v[8], wv[8];
        load8(p + c, v);
        load8(wp + c, wv);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc = fmaf(v[j], wv[j], acc);
      }
    }
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  if (live && q == 0) {
    const float alpha = *alpha_p;
    float v = acc + bias[0];
    dst[pix] = v >= 0.f ? v : alpha * v;
  }
}

// ---- colour: constants exactly as core/model_fusion.py:69-111 -------------------------------------------------
__device__ __forceinline__ void rgb_to_ycc(float r, float g, float b, float& y, float& cr, float& cb) {
  y = 0.299f * r + 0.587f * g + 0.114f * b;
  cr = (r - y) * 0.713f + 0.5f;
  cb = (b - y) * 0.564f + 0.5f;
}
__device__ __forceinline__ void ycc_to_rgb(float y, float cr, float cb, float& r, float& g, float& b) {
  const float c1 = cr - 0.5f, c2 = cb - 0.5f;       // (im + bias) . mat, mat rows: [1,1,1],[1.403,-.714,0],[0,-.344,1.773]
  r = y * 1.0f + c1 * 1.403f + c2 * 0.0f;
  g = y * 1.0f + c1 * -0.714f + c2 * -0.344f;
  b = y * 1.0f + c1 * 0.0f + c2 * 1.773f;
}

__global__ void rgb2ycrcb_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int64_t HW) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * HW) return;
  const int64_t b = i / HW, p = i - b * HW;
  const float* s = in + b * 3 * HW + p;
  float y, cr, cb;
  rgb_to_ycc(s[0], s[HW], s[2 * HW], y, cr, cb);
  float* d = out + b * 3 * HW + p;
  d[0] = y; d[HW] = cr; d[2 * HW] = cb;
}

__global__ void ycrcb2rgb_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int64_t HW) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * HW) return;
  const int64_t b = i / HW, p = i - b * HW;
  const float* s = in + b * 3 * HW + p;
  float r, g, bl;
  ycc_to_rgb(s[0], s[HW], s[2 * HW], r, g, bl);
  float* d = out + b * 3 * HW + p;
  d[0] = r; d[HW] = g; d[2 * HW] = bl;
}

// fused: YCrCb(vis) with Y replaced by the fused plane -> RGB -> optional clamp
__global__ void recompose_rgb_kernel(const float* __restrict__ fused, const float* __restrict__ vis,
                                     float* __restrict__ out, int clamp01, int B, int64_t HW) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * HW) return;
  const int64_t b = i / HW, p = i - b * HW;
  const float* s = vis + b * 3 * HW + p;
  float y, cr, cb, r, g, bl;
  rgb_to_ycc(s[0], s[HW], s[2 * HW], y, cr, cb);
  ycc_to_rgb(fused[i], cr, cb, r, g, bl);
  if (clamp01) {
    r = fminf(fmaxf(r, 0.f), 1.f); g = fminf(fmaxf(g, 0.f), 1.f); bl = fminf(fmaxf(bl, 0.f), 1.f);
  }
  float* d = out + b * 3 * HW + p;
  d[0] = r; d[HW] = g; d[2 * HW] = bl;
}

}  // namespace segmif

using namespace segmif;

extern "C" int segmif_conv3x3_in1_fwd(const float* plane, int64_t bstride, const float* w, const float* bias,
                                      const float* prelu_alpha, void* dst, int ld_dst, int dst_coff, int B, int H,
                                      int W, int Cout, segmif_stream_t stream) {
  SEGMIF_REQUIRE(plane && w && bias && prelu_alpha && dst, "conv3x3_in1: null pointer");
  SEGMIF_REQUIRE(Cout % 8 == 0 && ld_dst % 8 == 0 && dst_coff % 8 == 0, "conv3x3_in1: channel counts must be multiples of 8");
  const int64_t units = (int64_t)B * ((H + kIn1Rows - 1) / kIn1Rows) * ((W + 15) / 16) * ((Cout / 8 + 7) / 8);     // one warp each
  if (units == 0) return SEGMIF_OK;
  SEGMIF_REQUIRE(units < (1ll << 31), "conv3x3_in1: too many pixel blocks for the 32-bit unit index");
  conv3x3_in1_kernel<<<(unsigned)ceil_div(units, 8), 256, 0, as_stream(stream)>>>(
      plane, bstride, w, bias, prelu_alpha, (bf16*)dst, ld_dst, dst_coff, B, H, W, Cout);
  return check_launch("segmif_conv3x3_in1_fwd");
}

extern "C" int segmif_conv3x3_out1_fwd(const void* src, int ld_src, const float* w, const float* bias,
                                       const float* prelu_alpha, float* dst, int B, int H, int W, int Cin,
                                       segmif_stream_t stream) {
  SEGMIF_REQUIRE(src && w && bias && prelu_alpha && dst, "conv3x3_out1: null pointer");
  SEGMIF_REQUIRE(Cin % 32 == 0 && ld_src % 8 == 0, "conv3x3_out1: Cin must be a multiple of 32");
  SEGMIF_REQUIRE((int64_t)B * H * W < (1ll << 31), "conv3x3_out1: B*H*W must be below 2^31");
  const int64_t total = (int64_t)B * H * W * 4;
  if (total == 0) return SEGMIF_OK;
  if (Cin == 32 && ((uintptr_t)src & 15) == 0) {
    const int64_t threads = (int64_t)B * H * ((W + kOut1Run - 1) / kOut1Run) * 4;
    conv3x3_out1_c32_kernel<<<(unsigned)ceil_div(threads, 256), 256, 0, as_stream(stream)>>>(
        (const bf16*)src, ld_src, w, bias, prelu_alpha, dst, B, H, W);
    return check_launch("segmif_conv3x3_out1_fwd");
  }
  conv3x3_out1_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, as_stream(stream)>>>(
      (const bf16*)src, ld_src, w, bias, prelu_alpha, dst, B, H, W, Cin);
  return check_launch("segmif_conv3x3_out1_fwd");
}

extern "C" int segmif_rgb2ycrcb(const float* rgb, float* ycc, int B, int64_t HW, segmif_stream_t stream) {
  SEGMIF_REQUIRE(rgb && ycc, "rgb2ycrcb: null pointer");
  if (B * HW == 0) return SEGMIF_OK;
  rgb2ycrcb_kernel<<<(unsigned)ceil_div(B * HW, 256), 256, 0, as_stream(stream)>>>(rgb, ycc, B, HW);
  return check_launch("segmif_rgb2ycrcb");
}

extern "C" int segmif_ycrcb2rgb(const float* ycc, float* rgb, int B, int64_t HW, segmif_stream_t stream) {
  SEGMIF_REQUIRE(rgb && ycc, "ycrcb2rgb: null pointer");
  if (B * HW == 0) return SEGMIF_OK;
  ycrcb2rgb_kernel<<<(unsigned)ceil_div(B * HW, 256), 256, 0, as_stream(stream)>>>(ycc, rgb, B, HW);
  return check_launch("segmif_ycrcb2rgb");
}

extern "C" int segmif_recompose_rgb(const float* fused_y, const float* vis_rgb, float* rgb_out, int clamp01, int B,
                                    int64_t HW, segmif_stream_t stream) {
  SEGMIF_REQUIRE(fused_y && vis_rgb && rgb_out, "recompose_rgb: null pointer");
  if (B * HW == 0) return SEGMIF_OK;
  recompose_rgb_kernel<<<(unsigned)ceil_div(B * HW, 256), 256, 0, as_stream(stream)>>>(fused_y, vis_rgb, rgb_out,
                                                                                       clamp01, B, HW);
  return check_launch("segmif_recompose_rgb");
}

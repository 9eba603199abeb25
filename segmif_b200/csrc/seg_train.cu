// Training-side kernels of the segmentation network (train.py:115-245 train_seg, and the CE term of train_fusion):
//   upsample_ce_bwd      d logits of CrossEntropyLoss(ignore_index) o bilinear upsample (core/model_fusion.py:1095-1096)
//   bilinear_nhwc_bwd    adjoint of F.interpolate(bilinear, align_corners=False) (core/segformer_head.py:67-73)
//   bn_* / channel_scale train-mode BatchNorm2d + ReLU of linear_fuse and Dropout2d (core/segformer_head.py:50-57,77-79)
//   dwconv3x3 / dwconv3x3_gelu_bwd   Mix-FFN depthwise conv + GELU backward (core/mix_transformer.py:46-53,381-387)
//   col2im               adjoint of the im2col used for the strided patch-embedding / spatial-reduction convolutions
//   channel_affine, recompose_rgb_bwd, cast, scale_add_rows   small elementwise pieces around them
// All gathers are written in "pull" form (each output element sums its contributions), so they are deterministic and
// need no atomics; per-channel reductions use shared-memory partials + one global atomic per block and channel.
#include <algorithm>
#include <math.h>

#include "common.cuh"

namespace segmif {

__device__ __forceinline__ void bl_src2(int dst, float scale, int in_size, int& i0, int& i1, float& l0, float& l1) {
  float s = scale * ((float)dst + 0.5f) - 0.5f;
  s = s < 0.f ? 0.f : s;
  i0 = (int)s;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  l1 = s - (float)i0;
  l0 = 1.f - l1;
}

// output rows/cols whose bilinear taps can touch source index i: [lo, hi]
__device__ __forceinline__ void footprint(int i, float scale, int out_size, int& lo, int& hi) {
  const float inv = 1.f / scale;
  lo = max(0, (int)floorf(((float)i - 1.f + 0.5f) * inv - 0.5f) - 1);
  hi = min(out_size - 1, (int)ceilf(((float)i + 1.f + 0.5f) * inv - 0.5f) + 1);
}

// ------------------------------------------------------------------------------------------------ CE backward
// one thread per low-resolution logit pixel; gout = upstream gradient of the mean loss, cnt = number of valid labels
__global__ void __launch_bounds__(128) upsample_ce_bwd_kernel(const float* __restrict__ logits, int B, int h, int w, int nc,
                                                              const int64_t* __restrict__ labels, int H, int W,
                                                              int ignore_index, float sy, float sx,
                                                              const float* __restrict__ gout, const float* __restrict__ cnt,
                                                              float* __restrict__ dlogits) {
  const int64_t idx = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (idx >= (int64_t)B * h * w) return;
  const int x = (int)(idx % w), y = (int)((idx / w) % h);
  const int64_t b = idx / ((int64_t)w * h);
  float acc[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) acc[c] = 0.f;
  int Y0, Y1, X0, X1;
  footprint(y, sy, H, Y0, Y1);
  footprint(x, sx, W, X0, X1);
  const float* base = logits + b * h * w * nc;
  for (int Y = Y0; Y <= Y1; ++Y) {
    int y0, y1; float hy0, hy1;
    bl_src2(Y, sy, h, y0, y1, hy0, hy1);
    const float wy = (y0 == y ? hy0 : 0.f) + (y1 == y ? hy1 : 0.f);
    if (wy == 0.f) continue;
    for (int X = X0; X <= X1; ++X) {
      int x0, x1; float wx0, wx1;
      bl_src2(X, sx, w, x0, x1, wx0, wx1);
      const float wx = (x0 == x ? wx0 : 0.f) + (x1 == x ? wx1 : 0.f);
      if (wx == 0.f) continue;
      const int64_t lab = labels[(b * H + Y) * W + X];
      if (lab == ignore_index || lab < 0 || lab >= nc) continue;
      const float* p00 = base + ((int64_t)y0 * w + x0) * nc;
      const float* p01 = base + ((int64_t)y0 * w + x1) * nc;
      const float* p10 = base + ((int64_t)y1 * w + x0) * nc;
      const float* p11 = base + ((int64_t)y1 * w + x1) * nc;
      float vals[32], m = -INFINITY;
      for (int c = 0; c < nc; ++c) {
        vals[c] = hy0 * (wx0 * p00[c] + wx1 * p01[c]) + hy1 * (wx0 * p10[c] + wx1 * p11[c]);
        m = fmaxf(m, vals[c]);
      }
      float se = 0.f;
      for (int c = 0; c < nc; ++c) { vals[c] = expf(vals[c] - m); se += vals[c]; }
      const float wt = wy * wx, inv = 1.f / se;
      for (int c = 0; c < nc; ++c) acc[c] += wt * (vals[c] * inv - (c == lab ? 1.f : 0.f));
    }
  }
  const float gs = gout[0] / fmaxf(cnt[0], 1.f);
  for (int c = 0; c < nc; ++c) dlogits[idx * nc + c] = gs * acc[c];
}

// ------------------------------------------------------------------------------------------------ bilinear backward
// dsrc[b,y,x,c] = sum over destination pixels of weight * ddst; one thread = 8 channels of one source pixel
__global__ void __launch_bounds__(256) bilinear_nhwc_bwd_kernel(const bf16* __restrict__ ddst, int ld_dst, int H, int W,
                                                                bf16* __restrict__ dsrc, int B, int h, int w, int C,
                                                                float sy, float sx) {
  const int g8 = C >> 3;
  const int64_t n = (int64_t)B * h * w * g8;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const int grp = (int)(i % g8);
    const int64_t pix = i / g8;
    const int x = (int)(pix % w), y = (int)((pix / w) % h);
    const int64_t b = pix / ((int64_t)w * h);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int Y0, Y1, X0, X1;
    footprint(y, sy, H, Y0, Y1);
    footprint(x, sx, W, X0, X1);
    for (int Y = Y0; Y <= Y1; ++Y) {
      int y0, y1; float hy0, hy1;
      bl_src2(Y, sy, h, y0, y1, hy0, hy1);
      const float wy = (y0 == y ? hy0 : 0.f) + (y1 == y ? hy1 : 0.f);
      if (wy == 0.f) continue;
      for (int X = X0; X <= X1; ++X) {
        int x0, x1; float wx0, wx1;
        bl_src2(X, sx, w, x0, x1, wx0, wx1);
        const float wx = (x0 == x ? wx0 : 0.f) + (x1 == x ? wx1 : 0.f);
        if (wx == 0.f) continue;
        float v[8];
        load8(ddst + ((b * H + Y) * W + X) * ld_dst + grp * 8, v);
        const float wt = wy * wx;
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(wt, v[e], acc[e]);
      }
    }
    store8(dsrc + pix * C + grp * 8, acc);
  }
}

// ------------------------------------------------------------------------------------------------ BatchNorm (train)
// column sums and sums of squares of z [rows, C] (C % 8 == 0, C <= 512) into double accumulators [2][C]
__global__ void __launch_bounds__(256) col_moments_kernel(const bf16* __restrict__ z, int64_t rows, int C,
                                                          double* __restrict__ acc) {
  __shared__ float s1[512], s2[512];
  const int g = C >> 3, rpi = 256 / g;
  const bool active = threadIdx.x < rpi * g;
  const int grp = threadIdx.x % g, rin = threadIdx.x / g;
  float a1[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, a2[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (active)
    for (int64_t r = (int64_t)blockIdx.x * rpi + rin; r < rows; r += (int64_t)gridDim.x * rpi) {
      float v[8];
      load8(z + r * C + grp * 8, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) { a1[i] += v[i]; a2[i] = fmaf(v[i], v[i], a2[i]); }
    }
  for (int i = threadIdx.x; i < C; i += 256) { s1[i] = 0.f; s2[i] = 0.f; }
  __syncthreads();
  if (active) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { atomicAdd(&s1[grp * 8 + i], a1[i]); atomicAdd(&s2[grp * 8 + i], a2[i]); }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += 256) { atomicAdd(acc + i, (double)s1[i]); atomicAdd(acc + C + i, (double)s2[i]); }
}

// stats[0][c] = mean, stats[1][c] = rstd; running stats updated as nn.BatchNorm2d does (unbiased variance, momentum)
__global__ void bn_finalize_kernel(const double* __restrict__ acc, int64_t rows, int C, float eps, float momentum,
                                   float* __restrict__ stats, float* __restrict__ running_mean,
                                   float* __restrict__ running_var) {
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
    const double mean = acc[c] / (double)rows;
    double var = acc[C + c] / (double)rows - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    stats[c] = (float)mean;
    stats[C + c] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
    if (running_var) running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)(rows > 1 ? var * (double)rows / (double)(rows - 1) : var);
  }
}

// y = relu(gamma * (z - mean) * rstd + beta)
__global__ void __launch_bounds__(256) bn_relu_apply_kernel(const bf16* __restrict__ z, const float* __restrict__ stats,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            bf16* __restrict__ y, int64_t rows, int C) {
  const int g = C >> 3;
  const int64_t n = rows * g;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const int grp = (int)(i % g);
    float v[8];
    load8(z + i * 8, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = grp * 8 + e;
      v[e] = fmaxf(gamma[c] * (v[e] - stats[c]) * stats[C + c] + beta[c], 0.f);
    }
    store8(y + i * 8, v);
  }
}

// sums[0][c] = sum dyr, sums[1][c] = sum dyr * zhat with dyr = dy * 1[y > 0]
__global__ void __launch_bounds__(256) bn_relu_bwd_reduce_kernel(const bf16* __restrict__ z, const bf16* __restrict__ y,
                                                                 const bf16* __restrict__ dy, const float* __restrict__ stats,
                                                                 int64_t rows, int C, double* __restrict__ sums) {
  __shared__ float s1[512], s2[512];
  const int g = C >> 3, rpi = 256 / g;
  const bool active = threadIdx.x < rpi * g;
  const int grp = threadIdx.x % g, rin = threadIdx.x / g;
  float a1[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, a2[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (active)
    for (int64_t r = (int64_t)blockIdx.x * rpi + rin; r < rows; r += (int64_t)gridDim.x * rpi) {
      float vz[8], vy[8], vd[8];
      load8(z + r * C + grp * 8, vz);
      load8(y + r * C + grp * 8, vy);
      load8(dy + r * C + grp * 8, vd);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = grp * 8 + i;
        const float d = vy[i] > 0.f ? vd[i] : 0.f;
        a1[i] += d;
        a2[i] = fmaf(d, (vz[i] - stats[c]) * stats[C + c], a2[i]);
      }
    }
  for (int i = threadIdx.x; i < C; i += 256) { s1[i] = 0.f; s2[i] = 0.f; }
  __syncthreads();
  if (active) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { atomicAdd(&s1[grp * 8 + i], a1[i]); atomicAdd(&s2[grp * 8 + i], a2[i]); }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += 256) { atomicAdd(sums + i, (double)s1[i]); atomicAdd(sums + C + i, (double)s2[i]); }
}

// dz = gamma * rstd * (dyr - mean(dyr) - zhat * mean(dyr * zhat));  dgamma += sum dyr zhat, dbeta += sum dyr (block 0)
__global__ void __launch_bounds__(256) bn_relu_bwd_apply_kernel(const bf16* __restrict__ z, const bf16* __restrict__ y,
                                                                const bf16* __restrict__ dy, const float* __restrict__ stats,
                                                                const float* __restrict__ gamma, const double* __restrict__ sums,
                                                                bf16* __restrict__ dz, int64_t rows, int C,
                                                                float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                float corr) {
  const int g = C >> 3;
  const int64_t n = rows * g;
  const float inv_rows = corr / (float)rows;      // corr = 0: eval-mode BatchNorm (running statistics are constants)
  if (blockIdx.x == 0)
    for (int c = threadIdx.x; c < C; c += 256) {
      if (dgamma) dgamma[c] += (float)sums[C + c];
      if (dbeta) dbeta[c] += (float)sums[c];
    }
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const int grp = (int)(i % g);
    float vz[8], vy[8], vd[8];
    load8(z + i * 8, vz);
    load8(y + i * 8, vy);
    load8(dy + i * 8, vd);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = grp * 8 + e;
      const float zhat = (vz[e] - stats[c]) * stats[C + c];
      const float d = vy[e] > 0.f ? vd[e] : 0.f;
      vd[e] = gamma[c] * stats[C + c] * (d - (float)sums[c] * inv_rows - zhat * (float)sums[C + c] * inv_rows);
    }
    store8(dz + i * 8, vd);
  }
}

// y[b, p, c] = x[b, p, c] * scale[b, c]   (Dropout2d forward and backward; x, y bf16 [B, HW, C])
__global__ void __launch_bounds__(256) channel_scale_kernel(const bf16* __restrict__ x, const float* __restrict__ scale,
                                                            bf16* __restrict__ y, int64_t HW, int C, int64_t n8) {
  const int g = C >> 3;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n8; i += (int64_t)gridDim.x * 256) {
    const int grp = (int)(i % g);
    const int64_t b = (i / g) / HW;
    float v[8];
    load8(x + i * 8, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] *= scale[b * C + grp * 8 + e];
    store8(y + i * 8, v);
  }
}

// ------------------------------------------------------------------------------------------------ depthwise conv
// Strip-walking layout shared by the two kernels below.  A block owns one tile of `tcg` channel PAIRS (<= 64 pairs = 128
// channels = 256 B per pixel) and a strip of S = 256 / tcg adjacent image columns, and walks down the rows of its row
// range: thread (pair, column) loads its 3x3 neighbourhood as nine 4-byte loads (a warp reads whole 128-byte lines), so
// the three rows of the strip (+1 halo column each side) stay in L1 between consecutive rows and every input line comes
// from L2 about 1.5x instead of 9x.  Two channels per thread keep the backward's 9x2 + 2 weight-gradient accumulators in
// registers at three resident blocks per SM.
struct DwStrip {
  int c, col, row0, row1, b;
  bool live;
};
__device__ __forceinline__ DwStrip dw_strip(int tcg, int S, int strips_x, int rchunks, int H, int W) {
  DwStrip d;
  const int pair = threadIdx.x % tcg, slot = threadIdx.x / tcg;
  d.c = (blockIdx.x * tcg + pair) * 2;
  int u = blockIdx.y;
  const int rc = u % rchunks; u /= rchunks;
  const int sx = u % strips_x;
  d.b = u / strips_x;
  d.col = sx * S + slot;
  const int rows_per = (H + rchunks - 1) / rchunks;
  d.row0 = rc * rows_per;
  d.row1 = min(H, d.row0 + rows_per);
  d.live = slot < S && d.col < W;
  return d;
}
__device__ __forceinline__ float2 ld_bf2(const bf16* p) {
  const uint32_t u = *reinterpret_cast<const uint32_t*>(p);
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
__device__ __forceinline__ void st_bf2(bf16* p, float a, float b) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}

// one image row of the thread's 3-column window (zeros outside the image)
struct DwRow { float2 l, m, r; };
__device__ __forceinline__ DwRow dw_load_row(const bf16* __restrict__ x, int b, int yy, int H, int W, int C, int col, int c,
                                             bool hasl, bool hasr) {
  DwRow o;
  o.l = o.m = o.r = make_float2(0.f, 0.f);
  if ((unsigned)yy < (unsigned)H) {
    const bf16* row = x + ((size_t)(b * H + yy) * W + col) * C + c;
    o.m = ld_bf2(row);
    if (hasl) o.l = ld_bf2(row - C);
    if (hasr) o.r = ld_bf2(row + C);
  }
  return o;
}

// y = dwconv3x3(x) (+ bias), pixel-major bf16, w fp32 [9][C]; flip != 0 uses the transposed taps (data gradient).
// The three window rows live in registers and slide down; the row two below is requested before the current row's
// arithmetic, so its L2 latency overlaps it (the first version reloaded all nine taps per row and was latency-bound).
__global__ void __launch_bounds__(256) dwconv3x3_kernel(const bf16* __restrict__ x, const float* __restrict__ w9c,
                                                        const float* __restrict__ bias, bf16* __restrict__ y, int H, int W,
                                                        int C, int flip, int tcg, int S, int strips_x, int rchunks) {
  const DwStrip d = dw_strip(tcg, S, strips_x, rchunks, H, W);
  if (!d.live) return;
  float2 wr[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) wr[t] = *reinterpret_cast<const float2*>(w9c + (flip ? 8 - t : t) * C + d.c);
  const float2 bv = bias ? *reinterpret_cast<const float2*>(bias + d.c) : make_float2(0.f, 0.f);
  const bool hasl = d.col > 0, hasr = d.col + 1 < W;
  DwRow ra = dw_load_row(x, d.b, d.row0 - 1, H, W, C, d.col, d.c, hasl, hasr);
  DwRow rb = dw_load_row(x, d.b, d.row0, H, W, C, d.col, d.c, hasl, hasr);
  DwRow rc = dw_load_row(x, d.b, d.row0 + 1, H, W, C, d.col, d.c, hasl, hasr);
  for (int r = d.row0; r < d.row1; ++r) {
    const DwRow rn = dw_load_row(x, d.b, r + 2, H, W, C, d.col, d.c, hasl, hasr);
    float2 acc = bv;
#define SEGMIF_DW_ROW(R, K)                                                                             \
    acc.x = fmaf(wr[K].x, R.l.x, acc.x); acc.y = fmaf(wr[K].y, R.l.y, acc.y);                          \
    acc.x = fmaf(wr[K + 1].x, R.m.x, acc.x); acc.y = fmaf(wr[K + 1].y, R.m.y, acc.y);                  \
    acc.x = fmaf(wr[K + 2].x, R.r.x, acc.x); acc.y = fmaf(wr[K + 2].y, R.r.y, acc.y);
    SEGMIF_DW_ROW(ra, 0) SEGMIF_DW_ROW(rb, 3) SEGMIF_DW_ROW(rc, 6)
#undef SEGMIF_DW_ROW
    st_bf2(y + ((size_t)(d.b * H + r) * W + d.col) * C + d.c, acc.x, acc.y);
    ra = rb; rb = rc; rc = rn;
  }
}

// z = dwconv(x) + b;  dz = dy * gelu'(z);  part[blockIdx.y][t][c] = sum dz * x(p + t) (t < 9), [9][c] = sum dz
__global__ void __launch_bounds__(256, 3) dwconv3x3_gelu_bwd_kernel(const bf16* __restrict__ x, const float* __restrict__ w9c,
                                                                 const float* __restrict__ bias, const bf16* __restrict__ dy,
                                                                 bf16* __restrict__ dz, int H, int W, int C,
                                                                 float* __restrict__ part, int tcg, int S, int strips_x,
                                                                 int rchunks) {
  __shared__ float sacc[10 * 128];
  const DwStrip d = dw_strip(tcg, S, strips_x, rchunks, H, W);
  float2 aw[9], ab = make_float2(0.f, 0.f);
#pragma unroll
  for (int t = 0; t < 9; ++t) aw[t] = make_float2(0.f, 0.f);
  if (d.live) {
    float2 wr[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) wr[t] = *reinterpret_cast<const float2*>(w9c + t * C + d.c);
    const float2 bv = *reinterpret_cast<const float2*>(bias + d.c);
    const bool hasl = d.col > 0, hasr = d.col + 1 < W;
    DwRow ra = dw_load_row(x, d.b, d.row0 - 1, H, W, C, d.col, d.c, hasl, hasr);
    DwRow rb = dw_load_row(x, d.b, d.row0, H, W, C, d.col, d.c, hasl, hasr);
    DwRow rc = dw_load_row(x, d.b, d.row0 + 1, H, W, C, d.col, d.c, hasl, hasr);
    size_t o = ((size_t)(d.b * H + d.row0) * W + d.col) * C + d.c;
    const size_t ostep = (size_t)W * C;
    float2 g = d.row0 < d.row1 ? ld_bf2(dy + o) : make_float2(0.f, 0.f);
    for (int r = d.row0; r < d.row1; ++r, o += ostep) {
      const DwRow rn = dw_load_row(x, d.b, r + 2, H, W, C, d.col, d.c, hasl, hasr);
      const float2 gn = r + 1 < d.row1 ? ld_bf2(dy + o + ostep) : make_float2(0.f, 0.f);
      float2 z = bv;
#define SEGMIF_DW_ROW(R, K)                                                                             \
      z.x = fmaf(wr[K].x, R.l.x, z.x); z.y = fmaf(wr[K].y, R.l.y, z.y);                                \
      z.x = fmaf(wr[K + 1].x, R.m.x, z.x); z.y = fmaf(wr[K + 1].y, R.m.y, z.y);                        \
      z.x = fmaf(wr[K + 2].x, R.r.x, z.x); z.y = fmaf(wr[K + 2].y, R.r.y, z.y);
      SEGMIF_DW_ROW(ra, 0) SEGMIF_DW_ROW(rb, 3) SEGMIF_DW_ROW(rc, 6)
#undef SEGMIF_DW_ROW
      g.x *= 0.5f * (1.0f + erff(z.x * 0.70710678118654752440f)) + z.x * 0.39894228040143267794f * __expf(-0.5f * z.x * z.x);
      g.y *= 0.5f * (1.0f + erff(z.y * 0.70710678118654752440f)) + z.y * 0.39894228040143267794f * __expf(-0.5f * z.y * z.y);
      ab.x += g.x; ab.y += g.y;
#define SEGMIF_DW_ROW(R, K)                                                                             \
      aw[K].x = fmaf(g.x, R.l.x, aw[K].x); aw[K].y = fmaf(g.y, R.l.y, aw[K].y);                        \
      aw[K + 1].x = fmaf(g.x, R.m.x, aw[K + 1].x); aw[K + 1].y = fmaf(g.y, R.m.y, aw[K + 1].y);        \
      aw[K + 2].x = fmaf(g.x, R.r.x, aw[K + 2].x); aw[K + 2].y = fmaf(g.y, R.r.y, aw[K + 2].y);
      SEGMIF_DW_ROW(ra, 0) SEGMIF_DW_ROW(rb, 3) SEGMIF_DW_ROW(rc, 6)
#undef SEGMIF_DW_ROW
      st_bf2(dz + o, g.x, g.y);
      ra = rb; rb = rc; rc = rn; g = gn;
    }
  }
  const int tile_c = tcg * 2;
  for (int i = threadIdx.x; i < 10 * tile_c; i += 256) sacc[i] = 0.f;
  __syncthreads();
  if (d.live) {
    const int lc = (threadIdx.x % tcg) * 2;
#pragma unroll
    for (int t = 0; t < 9; ++t) { atomicAdd(&sacc[t * tile_c + lc], aw[t].x); atomicAdd(&sacc[t * tile_c + lc + 1], aw[t].y); }
    atomicAdd(&sacc[9 * tile_c + lc], ab.x);
    atomicAdd(&sacc[9 * tile_c + lc + 1], ab.y);
  }
  __syncthreads();
  float* dst = part + (size_t)blockIdx.y * 10 * C + blockIdx.x * tile_c;
  for (int i = threadIdx.x; i < 10 * tile_c; i += 256) dst[(i / tile_c) * C + (i % tile_c)] = sacc[i];
}

// dw9c[i] += sum_u part[u][i] (i < 9C), dbias[c] += sum_u part[u][9C + c]: fixed summation order, no atomics
__global__ void __launch_bounds__(256) dwconv_part_reduce_kernel(const float* __restrict__ part, int nparts, int C,
                                                                 float* __restrict__ dw9c, float* __restrict__ dbias) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= 10 * C) return;
  float s = 0.f;
  for (int u = 0; u < nparts; ++u) s += part[(size_t)u * 10 * C + i];
  if (i < 9 * C) dw9c[i] += s; else dbias[i - 9 * C] += s;
}

struct DwGeom { int tcg, S, strips_x, rchunks, ctiles, units; };
static DwGeom dw_geom(int B, int H, int W, int C) {
  DwGeom g;
  const int pairs = C / 2;
  g.tcg = 64;
  while (pairs % g.tcg) g.tcg >>= 1;                      // C % 8 == 0 -> tcg >= 4
  g.S = 256 / g.tcg;
  g.strips_x = (W + g.S - 1) / g.S;
  g.ctiles = pairs / g.tcg;
  const int base = B * g.strips_x * g.ctiles;
  g.rchunks = std::max(1, std::min(H / 8, (6 * 148 + base - 1) / base));      // ~six blocks per SM, >= 8 rows per walk
  g.units = B * g.strips_x * g.rchunks;
  return g;
}

// ------------------------------------------------------------------------------------------------ col2im
// dcol [B*Ho*Wo, ldc] with column (c*k + ky)*k + kx  ->  dx[b, y, x, c] (pixel-major, fp32 or bf16)
template <typename TO>
__global__ void __launch_bounds__(256) col2im_kernel(const bf16* __restrict__ dcol, int ldc, TO* __restrict__ dx, int B,
                                                     int H, int W, int C, int k, int s, int p, int Ho, int Wo) {
  const int64_t n = (int64_t)B * H * W * C;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const int c = (int)(i % C);
    const int64_t pix = i / C;
    const int x = (int)(pix % W), y = (int)((pix / W) % H);
    const int64_t b = pix / ((int64_t)W * H);
    float acc = 0.f;
    for (int ky = 0; ky < k; ++ky) {
      const int ty = y + p - ky;
      if (ty < 0 || ty % s != 0) continue;
      const int oy = ty / s;
      if (oy >= Ho) continue;
      for (int kx = 0; kx < k; ++kx) {
        const int tx = x + p - kx;
        if (tx < 0 || tx % s != 0) continue;
        const int ox = tx / s;
        if (ox >= Wo) continue;
        acc += __bfloat162float(dcol[((b * Ho + oy) * Wo + ox) * ldc + (c * k + ky) * k + kx]);
      }
    }
    st_from_float(dx + i, acc);
  }
}

// ------------------------------------------------------------------------------------------------ small pieces
// y[b,c,p] = x[b,c,p] * scale[c] + shift[c]   (Network3's input normalisation as its own op in training)
__global__ void __launch_bounds__(256) channel_affine_nchw_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                                                  const float* __restrict__ shift, float* __restrict__ y,
                                                                  int C, int64_t HW, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const int c = (int)((i / HW) % C);
    y[i] = x[i] * scale[c] + (shift ? shift[c] : 0.f);
  }
}

// rgb = clamp(ycrcb2rgb([fused, cr, cb])) -> d fused = sum_c 1[0 < rgb_c < 1] * drgb_c   (Y feeds R, G, B with weight 1)
__global__ void __launch_bounds__(256) recompose_rgb_bwd_kernel(const float* __restrict__ rgb, const float* __restrict__ drgb,
                                                                float* __restrict__ dfused, int clamp01, int64_t HW,
                                                                int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const int64_t b = i / HW, p = i - b * HW;
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = rgb[(b * 3 + c) * HW + p];
      if (!clamp01 || (v > 0.f && v < 1.f)) acc += drgb[(b * 3 + c) * HW + p];
    }
    dfused[i] = acc;
  }
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) cast_kernel(const TI* __restrict__ x, TO* __restrict__ y, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256)
    st_from_float(y + i, ld_as_float(x + i));
}

// out[r, c] = x[r, c] + scale[r / rows_per_sample] * y[r, c]   (DropPath residual; x, out fp32, y bf16 or fp32)
template <typename TY>
__global__ void __launch_bounds__(256) scale_add_rows_kernel(const float* __restrict__ x, const TY* __restrict__ y,
                                                             const float* __restrict__ scale, float* __restrict__ out,
                                                             int64_t rows_per_sample, int C, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const int64_t r = i / C;
    out[i] = x[i] + (scale ? scale[r / rows_per_sample] : 1.f) * ld_as_float(y + i);
  }
}

// y[r, c] = bf16(scale[r / rows_per_sample] * x[r, c]): the branch gradient of a DropPath residual, cast for the GEMMs
__global__ void __launch_bounds__(256) scale_cast_rows_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                                              bf16* __restrict__ y, int64_t rows_per_sample, int C, int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    const float s = scale[(i * 4 / C) / rows_per_sample];
    uint2 o;
    o.x = pack_bf16x2(v.x * s, v.y * s);
    o.y = pack_bf16x2(v.z * s, v.w * s);
    reinterpret_cast<uint2*>(y)[i] = o;
  }
}

}  // namespace segmif

using namespace segmif;

static int grid_n(int64_t items, int per_block) {
  return (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(items, per_block), 148 * 16));
}

extern "C" int segmif_upsample_ce_bwd(const float* logits, int B, int h, int w, int nc, const int64_t* labels, int H, int W,
                                      int ignore_index, const float* gout, const float* count, float* dlogits,
                                      segmif_stream_t stream) {
  SEGMIF_REQUIRE(logits && labels && gout && count && dlogits, "upsample_ce_bwd: null pointer");
  SEGMIF_REQUIRE(nc > 0 && nc <= 32, "upsample_ce_bwd: nc=%d unsupported (1..32)", nc);
  const int64_t n = (int64_t)B * h * w;
  SEGMIF_REQUIRE(n > 0 && H > 0 && W > 0, "upsample_ce_bwd: empty input");
  upsample_ce_bwd_kernel<<<(unsigned)ceil_div(n, 128), 128, 0, as_stream(stream)>>>(logits, B, h, w, nc, labels, H, W, ignore_index,
                                                                                      (float)h / (float)H, (float)w / (float)W, gout, count, dlogits);
  return check_launch("segmif_upsample_ce_bwd");
}

extern "C" int segmif_bilinear_nhwc_bwd(const void* ddst, int ld_dst, int dst_coff, int H, int W, void* dsrc, int B, int h,
                                        int w, int C, segmif_stream_t stream) {
  SEGMIF_REQUIRE(ddst && dsrc, "bilinear_nhwc_bwd: null pointer");
  SEGMIF_REQUIRE(C % 8 == 0 && ld_dst % 8 == 0 && dst_coff % 8 == 0, "bilinear_nhwc_bwd: C, pitch and offset must be multiples of 8");
  SEGMIF_REQUIRE(B > 0 && h > 0 && w > 0 && H > 0 && W > 0, "bilinear_nhwc_bwd: empty input");
  const int64_t n = (int64_t)B * h * w * (C >> 3);
  bilinear_nhwc_bwd_kernel<<<grid_n(n, 256), 256, 0, as_stream(stream)>>>((const bf16*)ddst + dst_coff, ld_dst, H, W, (bf16*)dsrc, B, h, w, C,
                                                                           (float)h / (float)H, (float)w / (float)W);
  return check_launch("segmif_bilinear_nhwc_bwd");
}

/* workspace: 2*C doubles (zeroed by the caller) */
extern "C" int segmif_bn_train_fwd(const void* z, int64_t rows, int C, const float* gamma, const float* beta, float eps,
                                   float momentum, float* running_mean, float* running_var, double* workspace, float* stats,
                                   void* y, segmif_stream_t stream) {
  SEGMIF_REQUIRE(z && gamma && beta && workspace && stats && y && rows > 0, "bn_train_fwd: bad arguments");
  SEGMIF_REQUIRE(C % 8 == 0 && C > 0 && C <= 512, "bn_train_fwd: C=%d must be a multiple of 8, <= 512", C);
  cudaStream_t st = as_stream(stream);
  const int rpi = 256 / (C >> 3);
  col_moments_kernel<<<grid_n(rows, rpi * 8), 256, 0, st>>>((const bf16*)z, rows, C, workspace);
  bn_finalize_kernel<<<1, 256, 0, st>>>(workspace, rows, C, eps, momentum, stats, running_mean, running_var);
  bn_relu_apply_kernel<<<grid_n(rows * (C >> 3), 1024), 256, 0, st>>>((const bf16*)z, stats, gamma, beta, (bf16*)y, rows, C);
  return check_launch("segmif_bn_train_fwd");
}

extern "C" int segmif_bn_train_bwd(const void* z, const void* y, const void* dy, const float* stats, const float* gamma,
                                   int64_t rows, int C, double* workspace, void* dz, float* dgamma, float* dbeta,
                                   segmif_stream_t stream) {
  SEGMIF_REQUIRE(z && y && dy && stats && gamma && workspace && dz && rows > 0, "bn_train_bwd: bad arguments");
  SEGMIF_REQUIRE(C % 8 == 0 && C > 0 && C <= 512, "bn_train_bwd: C=%d must be a multiple of 8, <= 512", C);
  cudaStream_t st = as_stream(stream);
  const int rpi = 256 / (C >> 3);
  bn_relu_bwd_reduce_kernel<<<grid_n(rows, rpi * 8), 256, 0, st>>>((const bf16*)z, (const bf16*)y, (const bf16*)dy, stats, rows, C, workspace);
  bn_relu_bwd_apply_kernel<<<grid_n(rows * (C >> 3), 1024), 256, 0, st>>>((const bf16*)z, (const bf16*)y, (const bf16*)dy, stats, gamma, workspace,
                                                                            (bf16*)dz, rows, C, dgamma, dbeta, 1.f);
  return check_launch("segmif_bn_train_bwd");
}

// eval-mode BatchNorm2d + ReLU with gradients (the reference keeps training after val_segformer() left the model in
// eval(), train.py:232-236: running statistics, no batch coupling): stats = {running_mean, rsqrt(running_var + eps)}
__global__ void bn_eval_stats_kernel(const float* __restrict__ running_mean, const float* __restrict__ running_var, float eps,
                                     int C, float* __restrict__ stats) {
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
    stats[c] = running_mean[c];
    stats[C + c] = (float)(1.0 / sqrt((double)running_var[c] + (double)eps));
  }
}

extern "C" int segmif_bn_eval_fwd(const void* z, int64_t rows, int C, const float* gamma, const float* beta, float eps,
                                  const float* running_mean, const float* running_var, float* stats, void* y,
                                  segmif_stream_t stream) {
  SEGMIF_REQUIRE(z && gamma && beta && running_mean && running_var && stats && y && rows > 0, "bn_eval_fwd: bad arguments");
  SEGMIF_REQUIRE(C % 8 == 0 && C > 0 && C <= 512, "bn_eval_fwd: C=%d must be a multiple of 8, <= 512", C);
  cudaStream_t st = as_stream(stream);
  bn_eval_stats_kernel<<<1, 256, 0, st>>>(running_mean, running_var, eps, C, stats);
  bn_relu_apply_kernel<<<grid_n(rows * (C >> 3), 1024), 256, 0, st>>>((const bf16*)z, stats, gamma, beta, (bf16*)y, rows, C);
  return check_launch("segmif_bn_eval_fwd");
}

extern "C" int segmif_bn_eval_bwd(const void* z, const void* y, const void* dy, const float* stats, const float* gamma,
                                  int64_t rows, int C, double* workspace, void* dz, float* dgamma, float* dbeta,
                                  segmif_stream_t stream) {
  SEGMIF_REQUIRE(z && y && dy && stats && gamma && workspace && dz && rows > 0, "bn_eval_bwd: bad arguments");
  SEGMIF_REQUIRE(C % 8 == 0 && C > 0 && C <= 512, "bn_eval_bwd: C=%d must be a multiple of 8, <= 512", C);
  cudaStream_t st = as_stream(stream);
  const int rpi = 256 / (C >> 3);
  bn_relu_bwd_reduce_kernel<<<grid_n(rows, rpi * 8), 256, 0, st>>>((const bf16*)z, (const bf16*)y, (const bf16*)dy, stats, rows, C, workspace);
  bn_relu_bwd_apply_kernel<<<grid_n(rows * (C >> 3), 1024), 256, 0, st>>>((const bf16*)z, (const bf16*)y, (const bf16*)dy, stats, gamma, workspace,
                                                                            (bf16*)dz, rows, C, dgamma, dbeta, 0.f);
  return check_launch("segmif_bn_eval_bwd");
}

extern "C" int segmif_channel_scale(const void* x, const float* scale, void* y, int B, int64_t HW, int C,
                                    segmif_stream_t stream) {
  SEGMIF_REQUIRE(x && scale && y && B > 0 && HW > 0 && C % 8 == 0, "channel_scale: bad arguments");
  const int64_t n8 = (int64_t)B * HW * (C >> 3);
  channel_scale_kernel<<<grid_n(n8, 1024), 256, 0, as_stream(stream)>>>((const bf16*)x, scale, (bf16*)y, HW, C, n8);
  return check_launch("segmif_channel_scale");
}

extern "C" int segmif_dwconv3x3(const void* x, const float* w9c, const float* bias, void* y, int B, int H, int W, int C,
                                int flip, segmif_stream_t stream) {
  SEGMIF_REQUIRE(x && w9c && y && C % 8 == 0 && B > 0 && H > 0 && W > 0, "dwconv3x3: bad arguments");
  if (dwconv_tma_ok(B, H, W, C)) return dwconv_tma_fwd(x, w9c, bias, y, B, H, W, C, flip, 0, as_stream(stream));
  const DwGeom g = dw_geom(B, H, W, C);
  SEGMIF_REQUIRE(g.units <= 65535, "dwconv3x3: %d strips exceed the grid's y extent", g.units);
  dwconv3x3_kernel<<<dim3(g.ctiles, g.units), 256, 0, as_stream(stream)>>>((const bf16*)x, w9c, bias, (bf16*)y, H, W, C, flip, g.tcg,
                                                                           g.S, g.strips_x, g.rchunks);
  return check_launch("segmif_dwconv3x3");
}

extern "C" int64_t segmif_dwconv3x3_gelu_bwd_workspace(int B, int H, int W, int C) {
  if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || C % 8) return 0;
  if (dwconv_tma_ok(B, H, W, C)) return dwconv_tma_bwd_workspace(B, H, W, C);
  return (int64_t)dw_geom(B, H, W, C).units * 10 * C;
}

extern "C" int segmif_dwconv3x3_gelu_bwd(const void* x, const float* w9c, const float* bias, const void* dy, void* dz, int B,
                                         int H, int W, int C, float* dw9c, float* dbias, float* workspace,
                                         segmif_stream_t stream) {
  SEGMIF_REQUIRE(x && w9c && bias && dy && dz && dw9c && dbias && workspace, "dwconv3x3_gelu_bwd: null pointer");
  SEGMIF_REQUIRE(C % 8 == 0 && C > 0 && B > 0 && H > 0 && W > 0, "dwconv3x3_gelu_bwd: C=%d must be a multiple of 8", C);
  if (dwconv_tma_ok(B, H, W, C)) return dwconv_tma_gelu_bwd(x, w9c, bias, dy, dz, B, H, W, C, dw9c, dbias, workspace, as_stream(stream));
  const DwGeom g = dw_geom(B, H, W, C);
  SEGMIF_REQUIRE(g.units <= 65535, "dwconv3x3_gelu_bwd: %d strips exceed the grid's y extent", g.units);
  cudaStream_t st = as_stream(stream);
  dwconv3x3_gelu_bwd_kernel<<<dim3(g.ctiles, g.units), 256, 0, st>>>((const bf16*)x, w9c, bias, (const bf16*)dy, (bf16*)dz, H, W, C,
                                                                     workspace, g.tcg, g.S, g.strips_x, g.rchunks);
  dwconv_part_reduce_kernel<<<(10 * C + 255) / 256, 256, 0, st>>>(workspace, g.units, C, dw9c, dbias);
  return check_launch("segmif_dwconv3x3_gelu_bwd");
}

extern "C" int segmif_col2im(const void* dcol, int ldc, void* dx, int dx_dtype, int B, int H, int W, int C, int k, int stride,
                             int pad, segmif_stream_t stream) {
  SEGMIF_REQUIRE(dcol && dx && B > 0 && H > 0 && W > 0 && C > 0 && k > 0 && stride > 0, "col2im: bad arguments");
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  SEGMIF_REQUIRE(ldc >= C * k * k, "col2im: column pitch %d < C*k*k", ldc);
  const int64_t n = (int64_t)B * H * W * C;
  if (dx_dtype == SEGMIF_F32)
    col2im_kernel<float><<<grid_n(n, 512), 256, 0, as_stream(stream)>>>((const bf16*)dcol, ldc, (float*)dx, B, H, W, C, k, stride, pad, Ho, Wo);
  else
    col2im_kernel<bf16><<<grid_n(n, 512), 256, 0, as_stream(stream)>>>((const bf16*)dcol, ldc, (bf16*)dx, B, H, W, C, k, stride, pad, Ho, Wo);
  return check_launch("segmif_col2im");
}

extern "C" int segmif_channel_affine_nchw(const float* x, const float* scale, const float* shift, float* y, int B, int C,
                                          int64_t HW, segmif_stream_t stream) {
  SEGMIF_REQUIRE(x && scale && y && B > 0 && C > 0 && HW > 0, "channel_affine_nchw: bad arguments");
  const int64_t n = (int64_t)B * C * HW;
  channel_affine_nchw_kernel<<<grid_n(n, 1024), 256, 0, as_stream(stream)>>>(x, scale, shift, y, C, HW, n);
  return check_launch("segmif_channel_affine_nchw");
}

extern "C" int segmif_recompose_rgb_bwd(const float* rgb, const float* drgb, float* dfused, int clamp01, int B, int64_t HW,
                                        segmif_stream_t stream) {
  SEGMIF_REQUIRE(rgb && drgb && dfused && B > 0 && HW > 0, "recompose_rgb_bwd: bad arguments");
  const int64_t n = (int64_t)B * HW;
  recompose_rgb_bwd_kernel<<<grid_n(n, 1024), 256, 0, as_stream(stream)>>>(rgb, drgb, dfused, clamp01, HW, n);
  return check_launch("segmif_recompose_rgb_bwd");
}

extern "C" int segmif_cast(const void* x, int x_dtype, void* y, int y_dtype, int64_t n, segmif_stream_t stream) {
  SEGMIF_REQUIRE(x && y && n > 0, "cast: bad arguments");
  cudaStream_t st = as_stream(stream);
  const int grid = grid_n(n, 1024);
  if (x_dtype == SEGMIF_F32 && y_dtype == SEGMIF_BF16) cast_kernel<float, bf16><<<grid, 256, 0, st>>>((const float*)x, (bf16*)y, n);
  else if (x_dtype == SEGMIF_BF16 && y_dtype == SEGMIF_F32) cast_kernel<bf16, float><<<grid, 256, 0, st>>>((const bf16*)x, (float*)y, n);
  else { set_error("cast: unsupported dtype pair (%d -> %d)", x_dtype, y_dtype); return SEGMIF_ERR_INVALID; }
  return check_launch("segmif_cast");
}

extern "C" int segmif_scale_cast_rows(const float* x, const float* scale, void* y, int64_t rows, int64_t rows_per_sample, int C,
                                      segmif_stream_t stream) {
  SEGMIF_REQUIRE(x && scale && y && rows > 0 && rows_per_sample > 0 && C > 0 && C % 4 == 0, "scale_cast_rows: bad arguments");
  const int64_t n4 = rows * C / 4;
  scale_cast_rows_kernel<<<grid_n(n4, 1024), 256, 0, as_stream(stream)>>>(x, scale, (bf16*)y, rows_per_sample, C, n4);
  return check_launch("segmif_scale_cast_rows");
}

extern "C" int segmif_scale_add_rows(const float* x, const void* y, int y_dtype, const float* scale, float* out,
                                     int64_t rows, int64_t rows_per_sample, int C, segmif_stream_t stream) {
  SEGMIF_REQUIRE(x && y && out && rows > 0 && rows_per_sample > 0 && C > 0, "scale_add_rows: bad arguments");
  const int64_t n = rows * C;
  cudaStream_t st = as_stream(stream);
  if (y_dtype == SEGMIF_BF16) scale_add_rows_kernel<bf16><<<grid_n(n, 1024), 256, 0, st>>>(x, (const bf16*)y, scale, out, rows_per_sample, C, n);
  else scale_add_rows_kernel<float><<<grid_n(n, 1024), 256, 0, st>>>(x, (const float*)y, scale, out, rows_per_sample, C, n);
  return check_launch("segmif_scale_add_rows");
}

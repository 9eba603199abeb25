// Fusion-loss forward kernels: SSIM (11x11 Gaussian, separable in smem), LapLoss/LapLoss2 (3/5/7 DoG
// residuals from one halo tile), soft-histogram patch Entropy (lane == bin), Sobel+L1, MSE/L1 and the
// upsample-fused cross entropy.  fp32 throughout (sigma^2 = E[x^2]-mu^2 cancels).
// Every kernel writes one partial per block; finalize_kernel reduces them in fp64 in a fixed order, so
// results are deterministic run to run.
#include <algorithm>
#include <math.h>

#include "common.cuh"

namespace segmif {

__device__ __forceinline__ float block_sum_256(float v, float* sred) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sred[warp] = v;
  __syncthreads();
  float r = 0.f;
  if (warp == 0) {
    r = lane < (blockDim.x >> 5) ? sred[lane] : 0.f;
    r = warp_sum(r);
  }
  return r;     // valid in warp 0 (all lanes)
}

// partials: [groups][nblocks][nout] -> sums[groups][nout] (double accumulate, fixed order)
__global__ void __launch_bounds__(256) finalize_kernel(const float* __restrict__ partials, int nblocks, int nout,
                                                       double* __restrict__ sums) {
  __shared__ double sh[256];
  const int grp = blockIdx.x;
  for (int o = 0; o < nout; ++o) {
    double s = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += 256) s += (double)partials[((int64_t)grp * nblocks + i) * nout + o];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
      if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
      __syncthreads();
    }
    if (threadIdx.x == 0) sums[grp * nout + o] = sh[0];
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ SSIM
struct Gauss11 { float g[11]; };

// Tile = 64 x 32 outputs.  Both passes are register blocked: a thread produces 16 (horizontal) / 8 (vertical)
// consecutive outputs from a sliding window, so each shared-memory value is loaded once per thread instead of 11 times
// (the first version was LSU-bound at 8 % of the HBM peak).
constexpr int kSsimTX = 64, kSsimTY = 32, kSsimR = 5;
constexpr int kSsimIW = kSsimTX + 2 * kSsimR;      // 74 input columns
constexpr int kSsimIH = kSsimTY + 2 * kSsimR;      // 42 input rows
constexpr int kSsimIP = kSsimIW + 1;               // odd pitch: rows hit different banks
constexpr size_t kSsimSmem = (size_t)(2 * kSsimIH * kSsimIP + 5 * kSsimIH * kSsimTX) * sizeof(float);

__global__ void __launch_bounds__(256) ssim_kernel(const float* __restrict__ a, const float* __restrict__ b, int H,
                                                   int W, Gauss11 win, float* __restrict__ partials) {
  extern __shared__ float ssm[];
  float* sa = ssm;                                  // [42][75]
  float* sb = sa + kSsimIH * kSsimIP;
  float* hz = sb + kSsimIH * kSsimIP;               // [5][42][64]
  __shared__ float sred[8];
  const int bx = blockIdx.x * kSsimTX, by = blockIdx.y * kSsimTY;
  const int64_t img = blockIdx.z;
  const float* pa = a + img * H * W;
  const float* pb = b + img * H * W;
  for (int i = threadIdx.x; i < kSsimIH * kSsimIW; i += 256) {
    const int r = i / kSsimIW, c = i - r * kSsimIW;
    const int y = by + r - kSsimR, x = bx + c - kSsimR;
    const bool ok = (unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W;
    sa[r * kSsimIP + c] = ok ? pa[(int64_t)y * W + x] : 0.f;
    sb[r * kSsimIP + c] = ok ? pb[(int64_t)y * W + x] : 0.f;
  }
  __syncthreads();
  // horizontal: item = (column block of 16, row); rows vary fastest across lanes -> conflict-free (odd pitch)
  for (int item = threadIdx.x; item < (kSsimTX / 16) * kSsimIH; item += 256) {
    const int r = item % kSsimIH, c0 = (item / kSsimIH) * 16;
    float m1[16], m2[16], s11[16], s22[16], s12[16];
#pragma unroll
    for (int o = 0; o < 16; ++o) { m1[o] = m2[o] = s11[o] = s22[o] = s12[o] = 0.f; }
#pragma unroll
    for (int k = 0; k < 26; ++k) {
      const float x = sa[r * kSsimIP + c0 + k], y = sb[r * kSsimIP + c0 + k];
      const float xx = x * x, yy = y * y, xy = x * y;
#pragma unroll
      for (int o = 0; o < 16; ++o) {
        const int t = k - o;                         // tap index of output o for input k
        if (t >= 0 && t < 11) {
          const float g = win.g[t];
          m1[o] = fmaf(g, x, m1[o]); m2[o] = fmaf(g, y, m2[o]);
          s11[o] = fmaf(g, xx, s11[o]); s22[o] = fmaf(g, yy, s22[o]); s12[o] = fmaf(g, xy, s12[o]);
        }
      }
    }
#pragma unroll
    for (int o = 0; o < 16; ++o) {
      const int idx = r * kSsimTX + c0 + o;
      hz[idx] = m1[o]; hz[kSsimIH * kSsimTX + idx] = m2[o]; hz[2 * kSsimIH * kSsimTX + idx] = s11[o];
      hz[3 * kSsimIH * kSsimTX + idx] = s22[o]; hz[4 * kSsimIH * kSsimTX + idx] = s12[o];
    }
  }
  __syncthreads();
  // vertical: thread = (column, block of 8 rows); columns vary fastest across lanes -> conflict-free
  float local = 0.f;
  {
    const int c = threadIdx.x & 63, r0 = (threadIdx.x >> 6) * 8;
    float acc[5][8];
#pragma unroll
    for (int q = 0; q < 5; ++q)
#pragma unroll
      for (int o = 0; o < 8; ++o) acc[q][o] = 0.f;
#pragma unroll
    for (int k = 0; k < 18; ++k) {
      float v[5];
#pragma unroll
      for (int q = 0; q < 5; ++q) v[q] = hz[q * kSsimIH * kSsimTX + (r0 + k) * kSsimTX + c];
#pragma unroll
      for (int o = 0; o < 8; ++o) {
        const int t = k - o;
        if (t >= 0 && t < 11) {
          const float g = win.g[t];
#pragma unroll
          for (int q = 0; q < 5; ++q) acc[q][o] = fmaf(g, v[q], acc[q][o]);
        }
      }
    }
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      if (by + r0 + o < H && bx + c < W) {
        const float m1 = acc[0][o], m2 = acc[1][o];
        const float mu1_sq = m1 * m1, mu2_sq = m2 * m2, mu12 = m1 * m2;
        const float sg1 = acc[2][o] - mu1_sq, sg2 = acc[3][o] - mu2_sq, sg12 = acc[4][o] - mu12;
        const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
        local += ((2.f * mu12 + C1) * (2.f * sg12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sg1 + sg2 + C2));
      }
    }
  }
  const float tot = block_sum_256(local, sred);
  if (threadIdx.x == 0) partials[((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = tot;
}

// ------------------------------------------------------------------------------------------------ Laplacian
// The reference's 2-D kernel exp(-(dx^2+dy^2)/(2 s^2)) / sum is the outer product of the normalised 1-D Gaussian with
// itself, so every scale is filtered separably.  Tile = 64 x 32 outputs; images are processed one after the other
// through the same shared-memory planes, their residuals (3 scales x 8 rows) stay in registers.
struct LapKernels { float g3[3], g5[5], g7[7]; };

constexpr int kLapTX = 64, kLapTY = 32, kLapR = 3;
constexpr int kLapIW = kLapTX + 2 * kLapR;        // 70
constexpr int kLapIH = kLapTY + 2 * kLapR;        // 38
constexpr int kLapIP = kLapIW + 1;                // 71 (odd)
constexpr size_t kLapSmem = (size_t)(kLapIH * kLapIP + 3 * kLapIH * kLapTX) * sizeof(float);

// NIMG = 3: LapLoss2 (input, ir, vis -> target = max(res ir, res vis));  NIMG = 2: LapLoss (input, target)
template <int NIMG>
__global__ void __launch_bounds__(256) laploss_kernel(const float* __restrict__ inp, const float* __restrict__ p1,
                                                      const float* __restrict__ p2, int H, int W, LapKernels ker,
                                                      float* __restrict__ partials) {
  extern __shared__ float lsm[];
  float* tile = lsm;                                // [38][71]
  float* hz = tile + kLapIH * kLapIP;               // [3 scales][38][64]
  __shared__ float sred[8];
  const int bx = blockIdx.x * kLapTX, by = blockIdx.y * kLapTY;
  const int64_t off = (int64_t)blockIdx.z * H * W;
  const int c = threadIdx.x & 63, r0 = (threadIdx.x >> 6) * 8;
  float res[NIMG][3][8];
#pragma unroll
  for (int im = 0; im < NIMG; ++im) {
    const float* src = (im == 0 ? inp : im == 1 ? p1 : p2) + off;
    __syncthreads();                                // previous image's planes are fully consumed
    for (int i = threadIdx.x; i < kLapIH * kLapIW; i += 256) {
      const int r = i / kLapIW, cc = i - r * kLapIW;
      const int y = by + r - kLapR, x = bx + cc - kLapR;
      tile[r * kLapIP + cc] = ((unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W) ? src[(int64_t)y * W + x] : 0.f;
    }
    __syncthreads();
    // horizontal pass of the three scales: item = (column block of 16, row), rows fastest across lanes
    for (int item = threadIdx.x; item < (kLapTX / 16) * kLapIH; item += 256) {
      const int r = item % kLapIH, c0 = (item / kLapIH) * 16;
      float h3[16], h5[16], h7[16];
#pragma unroll
      for (int o = 0; o < 16; ++o) { h3[o] = h5[o] = h7[o] = 0.f; }
#pragma unroll
      for (int k = 0; k < 22; ++k) {                 // input column c0 + k covers output o at tap k - o (7-tap frame)
        const float x = tile[r * kLapIP + c0 + k];
#pragma unroll
        for (int o = 0; o < 16; ++o) {
          const int t = k - o;
          if (t >= 0 && t < 7) h7[o] = fmaf(ker.g7[t], x, h7[o]);
          if (t >= 1 && t < 6) h5[o] = fmaf(ker.g5[t - 1], x, h5[o]);
          if (t >= 2 && t < 5) h3[o] = fmaf(ker.g3[t - 2], x, h3[o]);
        }
      }
#pragma unroll
      for (int o = 0; o < 16; ++o) {
        const int idx = r * kLapTX + c0 + o;
        hz[idx] = h3[o]; hz[kLapIH * kLapTX + idx] = h5[o]; hz[2 * kLapIH * kLapTX + idx] = h7[o];
      }
    }
    __syncthreads();
    // vertical pass: thread = (column, 8 rows); residual = centre - blurred
    float v3[8], v5[8], v7[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) { v3[o] = v5[o] = v7[o] = 0.f; }
#pragma unroll
    for (int k = 0; k < 14; ++k) {
      const float a3 = hz[(r0 + k) * kLapTX + c], a5 = hz[kLapIH * kLapTX + (r0 + k) * kLapTX + c],
                  a7 = hz[2 * kLapIH * kLapTX + (r0 + k) * kLapTX + c];
#pragma unroll
      for (int o = 0; o < 8; ++o) {
        const int t = k - o;
        if (t >= 0 && t < 7) v7[o] = fmaf(ker.g7[t], a7, v7[o]);
        if (t >= 1 && t < 6) v5[o] = fmaf(ker.g5[t - 1], a5, v5[o]);
        if (t >= 2 && t < 5) v3[o] = fmaf(ker.g3[t - 2], a3, v3[o]);
      }
    }
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      const float centre = tile[(r0 + o + kLapR) * kLapIP + c + kLapR];
      res[im][0][o] = centre - v3[o]; res[im][1][o] = centre - v5[o]; res[im][2][o] = centre - v7[o];
    }
  }
  float l3 = 0.f, l5 = 0.f, l7 = 0.f;
#pragma unroll
  for (int o = 0; o < 8; ++o) {
    if (by + r0 + o < H && bx + c < W) {
      float t3 = res[1][0][o], t5 = res[1][1][o], t7 = res[1][2][o];
      if (NIMG == 3) { t3 = fmaxf(t3, res[NIMG - 1][0][o]); t5 = fmaxf(t5, res[NIMG - 1][1][o]); t7 = fmaxf(t7, res[NIMG - 1][2][o]); }
      l3 += fabsf(res[0][0][o] - t3); l5 += fabsf(res[0][1][o] - t5); l7 += fabsf(res[0][2][o] - t7);
    }
  }
  const int64_t blk = ((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  const float t3 = block_sum_256(l3, sred);
  const float t5 = block_sum_256(l5, sred);
  const float t7 = block_sum_256(l7, sred);
  if (threadIdx.x == 0) { partials[blk * 3 + 0] = t3; partials[blk * 3 + 1] = t5; partials[blk * 3 + 2] = t7; }
}

// ------------------------------------------------------------------------------------------------ Entropy
// A warp walks 32-float wide row segments: P rows x 32 columns hold 32/P patches; lane k is histogram bin k.
struct Bins32 { float b[32]; };

template <int P>
__global__ void __launch_bounds__(256) entropy_kernel(const float* __restrict__ img, int B, int H, int W, Bins32 bins,
                                                      float* __restrict__ partials) {
  __shared__ float sred[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int segs_x = (W + 31) / 32;
  const int prow = H / P;
  const int64_t nseg = (int64_t)B * prow * segs_x;
  const float mybin = bins.b[lane];
  const float inv_sigma = 1.0f / 0.01f;
  const float nhl2e = -0.5f * 1.4426950408889634f;
  float total = 0.f;
  for (int64_t sidx = (int64_t)blockIdx.x * nwarp + warp; sidx < nseg; sidx += (int64_t)gridDim.x * nwarp) {
    const int sx = (int)(sidx % segs_x);
    const int py = (int)((sidx / segs_x) % prow);
    const int64_t b = sidx / ((int64_t)segs_x * prow);
    const int x = sx * 32 + lane;
    float v[P];
#pragma unroll
    for (int dy = 0; dy < P; ++dy) v[dy] = x < W ? img[(b * H + (int64_t)py * P + dy) * W + x] : 0.f;
#pragma unroll
    for (int q = 0; q < 32 / P; ++q) {
      if (sx * 32 + q * P >= W) break;              // warp-uniform
      float acc = 0.f;
#pragma unroll
      for (int dy = 0; dy < P; ++dy)
#pragma unroll
        for (int dx = 0; dx < P; ++dx) {
          const float val = __shfl_sync(0xffffffffu, v[dy], q * P + dx);
          const float r = (val - mybin) * inv_sigma;
          acc += exp2f(nhl2e * (r * r));          // exp(-r^2/2) through MUFU.EX2: the loop is MUFU-issue bound
        }
      float pdf = acc / (float)(P * P);
      const float norm = warp_sum(pdf) + 1e-40f;
      pdf = pdf / norm + 1e-40f;
      total -= warp_sum(pdf * logf(pdf));
    }
  }
  if (lane != 0) total = 0.f;
  const float t = block_sum_256(total, sred);
  if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

// ------------------------------------------------------------------------------------------------ Sobel + L1
__device__ __forceinline__ float sobel_at(const float* p, int y, int x, int H, int W) {
  float n[3][3];
#pragma unroll
  for (int dy = 0; dy < 3; ++dy)
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const int yy = y + dy - 1, xx = x + dx - 1;
      n[dy][dx] = ((unsigned)yy < (unsigned)H && (unsigned)xx < (unsigned)W) ? p[(int64_t)yy * W + xx] : 0.f;
    }
  const float gx = -n[0][0] + n[0][2] - 2.f * n[1][0] + 2.f * n[1][2] - n[2][0] + n[2][2];
  const float gy = n[0][0] + 2.f * n[0][1] + n[0][2] - n[2][0] - 2.f * n[2][1] - n[2][2];
  return fabsf(gx) + fabsf(gy);
}

__global__ void __launch_bounds__(256) sobel_l1_kernel(const float* __restrict__ x, const float* __restrict__ y, int B,
                                                       int H, int W, float* __restrict__ partials) {
  __shared__ float sred[8];
  float l1 = 0.f, lg = 0.f;
  const int64_t n = (int64_t)B * H * W;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const int px = (int)(i % W), py = (int)((i / W) % H);
    const int64_t b = i / ((int64_t)W * H);
    const float* xp = x + b * H * W;
    const float* yp = y + b * H * W;
    l1 += fabsf(xp[(int64_t)py * W + px] - yp[(int64_t)py * W + px]);
    lg += fabsf(sobel_at(xp, py, px, H, W) - sobel_at(yp, py, px, H, W));
  }
  const float a = block_sum_256(l1, sred);
  const float g = block_sum_256(lg, sred);
  if (threadIdx.x == 0) { partials[blockIdx.x * 2] = a; partials[blockIdx.x * 2 + 1] = g; }
}

__global__ void __launch_bounds__(256) mse_l1_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                     int64_t n, float* __restrict__ partials) {
  __shared__ float sred[8];
  float s2 = 0.f, s1 = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const float d = x[i] - y[i];
    s2 = fmaf(d, d, s2);
    s1 += fabsf(d);
  }
  const float a = block_sum_256(s2, sred);
  const float b = block_sum_256(s1, sred);
  if (threadIdx.x == 0) { partials[blockIdx.x * 2] = a; partials[blockIdx.x * 2 + 1] = b; }
}

// ------------------------------------------------------------------------------------------------ upsample + CE
__device__ __forceinline__ void bl_src(int dst, float scale, int in_size, int& i0, int& i1, float& l0, float& l1) {
  float s = scale * ((float)dst + 0.5f) - 0.5f;
  s = s < 0.f ? 0.f : s;
  i0 = (int)s;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  l1 = s - (float)i0;
  l0 = 1.f - l1;
}

__global__ void __launch_bounds__(256) upsample_ce_kernel(const float* __restrict__ logits, int B, int h, int w, int nc,
                                                          const int64_t* __restrict__ labels, int H, int W,
                                                          int ignore_index, float sy, float sx,
                                                          float* __restrict__ partials) {
  __shared__ float sred[8];
  float loss = 0.f, cnt = 0.f;
  const int64_t n = (int64_t)B * H * W;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const int64_t lab = labels[i];
    if (lab == ignore_index || lab < 0 || lab >= nc) continue;
    const int X = (int)(i % W), Y = (int)((i / W) % H);
    const int64_t b = i / ((int64_t)W * H);
    int y0, y1, x0, x1;
    float hy0, hy1, wx0, wx1;
    bl_src(Y, sy, h, y0, y1, hy0, hy1);
    bl_src(X, sx, w, x0, x1, wx0, wx1);
    const float* base = logits + b * h * w * nc;
    const float* p00 = base + ((int64_t)y0 * w + x0) * nc;
    const float* p01 = base + ((int64_t)y0 * w + x1) * nc;
    const float* p10 = base + ((int64_t)y1 * w + x0) * nc;
    const float* p11 = base + ((int64_t)y1 * w + x1) * nc;
    float m = -INFINITY, picked = 0.f;
    float vals[32];
    for (int c = 0; c < nc; ++c) {
      const float v = hy0 * (wx0 * p00[c] + wx1 * p01[c]) + hy1 * (wx0 * p10[c] + wx1 * p11[c]);
      vals[c] = v;
      m = fmaxf(m, v);
      if (c == lab) picked = v;
    }
    float se = 0.f;
    for (int c = 0; c < nc; ++c) se += expf(vals[c] - m);
    loss += (m + logf(se)) - picked;
    cnt += 1.f;
  }
  const float a = block_sum_256(loss, sred);
  const float c = block_sum_256(cnt, sred);
  if (threadIdx.x == 0) { partials[blockIdx.x * 2] = a; partials[blockIdx.x * 2 + 1] = c; }
}

// scale sums -> outputs (tiny, one thread)
__global__ void loss_epilogue_kernel(const double* __restrict__ sums, int mode, int ngroups, double inv_n,
                                     float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  switch (mode) {
    case 0:  // mean per group (ssim): out[g] = sums[g] * inv_n
      for (int g = 0; g < ngroups; ++g) out[g] = (float)(sums[g] * inv_n);
      break;
    case 1:  // laplacian: 10*(l3+l5)+l7, each a mean  (fp32 combine like the reference)
    {
      const float l3 = (float)(sums[0] * inv_n), l5 = (float)(sums[1] * inv_n), l7 = (float)(sums[2] * inv_n);
      out[0] = 10.f * (l3 + l5) + l7;
      break;
    }
    case 2:  // plain sum (entropy)
      out[0] = (float)sums[0];
      break;
    case 3:  // two means
      out[0] = (float)(sums[0] * inv_n);
      out[1] = (float)(sums[1] * inv_n);
      break;
    case 4:  // ratio (cross entropy: sum / count)
      out[0] = (float)(sums[0] / sums[1]);
      break;
  }
}

static inline double* sums_area(float* workspace) { return reinterpret_cast<double*>(workspace); }
static inline float* partial_area(float* workspace) { return workspace + 64; }   // 256 bytes reserved for sums

static int finish(float* workspace, int groups, int nblocks, int nout, int mode, double inv_n, float* out, cudaStream_t st,
                  const char* what) {
  finalize_kernel<<<groups, 256, 0, st>>>(partial_area(workspace), nblocks, nout, sums_area(workspace));
  loss_epilogue_kernel<<<1, 32, 0, st>>>(sums_area(workspace), mode, groups, inv_n, out);
  return check_launch(what);
}

}  // namespace segmif

using namespace segmif;

extern "C" size_t segmif_loss_workspace_bytes(int B, int H, int W) {
  const size_t tiles = (size_t)B * ((H + 31) / 32) * ((W + 63) / 64);
  const size_t blocks = tiles > 4096 ? tiles : 4096;
  return 256 + blocks * 3 * sizeof(float);
}

static Gauss11 make_gauss11() {
  // pytorch_ssim/__init__.py:8-10: exp(-(x-5)^2 / (2*1.5^2)) as python floats, stored to fp32, normalised in fp32
  Gauss11 w;
  float tmp[11], s = 0.f;
  for (int i = 0; i < 11; ++i) { tmp[i] = (float)exp(-(double)((i - 5) * (i - 5)) / (2.0 * 1.5 * 1.5)); }
  for (int i = 0; i < 11; ++i) s += tmp[i];
  for (int i = 0; i < 11; ++i) w.g[i] = tmp[i] / s;
  return w;
}

extern "C" int segmif_ssim_fwd(const float* img1, const float* img2, int B, int H, int W, int per_image,
                               float* workspace, float* out, segmif_stream_t stream) {
  SEGMIF_REQUIRE(img1 && img2 && workspace && out, "ssim: null pointer");
  SEGMIF_REQUIRE(B > 0 && H > 0 && W > 0, "ssim: empty input");
  SEGMIF_REQUIRE(!per_image || B <= 32, "ssim: per-image mode supports at most 32 images per call");
  static const Gauss11 win = make_gauss11();
  dim3 grid((W + kSsimTX - 1) / kSsimTX, (H + kSsimTY - 1) / kSsimTY, B);
  cudaStream_t st = as_stream(stream);
  static bool cfg = false;
  if (!cfg) { cudaFuncSetAttribute(ssim_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSsimSmem); cfg = true; }
  ssim_kernel<<<grid, 256, kSsimSmem, st>>>(img1, img2, H, W, win, partial_area(workspace));
  int rc = check_launch("segmif_ssim_fwd");
  if (rc) return rc;
  const int per = grid.x * grid.y;
  if (per_image) return finish(workspace, B, per, 1, 0, 1.0 / ((double)H * W), out, st, "segmif_ssim_fwd");
  return finish(workspace, 1, per * B, 1, 0, 1.0 / ((double)B * H * W), out, st, "segmif_ssim_fwd");
}

static LapKernels make_lap_kernels() {
  // lap_loss.py:39-60: exp(-(dx^2+dy^2)/(2*sigma^2)) normalised to sum 1 == outer product of n[i] = g[i] / sum(g)
  LapKernels k;
  const int sizes[3] = {3, 5, 7};
  float* dst[3] = {k.g3, k.g5, k.g7};
  for (int s = 0; s < 3; ++s) {
    const int n = sizes[s];
    const double mean = (n - 1) / 2.0;
    double g[7], sum = 0.0;
    for (int i = 0; i < n; ++i) { g[i] = exp(-(i - mean) * (i - mean) / 8.0); sum += g[i]; }
    for (int i = 0; i < n; ++i) dst[s][i] = (float)(g[i] / sum);
  }
  return k;
}
extern "C" int segmif_laploss2_fwd(const float* inp, const float* ir, const float* vis, int B, int H, int W,
                                   float* workspace, float* out, segmif_stream_t stream) {
  SEGMIF_REQUIRE(inp && ir && vis && workspace && out, "laploss2: null pointer");
  SEGMIF_REQUIRE(B > 0 && H > 0 && W > 0, "laploss2: empty input");
  static const LapKernels ker = make_lap_kernels();
  dim3 grid((W + kLapTX - 1) / kLapTX, (H + kLapTY - 1) / kLapTY, B);
  cudaStream_t st = as_stream(stream);
  static bool cfg = false;
  if (!cfg) { cudaFuncSetAttribute(laploss_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLapSmem); cfg = true; }
  laploss_kernel<3><<<grid, 256, kLapSmem, st>>>(inp, ir, vis, H, W, ker, partial_area(workspace));
  int rc = check_launch("segmif_laploss2_fwd");
  if (rc) return rc;
  return finish(workspace, 1, grid.x * grid.y * B, 3, 1, 1.0 / ((double)B * H * W), out, st, "segmif_laploss2_fwd");
}

extern "C" int segmif_laploss_fwd(const float* inp, const float* target, int B, int H, int W, float* workspace,
                                  float* out, segmif_stream_t stream) {
  SEGMIF_REQUIRE(inp && target && workspace && out, "laploss: null pointer");
  SEGMIF_REQUIRE(B > 0 && H > 0 && W > 0, "laploss: empty input");
  static const LapKernels ker = make_lap_kernels();
  dim3 grid((W + kLapTX - 1) / kLapTX, (H + kLapTY - 1) / kLapTY, B);
  cudaStream_t st = as_stream(stream);
  static bool cfg = false;
  if (!cfg) { cudaFuncSetAttribute(laploss_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLapSmem); cfg = true; }
  laploss_kernel<2><<<grid, 256, kLapSmem, st>>>(inp, target, nullptr, H, W, ker, partial_area(workspace));
  int rc = check_launch("segmif_laploss_fwd");
  if (rc) return rc;
  return finish(workspace, 1, grid.x * grid.y * B, 3, 1, 1.0 / ((double)B * H * W), out, st, "segmif_laploss_fwd");
}

extern "C" int segmif_entropy_fwd(const float* img, int B, int H, int W, int patch, float* workspace, float* out,
                                  segmif_stream_t stream) {
  SEGMIF_REQUIRE(img && workspace && out, "entropy: null pointer");
  SEGMIF_REQUIRE(patch == 2 || patch == 4 || patch == 8 || patch == 16, "entropy: patch size %d unsupported (2,4,8,16)", patch);
  SEGMIF_REQUIRE(H % patch == 0 && W % patch == 0 && B > 0, "entropy: H and W must be multiples of the patch size");
  Bins32 bins;   // torch.linspace(0, 1, 32) in fp32: symmetric two-sided formula
  const float step = 1.0f / 31.0f;
  for (int i = 0; i < 32; ++i) bins.b[i] = i < 16 ? 0.0f + step * (float)i : 1.0f - step * (float)(31 - i);
  const int nblocks = 148 * 8;
  cudaStream_t st = as_stream(stream);
  float* part = partial_area(workspace);
  switch (patch) {
    case 2: entropy_kernel<2><<<nblocks, 256, 0, st>>>(img, B, H, W, bins, part); break;
    case 4: entropy_kernel<4><<<nblocks, 256, 0, st>>>(img, B, H, W, bins, part); break;
    case 8: entropy_kernel<8><<<nblocks, 256, 0, st>>>(img, B, H, W, bins, part); break;
    default: entropy_kernel<16><<<nblocks, 256, 0, st>>>(img, B, H, W, bins, part); break;
  }
  int rc = check_launch("segmif_entropy_fwd");
  if (rc) return rc;
  return finish(workspace, 1, nblocks, 1, 2, 1.0, out, st, "segmif_entropy_fwd");
}

extern "C" int segmif_sobel_l1_fwd(const float* x, const float* y, int B, int H, int W, float* workspace, float* out,
                                   segmif_stream_t stream) {
  SEGMIF_REQUIRE(x && y && workspace && out, "sobel_l1: null pointer");
  SEGMIF_REQUIRE(B > 0 && H > 0 && W > 0, "sobel_l1: empty input");
  const int64_t n = (int64_t)B * H * W;
  const int nblocks = (int)std::min<int64_t>(ceil_div(n, 256), 148 * 8);
  cudaStream_t st = as_stream(stream);
  sobel_l1_kernel<<<nblocks, 256, 0, st>>>(x, y, B, H, W, partial_area(workspace));
  int rc = check_launch("segmif_sobel_l1_fwd");
  if (rc) return rc;
  return finish(workspace, 1, nblocks, 2, 3, 1.0 / (double)n, out, st, "segmif_sobel_l1_fwd");
}

extern "C" int segmif_mse_l1_fwd(const float* x, const float* y, int64_t n, float* workspace, float* out,
                                 segmif_stream_t stream) {
  SEGMIF_REQUIRE(x && y && workspace && out && n > 0, "mse_l1: bad arguments");
  const int nblocks = (int)std::min<int64_t>(ceil_div(n, 256), 148 * 8);
  cudaStream_t st = as_stream(stream);
  mse_l1_kernel<<<nblocks, 256, 0, st>>>(x, y, n, partial_area(workspace));
  int rc = check_launch("segmif_mse_l1_fwd");
  if (rc) return rc;
  return finish(workspace, 1, nblocks, 2, 3, 1.0 / (double)n, out, st, "segmif_mse_l1_fwd");
}

extern "C" int segmif_upsample_ce_fwd(const float* logits, int B, int h, int w, int nc, const int64_t* labels, int H,
                                      int W, int ignore_index, float* workspace, float* out, segmif_stream_t stream) {
  SEGMIF_REQUIRE(logits && labels && workspace && out, "upsample_ce: null pointer");
  SEGMIF_REQUIRE(nc > 0 && nc <= 32, "upsample_ce: nc=%d unsupported (1..32)", nc);
  const int64_t n = (int64_t)B * H * W;
  SEGMIF_REQUIRE(n > 0, "upsample_ce: empty input");
  const int nblocks = (int)std::min<int64_t>(ceil_div(n, 256), 148 * 8);
  cudaStream_t st = as_stream(stream);
  upsample_ce_kernel<<<nblocks, 256, 0, st>>>(logits, B, h, w, nc, labels, H, W, ignore_index, (float)h / (float)H,
                                              (float)w / (float)W, partial_area(workspace));
  int rc = check_launch("segmif_upsample_ce_fwd");
  if (rc) return rc;
  return finish(workspace, 1, nblocks, 2, 4, 1.0, out, st, "segmif_upsample_ce_fwd");
}

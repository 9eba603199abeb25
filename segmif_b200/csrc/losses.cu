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

__global__ void __launch_bounds__(256) ssim_kernel(const float* __restrict__ a, const float* __restrict__ b, int H,
                                                   int W, Gauss11 win, float* __restrict__ partials) {
  constexpr int T = 32, R = 5, TW = T + 2 * R;   // 42
  __shared__ float sa[TW][TW + 1], sb[TW][TW + 1];
  __shared__ float hz[5][TW][T + 1];
  __shared__ float sred[8];
  const int bx = blockIdx.x * T, by = blockIdx.y * T;
  const int64_t img = blockIdx.z;
  const float* pa = a + img * H * W;
  const float* pb = b + img * H * W;
  for (int i = threadIdx.x; i < TW * TW; i += 256) {
    const int r = i / TW, c = i % TW;
    const int y = by + r - R, x = bx + c - R;
    const bool ok = (unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W;
    sa[r][c] = ok ? pa[(int64_t)y * W + x] : 0.f;
    sb[r][c] = ok ? pb[(int64_t)y * W + x] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < TW * T; i += 256) {
    const int r = i / T, c = i % T;
    float m1 = 0.f, m2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float x = sa[r][c + k], y = sb[r][c + k], g = win.g[k];
      m1 = fmaf(g, x, m1); m2 = fmaf(g, y, m2);
      s11 = fmaf(g, x * x, s11); s22 = fmaf(g, y * y, s22); s12 = fmaf(g, x * y, s12);
    }
    hz[0][r][c] = m1; hz[1][r][c] = m2; hz[2][r][c] = s11; hz[3][r][c] = s22; hz[4][r][c] = s12;
  }
  __syncthreads();
  float local = 0.f;
  for (int i = threadIdx.x; i < T * T; i += 256) {
    const int r = i / T, c = i % T;
    if (by + r >= H || bx + c >= W) continue;
    float m1 = 0.f, m2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float g = win.g[k];
      m1 = fmaf(g, hz[0][r + k][c], m1); m2 = fmaf(g, hz[1][r + k][c], m2);
      s11 = fmaf(g, hz[2][r + k][c], s11); s22 = fmaf(g, hz[3][r + k][c], s22); s12 = fmaf(g, hz[4][r + k][c], s12);
    }
    const float mu1_sq = m1 * m1, mu2_sq = m2 * m2, mu12 = m1 * m2;
    const float sg1 = s11 - mu1_sq, sg2 = s22 - mu2_sq, sg12 = s12 - mu12;
    const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
    local += ((2.f * mu12 + C1) * (2.f * sg12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sg1 + sg2 + C2));
  }
  const float tot = block_sum_256(local, sred);
  if (threadIdx.x == 0) partials[((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = tot;
}

// ------------------------------------------------------------------------------------------------ Laplacian
struct LapKernels { float k3[9], k5[25], k7[49]; };

template <int K>
__device__ __forceinline__ float lap_residual(const float (*t)[39], int r, int c, const float* ker) {
  // t has a 3-pixel halo; residual = centre - (G_K * img)
  constexpr int R = K / 2;
  float s = 0.f;
#pragma unroll
  for (int dy = 0; dy < K; ++dy)
#pragma unroll
    for (int dx = 0; dx < K; ++dx) s = fmaf(ker[dy * K + dx], t[r + 3 - R + dy][c + 3 - R + dx], s);
  return t[r + 3][c + 3] - s;
}

// NIMG = 3: LapLoss2 (input, ir, vis -> target = max(res ir, res vis));  NIMG = 2: LapLoss (input, target)
template <int NIMG>
__global__ void __launch_bounds__(256) laploss_kernel(const float* __restrict__ inp, const float* __restrict__ p1,
                                                      const float* __restrict__ p2, int H, int W, LapKernels ker,
                                                      float* __restrict__ partials) {
  constexpr int T = 32, R = 3, TW = T + 2 * R;   // 38
  __shared__ float s0[TW][39], s1[TW][39], s2[NIMG == 3 ? TW : 1][39];
  __shared__ float sred[8];
  const int bx = blockIdx.x * T, by = blockIdx.y * T;
  const int64_t off = (int64_t)blockIdx.z * H * W;
  for (int i = threadIdx.x; i < TW * TW; i += 256) {
    const int r = i / TW, c = i % TW;
    const int y = by + r - R, x = bx + c - R;
    const bool ok = (unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W;
    const int64_t o = off + (int64_t)y * W + x;
    s0[r][c] = ok ? inp[o] : 0.f;
    s1[r][c] = ok ? p1[o] : 0.f;
    if (NIMG == 3) s2[r][c] = ok ? p2[o] : 0.f;
  }
  __syncthreads();
  float l3 = 0.f, l5 = 0.f, l7 = 0.f;
  for (int i = threadIdx.x; i < T * T; i += 256) {
    const int r = i / T, c = i % T;
    if (by + r >= H || bx + c >= W) continue;
    {
      const float a = lap_residual<3>(s0, r, c, ker.k3);
      float t = lap_residual<3>(s1, r, c, ker.k3);
      if (NIMG == 3) t = fmaxf(t, lap_residual<3>(s2, r, c, ker.k3));
      l3 += fabsf(a - t);
    }
    {
      const float a = lap_residual<5>(s0, r, c, ker.k5);
      float t = lap_residual<5>(s1, r, c, ker.k5);
      if (NIMG == 3) t = fmaxf(t, lap_residual<5>(s2, r, c, ker.k5));
      l5 += fabsf(a - t);
    }
    {
      const float a = lap_residual<7>(s0, r, c, ker.k7);
      float t = lap_residual<7>(s1, r, c, ker.k7);
      if (NIMG == 3) t = fmaxf(t, lap_residual<7>(s2, r, c, ker.k7));
      l7 += fabsf(a - t);
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
    case 4:  // ratio (cross entropy: sum / count); out[1] = count (the backward's normaliser)
      out[0] = (float)(sums[0] / sums[1]);
      out[1] = (float)sums[1];
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
  const size_t tiles = (size_t)B * ((H + 31) / 32) * ((W + 31) / 32);
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
  dim3 grid((W + 31) / 32, (H + 31) / 32, B);
  cudaStream_t st = as_stream(stream);
  ssim_kernel<<<grid, 256, 0, st>>>(img1, img2, H, W, win, partial_area(workspace));
  int rc = check_launch("segmif_ssim_fwd");
  if (rc) return rc;
  const int per = grid.x * grid.y;
  if (per_image) return finish(workspace, B, per, 1, 0, 1.0 / ((double)H * W), out, st, "segmif_ssim_fwd");
  return finish(workspace, 1, per * B, 1, 0, 1.0 / ((double)B * H * W), out, st, "segmif_ssim_fwd");
}

static LapKernels make_lap_kernels() {
  // lap_loss.py:39-60: fp32 exp of -(dx^2+dy^2)/(2*sigma^2), times 1/(2 pi sigma^2), normalised by its fp32 sum
  LapKernels k;
  const int sizes[3] = {3, 5, 7};
  float* dst[3] = {k.k3, k.k5, k.k7};
  for (int s = 0; s < 3; ++s) {
    const int n = sizes[s];
    const float mean = (n - 1) / 2.0f, var = 4.0f;
    float sum = 0.f;
    for (int y = 0; y < n; ++y)
      for (int x = 0; x < n; ++x) {
        const float d2 = (x - mean) * (x - mean) + (y - mean) * (y - mean);
        const float e = expf(-d2 / (2.f * var));
        const float v = (float)(1.0 / (2.0 * 3.14159265358979323846 * 4.0)) * e;
        dst[s][y * n + x] = v;
        sum += v;
      }
    for (int i = 0; i < n * n; ++i) dst[s][i] /= sum;
  }
  return k;
}

extern "C" int segmif_laploss2_fwd(const float* inp, const float* ir, const float* vis, int B, int H, int W,
                                   float* workspace, float* out, segmif_stream_t stream) {
  SEGMIF_REQUIRE(inp && ir && vis && workspace && out, "laploss2: null pointer");
  SEGMIF_REQUIRE(B > 0 && H > 0 && W > 0, "laploss2: empty input");
  static const LapKernels ker = make_lap_kernels();
  dim3 grid((W + 31) / 32, (H + 31) / 32, B);
  cudaStream_t st = as_stream(stream);
  laploss_kernel<3><<<grid, 256, 0, st>>>(inp, ir, vis, H, W, ker, partial_area(workspace));
  int rc = check_launch("segmif_laploss2_fwd");
  if (rc) return rc;
  return finish(workspace, 1, grid.x * grid.y * B, 3, 1, 1.0 / ((double)B * H * W), out, st, "segmif_laploss2_fwd");
}

extern "C" int segmif_laploss_fwd(const float* inp, const float* target, int B, int H, int W, float* workspace,
                                  float* out, segmif_stream_t stream) {
  SEGMIF_REQUIRE(inp && target && workspace && out, "laploss: null pointer");
  SEGMIF_REQUIRE(B > 0 && H > 0 && W > 0, "laploss: empty input");
  static const LapKernels ker = make_lap_kernels();
  dim3 grid((W + 31) / 32, (H + 31) / 32, B);
  cudaStream_t st = as_stream(stream);
  laploss_kernel<2><<<grid, 256, 0, st>>>(inp, target, nullptr, H, W, ker, partial_area(workspace));
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

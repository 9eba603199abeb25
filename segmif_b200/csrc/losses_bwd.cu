// Fusion-loss backward kernels (gradient with respect to the first image argument, fp32 planes [B,1,H,W]):
// MSE / L1, Sobel + L1 (core/loss.py:471-475), SSIM (pytorch_ssim/__init__.py:19-43), LapLoss / LapLoss2
// (lap_loss.py:93-118) and the soft-histogram patch entropy (core/Entropy.py:15-56).
// Every window is symmetric, so the adjoint of "filter then pointwise" is "pointwise derivative maps, then the same
// filter": each kernel recomputes the forward quantities of its tile plus a halo in shared memory, forms the
// derivative maps there and filters them again -- one read of the inputs, one write (or read-modify-write when
// `accumulate` is set, which lets several loss terms add into one gradient plane without extra passes).
// `gout` is the DEVICE scalar (or vector) of upstream gradients, so no host synchronisation is needed.
#include <algorithm>
#include <math.h>

#include "common.cuh"

namespace segmif {

__device__ __forceinline__ float sgnf(float v) { return (float)(v > 0.f) - (float)(v < 0.f); }

// ------------------------------------------------------------------------------------------------ MSE / L1
__global__ void __launch_bounds__(256) mse_l1_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                         int64_t n, const float* __restrict__ gout, float inv_n,
                                                         float* __restrict__ dx, int accumulate) {
  const float g2 = gout[0] * 2.f * inv_n, g1 = gout[1] * inv_n;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const float d = x[i] - y[i];
    const float v = g2 * d + g1 * sgnf(d);
    dx[i] = accumulate ? dx[i] + v : v;
  }
}

// ------------------------------------------------------------------------------------------------ Sobel + L1
__global__ void __launch_bounds__(256) sobel_l1_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                           int H, int W, const float* __restrict__ gout, float inv_n,
                                                           float* __restrict__ dx, int accumulate) {
  constexpr int T = 32, TW = T + 4, EW = T + 2;
  __shared__ float sx[TW][TW + 1], sy[TW][TW + 1];
  __shared__ float sa[EW][EW + 1], sb[EW][EW + 1];
  const int bx = blockIdx.x * T, by = blockIdx.y * T;
  const int64_t off = (int64_t)blockIdx.z * H * W;
  for (int i = threadIdx.x; i < TW * TW; i += 256) {
    const int r = i / TW, c = i % TW;
    const int yy = by + r - 2, xx = bx + c - 2;
    const bool ok = (unsigned)yy < (unsigned)H && (unsigned)xx < (unsigned)W;
    sx[r][c] = ok ? x[off + (int64_t)yy * W + xx] : 0.f;
    sy[r][c] = ok ? y[off + (int64_t)yy * W + xx] : 0.f;
  }
  __syncthreads();
  const float g_l1 = gout[0] * inv_n, g_sb = gout[1] * inv_n;
  // a = e * sign(gx(x)), b = e * sign(gy(x)) with e = sign(S(x) - S(y)) on the tile plus a 1-pixel ring
  for (int i = threadIdx.x; i < EW * EW; i += 256) {
    const int r = i / EW, c = i % EW;
    const int yy = by + r - 1, xx = bx + c - 1;
    float a = 0.f, b = 0.f;
    if ((unsigned)yy < (unsigned)H && (unsigned)xx < (unsigned)W) {
      float s[2], sgx = 0.f, sgy = 0.f;
#pragma unroll
      for (int im = 0; im < 2; ++im) {
        const float(*t)[TW + 1] = im == 0 ? sx : sy;
        const float gx = -t[r][c] + t[r][c + 2] - 2.f * t[r + 1][c] + 2.f * t[r + 1][c + 2] - t[r + 2][c] + t[r + 2][c + 2];
        const float gy = t[r][c] + 2.f * t[r][c + 1] + t[r][c + 2] - t[r + 2][c] - 2.f * t[r + 2][c + 1] - t[r + 2][c + 2];
        s[im] = fabsf(gx) + fabsf(gy);
        if (im == 0) { sgx = sgnf(gx); sgy = sgnf(gy); }
      }
      const float e = sgnf(s[0] - s[1]);
      a = e * sgx;
      b = e * sgy;
    }
    sa[r][c] = a;
    sb[r][c] = b;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T * T; i += 256) {
    const int r = i / T, c = i % T;
    const int yy = by + r, xx = bx + c;
    if (yy >= H || xx >= W) continue;
    // dx(q) = sum_d kx[d] a(q - d) + ky[d] b(q - d);  (q - d) with d = (dy, dx) in [-1,1]^2 sits at sa[r+1-dy][c+1-dx]
    const float gxs = -sa[r + 2][c + 2] + sa[r + 2][c] - 2.f * sa[r + 1][c + 2] + 2.f * sa[r + 1][c] - sa[r][c + 2] + sa[r][c];
    const float gys = sb[r + 2][c + 2] + 2.f * sb[r + 2][c + 1] + sb[r + 2][c] - sb[r][c + 2] - 2.f * sb[r][c + 1] - sb[r][c];
    const float v = g_l1 * sgnf(sx[r + 2][c + 2] - sy[r + 2][c + 2]) + g_sb * (gxs + gys);
    const int64_t o = off + (int64_t)yy * W + xx;
    dx[o] = accumulate ? dx[o] + v : v;
  }
}

// ------------------------------------------------------------------------------------------------ SSIM
struct Gauss11b { float g[11]; };

__global__ void __launch_bounds__(256) ssim_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, int H,
                                                       int W, Gauss11b win, const float* __restrict__ gout,
                                                       int per_image, float inv_n, float* __restrict__ da,
                                                       int accumulate) {
  constexpr int T = 32, R = 5, IW = T + 4 * R, MW = T + 2 * R;    // 52 input columns, 42 derivative-map columns
  extern __shared__ float sm[];
  float(*sa)[IW + 1] = reinterpret_cast<float(*)[IW + 1]>(sm);                                   // [52][53]
  float(*sb)[IW + 1] = reinterpret_cast<float(*)[IW + 1]>(sm + IW * (IW + 1));
  float* hz = sm + 2 * IW * (IW + 1);                                                            // [5][52][43]
  float* abg = hz + 5 * IW * (MW + 1);                                                           // [3][42][43]
  float* h2 = hz;                                                                                // [3][42][33] (hz is dead by then)
  const int bx = blockIdx.x * T, by = blockIdx.y * T;
  const int64_t off = (int64_t)blockIdx.z * H * W;
  for (int i = threadIdx.x; i < IW * IW; i += 256) {
    const int r = i / IW, c = i % IW;
    const int y = by + r - 2 * R, x = bx + c - 2 * R;
    const bool ok = (unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W;
    sa[r][c] = ok ? a[off + (int64_t)y * W + x] : 0.f;
    sb[r][c] = ok ? b[off + (int64_t)y * W + x] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < IW * MW; i += 256) {
    const int r = i / MW, c = i % MW;
    float m1 = 0.f, m2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float x = sa[r][c + k], y = sb[r][c + k], g = win.g[k];
      m1 = fmaf(g, x, m1); m2 = fmaf(g, y, m2);
      s11 = fmaf(g, x * x, s11); s22 = fmaf(g, y * y, s22); s12 = fmaf(g, x * y, s12);
    }
    float* h = hz + r * (MW + 1) + c;
    h[0] = m1; h[IW * (MW + 1)] = m2; h[2 * IW * (MW + 1)] = s11; h[3 * IW * (MW + 1)] = s22; h[4 * IW * (MW + 1)] = s12;
  }
  __syncthreads();
  // derivative maps dS/dmu1, dS/dE[x^2], dS/dE[xy] on the tile plus a 5-pixel ring (zero outside the image)
  for (int i = threadIdx.x; i < MW * MW; i += 256) {
    const int r = i / MW, c = i % MW;
    const int y = by + r - R, x = bx + c - R;
    float al = 0.f, be = 0.f, ga = 0.f;
    if ((unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W) {
      float m1 = 0.f, m2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
#pragma unroll
      for (int k = 0; k < 11; ++k) {
        const float g = win.g[k];
        const float* h = hz + (r + k) * (MW + 1) + c;
        m1 = fmaf(g, h[0], m1); m2 = fmaf(g, h[IW * (MW + 1)], m2);
        s11 = fmaf(g, h[2 * IW * (MW + 1)], s11); s22 = fmaf(g, h[3 * IW * (MW + 1)], s22); s12 = fmaf(g, h[4 * IW * (MW + 1)], s12);
      }
      const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
      const float mu1_sq = m1 * m1, mu2_sq = m2 * m2, mu12 = m1 * m2;
      const float A1 = 2.f * mu12 + C1, A2 = 2.f * (s12 - mu12) + C2;
      const float B1 = mu1_sq + mu2_sq + C1, B2 = (s11 - mu1_sq) + (s22 - mu2_sq) + C2;
      const float inv = 1.f / (B1 * B2);
      const float S = A1 * A2 * inv;
      al = 2.f * m2 * (A2 - A1) * inv - 2.f * m1 * S * (1.f / B1 - 1.f / B2);
      be = -S / B2;
      ga = 2.f * A1 * inv;
    }
    float* o = abg + r * (MW + 1) + c;
    o[0] = al; o[MW * (MW + 1)] = be; o[2 * MW * (MW + 1)] = ga;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < MW * T; i += 256) {
    const int r = i / T, c = i % T;
    float v0 = 0.f, v1 = 0.f, v2 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float g = win.g[k];
      const float* s = abg + r * (MW + 1) + c + k;
      v0 = fmaf(g, s[0], v0); v1 = fmaf(g, s[MW * (MW + 1)], v1); v2 = fmaf(g, s[2 * MW * (MW + 1)], v2);
    }
    float* o = h2 + r * (T + 1) + c;
    o[0] = v0; o[MW * (T + 1)] = v1; o[2 * MW * (T + 1)] = v2;
  }
  __syncthreads();
  const float gs = gout[per_image ? blockIdx.z : 0] * inv_n;
  for (int i = threadIdx.x; i < T * T; i += 256) {
    const int r = i / T, c = i % T;
    const int y = by + r, x = bx + c;
    if (y >= H || x >= W) continue;
    float v0 = 0.f, v1 = 0.f, v2 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float g = win.g[k];
      const float* s = h2 + (r + k) * (T + 1) + c;
      v0 = fmaf(g, s[0], v0); v1 = fmaf(g, s[MW * (T + 1)], v1); v2 = fmaf(g, s[2 * MW * (T + 1)], v2);
    }
    const float v = gs * (v0 + 2.f * sa[r + 2 * R][c + 2 * R] * v1 + sb[r + 2 * R][c + 2 * R] * v2);
    const int64_t o = off + (int64_t)y * W + x;
    da[o] = accumulate ? da[o] + v : v;
  }
}

// ------------------------------------------------------------------------------------------------ Laplacian
struct LapKernelsB { float k3[9], k5[25], k7[49]; };

template <int K, int PITCH>
__device__ __forceinline__ float lap_res(const float* t, int r, int c, const float* ker) {
  // t points at a tile whose element (r, c) is the centre; residual = centre - (G_K * img)
  constexpr int R = K / 2;
  float s = 0.f;
#pragma unroll
  for (int dy = 0; dy < K; ++dy)
#pragma unroll
    for (int dx = 0; dx < K; ++dx) s = fmaf(ker[dy * K + dx], t[(r - R + dy) * PITCH + c - R + dx], s);
  return t[r * PITCH + c] - s;
}

template <int NIMG>
__global__ void __launch_bounds__(256) laploss_bwd_kernel(const float* __restrict__ inp, const float* __restrict__ p1,
                                                          const float* __restrict__ p2, int H, int W, LapKernelsB ker,
                                                          const float* __restrict__ gout, float inv_n,
                                                          float* __restrict__ dinp, int accumulate) {
  constexpr int T = 32, R = 3, IW = T + 4 * R, EW = T + 2 * R, IP = IW + 1, EP = EW + 1;   // 44 / 38
  __shared__ float s0[IW * IP], s1[IW * IP], s2[NIMG == 3 ? IW * IP : 1];
  __shared__ float e[3][EW * EP];
  const int bx = blockIdx.x * T, by = blockIdx.y * T;
  const int64_t off = (int64_t)blockIdx.z * H * W;
  for (int i = threadIdx.x; i < IW * IW; i += 256) {
    const int r = i / IW, c = i % IW;
    const int y = by + r - 2 * R, x = bx + c - 2 * R;
    const bool ok = (unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W;
    const int64_t o = off + (int64_t)y * W + x;
    s0[r * IP + c] = ok ? inp[o] : 0.f;
    s1[r * IP + c] = ok ? p1[o] : 0.f;
    if (NIMG == 3) s2[r * IP + c] = ok ? p2[o] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < EW * EW; i += 256) {
    const int r = i / EW, c = i % EW;
    const int y = by + r - R, x = bx + c - R;
    float e3 = 0.f, e5 = 0.f, e7 = 0.f;
    if ((unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W) {
      const int rr = r + R, cc = c + R;      // position inside the input tile
      {
        const float a = lap_res<3, IP>(s0, rr, cc, ker.k3);
        float t = lap_res<3, IP>(s1, rr, cc, ker.k3);
        if (NIMG == 3) t = fmaxf(t, lap_res<3, IP>(s2, rr, cc, ker.k3));
        e3 = 10.f * sgnf(a - t);
      }
      {
        const float a = lap_res<5, IP>(s0, rr, cc, ker.k5);
        float t = lap_res<5, IP>(s1, rr, cc, ker.k5);
        if (NIMG == 3) t = fmaxf(t, lap_res<5, IP>(s2, rr, cc, ker.k5));
        e5 = 10.f * sgnf(a - t);
      }
      {
        const float a = lap_res<7, IP>(s0, rr, cc, ker.k7);
        float t = lap_res<7, IP>(s1, rr, cc, ker.k7);
        if (NIMG == 3) t = fmaxf(t, lap_res<7, IP>(s2, rr, cc, ker.k7));
        e7 = sgnf(a - t);
      }
    }
    e[0][r * EP + c] = e3; e[1][r * EP + c] = e5; e[2][r * EP + c] = e7;
  }
  __syncthreads();
  const float gs = gout[0] * inv_n;
  for (int i = threadIdx.x; i < T * T; i += 256) {
    const int r = i / T, c = i % T;
    const int y = by + r, x = bx + c;
    if (y >= H || x >= W) continue;
    // adjoint of (I - G_k) applied to e_k (G_k symmetric): the residual operator again
    const float v = lap_res<3, EP>(e[0], r + R, c + R, ker.k3) + lap_res<5, EP>(e[1], r + R, c + R, ker.k5) +
                    lap_res<7, EP>(e[2], r + R, c + R, ker.k7);
    const int64_t o = off + (int64_t)y * W + x;
    dinp[o] = accumulate ? dinp[o] + gs * v : gs * v;
  }
}

// ------------------------------------------------------------------------------------------------ Entropy
struct Bins32b { float b[32]; };

template <int P>
__global__ void __launch_bounds__(256) entropy_bwd_kernel(const float* __restrict__ img, int B, int H, int W,
                                                          Bins32b bins, const float* __restrict__ gout,
                                                          float* __restrict__ dimg, int accumulate) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int segs_x = (W + 31) / 32;
  const int prow = H / P;
  const int64_t nseg = (int64_t)B * prow * segs_x;
  const float mybin = bins.b[lane];
  const float inv_sigma = 1.0f / 0.01f;
  const float nhl2e = -0.5f * 1.4426950408889634f;
  const float g = gout[0];
  for (int64_t sidx = (int64_t)blockIdx.x * nwarp + warp; sidx < nseg; sidx += (int64_t)gridDim.x * nwarp) {
    const int sx = (int)(sidx % segs_x);
    const int py = (int)((sidx / segs_x) % prow);
    const int64_t b = sidx / ((int64_t)segs_x * prow);
    const int x = sx * 32 + lane;
    float v[P], dv[P];
#pragma unroll
    for (int dy = 0; dy < P; ++dy) {
      v[dy] = x < W ? img[(b * H + (int64_t)py * P + dy) * W + x] : 0.f;
      dv[dy] = 0.f;
    }
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
          acc += exp2f(nhl2e * (r * r));
        }
      // forward: q_k = acc / P^2, S = sum_k q_k, p_k = q_k / (S + eps) + eps, H = -sum p log p
      const float qk = acc / (float)(P * P);
      const float norm = warp_sum(qk) + 1e-40f;
      const float pk = qk / norm + 1e-40f;
      const float dHdp = -(logf(pk) + 1.f);
      // lambda_k = dH/dq_k = dHdp_k / norm - (sum_j dHdp_j q_j) / norm^2
      const float lam = (dHdp - warp_sum(dHdp * qk) / norm) / norm;
      const float coef = lam * (-inv_sigma) / (float)(P * P);
#pragma unroll
      for (int dy = 0; dy < P; ++dy)
#pragma unroll
        for (int dx = 0; dx < P; ++dx) {
          const float val = __shfl_sync(0xffffffffu, v[dy], q * P + dx);
          const float r = (val - mybin) * inv_sigma;
          // d/dv exp(-r^2/2) = exp(-r^2/2) * (-r) / sigma
          const float t = warp_sum(coef * r * exp2f(nhl2e * (r * r)));
          if (lane == q * P + dx) dv[dy] = t;
        }
    }
    if (x < W) {
#pragma unroll
      for (int dy = 0; dy < P; ++dy) {
        const int64_t o = (b * H + (int64_t)py * P + dy) * W + x;
        dimg[o] = accumulate ? dimg[o] + g * dv[dy] : g * dv[dy];
      }
    }
  }
}

}  // namespace segmif

using namespace segmif;

static Gauss11b make_gauss11b() {
  Gauss11b w;
  float tmp[11], s = 0.f;
  for (int i = 0; i < 11; ++i) tmp[i] = (float)exp(-(double)((i - 5) * (i - 5)) / (2.0 * 1.5 * 1.5));
  for (int i = 0; i < 11; ++i) s += tmp[i];
  for (int i = 0; i < 11; ++i) w.g[i] = tmp[i] / s;
  return w;
}

static LapKernelsB make_lap_kernels_b() {
  LapKernelsB k;
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

extern "C" int segmif_mse_l1_bwd(const float* x, const float* y, int64_t n, const float* gout2, float* dx,
                                 int accumulate, segmif_stream_t stream) {
  SEGMIF_REQUIRE(x && y && gout2 && dx && n > 0, "mse_l1_bwd: bad arguments");
  const int nblocks = (int)std::min<int64_t>(ceil_div(n, 256), 148 * 8);
  mse_l1_bwd_kernel<<<nblocks, 256, 0, as_stream(stream)>>>(x, y, n, gout2, (float)(1.0 / (double)n), dx, accumulate);
  return check_launch("segmif_mse_l1_bwd");
}

extern "C" int segmif_sobel_l1_bwd(const float* x, const float* y, int B, int H, int W, const float* gout2, float* dx,
                                   int accumulate, segmif_stream_t stream) {
  SEGMIF_REQUIRE(x && y && gout2 && dx, "sobel_l1_bwd: null pointer");
  SEGMIF_REQUIRE(B > 0 && H > 0 && W > 0, "sobel_l1_bwd: empty input");
  dim3 grid((W + 31) / 32, (H + 31) / 32, B);
  sobel_l1_bwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, y, H, W, gout2, (float)(1.0 / ((double)B * H * W)), dx, accumulate);
  return check_launch("segmif_sobel_l1_bwd");
}

extern "C" int segmif_ssim_bwd(const float* img1, const float* img2, int B, int H, int W, int per_image,
                               const float* gout, float* dimg1, int accumulate, segmif_stream_t stream) {
  SEGMIF_REQUIRE(img1 && img2 && gout && dimg1, "ssim_bwd: null pointer");
  SEGMIF_REQUIRE(B > 0 && H > 0 && W > 0, "ssim_bwd: empty input");
  static const Gauss11b win = make_gauss11b();
  const size_t smem = (size_t)(2 * 52 * 53 + 5 * 52 * 43 + 3 * 42 * 43) * sizeof(float);
  static bool cfg = false;
  if (!cfg) {
    cfg = true;
    cudaError_t e = cudaFuncSetAttribute((const void*)ssim_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("ssim_bwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return SEGMIF_ERR_CUDA; }
  }
  dim3 grid((W + 31) / 32, (H + 31) / 32, B);
  const double n = per_image ? (double)H * W : (double)B * H * W;
  ssim_bwd_kernel<<<grid, 256, smem, as_stream(stream)>>>(img1, img2, H, W, win, gout, per_image, (float)(1.0 / n), dimg1, accumulate);
  return check_launch("segmif_ssim_bwd");
}

extern "C" int segmif_laploss2_bwd(const float* inp, const float* ir, const float* vis, int B, int H, int W,
                                   const float* gout, float* dinp, int accumulate, segmif_stream_t stream) {
  SEGMIF_REQUIRE(inp && ir && vis && gout && dinp, "laploss2_bwd: null pointer");
  SEGMIF_REQUIRE(B > 0 && H > 0 && W > 0, "laploss2_bwd: empty input");
  static const LapKernelsB ker = make_lap_kernels_b();
  dim3 grid((W + 31) / 32, (H + 31) / 32, B);
  laploss_bwd_kernel<3><<<grid, 256, 0, as_stream(stream)>>>(inp, ir, vis, H, W, ker, gout, (float)(1.0 / ((double)B * H * W)), dinp, accumulate);
  return check_launch("segmif_laploss2_bwd");
}

extern "C" int segmif_laploss_bwd(const float* inp, const float* target, int B, int H, int W, const float* gout,
                                  float* dinp, int accumulate, segmif_stream_t stream) {
  SEGMIF_REQUIRE(inp && target && gout && dinp, "laploss_bwd: null pointer");
  SEGMIF_REQUIRE(B > 0 && H > 0 && W > 0, "laploss_bwd: empty input");
  static const LapKernelsB ker = make_lap_kernels_b();
  dim3 grid((W + 31) / 32, (H + 31) / 32, B);
  laploss_bwd_kernel<2><<<grid, 256, 0, as_stream(stream)>>>(inp, target, nullptr, H, W, ker, gout, (float)(1.0 / ((double)B * H * W)), dinp, accumulate);
  return check_launch("segmif_laploss_bwd");
}

extern "C" int segmif_entropy_bwd(const float* img, int B, int H, int W, int patch, const float* gout, float* dimg,
                                  int accumulate, segmif_stream_t stream) {
  SEGMIF_REQUIRE(img && gout && dimg, "entropy_bwd: null pointer");
  SEGMIF_REQUIRE(patch == 2 || patch == 4 || patch == 8 || patch == 16, "entropy_bwd: patch size %d unsupported (2,4,8,16)", patch);
  SEGMIF_REQUIRE(H % patch == 0 && W % patch == 0 && B > 0, "entropy_bwd: H and W must be multiples of the patch size");
  Bins32b bins;
  const float step = 1.0f / 31.0f;
  for (int i = 0; i < 32; ++i) bins.b[i] = i < 16 ? 0.0f + step * (float)i : 1.0f - step * (float)(31 - i);
  const int nblocks = 148 * 8;
  cudaStream_t st = as_stream(stream);
  switch (patch) {
    case 2: entropy_bwd_kernel<2><<<nblocks, 256, 0, st>>>(img, B, H, W, bins, gout, dimg, accumulate); break;
    case 4: entropy_bwd_kernel<4><<<nblocks, 256, 0, st>>>(img, B, H, W, bins, gout, dimg, accumulate); break;
    case 8: entropy_bwd_kernel<8><<<nblocks, 256, 0, st>>>(img, B, H, W, bins, gout, dimg, accumulate); break;
    default: entropy_bwd_kernel<16><<<nblocks, 256, 0, st>>>(img, B, H, W, bins, gout, dimg, accumulate); break;
  }
  return check_launch("segmif_entropy_bwd");
}

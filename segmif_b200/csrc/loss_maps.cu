// Map-valued pieces of the fusion-loss composites that the live training loop never builds but core/loss.py exports
// (Fusionloss, Fusionloss4, Fusionloss_add, new_loss_sobel / Total_fusion_loss[2], IQALoss -- core/loss.py:389-633):
// the Sobel gradient-magnitude MAP (Sobelxy.forward, core/loss.py:634-650) with its backward, and the few element-wise
// combinations those composites apply to maps before a reduction (max, a*x + b*y, |a + b*x|, x*y).  HBM-bound fp32.
#include "common.cuh"

namespace segmif {

// cross-correlation with kx = [[-1,0,1],[-2,0,2],[-1,0,1]], ky = [[1,2,1],[0,0,0],[-1,-2,-1]], zero padding
__device__ __forceinline__ float px(const float* __restrict__ img, int H, int W, int y, int x) {
  return ((unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W) ? img[(int64_t)y * W + x] : 0.f;
}
__device__ __forceinline__ void sobel_at(const float* __restrict__ img, int H, int W, int y, int x, float& gx, float& gy) {
  const float a = px(img, H, W, y - 1, x - 1), b = px(img, H, W, y - 1, x), c = px(img, H, W, y - 1, x + 1);
  const float d = px(img, H, W, y, x - 1), f = px(img, H, W, y, x + 1);
  const float g = px(img, H, W, y + 1, x - 1), h = px(img, H, W, y + 1, x), i = px(img, H, W, y + 1, x + 1);
  gx = (c - a) + 2.f * (f - d) + (i - g);
  gy = (a - g) + 2.f * (b - h) + (c - i);
}

__global__ void __launch_bounds__(256) sobel_map_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int H, int W) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * H * W) return;
  const int xx = (int)(i % W), yy = (int)((i / W) % H);
  const int64_t b = i / ((int64_t)W * H);
  float gx, gy;
  sobel_at(x + b * H * W, H, W, yy, xx, gx, gy);
  out[i] = fabsf(gx) + fabsf(gy);
}

__device__ __forceinline__ float sgn(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }

// dx[p] = sum over the 3x3 neighbours q of dout[q] * (sign(Gx[q]) kx[p - q] + sign(Gy[q]) ky[p - q])
__global__ void __launch_bounds__(256) sobel_map_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dout,
                                                            float* __restrict__ dx, int B, int H, int W, int accumulate) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * H * W) return;
  const int xx = (int)(i % W), yy = (int)((i / W) % H);
  const int64_t b = i / ((int64_t)W * H);
  const float* img = x + b * H * W;
  const float* go = dout + b * H * W;
  const float kx[3][3] = {{-1.f, 0.f, 1.f}, {-2.f, 0.f, 2.f}, {-1.f, 0.f, 1.f}};
  const float ky[3][3] = {{1.f, 2.f, 1.f}, {0.f, 0.f, 0.f}, {-1.f, -2.f, -1.f}};
  float acc = 0.f;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
    for (int dxo = -1; dxo <= 1; ++dxo) {
      const int qy = yy + dy, qx = xx + dxo;
      if ((unsigned)qy >= (unsigned)H || (unsigned)qx >= (unsigned)W) continue;
      float gx, gy;
      sobel_at(img, H, W, qy, qx, gx, gy);
      // x[p] enters G[q] with kernel index (1 + py - qy, 1 + px - qx) = (1 - dy, 1 - dxo)
      acc += go[(int64_t)qy * W + qx] * (sgn(gx) * kx[1 - dy][1 - dxo] + sgn(gy) * ky[1 - dy][1 - dxo]);
    }
  dx[i] = accumulate ? dx[i] + acc : acc;
}

// mode 0: a*x + b*y   1: max(x, y)   2: |a + b*x|   3: x*y   4: b * sign(a + b*x) * y  (backward of mode 2, y = upstream)
__global__ void __launch_bounds__(256) ew2_kernel(const float* __restrict__ x, const float* __restrict__ y, float a, float b, int mode,
                                                  float* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float xv = x[i], yv = y ? y[i] : 0.f;
  float r;
  switch (mode) {
    case 0: r = __fadd_rn(__fmul_rn(a, xv), __fmul_rn(b, yv)); break;     // as torch evaluates a*x + b*y: no FMA contraction
    case 1: r = fmaxf(xv, yv); break;
    case 2: r = fabsf(a + b * xv); break;
    case 3: r = xv * yv; break;
    default: r = b * sgn(a + b * xv) * yv; break;
  }
  out[i] = r;
}

}  // namespace segmif

using namespace segmif;

extern "C" int segmif_sobel_map_fwd(const float* x, float* out, int B, int H, int W, segmif_stream_t stream) {
  SEGMIF_REQUIRE(x && out, "sobel_map: null pointer");
  const int64_t n = (int64_t)B * H * W;
  if (n == 0) return SEGMIF_OK;
  sobel_map_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(x, out, B, H, W);
  return check_launch("segmif_sobel_map_fwd");
}

extern "C" int segmif_sobel_map_bwd(const float* x, const float* dout, float* dx, int B, int H, int W, int accumulate,
                                    segmif_stream_t stream) {
  SEGMIF_REQUIRE(x && dout && dx, "sobel_map_bwd: null pointer");
  const int64_t n = (int64_t)B * H * W;
  if (n == 0) return SEGMIF_OK;
  sobel_map_bwd_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(x, dout, dx, B, H, W, accumulate);
  return check_launch("segmif_sobel_map_bwd");
}

extern "C" int segmif_ew2(const float* x, const float* y, float a, float b, int mode, float* out, int64_t n,
                          segmif_stream_t stream) {
  SEGMIF_REQUIRE(x && out && mode >= 0 && mode <= 4, "ew2: bad arguments");
  SEGMIF_REQUIRE(y || mode == 2, "ew2: this mode needs two inputs");
  if (n == 0) return SEGMIF_OK;
  ew2_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(x, y, a, b, mode, out, n);
  return check_launch("segmif_ew2");
}

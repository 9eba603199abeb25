// fp32 companions of the strict-precision mode (see gemm_split_tc.cu): everything between two tensor-core contractions
// that the default path keeps in bf16 -- operand splitting, im2col for the strided convolutions, the attention core,
// depthwise 3x3 + GELU, the Gram / context pass of the hierarchical interactive attention -- evaluated on fp32 tensors
// with fp32 (Gram: fp64) accumulation, so that the whole inference path reproduces the reference's fp32 numbers to
// rounding-order noise.  All HBM-bound or small; plain coalesced SIMT kernels.
#include <math.h>

#include "common.cuh"

namespace segmif {

__device__ __forceinline__ void split3f(float v, bf16& h, bf16& m, bf16& l) {
  h = __float2bfloat16_rn(v);
  const float r1 = v - __bfloat162float(h);
  m = __float2bfloat16_rn(r1);
  const float r2 = r1 - __bfloat162float(m);
  l = __float2bfloat16_rn(r2);
}
__device__ __forceinline__ uint2 pack4(bf16 a, bf16 b, bf16 c, bf16 d) {
  uint2 u;
  u.x = (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
  u.y = (uint32_t)__bfloat16_as_ushort(c) | ((uint32_t)__bfloat16_as_ushort(d) << 16);
  return u;
}
__device__ __forceinline__ void store_split4(bf16* base, int64_t plane, const float4 v) {
  bf16 h[4], m[4], l[4];
  split3f(v.x, h[0], m[0], l[0]); split3f(v.y, h[1], m[1], l[1]);
  split3f(v.z, h[2], m[2], l[2]); split3f(v.w, h[3], m[3], l[3]);
  *reinterpret_cast<uint2*>(base) = pack4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint2*>(base + plane) = pack4(m[0], m[1], m[2], m[3]);
  *reinterpret_cast<uint2*>(base + 2 * plane) = pack4(l[0], l[1], l[2], l[3]);
}

// ---- x fp32 [rows, ld_x] slice -> three bf16 planes (optionally ReLU first, optionally also the fp32 result) ------
__global__ void __launch_bounds__(256) split3_kernel(const float* __restrict__ x, int ld_x, int coff_x, int64_t rows, int C4,
                                                     int relu, float* __restrict__ y, int ld_y, int coff_y,
                                                     bf16* __restrict__ planes, int ld_p, int coff_p, int64_t plane) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * C4) return;
  const int64_t r = i / C4;
  const int c = (int)(i - r * C4) * 4;
  float4 v = *reinterpret_cast<const float4*>(x + r * ld_x + coff_x + c);
  if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
  if (y) *reinterpret_cast<float4*>(y + r * ld_y + coff_y + c) = v;
  if (planes) store_split4(planes + r * ld_p + coff_p + c, plane, v);
}

// ---- im2col of a pixel-major fp32 tensor straight into split planes: column (ky*k + kx)*C + c ----------------------
__global__ void __launch_bounds__(256) im2col_split3_kernel(const float* __restrict__ x, int ld_x, int coff_x, int B, int H, int W,
                                                            int C4, int k, int stride, int pad, int Ho, int Wo,
                                                            bf16* __restrict__ planes, int64_t plane) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)B * Ho * Wo * k * k * C4;
  if (i >= total) return;
  const int c = (int)(i % C4) * 4;
  int64_t t = i / C4;
  const int tap = (int)(t % (k * k));
  t /= k * k;
  const int ox = (int)(t % Wo);
  t /= Wo;
  const int oy = (int)(t % Ho);
  const int64_t b = t / Ho;
  const int iy = oy * stride - pad + tap / k, ix = ox * stride - pad + tap % k;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if ((unsigned)iy < (unsigned)H && (unsigned)ix < (unsigned)W)
    v = *reinterpret_cast<const float4*>(x + ((b * H + iy) * W + ix) * ld_x + coff_x + c);
  const int64_t row = (b * Ho + oy) * Wo + ox;
  const int Kc = k * k * C4 * 4;
  store_split4(planes + row * Kc + tap * (C4 * 4) + c, plane, v);
}

// ---- attention core in fp32: softmax(q k^T * scale) v, one query row per thread, K/V tiles broadcast from smem --------
template <int D>
__global__ void __launch_bounds__(128) sr_attention_f32_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ k,
                                                               const float* __restrict__ v, int ldkv, float* __restrict__ out,
                                                               int ldo, int N, int Nk, float scale) {
  constexpr int KT = 32;
  __shared__ __align__(16) float sK[KT][D];
  __shared__ __align__(16) float sV[KT][D];
  const int b = blockIdx.z, h = blockIdx.y;
  const int n = blockIdx.x * 128 + threadIdx.x;
  const bool live = n < N;
  float qr[D], acc[D];
  const float* qp = q + ((int64_t)b * N + (live ? n : 0)) * ldq + h * D;
#pragma unroll
  for (int d = 0; d < D; d += 4) {
    const float4 t = *reinterpret_cast<const float4*>(qp + d);
    qr[d] = t.x; qr[d + 1] = t.y; qr[d + 2] = t.z; qr[d + 3] = t.w;
    acc[d] = acc[d + 1] = acc[d + 2] = acc[d + 3] = 0.f;
  }
  float mx = -INFINITY, l = 0.f;
  for (int j0 = 0; j0 < Nk; j0 += KT) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < KT * (D / 4); idx += 128) {
      const int jj = idx / (D / 4), d4 = (idx % (D / 4)) * 4;
      float4 kv4 = make_float4(0.f, 0.f, 0.f, 0.f), vv4 = kv4;
      if (j0 + jj < Nk) {
        const int64_t o = ((int64_t)b * Nk + j0 + jj) * ldkv + h * D + d4;
        kv4 = *reinterpret_cast<const float4*>(k + o);
        vv4 = *reinterpret_cast<const float4*>(v + o);
      }
      *reinterpret_cast<float4*>(&sK[jj][d4]) = kv4;
      *reinterpret_cast<float4*>(&sV[jj][d4]) = vv4;
    }
    __syncthreads();
    const int jn = Nk - j0 < KT ? Nk - j0 : KT;
    for (int jj = 0; jj < jn; ++jj) {
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int d = 0; d < D; d += 4) {
        const float4 kk = *reinterpret_cast<const float4*>(&sK[jj][d]);
        s0 = fmaf(qr[d], kk.x, s0); s1 = fmaf(qr[d + 1], kk.y, s1);
        s2 = fmaf(qr[d + 2], kk.z, s2); s3 = fmaf(qr[d + 3], kk.w, s3);
      }
      const float s = ((s0 + s1) + (s2 + s3)) * scale;
      if (s > mx) {
        const float f = expf(mx - s);          // first key: exp(-inf) = 0
        l *= f;
#pragma unroll
        for (int d = 0; d < D; ++d) acc[d] *= f;
        mx = s;
      }
      const float p = expf(s - mx);
      l += p;
#pragma unroll
      for (int d = 0; d < D; d += 4) {
        const float4 vv = *reinterpret_cast<const float4*>(&sV[jj][d]);
        acc[d] = fmaf(p, vv.x, acc[d]); acc[d + 1] = fmaf(p, vv.y, acc[d + 1]);
        acc[d + 2] = fmaf(p, vv.z, acc[d + 2]); acc[d + 3] = fmaf(p, vv.w, acc[d + 3]);
      }
    }
  }
  if (live) {
    const float inv = 1.f / l;
    float* op = out + ((int64_t)b * N + n) * ldo + h * D;
#pragma unroll
    for (int d = 0; d < D; d += 4)
      *reinterpret_cast<float4*>(op + d) = make_float4(acc[d] * inv, acc[d + 1] * inv, acc[d + 2] * inv, acc[d + 3] * inv);
  }
}

// ---- depthwise 3x3 (pad 1, bias) + GELU(erf), fp32 pixel-major -------------------------------------------------------
__global__ void __launch_bounds__(256) dwconv3x3_gelu_f32_kernel(const float* __restrict__ x, const float* __restrict__ w9c,
                                                                 const float* __restrict__ bias, float* __restrict__ y, int B, int H,
                                                                 int W, int C4, int gelu) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * H * W * C4) return;
  const int c = (int)(i % C4) * 4;
  int64_t t = i / C4;
  const int px = (int)(t % W);
  t /= W;
  const int py = (int)(t % H);
  const int64_t b = t / H;
  const int C = C4 * 4;
  float4 a = *reinterpret_cast<const float4*>(bias + c);
  // accumulate in the order of a direct convolution: taps row-major
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = py + ky - 1;
    if ((unsigned)iy >= (unsigned)H) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = px + kx - 1;
      if ((unsigned)ix >= (unsigned)W) continue;
      const float4 v = *reinterpret_cast<const float4*>(x + ((b * H + iy) * W + ix) * C + c);
      const float4 wv = *reinterpret_cast<const float4*>(w9c + (ky * 3 + kx) * C + c);
      a.x = fmaf(v.x, wv.x, a.x); a.y = fmaf(v.y, wv.y, a.y); a.z = fmaf(v.z, wv.z, a.z); a.w = fmaf(v.w, wv.w, a.w);
    }
  }
  if (gelu) { a.x = gelu_erf(a.x); a.y = gelu_erf(a.y); a.z = gelu_erf(a.z); a.w = gelu_erf(a.w); }
  *reinterpret_cast<float4*>(y + ((b * H + py) * W + px) * C + c) = a;
}

// ---- Gram matrix of a 64-channel fp32 slice over the pixels of one image chunk, fp64 accumulation -----------------
// partials double [B, nchunk, 64, 64]; thread (ty, tx) of a 16 x 16 CTA owns the 4 x 4 block (4 ty.., 4 tx..).
__global__ void __launch_bounds__(256) gram64_f64_kernel(const float* __restrict__ p, int ld, int coff, int64_t HW, int relu,
                                                         double* __restrict__ partials, int nchunk) {
  __shared__ __align__(16) float tile[32][64];
  const int chunk = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int64_t per = (HW + nchunk - 1) / nchunk;
  const int64_t p0 = chunk * per, p1 = p0 + per < HW ? p0 + per : HW;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  for (int64_t base = p0; base < p1; base += 32) {
    __syncthreads();
    for (int idx = tid; idx < 32 * 16; idx += 256) {
      const int r = idx >> 4, c4 = (idx & 15) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (base + r < p1) v = *reinterpret_cast<const float4*>(p + ((int64_t)b * HW + base + r) * ld + coff + c4);
      if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      *reinterpret_cast<float4*>(&tile[r][c4]) = v;
    }
    __syncthreads();
    float f[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) f[i][j] = 0.f;
#pragma unroll 8
    for (int r = 0; r < 32; ++r) {
      const float4 a = *reinterpret_cast<const float4*>(&tile[r][ty * 4]);
      const float4 c = *reinterpret_cast<const float4*>(&tile[r][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, cv[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) f[i][j] = fmaf(av[i], cv[j], f[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] += (double)f[i][j];
  }
  double* o = partials + ((int64_t)b * nchunk + chunk) * 4096;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[(ty * 4 + i) * 64 + tx * 4 + j] = acc[i][j];
}

// ---- contexts + folded matrices from the Gram partials, fp64 inside (one CTA per (stream, image)) -------------------
// ctx[b][s][h][i][j] = softmax_i( scale * sum_c (Wk G_s)[h8+i, c] Wv[h8+j, c] ),  G_s = sum of the partials of stream s.
__global__ void __launch_bounds__(256) ffm_ctx_f64_kernel(const double* __restrict__ partials, int nchunk, int64_t stream_stride,
                                                          int64_t batch_stride, const float* __restrict__ wkv,
                                                          float* __restrict__ ctx_out) {
  extern __shared__ double sm[];
  double* G = sm;            // 4096
  double* T = sm + 4096;     // 4096
  double* lg = sm + 8192;    // 512
  const int s = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const double* pb = partials + (int64_t)s * stream_stride + (int64_t)b * batch_stride;
  for (int idx = tid; idx < 4096; idx += 256) {
    double a = 0.0;
    for (int c = 0; c < nchunk; ++c) a += pb[(int64_t)c * 4096 + idx];
    G[idx] = a;
  }
  __syncthreads();
  const float* Wk = wkv + (int64_t)s * 128 * 64;
  const float* Wv = Wk + 64 * 64;
  for (int idx = tid; idx < 4096; idx += 256) {
    const int r = idx >> 6, c = idx & 63;
    double a = 0.0;
    for (int k = 0; k < 64; ++k) a += (double)Wk[r * 64 + k] * G[k * 64 + c];
    T[idx] = a;
  }
  __syncthreads();
  const double scale = 0.35355339059327379;            // float32(8 ** -0.5) is applied by the reference; see below
  for (int idx = tid; idx < 512; idx += 256) {
    const int h = idx >> 6, i = (idx >> 3) & 7, j = idx & 7;
    double a = 0.0;
    for (int c = 0; c < 64; ++c) a += T[(h * 8 + i) * 64 + c] * (double)Wv[(h * 8 + j) * 64 + c];
    lg[idx] = a * (double)(float)scale;
  }
  __syncthreads();
  if (tid < 64) {
    const int h = tid >> 3, j = tid & 7;
    double* c = lg + h * 64 + j;
    double m = -INFINITY;
    for (int i = 0; i < 8; ++i) m = fmax(m, c[i * 8]);
    double ev[8], sum = 0.0;
    for (int i = 0; i < 8; ++i) { ev[i] = exp(c[i * 8] - m); sum += ev[i]; }
    for (int i = 0; i < 8; ++i) c[i * 8] = ev[i] / sum;
  }
  __syncthreads();
  for (int idx = tid; idx < 512; idx += 256) ctx_out[((int64_t)b * 3 + s) * 512 + idx] = (float)lg[idx];
}

// folded[b][m][o][c] fp32: m=0 Mz1 (ctx1, We1[:, :64]); 1 Mv1 (ctx3, We1[:, 64:]); 2 Mz2 (ctx2, We2[:, :64]); 3 Mv2 (ctx3, We2[:, 64:])
__global__ void __launch_bounds__(256) ffm_fold_f32_kernel(const float* __restrict__ ctx, const float* __restrict__ wend,
                                                           float* __restrict__ folded) {
  const int m = blockIdx.x, b = blockIdx.y;
  const int stream = m >> 1, is_v = m & 1;
  const float* cb = ctx + ((int64_t)b * 3 + (is_v ? 2 : stream)) * 512;
  for (int idx = threadIdx.x; idx < 4096; idx += 256) {
    const int o = idx >> 6, c = idx & 63;
    const int h = c >> 3, i = c & 7;
    const float* cx = cb + h * 64 + i * 8;
    const float* we = wend + (int64_t)stream * 64 * 128 + o * 128 + (is_v ? 64 : 0) + h * 8;
    double a = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) a += (double)cx[j] * (double)we[j];
    folded[((int64_t)b * 4 + m) * 4096 + idx] = (float)a;
  }
}

// ---- generic cross product  X^T Y  over the pixels of each image (Cx, Cy <= 64, multiples of 4), fp64 accumulation ------
// The ablation networks' attention modules (core/model_fusion.py:363-425, dim 32, head_dim 4) need k^T v for k, v of any
// width; partials double [B, nchunk, Cx, Cy].  One thread per output element group: thread t owns row i = t / (Cy/4),
// columns 4 (t % (Cy/4)) .. +3.
__global__ void __launch_bounds__(256) xty_f64_kernel(const float* __restrict__ x, int ldx, int coffx, int Cx,
                                                      const float* __restrict__ y, int ldy, int coffy, int Cy, int64_t HW,
                                                      double* __restrict__ partials, int nchunk) {
  __shared__ __align__(16) float sx[32][64];
  __shared__ __align__(16) float sy[32][64];
  const int chunk = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int cy4 = Cy >> 2;
  const bool active = tid < Cx * cy4;
  const int i = active ? tid / cy4 : 0, j4 = active ? (tid % cy4) * 4 : 0;
  const int64_t per = (HW + nchunk - 1) / nchunk;
  const int64_t p0 = chunk * per, p1 = p0 + per < HW ? p0 + per : HW;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int64_t base = p0; base < p1; base += 32) {
    __syncthreads();
    for (int idx = tid; idx < 32 * 16; idx += 256) {
      const int r = idx >> 4, c4 = (idx & 15) * 4;
      const bool row_ok = base + r < p1;
      float4 vx = make_float4(0.f, 0.f, 0.f, 0.f), vy = vx;
      if (row_ok && c4 < Cx) vx = *reinterpret_cast<const float4*>(x + ((int64_t)b * HW + base + r) * ldx + coffx + c4);
      if (row_ok && c4 < Cy) vy = *reinterpret_cast<const float4*>(y + ((int64_t)b * HW + base + r) * ldy + coffy + c4);
      *reinterpret_cast<float4*>(&sx[r][c4]) = vx;
      *reinterpret_cast<float4*>(&sy[r][c4]) = vy;
    }
    __syncthreads();
    float f[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
    for (int r = 0; r < 32; ++r) {
      const float a = sx[r][i];
      const float4 c = *reinterpret_cast<const float4*>(&sy[r][j4]);
      f[0] = fmaf(a, c.x, f[0]); f[1] = fmaf(a, c.y, f[1]); f[2] = fmaf(a, c.z, f[2]); f[3] = fmaf(a, c.w, f[3]);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] += (double)f[k];
  }
  if (active) {
    double* o = partials + (((int64_t)b * nchunk + chunk) * Cx + i) * Cy + j4;
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = acc[k];
  }
}

// ctx[b][h][i][j] = softmax over i of (scale * sum_chunks partial[b][.][h d + i][h d + j]);  also the block-diagonal weight
// wout[b][h d + j][h d + i] = ctx[b][h][i][j] that turns "q @ ctx per head" into one [C x C] product per image.
__global__ void __launch_bounds__(256) ctx_blockdiag_kernel(const double* __restrict__ partials, int nchunk, int C, int heads,
                                                            float scale, float* __restrict__ ctx, float* __restrict__ wout) {
  __shared__ double lg[64 * 8];           // heads * d * d <= 64 * 8 (C <= 64, d <= 8)
  const int b = blockIdx.x, tid = threadIdx.x;
  const int d = C / heads, per_head = d * d, n = heads * per_head;
  for (int idx = tid; idx < n; idx += 256) {
    const int h = idx / per_head, i = (idx % per_head) / d, j = idx % d;
    double a = 0.0;
    for (int c = 0; c < nchunk; ++c) a += partials[(((int64_t)b * nchunk + c) * C + h * d + i) * C + h * d + j];
    lg[idx] = (double)((float)a) * (double)scale;       // the reference forms k^T v in fp32, then multiplies by the fp32 scale
  }
  for (int idx = tid; idx < C * C; idx += 256) wout[(int64_t)b * C * C + idx] = 0.f;
  __syncthreads();
  if (tid < heads * d) {                                // one (head, column j): softmax over i (dim = -2)
    const int h = tid / d, j = tid % d;
    double m = -INFINITY;
    for (int i = 0; i < d; ++i) m = fmax(m, lg[h * per_head + i * d + j]);
    double s = 0.0;
    for (int i = 0; i < d; ++i) s += exp(lg[h * per_head + i * d + j] - m);
    for (int i = 0; i < d; ++i) {
      const float p = (float)(exp(lg[h * per_head + i * d + j] - m) / s);
      ctx[(int64_t)b * n + h * per_head + i * d + j] = p;
      wout[(int64_t)b * C * C + (h * d + j) * C + h * d + i] = p;
    }
  }
}

// out = x * sigmoid(x)  (AttentionModule, core/model_fusion.py:759-771: `out = sigmoid(x1); return out * x1`)
__global__ void __launch_bounds__(256) sigmoid_gate_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = x[i];
  out[i] = v * (1.0f / (1.0f + expf(-v)));
}

// ---- edge layers of the fusion network with fp32 activations ---------------------------------------------------------
// conv1_ir / conv1_vis: fp32 plane -> fp32 pixel-major [.., Cout]; one thread per (pixel, 4 channels)
__global__ void __launch_bounds__(256) conv3x3_in1_f32_kernel(const float* __restrict__ plane, int64_t bstride,
                                                              const float* __restrict__ w, const float* __restrict__ bias,
                                                              const float* __restrict__ alpha_p, float* __restrict__ dst, int ld_dst,
                                                              int dst_coff, int B, int H, int W, int C4) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * H * W * C4) return;
  const int c = (int)(i % C4) * 4;
  int64_t t = i / C4;
  const int px = (int)(t % W);
  t /= W;
  const int py = (int)(t % H);
  const int64_t b = t / H;
  const int C = C4 * 4;
  const float* src = plane + b * bstride;
  // cuDNN / ATen accumulate the taps and add the bias last; the order only moves the last ulp
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = py + ky - 1;
    if ((unsigned)iy >= (unsigned)H) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = px + kx - 1;
      if ((unsigned)ix >= (unsigned)W) continue;
      const float v = src[(int64_t)iy * W + ix];
      const float4 wv = *reinterpret_cast<const float4*>(w + (ky * 3 + kx) * C + c);
      a.x = fmaf(v, wv.x, a.x); a.y = fmaf(v, wv.y, a.y); a.z = fmaf(v, wv.z, a.z); a.w = fmaf(v, wv.w, a.w);
    }
  }
  const float4 bv = *reinterpret_cast<const float4*>(bias + c);
  const float alpha = *alpha_p;
  a.x += bv.x; a.y += bv.y; a.z += bv.z; a.w += bv.w;
  a.x = a.x >= 0.f ? a.x : alpha * a.x; a.y = a.y >= 0.f ? a.y : alpha * a.y;
  a.z = a.z >= 0.f ? a.z : alpha * a.z; a.w = a.w >= 0.f ? a.w : alpha * a.w;
  *reinterpret_cast<float4*>(dst + ((b * H + py) * W + px) * ld_dst + dst_coff + c) = a;
}

// conv22: fp32 pixel-major Cin channels -> fp32 plane; a quad of 4 lanes shares one pixel
__global__ void __launch_bounds__(256) conv3x3_out1_f32_kernel(const float* __restrict__ src, int ld_src, const float* __restrict__ w,
                                                               const float* __restrict__ bias, const float* __restrict__ alpha_p,
                                                               float* __restrict__ dst, int B, int H, int W, int Cin) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t pix = gid >> 2;
  const int q = (int)(gid & 3);
  const int64_t npix = (int64_t)B * H * W;
  const bool live = pix < npix;
  const int64_t pp = live ? pix : 0;
  const int X = (int)(pp % W), Y = (int)((pp / W) % H);
  const int64_t b = pp / ((int64_t)W * H);
  const int cpl = Cin >> 2;
  float acc = 0.f;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = Y + ky - 1;
    if ((unsigned)iy >= (unsigned)H) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = X + kx - 1;
      if ((unsigned)ix >= (unsigned)W) continue;
      const float* p = src + ((b * H + iy) * W + ix) * ld_src + q * cpl;
      const float* wp = w + (ky * 3 + kx) * Cin + q * cpl;
      for (int c = 0; c < cpl; c += 4) {
        const float4 v = *reinterpret_cast<const float4*>(p + c);
        const float4 wv = *reinterpret_cast<const float4*>(wp + c);
        acc = fmaf(v.x, wv.x, acc); acc = fmaf(v.y, wv.y, acc); acc = fmaf(v.z, wv.z, acc); acc = fmaf(v.w, wv.w, acc);
      }
    }
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  if (live && q == 0) {
    const float alpha = *alpha_p;
    const float v = acc + bias[0];
    dst[pix] = v >= 0.f ? v : alpha * v;
  }
}

}  // namespace segmif

using namespace segmif;

extern "C" int segmif_split3(const float* x, int ld_x, int coff_x, int64_t rows, int C, int relu, float* y, int ld_y, int coff_y,
                             void* planes, int ld_p, int coff_p, int64_t plane_stride, segmif_stream_t stream) {
  SEGMIF_REQUIRE(x && (y || planes), "split3: null pointer");
  SEGMIF_REQUIRE(C % 4 == 0 && ld_x % 4 == 0 && coff_x % 4 == 0 && ((uintptr_t)x & 15) == 0, "split3: C, ld_x, coff_x must be multiples of 4");
  SEGMIF_REQUIRE(!y || (ld_y % 4 == 0 && coff_y % 4 == 0 && ((uintptr_t)y & 15) == 0), "split3: fp32 output must be 16-byte aligned");
  SEGMIF_REQUIRE(!planes || (ld_p % 4 == 0 && coff_p % 4 == 0 && plane_stride % 4 == 0 && ((uintptr_t)planes & 7) == 0),
                 "split3: planes must be 8-byte aligned slices");
  const int64_t total = rows * (C / 4);
  if (total == 0) return SEGMIF_OK;
  split3_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, as_stream(stream)>>>(x, ld_x, coff_x, rows, C / 4, relu, y, ld_y, coff_y,
                                                                               (bf16*)planes, ld_p, coff_p, plane_stride);
  return check_launch("segmif_split3");
}

extern "C" int segmif_im2col_split3(const float* x, int ld_x, int coff_x, int B, int H, int W, int C, int k, int stride, int pad,
                                    void* planes, int64_t plane_stride, segmif_stream_t stream) {
  SEGMIF_REQUIRE(x && planes, "im2col_split3: null pointer");
  SEGMIF_REQUIRE(C % 4 == 0 && ld_x % 4 == 0 && coff_x % 4 == 0 && k > 0 && stride > 0 && pad >= 0 && plane_stride % 4 == 0,
                 "im2col_split3: C, ld_x, coff_x must be multiples of 4");
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  const int64_t total = (int64_t)B * Ho * Wo * k * k * (C / 4);
  if (total <= 0) return SEGMIF_OK;
  SEGMIF_REQUIRE(ceil_div(total, 256) < (1ll << 31), "im2col_split3: too large");
  im2col_split3_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, as_stream(stream)>>>(x, ld_x, coff_x, B, H, W, C / 4, k, stride, pad,
                                                                                      Ho, Wo, (bf16*)planes, plane_stride);
  return check_launch("segmif_im2col_split3");
}

extern "C" int segmif_sr_attention_f32_fwd(const float* q, int ldq, const float* k, const float* v, int ldkv, float* out, int ldo,
                                           int B, int heads, int N, int Nk, int D, float scale, segmif_stream_t stream) {
  SEGMIF_REQUIRE(q && k && v && out, "sr_attention_f32: null pointer");
  SEGMIF_REQUIRE(D == 32 || D == 64, "sr_attention_f32: head dim %d must be 32 or 64", D);
  SEGMIF_REQUIRE(ldq % 4 == 0 && ldkv % 4 == 0 && ldo % 4 == 0 && N > 0 && Nk > 0, "sr_attention_f32: pitches must be multiples of 4");
  SEGMIF_REQUIRE((((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out) & 15) == 0, "sr_attention_f32: pointers must be 16-byte aligned");
  dim3 grid((unsigned)ceil_div(N, 128), heads, B);
  if (D == 64)
    sr_attention_f32_kernel<64><<<grid, 128, 0, as_stream(stream)>>>(q, ldq, k, v, ldkv, out, ldo, N, Nk, scale);
  else
    sr_attention_f32_kernel<32><<<grid, 128, 0, as_stream(stream)>>>(q, ldq, k, v, ldkv, out, ldo, N, Nk, scale);
  return check_launch("segmif_sr_attention_f32_fwd");
}

extern "C" int segmif_dwconv3x3_f32_fwd(const float* x, const float* w9c, const float* bias, float* y, int B, int H, int W, int C,
                                        int gelu, segmif_stream_t stream) {
  SEGMIF_REQUIRE(x && w9c && bias && y, "dwconv_f32: null pointer");
  SEGMIF_REQUIRE(C % 4 == 0, "dwconv_f32: C=%d must be a multiple of 4", C);
  const int64_t total = (int64_t)B * H * W * (C / 4);
  if (total == 0) return SEGMIF_OK;
  dwconv3x3_gelu_f32_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, as_stream(stream)>>>(x, w9c, bias, y, B, H, W, C / 4, gelu);
  return check_launch("segmif_dwconv3x3_f32_fwd");
}

extern "C" int segmif_gram64_f64(const float* p, int ld, int coff, int B, int64_t HW, int relu, double* partials, int nchunk,
                                 segmif_stream_t stream) {
  SEGMIF_REQUIRE(p && partials && nchunk > 0 && B > 0, "gram64: bad arguments");
  SEGMIF_REQUIRE(ld % 4 == 0 && coff % 4 == 0 && ((uintptr_t)p & 15) == 0, "gram64: slice must be 16-byte aligned");
  gram64_f64_kernel<<<dim3(nchunk, B), 256, 0, as_stream(stream)>>>(p, ld, coff, HW, relu, partials, nchunk);
  return check_launch("segmif_gram64_f64");
}

extern "C" int segmif_ffm_ctx_f64_fwd(const double* partials, int nchunk, const float* wkv, const float* wend, float* folded,
                                      float* ctx_out, int B, segmif_stream_t stream) {
  SEGMIF_REQUIRE(partials && wkv && wend && folded && ctx_out && nchunk > 0 && B > 0, "ffm_ctx_f64: bad arguments");
  cudaStream_t st = as_stream(stream);
  const size_t smem = (size_t)(8192 + 512) * sizeof(double);
  static bool cfg = false;
  if (!cfg) {
    cudaError_t err = cudaFuncSetAttribute(ffm_ctx_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) { set_error("ffm_ctx_f64: cudaFuncSetAttribute: %s", cudaGetErrorString(err)); return SEGMIF_ERR_CUDA; }
    cfg = true;
  }
  // partials: double [3][B][nchunk][64][64]
  ffm_ctx_f64_kernel<<<dim3(3, B), 256, smem, st>>>(partials, nchunk, (int64_t)B * nchunk * 4096, (int64_t)nchunk * 4096, wkv, ctx_out);
  int rc = check_launch("segmif_ffm_ctx_f64_fwd");
  if (rc) return rc;
  ffm_fold_f32_kernel<<<dim3(4, B), 256, 0, st>>>(ctx_out, wend, folded);
  return check_launch("segmif_ffm_ctx_f64_fwd(fold)");
}

extern "C" int segmif_conv3x3_in1_f32_fwd(const float* plane, int64_t bstride, const float* w, const float* bias,
                                          const float* prelu_alpha, float* dst, int ld_dst, int dst_coff, int B, int H, int W,
                                          int Cout, segmif_stream_t stream) {
  SEGMIF_REQUIRE(plane && w && bias && prelu_alpha && dst, "conv3x3_in1_f32: null pointer");
  SEGMIF_REQUIRE(Cout % 4 == 0 && ld_dst % 4 == 0 && dst_coff % 4 == 0, "conv3x3_in1_f32: channel counts must be multiples of 4");
  const int64_t total = (int64_t)B * H * W * (Cout / 4);
  if (total == 0) return SEGMIF_OK;
  conv3x3_in1_f32_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, as_stream(stream)>>>(plane, bstride, w, bias, prelu_alpha, dst, ld_dst,
                                                                                        dst_coff, B, H, W, Cout / 4);
  return check_launch("segmif_conv3x3_in1_f32_fwd");
}

extern "C" int segmif_conv3x3_out1_f32_fwd(const float* src, int ld_src, const float* w, const float* bias, const float* prelu_alpha,
                                           float* dst, int B, int H, int W, int Cin, segmif_stream_t stream) {
  SEGMIF_REQUIRE(src && w && bias && prelu_alpha && dst, "conv3x3_out1_f32: null pointer");
  SEGMIF_REQUIRE(Cin % 16 == 0 && ld_src % 4 == 0, "conv3x3_out1_f32: Cin must be a multiple of 16");
  const int64_t total = (int64_t)B * H * W * 4;
  if (total == 0) return SEGMIF_OK;
  conv3x3_out1_f32_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, as_stream(stream)>>>(src, ld_src, w, bias, prelu_alpha, dst, B, H, W, Cin);
  return check_launch("segmif_conv3x3_out1_f32_fwd");
}

extern "C" int segmif_xty_f64(const float* x, int ldx, int coffx, int Cx, const float* y, int ldy, int coffy, int Cy, int B,
                              int64_t HW, double* partials, int nchunk, segmif_stream_t stream) {
  SEGMIF_REQUIRE(x && y && partials && nchunk > 0 && B > 0, "xty: bad arguments");
  SEGMIF_REQUIRE(Cx > 0 && Cy > 0 && Cx <= 64 && Cy <= 64 && Cx % 4 == 0 && Cy % 4 == 0 && Cx * (Cy / 4) <= 256,
                 "xty: Cx, Cy must be multiples of 4 with Cx * Cy <= 1024 (got %d x %d)", Cx, Cy);
  SEGMIF_REQUIRE(ldx % 4 == 0 && coffx % 4 == 0 && ldy % 4 == 0 && coffy % 4 == 0 && (((uintptr_t)x | (uintptr_t)y) & 15) == 0,
                 "xty: slices must be 16-byte aligned");
  xty_f64_kernel<<<dim3(nchunk, B), 256, 0, as_stream(stream)>>>(x, ldx, coffx, Cx, y, ldy, coffy, Cy, HW, partials, nchunk);
  return check_launch("segmif_xty_f64");
}

extern "C" int segmif_ctx_blockdiag(const double* partials, int nchunk, int C, int heads, float scale, float* ctx, float* wout,
                                    int B, segmif_stream_t stream) {
  SEGMIF_REQUIRE(partials && ctx && wout && nchunk > 0 && B > 0, "ctx_blockdiag: bad arguments");
  SEGMIF_REQUIRE(C > 0 && C <= 64 && heads > 0 && C % heads == 0 && C / heads <= 8, "ctx_blockdiag: C <= 64, head dim <= 8");
  ctx_blockdiag_kernel<<<B, 256, 0, as_stream(stream)>>>(partials, nchunk, C, heads, scale, ctx, wout);
  return check_launch("segmif_ctx_blockdiag");
}

extern "C" int segmif_sigmoid_gate(const float* x, float* out, int64_t n, segmif_stream_t stream) {
  SEGMIF_REQUIRE(x && out, "sigmoid_gate: null pointer");
  if (n == 0) return SEGMIF_OK;
  sigmoid_gate_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(x, out, n);
  return check_launch("segmif_sigmoid_gate");
}

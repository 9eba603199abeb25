// Depthwise 3x3 convolution of the Mix-FFN (core/mix_transformer.py:46-53,381-387) on TMA-staged halo tiles:
//   segmif_dwconv3x3_gelu_fwd   y = GELU(dwconv(x) + b)                      (Mlp.forward, inference and training forward)
//   segmif_dwconv3x3            y = dwconv(x) (+ b), optionally transposed taps (the data gradient)
//   segmif_dwconv3x3_gelu_bwd   dz = dy * GELU'(dwconv(x) + b), dW, db        (training)
// These are HBM-bound stencils (2-3 tensor passes, ~40 FLOP per element).  The first versions gathered the nine taps with
// per-thread global loads and sat at 1.2-1.3 TB/s: with ~0.5 KB in flight per warp the SMs could not cover the DRAM
// latency (Little's law wants ~50 KB per SM).  Here a persistent block owns one 64-channel tile (128 B per pixel) and
// walks 8 x 16-pixel output tiles; one elected thread requests the (8+2) x (16+2) halo tile -- and, for the backward,
// the 8 x 16 dy tile -- with cp.async.bulk.tensor into a multi-stage shared-memory ring (zero fill outside the image
// comes from the TMA unit), so 45-70 KB per block are in flight while the previous tile is consumed.  Each warp takes
// one tile row and slides a 3 x 3 register window along it: three 4-byte shared loads per output, lane = channel pair,
// so every shared access is one conflict-free 128-byte line and every global store a full line.
#include <algorithm>

#include "tc_common.cuh"

namespace segmif {

namespace {

constexpr int TW = 16, TH = 8, CT = 64;                       // output tile and channel tile
constexpr int HALO_BYTES = (TH + 2) * (TW + 2) * CT * 2;      // 23040
constexpr int DY_BYTES = TH * TW * CT * 2;                    // 16384
constexpr int FWD_STAGES = 3, BWD_STAGES = 2;

__device__ __forceinline__ float2 lds_bf2(const unsigned char* p) {
  const uint32_t u = *reinterpret_cast<const uint32_t*>(p);
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
__device__ __forceinline__ void stg_bf2(bf16* p, float a, float b) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}

// Packed fp32 pairs: sm_100's FFMA2 / FMUL2 process both channels of a lane in one issue slot (a scalar FFMA occupies the
// fma pipe for two cycles per warp, and these kernels are fma-issue bound once the loads are off the critical path).
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                     rc = *reinterpret_cast<unsigned long long*>(&c), rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 splat(float v) { return make_float2(v, v); }

// erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7), both channels at once; its exp(-x^2/2) is also the Gaussian
// density that GELU' needs.  cdf = Phi(x), e = exp(-x^2/2).
__device__ __forceinline__ void gelu_parts2(float2 x, float2& cdf, float2& e) {
  const float2 z = mul2(make_float2(fabsf(x.x), fabsf(x.y)), splat(0.70710678118654752440f));
  const float2 d = fma2(splat(0.3275911f), z, splat(1.0f));
  const float2 t = make_float2(__frcp_rn(d.x), __frcp_rn(d.y));
  float2 p = fma2(splat(1.061405429f), t, splat(-1.453152027f));
  p = fma2(p, t, splat(1.421413741f));
  p = fma2(p, t, splat(-0.284496736f));
  p = fma2(p, t, splat(0.254829592f));
  const float2 zz = mul2(mul2(z, z), splat(-1.4426950408889634f));
  e = make_float2(exp2f(zz.x), exp2f(zz.y));
  const float2 er = fma2(mul2(p, t), make_float2(-e.x, -e.y), splat(1.0f));          // erf(|x| / sqrt 2)
  cdf = fma2(splat(0.5f), make_float2(copysignf(er.x, x.x), copysignf(er.y, x.y)), splat(0.5f));
}

struct DwTiles {
  int tiles_x, tiles_y, B, H, W, C;
  __device__ __forceinline__ int count() const { return tiles_x * tiles_y * B; }
  __device__ __forceinline__ void at(int t, int& b, int& y0, int& x0) const {
    x0 = (t % tiles_x) * TW;
    t /= tiles_x;
    y0 = (t % tiles_y) * TH;
    b = t / tiles_y;
  }
};

// MODE 0: y = conv (+bias);  MODE 1: y = GELU(conv + bias)
template <int MODE>
__global__ void __launch_bounds__(256) dwconv_tma_fwd_kernel(const __grid_constant__ CUtensorMap tmX, const float* __restrict__ w9c,
                                                             const float* __restrict__ bias, bf16* __restrict__ y, DwTiles g,
                                                             int flip) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  __shared__ uint64_t full[FWD_STAGES];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = blockIdx.y * CT, c = c0 + lane * 2;
  const int ntiles = g.count();
  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&tmX);
#pragma unroll
    for (int s = 0; s < FWD_STAGES; ++s) tc::mbar_init(full + s, 1);
    tc::fence_barrier_init();
  }
  __syncthreads();
  auto issue = [&](int i) {                       // i-th tile of this block -> stage i % FWD_STAGES
    const int t = blockIdx.x + i * gridDim.x;
    if (t < ntiles) {
      int b, y0, x0;
      g.at(t, b, y0, x0);
      const int s = i % FWD_STAGES;
      tc::mbar_expect_tx(full + s, HALO_BYTES);
      tc::tma_load_4d(smem + s * HALO_BYTES, &tmX, full + s, c0, x0 - 1, y0 - 1, b);
    }
  };
  if (threadIdx.x == 0)
    for (int i = 0; i < FWD_STAGES - 1; ++i) issue(i);
  float2 wr[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) wr[t] = *reinterpret_cast<const float2*>(w9c + (flip ? 8 - t : t) * g.C + c);
  const float2 bv = bias ? *reinterpret_cast<const float2*>(bias + c) : make_float2(0.f, 0.f);
  int i = 0;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++i) {
    if (threadIdx.x == 0) issue(i + FWD_STAGES - 1);
    const int s = i % FWD_STAGES;
    tc::mbar_wait(full + s, (i / FWD_STAGES) & 1);
    int b, y0, x0;
    g.at(t, b, y0, x0);
    const int r = y0 + warp;
    if (r < g.H) {
      const unsigned char* base = smem + s * HALO_BYTES + (warp * (TW + 2)) * (CT * 2) + lane * 4;
      constexpr int RS = (TW + 2) * CT * 2;      // halo row pitch in bytes
      float2 a[3][3];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        a[ky][1] = lds_bf2(base + ky * RS);
        a[ky][2] = lds_bf2(base + ky * RS + CT * 2);
      }
      bf16* out = y + ((size_t)(b * g.H + r) * g.W + x0) * g.C + c;
#pragma unroll
      for (int col = 0; col < TW; ++col) {
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          a[ky][0] = a[ky][1];
          a[ky][1] = a[ky][2];
          a[ky][2] = lds_bf2(base + ky * RS + (col + 2) * (CT * 2));
        }
        float2 acc = bv;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) acc = fma2(wr[ky * 3 + kx], a[ky][kx], acc);
        if (MODE == 1) {
          float2 cdf, e;
          gelu_parts2(acc, cdf, e);
          acc = mul2(acc, cdf);
        }
        if (x0 + col < g.W) stg_bf2(out + (size_t)col * g.C, acc.x, acc.y);
      }
    }
    __syncthreads();                               // the stage may be refilled by the next iteration's request
  }
}

// dz = dy * GELU'(conv(x) + b);  part[blockIdx.x][t][c] = sum dz * x(p + t) (t < 9), part[..][9][c] = sum dz
__global__ void __launch_bounds__(256, 2) dwconv_tma_gelu_bwd_kernel(const __grid_constant__ CUtensorMap tmX,
                                                                     const __grid_constant__ CUtensorMap tmDy,
                                                                     const float* __restrict__ w9c, const float* __restrict__ bias,
                                                                     bf16* __restrict__ dz, float* __restrict__ part, DwTiles g) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  __shared__ uint64_t full[BWD_STAGES];
  __shared__ float sacc[10 * CT];
  constexpr int STAGE = HALO_BYTES + DY_BYTES;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = blockIdx.y * CT, c = c0 + lane * 2;
  const int ntiles = g.count();
  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&tmX);
    tc::prefetch_tmap(&tmDy);
#pragma unroll
    for (int s = 0; s < BWD_STAGES; ++s) tc::mbar_init(full + s, 1);
    tc::fence_barrier_init();
  }
  for (int k = threadIdx.x; k < 10 * CT; k += 256) sacc[k] = 0.f;
  __syncthreads();
  auto issue = [&](int i) {
    const int t = blockIdx.x + i * gridDim.x;
    if (t < ntiles) {
      int b, y0, x0;
      g.at(t, b, y0, x0);
      const int s = i % BWD_STAGES;
      tc::mbar_expect_tx(full + s, STAGE);
      tc::tma_load_4d(smem + s * STAGE, &tmX, full + s, c0, x0 - 1, y0 - 1, b);
      tc::tma_load_4d(smem + s * STAGE + HALO_BYTES, &tmDy, full + s, c0, x0, y0, b);
    }
  };
  if (threadIdx.x == 0)
    for (int i = 0; i < BWD_STAGES - 1; ++i) issue(i);
  float2 wr[9], aw[9], ab = make_float2(0.f, 0.f);
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    wr[t] = *reinterpret_cast<const float2*>(w9c + t * g.C + c);
    aw[t] = make_float2(0.f, 0.f);
  }
  const float2 bv = *reinterpret_cast<const float2*>(bias + c);
  int i = 0;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++i) {
    if (threadIdx.x == 0) issue(i + BWD_STAGES - 1);
    const int s = i % BWD_STAGES;
    tc::mbar_wait(full + s, (i / BWD_STAGES) & 1);
    int b, y0, x0;
    g.at(t, b, y0, x0);
    const int r = y0 + warp;
    if (r < g.H) {
      const unsigned char* base = smem + s * STAGE + (warp * (TW + 2)) * (CT * 2) + lane * 4;
      const unsigned char* dyb = smem + s * STAGE + HALO_BYTES + (warp * TW) * (CT * 2) + lane * 4;
      constexpr int RS = (TW + 2) * CT * 2;
      float2 a[3][3];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        a[ky][1] = lds_bf2(base + ky * RS);
        a[ky][2] = lds_bf2(base + ky * RS + CT * 2);
      }
      bf16* out = dz + ((size_t)(b * g.H + r) * g.W + x0) * g.C + c;
#pragma unroll 4
      for (int col = 0; col < TW; ++col) {
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          a[ky][0] = a[ky][1];
          a[ky][1] = a[ky][2];
          a[ky][2] = lds_bf2(base + ky * RS + (col + 2) * (CT * 2));
        }
        float2 z = bv;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) z = fma2(wr[ky * 3 + kx], a[ky][kx], z);
        float2 gy = lds_bf2(dyb + col * (CT * 2));           // zero outside the image (TMA fill): contributes nothing
        float2 cdf, e;
        gelu_parts2(z, cdf, e);
        gy = mul2(gy, fma2(mul2(z, splat(0.39894228040143267794f)), e, cdf));        // GELU'(z) = Phi(z) + z phi(z)
        ab.x += gy.x;
        ab.y += gy.y;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) aw[ky * 3 + kx] = fma2(gy, a[ky][kx], aw[ky * 3 + kx]);
        if (x0 + col < g.W) stg_bf2(out + (size_t)col * g.C, gy.x, gy.y);
      }
    }
    __syncthreads();
  }
  // fold the eight warps (same channels, different tile rows), then one plain store per block and channel
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    atomicAdd(&sacc[t * CT + lane * 2], aw[t].x);
    atomicAdd(&sacc[t * CT + lane * 2 + 1], aw[t].y);
  }
  atomicAdd(&sacc[9 * CT + lane * 2], ab.x);
  atomicAdd(&sacc[9 * CT + lane * 2 + 1], ab.y);
  __syncthreads();
  float* dst = part + (size_t)blockIdx.x * 10 * g.C + c0;
  for (int k = threadIdx.x; k < 10 * CT; k += 256) dst[(k / CT) * g.C + (k % CT)] = sacc[k];
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

int make_map(CUtensorMap* tm, const void* base, int B, int H, int W, int C, int bw, int bh, const char* what) {
  const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  const uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
  const uint32_t box[4] = {(uint32_t)CT, (uint32_t)bw, (uint32_t)bh, 1};
  return make_tmap_bf16(tm, base, 4, dims, strides, box, false, what, C == CT ? 256 : 128);
}

DwTiles make_tiles(int B, int H, int W, int C) {
  DwTiles g;
  g.tiles_x = (W + TW - 1) / TW; g.tiles_y = (H + TH - 1) / TH; g.B = B; g.H = H; g.W = W; g.C = C;
  return g;
}

int sm_count() {
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  return sms;
}

int bwd_grid_x(int B, int H, int W, int C) {
  const int spatial = ((W + TW - 1) / TW) * ((H + TH - 1) / TH) * B, ctiles = C / CT;
  return std::max(1, std::min(spatial, (2 * 148 + ctiles - 1) / ctiles));
}

}  // namespace

bool dwconv_tma_ok(int B, int H, int W, int C) {
  return C % CT == 0 && B > 0 && H > 0 && W > 0 && (int64_t)C / CT <= 65535 && (int64_t)B * H * W * C < (1ll << 40);
}

int dwconv_tma_fwd(const void* x, const float* w9c, const float* bias, void* y, int B, int H, int W, int C, int flip, int gelu,
                   cudaStream_t st) {
  CUtensorMap tmX;
  if (int rc = make_map(&tmX, x, B, H, W, C, TW + 2, TH + 2, "dwconv(x)")) return rc;
  const DwTiles g = make_tiles(B, H, W, C);
  const int ctiles = C / CT, spatial = g.tiles_x * g.tiles_y * B;
  const int gx = std::max(1, std::min(spatial, (3 * sm_count() + ctiles - 1) / ctiles));
  const size_t smem = (size_t)FWD_STAGES * HALO_BYTES + 128;
  static bool cfg = false;
  if (!cfg) {
    cudaError_t e0 = cudaFuncSetAttribute(dwconv_tma_fwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaError_t e1 = cudaFuncSetAttribute(dwconv_tma_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e0 != cudaSuccess || e1 != cudaSuccess) { set_error("dwconv: cudaFuncSetAttribute failed"); return SEGMIF_ERR_CUDA; }
    cfg = true;
  }
  if (gelu) dwconv_tma_fwd_kernel<1><<<dim3(gx, ctiles), 256, smem, st>>>(tmX, w9c, bias, (bf16*)y, g, flip);
  else dwconv_tma_fwd_kernel<0><<<dim3(gx, ctiles), 256, smem, st>>>(tmX, w9c, bias, (bf16*)y, g, flip);
  return check_launch("segmif_dwconv3x3 (tma)");
}

int64_t dwconv_tma_bwd_workspace(int B, int H, int W, int C) { return (int64_t)bwd_grid_x(B, H, W, C) * 10 * C; }

int dwconv_tma_gelu_bwd(const void* x, const float* w9c, const float* bias, const void* dy, void* dz, int B, int H, int W, int C,
                        float* dw9c, float* dbias, float* workspace, cudaStream_t st) {
  CUtensorMap tmX, tmDy;
  if (int rc = make_map(&tmX, x, B, H, W, C, TW + 2, TH + 2, "dwconv_bwd(x)")) return rc;
  if (int rc = make_map(&tmDy, dy, B, H, W, C, TW, TH, "dwconv_bwd(dy)")) return rc;
  const DwTiles g = make_tiles(B, H, W, C);
  const int ctiles = C / CT, gx = bwd_grid_x(B, H, W, C);
  const size_t smem = (size_t)BWD_STAGES * (HALO_BYTES + DY_BYTES) + 128;
  static bool cfg = false;
  if (!cfg) {
    if (cudaFuncSetAttribute(dwconv_tma_gelu_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      set_error("dwconv_bwd: cudaFuncSetAttribute failed");
      return SEGMIF_ERR_CUDA;
    }
    cfg = true;
  }
  dwconv_tma_gelu_bwd_kernel<<<dim3(gx, ctiles), 256, smem, st>>>(tmX, tmDy, w9c, bias, (bf16*)dz, workspace, g);
  dwconv_part_reduce_kernel<<<(10 * C + 255) / 256, 256, 0, st>>>(workspace, gx, C, dw9c, dbias);
  return check_launch("segmif_dwconv3x3_gelu_bwd (tma)");
}

}  // namespace segmif

// Training-side building blocks of the fusion network's backward pass (train.py:266-413 -> loss.backward()):
//   act_bwd        ReLU / PReLU derivative from the layer OUTPUT + bias gradient (column sums) + PReLU slope gradient
//   prelu_plane_bwd  the same for conv22's single fp32 output plane
//   colsum, add    column sums of a pixel-major slice; bf16 slice addition (DRDB residual in training mode)
//   layernorm_bwd  dx, dgamma, dbeta (+ column sums of dx) for nn.LayerNorm over the last dim
//   wgrad          weight gradient of a 3x3 (dilated) convolution or a linear layer as a tensor-core contraction over
//                  ALL pixels:  dW[co][tap][ci] = sum_p dY[p][co] * X[p + tap][ci]   (mma.sync m16n8k16, both operands
//                  pixel-major so both fragments come from ldmatrix.trans; 8x16-pixel tiles with the dilation halo
//                  staged once in shared memory and shared by the nine taps; per-chunk partials reduced in a fixed
//                  order by wgrad_reduce, which accumulates straight into the strided .grad tensor)
//   adamw_step     fused AdamW over a flat fp32 parameter / gradient / moment buffer (utils/optimizer.py:16-33)
// Activations and their gradients are bf16 pixel-major [pixels, ld] slices (ld = channel pitch, coff = offset).
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace segmif {

// ------------------------------------------------------------------------------------------------ act_bwd
// one thread = 8 consecutive channels of one pixel; C % 8 == 0
__global__ void __launch_bounds__(256) act_bwd_kernel(const bf16* __restrict__ y, int ldy, const bf16* __restrict__ dy,
                                                      int lddy, bf16* __restrict__ dz, int lddz, int64_t rows, int C,
                                                      int act, const float* __restrict__ alpha_p,
                                                      float* __restrict__ dbias, float* __restrict__ dalpha) {
  __shared__ float scol[256];      // C <= 256
  __shared__ float sred[8];
  const int g = C >> 3;                       // channel groups per row
  const int rpi = 256 / g;                    // rows per block iteration
  const bool active = threadIdx.x < rpi * g;
  const int grp = threadIdx.x % g, rin = threadIdx.x / g;
  const float alpha = act == SEGMIF_ACT_PRELU ? *alpha_p : 0.f;
  const float inv_alpha = act == SEGMIF_ACT_PRELU ? 1.f / alpha : 0.f;
  // The pre-activation is recovered from the stored OUTPUT (sign(y) == sign(z), z = y / alpha on the negative side), which
  // only holds for a slope > 0.  A trained slope that reaches <= 0 must not silently give wrong gradients: poison them.
  const float poison = (act == SEGMIF_ACT_PRELU && !(alpha > 0.f)) ? __int_as_float(0x7fc00000) : 0.f;
  float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float da = 0.f;
  if (active) {
    for (int64_t r = (int64_t)blockIdx.x * rpi + rin; r < rows; r += (int64_t)gridDim.x * rpi) {
      float vy[8], vd[8], o[8];
      load8(y + r * ldy + grp * 8, vy);
      load8(dy + r * lddy + grp * 8, vd);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const bool pos = vy[i] > 0.f;
        o[i] = (pos ? vd[i] : alpha * vd[i]) + poison;
        if (!pos) da = fmaf(vd[i], vy[i] * inv_alpha, da);     // z = y / alpha on the negative side
        cs[i] += o[i];
      }
      store8(dz + r * lddz + grp * 8, o);
    }
  }
  if (dbias != nullptr) {
    for (int i = threadIdx.x; i < C; i += 256) scol[i] = 0.f;
    __syncthreads();
    if (active) {
#pragma unroll
      for (int i = 0; i < 8; ++i) atomicAdd(&scol[grp * 8 + i], cs[i]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += 256) atomicAdd(dbias + i, scol[i]);
  }
  if (dalpha != nullptr && act == SEGMIF_ACT_PRELU) {
    da = warp_sum(da);
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = da;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int i = 0; i < 8; ++i) t += sred[i];
      atomicAdd(dalpha, t);
    }
  }
}

// conv22: out = prelu(z) fp32 plane, dout fp32 -> dz bf16 written to channel `coff` of a pixel-major [n, ld] tensor
__global__ void __launch_bounds__(256) prelu_plane_bwd_kernel(const float* __restrict__ out, const float* __restrict__ dout,
                                                              int64_t n, const float* __restrict__ alpha_p,
                                                              bf16* __restrict__ dz, int ld, float* __restrict__ dbias,
                                                              float* __restrict__ dalpha) {
  __shared__ float sred[2][8];
  const float alpha = *alpha_p, inv_alpha = 1.f / alpha;
  const float poison = !(alpha > 0.f) ? __int_as_float(0x7fc00000) : 0.f;     // see act_bwd_kernel: slope must stay > 0
  float sb = 0.f, sa = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const float o = out[i], d = dout[i];
    const bool pos = o > 0.f;
    const float v = (pos ? d : alpha * d) + poison;
    if (!pos) sa = fmaf(d, o * inv_alpha, sa);
    sb += v;
    dz[i * ld] = __float2bfloat16_rn(v);
  }
  sb = warp_sum(sb);
  sa = warp_sum(sa);
  if ((threadIdx.x & 31) == 0) { sred[0][threadIdx.x >> 5] = sb; sred[1][threadIdx.x >> 5] = sa; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float tb = 0.f, ta = 0.f;
    for (int i = 0; i < 8; ++i) { tb += sred[0][i]; ta += sred[1][i]; }
    if (dbias) atomicAdd(dbias, tb);
    if (dalpha) atomicAdd(dalpha, ta);
  }
}

// grid = (row chunks, 256-channel tiles): every thread walks its chunk's rows with 16-byte loads, the block folds its
// row lanes in shared memory and issues one global atomic per channel (<= 128 row chunks -> <= 128 adds per address)
__global__ void __launch_bounds__(256) colsum_kernel(const bf16* __restrict__ x, int ld, int64_t rows, int C,
                                                     float* __restrict__ out) {
  __shared__ float scol[256];
  const int c0 = blockIdx.y * 256;
  const int ct = min(256, C - c0);
  const int g = ct >> 3, rpi = 256 / g;
  const bool active = threadIdx.x < rpi * g;
  const int grp = threadIdx.x % g, rin = threadIdx.x / g;
  float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (active) {
    for (int64_t r = (int64_t)blockIdx.x * rpi + rin; r < rows; r += (int64_t)gridDim.x * rpi) {
      float v[8];
      load8(x + r * ld + c0 + grp * 8, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) cs[i] += v[i];
    }
  }
  for (int i = threadIdx.x; i < ct; i += 256) scol[i] = 0.f;
  __syncthreads();
  if (active) {
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(&scol[grp * 8 + i], cs[i]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ct; i += 256) atomicAdd(out + c0 + i, scol[i]);
}

__global__ void __launch_bounds__(256) add_bf16_kernel(const bf16* __restrict__ a, int lda, const bf16* __restrict__ b,
                                                       int ldb, bf16* __restrict__ o, int ldo, int64_t rows, int C) {
  const int g = C >> 3;
  const int64_t n = rows * g;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const int64_t r = i / g;
    const int grp = (int)(i - r * g);
    float va[8], vb[8];
    load8(a + r * lda + grp * 8, va);
    load8(b + r * ldb + grp * 8, vb);
#pragma unroll
    for (int k = 0; k < 8; ++k) va[k] += vb[k];
    store8(o + r * ldo + grp * 8, va);
  }
}

// ------------------------------------------------------------------------------------------------ layernorm_bwd
// warp per row, lane owns channels lane + 32*i.  x: pre-normalisation input [rows, C] (dense); dy / dx with pitches.
template <typename TX, typename TD, typename TO, int NC>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const TX* __restrict__ x, const TD* __restrict__ dy, int lddy,
                                                            const float* __restrict__ gamma, float eps,
                                                            TO* __restrict__ dx, int lddx, int64_t rows,
                                                            float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                            float* __restrict__ dxsum, int accumulate) {
  constexpr int C = NC * 32;
  __shared__ float sacc[3][C];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float gm[NC], ag[NC], ab[NC], as[NC];
#pragma unroll
  for (int i = 0; i < NC; ++i) { gm[i] = gamma[lane + 32 * i]; ag[i] = 0.f; ab[i] = 0.f; as[i] = 0.f; }
  for (int64_t r = (int64_t)blockIdx.x * 8 + warp; r < rows; r += (int64_t)gridDim.x * 8) {
    float xv[NC], dv[NC];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      xv[i] = ld_as_float(x + r * C + lane + 32 * i);
      dv[i] = ld_as_float(dy + r * lddy + lane + 32 * i);
      s += xv[i];
    }
    const float mean = warp_sum(s) * (1.f / C);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NC; ++i) { xv[i] -= mean; sq = fmaf(xv[i], xv[i], sq); }
    const float rstd = rsqrtf(warp_sum(sq) * (1.f / C) + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      xv[i] *= rstd;                               // xhat
      ag[i] = fmaf(dv[i], xv[i], ag[i]);
      ab[i] += dv[i];
      dv[i] *= gm[i];                              // dxhat
      s1 += dv[i];
      s2 = fmaf(dv[i], xv[i], s2);
    }
    s1 = warp_sum(s1) * (1.f / C);
    s2 = warp_sum(s2) * (1.f / C);
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const float v = rstd * (dv[i] - s1 - xv[i] * s2);
      as[i] += v;
      TO* o = dx + r * lddx + lane + 32 * i;
      st_from_float(o, accumulate ? ld_as_float(o) + v : v);      // accumulate: dx holds the residual-stream gradient
    }
  }
  for (int i = threadIdx.x; i < 3 * C; i += 256) (&sacc[0][0])[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    atomicAdd(&sacc[0][lane + 32 * i], ag[i]);
    atomicAdd(&sacc[1][lane + 32 * i], ab[i]);
    atomicAdd(&sacc[2][lane + 32 * i], as[i]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += 256) {
    if (dgamma) atomicAdd(dgamma + i, sacc[0][i]);
    if (dbeta) atomicAdd(dbeta + i, sacc[1][i]);
    if (dxsum) atomicAdd(dxsum + i, sacc[2][i]);
  }
}

// ------------------------------------------------------------------------------------------------ wgrad
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t (&r)[2], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n" : "=r"(r[0]), "=r"(r[1]) : "r"(saddr));
}

struct WgradArgs {
  const bf16* dy;      // [P, ldy], channel offset applied
  const bf16* x;       // [B, H, W, ldx], channel offset applied
  float* partials;     // [nchunk][Cout][NT][Cin]
  int ldy, ldx, B, H, W, Cin, Cout, dil;
  int64_t P;           // valid pixel count (guards the last rows of a flattened linear layer)
  int tiles_x, tiles_y;
};

constexpr int kWgTH = 8, kWgTW = 16, kWgThreads = 256;

// NT = 9: 3x3 taps with dilation `dil` ('same' padding);  NT = 1: linear layer / 1x1 conv
template <int NT>
__global__ void __launch_bounds__(kWgThreads) wgrad_kernel(WgradArgs a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int d = NT == 9 ? a.dil : 0;
  const int HXP = kWgTW + 2 * d, HROWS = kWgTH + 2 * d;
  const int x_bytes = HROWS * HXP * 128;
  const int stage_bytes = x_bytes + 128 * 64;          // halo tile of 64 channels + dY tile [128 px][32 ch]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ci0 = blockIdx.y * 64, co0 = blockIdx.z * 32;
  const int ci_valid = min(64, a.Cin - ci0);
  const int64_t ntiles = (int64_t)a.B * a.tiles_y * a.tiles_x;

  float acc[NT][2][4];
#pragma unroll
  for (int t = 0; t < NT; ++t)
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[t][m][j] = 0.f;

  auto issue_loads = [&](int64_t tile, int stage) {
    uint8_t* sX = smem_raw + stage * stage_bytes;
    uint8_t* sY = sX + x_bytes;
    const int tx = (int)(tile % a.tiles_x), ty = (int)((tile / a.tiles_x) % a.tiles_y);
    const int b = (int)(tile / ((int64_t)a.tiles_x * a.tiles_y));
    const int y0 = ty * kWgTH, x0 = tx * kWgTW;
    for (int i = tid; i < HROWS * HXP * 8; i += kWgThreads) {
      const int hp = i >> 3, chunk = i & 7;
      const int hy = hp / HXP, hx = hp - hy * HXP;
      const int yy = y0 + hy - d, xx = x0 + hx - d;
      const int64_t pix = ((int64_t)b * a.H + yy) * a.W + xx;
      const bool ok = (unsigned)yy < (unsigned)a.H && (unsigned)xx < (unsigned)a.W && pix < a.P && chunk * 8 < ci_valid;
      const bf16* src = ok ? a.x + pix * a.ldx + ci0 + chunk * 8 : a.x;
      cp_async16_cg(smem_u32(sX + hp * 128 + ((chunk ^ (hp & 7)) << 4)), src, ok ? 16 : 0);
    }
    for (int i = tid; i < 128 * 4; i += kWgThreads) {
      const int p = i >> 2, chunk = i & 3;
      const int yy = y0 + (p >> 4), xx = x0 + (p & 15);
      const int64_t pix = ((int64_t)b * a.H + yy) * a.W + xx;
      const bool ok = yy < a.H && xx < a.W && pix < a.P;
      const bf16* src = ok ? a.dy + pix * a.ldy + co0 + chunk * 8 : a.dy;
      cp_async16_cg(smem_u32(sY + p * 64 + ((chunk ^ ((p >> 1) & 3)) << 4)), src, ok ? 16 : 0);
    }
  };

  int stage = 0;
  int64_t tile = blockIdx.x;
  if (tile < ntiles) issue_loads(tile, 0);
  cp_async_commit();
  for (; tile < ntiles; tile += gridDim.x) {
    const int64_t next = tile + gridDim.x;
    if (next < ntiles) issue_loads(next, stage ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const uint8_t* sX = smem_raw + stage * stage_bytes;
    const uint8_t* sY = sX + x_bytes;
#pragma unroll 1
    for (int kr = 0; kr < kWgTH; ++kr) {             // one k-step = the 16 pixels of tile row kr
      uint32_t af[2][4];
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        const int p = kr * 16 + (lane & 7) + ((lane >> 4) << 3);
        const int chunk = m * 2 + ((lane >> 3) & 1);
        ldmatrix_x4_trans(af[m], smem_u32(sY + p * 64 + ((chunk ^ ((p >> 1) & 3)) << 4)));
      }
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const int ky = NT == 9 ? t / 3 : 0, kx = NT == 9 ? t % 3 : 0;
        const int hp = (kr + ky * d) * HXP + kx * d + (lane & 15);     // lanes 0-7: pixels 0-7, lanes 8-15: pixels 8-15
        uint32_t bfr[2];
        ldmatrix_x2_trans(bfr, smem_u32(sX + hp * 128 + ((warp ^ (hp & 7)) << 4)));
        mma_bf16_16816(acc[t][0], af[0], bfr[0], bfr[1]);
        mma_bf16_16816(acc[t][1], af[1], bfr[0], bfr[1]);
      }
    }
    __syncthreads();                                  // everyone is done with this stage before it is refilled
    stage ^= 1;
  }
  cp_async_wait<0>();
  // partial[chunk][co][tap][ci]
  const int g = lane >> 2, tq = lane & 3;
  float* out = a.partials + (int64_t)blockIdx.x * a.Cout * NT * a.Cin;
  const int ci = ci0 + warp * 8 + tq * 2;
  if (warp * 8 + tq * 2 < ci_valid) {
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int co = co0 + m * 16 + g + half * 8;
          *reinterpret_cast<float2*>(out + ((int64_t)co * NT + t) * a.Cin + ci) =
              make_float2(acc[t][m][half * 2], acc[t][m][half * 2 + 1]);
        }
  }
}

// ---- linear layers (NT = 1):  dW[co][ci] = sum_p dY[p][co] * X[p][ci]  as a 128 x 128 x (pixels) tile per block ----------
// The generic kernel above gives each block a 32 x 64 output tile, so for a linear layer it refills 24 KB of shared
// memory per 16 MMAs and re-reads X once per 32 output channels: it ran at 25-60 TFLOP/s, bound by the fill traffic.
// Here a block owns 128 output x 128 input channels, eight warps of 64 x 32 each (16 MMAs per six ldmatrix), and walks its
// contiguous pixel range in 64-pixel steps through a 3-stage cp.async ring.  Both operands are pixel-major, i.e. K-outer
// in shared memory, so both fragment loads are ldmatrix.trans; the 16-byte chunks of each 256-byte pixel row are XOR-
// swizzled with the pixel index so that the eight rows of an 8x8 matrix fall into distinct bank groups.
constexpr int kWlBM = 128, kWlBN = 128, kWlBK = 64, kWlStages = 3;

__global__ void __launch_bounds__(256, 2) wgrad_lin_kernel(WgradArgs a, int64_t pix_per_chunk) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  constexpr int ROW = 256;                                   // bytes per pixel row of either operand tile (128 channels)
  constexpr int OP_BYTES = kWlBK * ROW;                      // 16 KB
  constexpr int STAGE = 2 * OP_BYTES;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;                   // 2 x 4 warps: 64 co x 32 ci each
  const int ci0 = blockIdx.y * kWlBN, co0 = blockIdx.z * kWlBM;
  const int64_t p_begin = (int64_t)blockIdx.x * pix_per_chunk;
  const int64_t p_end = min(a.P, p_begin + pix_per_chunk);
  const int nsteps = p_end > p_begin ? (int)((p_end - p_begin + kWlBK - 1) / kWlBK) : 0;

  float acc[4][4][4];
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[m][n][j] = 0.f;

  auto issue = [&](int step) {
    if (step < nsteps) {
      uint8_t* sY = smem_raw + (step % kWlStages) * STAGE;
      uint8_t* sX = sY + OP_BYTES;
      const int64_t p0 = p_begin + (int64_t)step * kWlBK;
#pragma unroll
      for (int k = 0; k < 4; ++k) {                          // 64 px x 16 chunks per operand, 256 threads
        const int i = tid + k * 256;
        const int p = i >> 4, chunk = i & 15;
        const int64_t pix = p0 + p;
        const bool okp = pix < p_end;
        const bool oky = okp && co0 + chunk * 8 < a.Cout;
        const bool okx = okp && ci0 + chunk * 8 < a.Cin;
        const uint32_t off = (uint32_t)(p * ROW + ((chunk ^ (p & 7)) << 4));
        cp_async16_cg(smem_u32(sY + off), oky ? a.dy + pix * a.ldy + co0 + chunk * 8 : a.dy, oky ? 16 : 0);
        cp_async16_cg(smem_u32(sX + off), okx ? a.x + pix * a.ldx + ci0 + chunk * 8 : a.x, okx ? 16 : 0);
      }
    }
    cp_async_commit();
  };

  issue(0);
  issue(1);
  for (int step = 0; step < nsteps; ++step) {
    issue(step + 2);
    cp_async_wait<2>();
    __syncthreads();
    const uint8_t* sY = smem_raw + (step % kWlStages) * STAGE;
    const uint8_t* sX = sY + OP_BYTES;
#pragma unroll
    for (int kk = 0; kk < kWlBK / 16; ++kk) {
      uint32_t af[4][4], bfr[2][4];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int p = kk * 16 + (lane & 7) + ((lane >> 4) << 3);
        const int chunk = wm * 8 + m * 2 + ((lane >> 3) & 1);
        ldmatrix_x4_trans(af[m], smem_u32(sY + p * ROW + ((chunk ^ (p & 7)) << 4)));
      }
#pragma unroll
      for (int n2 = 0; n2 < 2; ++n2) {
        const int p = kk * 16 + (lane & 15);
        const int chunk = wn * 4 + n2 * 2 + (lane >> 4);
        ldmatrix_x4_trans(bfr[n2], smem_u32(sX + p * ROW + ((chunk ^ (p & 7)) << 4)));
      }
#pragma unroll
      for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int n = 0; n < 4; ++n) mma_bf16_16816(acc[m][n], af[m], bfr[n >> 1][(n & 1) * 2], bfr[n >> 1][(n & 1) * 2 + 1]);
    }
    __syncthreads();                                          // the stage is refilled two iterations later
  }
  cp_async_wait<0>();
  // partial[chunk][co][ci]
  const int g = lane >> 2, tq = lane & 3;
  float* out = a.partials + (int64_t)blockIdx.x * a.Cout * a.Cin;
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      const int ci = ci0 + wn * 32 + n * 8 + tq * 2;
      if (ci >= a.Cin) continue;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int co = co0 + wm * 64 + m * 16 + g + half * 8;
        if (co < a.Cout) *reinterpret_cast<float2*>(out + (int64_t)co * a.Cin + ci) = make_float2(acc[m][n][half * 2], acc[m][n][half * 2 + 1]);
      }
    }
}

// grad[co*s_co + tap*s_tap + ci*s_ci] += sum_chunk partial[chunk][co][tap][ci]   (fixed order: deterministic)
// A block reduces 32 consecutive outputs; its 8 warps split the chunks (warp w takes chunks w, w+8, ... with four
// independent accumulators), then the 8 warp sums are added in warp order.  (One thread per output walking all chunks
// serially was latency-bound: a 64x64 layer has 4096 outputs but ~300 chunks.)
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partials, int nchunk, int Cout, int NT,
                                                           int Cin, float* __restrict__ grad, int64_t s_co, int64_t s_tap,
                                                           int64_t s_ci, int co_take, int ci_take) {
  __shared__ float ssum[8][32];
  const int64_t n = (int64_t)Cout * NT * Cin;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t base = (int64_t)blockIdx.x * 32; base < n; base += (int64_t)gridDim.x * 32) {
    const int64_t i = base + lane;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (i < n) {
      int c = warp;
      for (; c + 24 < nchunk; c += 32) {
        a0 += partials[(int64_t)c * n + i];
        a1 += partials[(int64_t)(c + 8) * n + i];
        a2 += partials[(int64_t)(c + 16) * n + i];
        a3 += partials[(int64_t)(c + 24) * n + i];
      }
      for (; c < nchunk; c += 8) a0 += partials[(int64_t)c * n + i];
    }
    ssum[warp][lane] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (warp == 0 && i < n) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += ssum[w][lane];
      const int ci = (int)(i % Cin), tp = (int)((i / Cin) % NT), co = (int)(i / ((int64_t)Cin * NT));
      if (co < co_take && ci < ci_take) grad[co * s_co + tp * s_tap + ci * s_ci] += t;
    }
    __syncthreads();
  }
}

// few chunks (wide layers: many outputs, each a short sum): one thread per output, coalesced, two accumulators
__global__ void __launch_bounds__(256) wgrad_reduce_flat_kernel(const float* __restrict__ partials, int nchunk, int Cout, int NT,
                                                                int Cin, float* __restrict__ grad, int64_t s_co, int64_t s_tap,
                                                                int64_t s_ci, int co_take, int ci_take) {
  const int64_t n = (int64_t)Cout * NT * Cin;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const int ci = (int)(i % Cin), t = (int)((i / Cin) % NT), co = (int)(i / ((int64_t)Cin * NT));
    if (co >= co_take || ci >= ci_take) continue;
    float a0 = 0.f, a1 = 0.f;
    int c = 0;
    for (; c + 2 <= nchunk; c += 2) { a0 += partials[(int64_t)c * n + i]; a1 += partials[(int64_t)(c + 1) * n + i]; }
    if (c < nchunk) a0 += partials[(int64_t)c * n + i];
    grad[co * s_co + t * s_tap + ci * s_ci] += a0 + a1;
  }
}

// ------------------------------------------------------------------------------------------------ AdamW
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, int64_t n, float lr, float beta1, float beta2,
                                                    float eps, float wd, float bc1, float bc2_sqrt, float gscale) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const float gr = g[i] * gscale;
    float pv = p[i] * (1.f - lr * wd);
    const float mv = beta1 * m[i] + (1.f - beta1) * gr;
    const float vv = beta2 * v[i] + (1.f - beta2) * gr * gr;
    m[i] = mv;
    v[i] = vv;
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    pv -= (lr / bc1) * (mv / denom);
    p[i] = pv;
  }
}

}  // namespace segmif

using namespace segmif;

static int grid_for(int64_t items, int per_block) {
  return (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(items, per_block), 148 * 8));
}

extern "C" int segmif_act_bwd(const void* y, int ldy, int coffy, const void* dy, int lddy, int coffdy, void* dz, int lddz,
                              int coffdz, int64_t rows, int C, int act, const float* prelu_alpha, float* dbias,
                              float* dalpha, segmif_stream_t stream) {
  SEGMIF_REQUIRE(y && dy && dz && rows > 0, "act_bwd: bad arguments");
  SEGMIF_REQUIRE(C % 8 == 0 && C > 0 && C <= 256, "act_bwd: C=%d must be a multiple of 8, <= 256", C);
  SEGMIF_REQUIRE(ldy % 8 == 0 && lddy % 8 == 0 && lddz % 8 == 0 && coffy % 8 == 0 && coffdy % 8 == 0 && coffdz % 8 == 0,
                 "act_bwd: pitches/offsets must be multiples of 8");
  SEGMIF_REQUIRE(act == SEGMIF_ACT_RELU || (act == SEGMIF_ACT_PRELU && prelu_alpha), "act_bwd: ReLU or PReLU (with slope) only");
  const int rpi = 256 / (C >> 3);
  act_bwd_kernel<<<grid_for(rows, rpi * 4), 256, 0, as_stream(stream)>>>((const bf16*)y + coffy, ldy, (const bf16*)dy + coffdy, lddy,
                                                                          (bf16*)dz + coffdz, lddz, rows, C, act, prelu_alpha, dbias, dalpha);
  return check_launch("segmif_act_bwd");
}

extern "C" int segmif_prelu_plane_bwd(const float* out, const float* dout, int64_t n, const float* prelu_alpha, void* dz,
                                      int lddz, int coffdz, float* dbias, float* dalpha, segmif_stream_t stream) {
  SEGMIF_REQUIRE(out && dout && prelu_alpha && dz && n > 0 && lddz > 0, "prelu_plane_bwd: bad arguments");
  prelu_plane_bwd_kernel<<<grid_for(n, 1024), 256, 0, as_stream(stream)>>>(out, dout, n, prelu_alpha, (bf16*)dz + coffdz, lddz, dbias, dalpha);
  return check_launch("segmif_prelu_plane_bwd");
}

extern "C" int segmif_colsum(const void* x, int ld, int coff, int64_t rows, int C, float* out, segmif_stream_t stream) {
  SEGMIF_REQUIRE(x && out && rows > 0, "colsum: bad arguments");
  SEGMIF_REQUIRE(C % 8 == 0 && C > 0 && ld % 8 == 0 && coff % 8 == 0, "colsum: C, pitch and offset must be multiples of 8");
  const int ctiles = (C + 255) / 256;
  const int rpi = 256 / (std::min(C, 256) >> 3);
  // ~16 rows per thread, at most 8 blocks per SM: big tensors (the fusion net's full-resolution maps) need the whole GPU,
  // small ones few atomics per address
  const int chunks = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(rows, (int64_t)rpi * 16), std::max(1, 148 * 8 / ctiles)));
  colsum_kernel<<<dim3(chunks, ctiles), 256, 0, as_stream(stream)>>>((const bf16*)x + coff, ld, rows, C, out);
  return check_launch("segmif_colsum");
}

extern "C" int segmif_add_bf16(const void* a, int lda, int coffa, const void* b, int ldb, int coffb, void* out, int ldo,
                               int coffo, int64_t rows, int C, segmif_stream_t stream) {
  SEGMIF_REQUIRE(a && b && out && rows > 0, "add_bf16: bad arguments");
  SEGMIF_REQUIRE(C % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && ldo % 8 == 0 && coffa % 8 == 0 && coffb % 8 == 0 && coffo % 8 == 0,
                 "add_bf16: C, pitches and offsets must be multiples of 8");
  add_bf16_kernel<<<grid_for(rows * (C >> 3), 1024), 256, 0, as_stream(stream)>>>((const bf16*)a + coffa, lda, (const bf16*)b + coffb, ldb,
                                                                                   (bf16*)out + coffo, ldo, rows, C);
  return check_launch("segmif_add_bf16");
}

template <typename TX, typename TD, typename TO>
static int launch_ln_bwd(const void* x, const void* dy, int lddy, const float* gamma, float eps, void* dx, int lddx,
                         int64_t rows, int C, float* dgamma, float* dbeta, float* dxsum, int accumulate, cudaStream_t st) {
  const int grid = grid_for(rows, 32);
#define SEGMIF_LNB(NC)                                                                                                   \
  layernorm_bwd_kernel<TX, TD, TO, NC><<<grid, 256, 0, st>>>((const TX*)x, (const TD*)dy, lddy, gamma, eps, (TO*)dx, lddx, \
                                                             rows, dgamma, dbeta, dxsum, accumulate)
  switch (C) {
    case 64: SEGMIF_LNB(2); break;
    case 128: SEGMIF_LNB(4); break;
    case 320: SEGMIF_LNB(10); break;
    case 512: SEGMIF_LNB(16); break;
    default: set_error("layernorm_bwd: C=%d unsupported (64, 128, 320, 512)", C); return SEGMIF_ERR_INVALID;
  }
#undef SEGMIF_LNB
  return check_launch("segmif_layernorm_bwd");
}

extern "C" int segmif_layernorm_bwd(const void* x, int x_dtype, const void* dy, int dy_dtype, int lddy, int coffdy,
                                    const float* gamma, float eps, void* dx, int dx_dtype, int lddx, int coffdx,
                                    int64_t rows, int C, float* dgamma, float* dbeta, float* dxsum, int accumulate,
                                    segmif_stream_t stream) {
  SEGMIF_REQUIRE(x && dy && gamma && dx && rows > 0, "layernorm_bwd: bad arguments");
  cudaStream_t st = as_stream(stream);
  if (x_dtype == SEGMIF_BF16 && dy_dtype == SEGMIF_BF16 && dx_dtype == SEGMIF_BF16)
    return launch_ln_bwd<bf16, bf16, bf16>(x, (const bf16*)dy + coffdy, lddy, gamma, eps, (bf16*)dx + coffdx, lddx, rows, C, dgamma, dbeta, dxsum, accumulate, st);
  if (x_dtype == SEGMIF_F32 && dy_dtype == SEGMIF_F32 && dx_dtype == SEGMIF_F32)
    return launch_ln_bwd<float, float, float>(x, (const float*)dy + coffdy, lddy, gamma, eps, (float*)dx + coffdx, lddx, rows, C, dgamma, dbeta, dxsum, accumulate, st);
  if (x_dtype == SEGMIF_F32 && dy_dtype == SEGMIF_BF16 && dx_dtype == SEGMIF_F32)
    return launch_ln_bwd<float, bf16, float>(x, (const bf16*)dy + coffdy, lddy, gamma, eps, (float*)dx + coffdx, lddx, rows, C, dgamma, dbeta, dxsum, accumulate, st);
  set_error("layernorm_bwd: unsupported dtype combination (x %d, dy %d, dx %d)", x_dtype, dy_dtype, dx_dtype);
  return SEGMIF_ERR_INVALID;
}

extern "C" size_t segmif_wgrad_workspace_bytes(int nchunk, int Cout, int taps, int Cin) {
  return (size_t)nchunk * Cout * taps * Cin * sizeof(float);
}

static bool wgrad_use_tc() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("SEGMIF_WGRAD_TC"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

/* number of pixel chunks (= per-block partial buffers) segmif_wgrad should be called with for this problem */
extern "C" int segmif_wgrad_chunks(int B, int H, int W, int64_t P, int Cin, int Cout, int taps, int dil) {
  if (taps == 1) {
    const int64_t tiles = (int64_t)((Cin + kWlBN - 1) / kWlBN) * ((Cout + kWlBM - 1) / kWlBM);
    return (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(P, (int64_t)kWlBK), (2 * 148) / std::max<int64_t>(1, tiles)));
  }
  if (wgrad_use_tc() && wgrad_tc_ok(B, H, W, Cin, Cout, taps, dil, 8, 8)) return wgrad_tc_chunks(B, H, W, Cin, Cout);
  const int64_t ntiles = (int64_t)B * ((H + kWgTH - 1) / kWgTH) * ((W + kWgTW - 1) / kWgTW);
  return (int)std::max<int64_t>(1, std::min<int64_t>(ntiles, (2 * 148) / std::max(1, ((Cin + 63) / 64) * (Cout / 32))));
}

extern "C" int segmif_wgrad(const void* dy, int ldy, int coffy, const void* x, int ldx, int coffx, int B, int H, int W,
                            int64_t P, int Cin, int Cout, int taps, int dil, float* workspace, int nchunk, float* grad,
                            int64_t s_co, int64_t s_tap, int64_t s_ci, int co_take, int ci_take, segmif_stream_t stream) {
  SEGMIF_REQUIRE(dy && x && workspace && grad, "wgrad: null pointer");
  SEGMIF_REQUIRE(taps == 1 || taps == 9, "wgrad: taps=%d must be 1 (linear) or 9 (3x3)", taps);
  SEGMIF_REQUIRE(taps == 1 || dil == 1 || dil == 2, "wgrad: dilation must be 1 or 2");
  SEGMIF_REQUIRE(Cout % 32 == 0 && Cout > 0 && Cin % 8 == 0 && Cin > 0, "wgrad: Cout must be a multiple of 32, Cin of 8");
  SEGMIF_REQUIRE(ldy % 8 == 0 && ldx % 8 == 0 && coffy % 8 == 0 && coffx % 8 == 0, "wgrad: pitches/offsets must be multiples of 8");
  SEGMIF_REQUIRE(B > 0 && H > 0 && W > 0 && nchunk > 0 && P > 0 && P <= (int64_t)B * H * W, "wgrad: bad sizes");
  SEGMIF_REQUIRE(co_take > 0 && co_take <= Cout && ci_take > 0 && ci_take <= Cin, "wgrad: co_take / ci_take out of range");
  cudaStream_t st = as_stream(stream);
  WgradArgs a;
  a.dy = (const bf16*)dy + coffy; a.x = (const bf16*)x + coffx; a.partials = workspace;
  a.ldy = ldy; a.ldx = ldx; a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.dil = taps == 9 ? dil : 0; a.P = P;
  a.tiles_x = (W + kWgTW - 1) / kWgTW; a.tiles_y = (H + kWgTH - 1) / kWgTH;
  const int d = a.dil;
  const size_t smem = 2 * (size_t)((kWgTH + 2 * d) * (kWgTW + 2 * d) * 128 + 128 * 64);
  static bool cfg1 = false, cfg9 = false;
  const size_t smem_max = 2 * (size_t)((kWgTH + 4) * (kWgTW + 4) * 128 + 128 * 64);
  if (taps == 9 && !cfg9) {
    cfg9 = true;
    cudaError_t e = cudaFuncSetAttribute((const void*)wgrad_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
    if (e != cudaSuccess) { set_error("wgrad: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return SEGMIF_ERR_CUDA; }
  }
  if (taps == 1 && !cfg1) {
    cfg1 = true;
    cudaError_t e = cudaFuncSetAttribute((const void*)wgrad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
    if (e != cudaSuccess) { set_error("wgrad: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return SEGMIF_ERR_CUDA; }
  }
  if (taps == 1) {
    static bool cfgl = false;
    constexpr int smem_lin = kWlStages * 2 * kWlBK * 256;
    if (!cfgl) {
      cfgl = true;
      cudaError_t e = cudaFuncSetAttribute((const void*)wgrad_lin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_lin);
      if (e != cudaSuccess) { set_error("wgrad: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return SEGMIF_ERR_CUDA; }
    }
    const int64_t ppc = ceil_div(ceil_div(P, (int64_t)nchunk), (int64_t)kWlBK) * kWlBK;      // whole 64-pixel steps per chunk
    dim3 grid(nchunk, (Cin + kWlBN - 1) / kWlBN, (Cout + kWlBM - 1) / kWlBM);
    wgrad_lin_kernel<<<grid, 256, smem_lin, st>>>(a, ppc);
  } else if (wgrad_use_tc() && wgrad_tc_ok(B, H, W, Cin, Cout, taps, dil, ldy, ldx)) {
    int rc = wgrad_tc(a.dy, ldy, a.x, ldx, B, H, W, Cin, Cout, dil, workspace, nchunk, st);
    if (rc) return rc;
  } else {
    dim3 grid(nchunk, (Cin + 63) / 64, Cout / 32);
    wgrad_kernel<9><<<grid, kWgThreads, smem, st>>>(a);
  }
  int rc = check_launch("segmif_wgrad");
  if (rc) return rc;
  const int64_t n = (int64_t)Cout * taps * Cin;
  if (nchunk >= 32)
    wgrad_reduce_kernel<<<grid_for(n, 32), 256, 0, st>>>(workspace, nchunk, Cout, taps, Cin, grad, s_co, s_tap, s_ci, co_take, ci_take);
  else
    wgrad_reduce_flat_kernel<<<grid_for(n, 256), 256, 0, st>>>(workspace, nchunk, Cout, taps, Cin, grad, s_co, s_tap, s_ci, co_take, ci_take);
  return check_launch("segmif_wgrad(reduce)");
}

/* Linear layer: weight gradient (+= into grad[co*s_co + ci*s_ci]) and, when dbias is given, the bias gradient (+= column sums
 * of dY) from ONE pass over dY and X on tcgen05 (wgrad_lin_tc.cu).  nchunk / workspace as reported by
 * segmif_wgrad_lin_chunks.  SEGMIF_WGRAD_LIN_TC=0 (or shapes TMA cannot address) selects the mma.sync kernel + colsum. */
static bool wgrad_lin_use_tc() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("SEGMIF_WGRAD_LIN_TC"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

extern "C" int segmif_wgrad_lin_chunks(int64_t P, int Cin, int Cout) {
  if (wgrad_lin_use_tc() && wgrad_lin_tc_ok(P, Cin, Cout, 8, 8)) {
    int nchunk, bn;
    int64_t per;
    wgrad_lin_tc_plan(P, Cin, Cout, &nchunk, &per, &bn);
    return nchunk;
  }
  return segmif_wgrad_chunks(1, (int)((P + 15) / 16), 16, P, Cin, Cout, 1, 1);
}

extern "C" int segmif_wgrad_lin(const void* dy, int ldy, int coffy, const void* x, int ldx, int coffx, int64_t P, int Cin, int Cout,
                                float* workspace, int nchunk, float* grad, int64_t s_co, int64_t s_ci, int co_take, int ci_take,
                                float* dbias, segmif_stream_t stream) {
  SEGMIF_REQUIRE(dy && x && workspace && grad && P > 0 && nchunk > 0, "wgrad_lin: bad arguments");
  SEGMIF_REQUIRE(co_take > 0 && co_take <= Cout && ci_take > 0 && ci_take <= Cin, "wgrad_lin: co_take / ci_take out of range");
  SEGMIF_REQUIRE(ldy % 8 == 0 && ldx % 8 == 0 && coffy % 8 == 0 && coffx % 8 == 0, "wgrad_lin: pitches/offsets must be multiples of 8");
  cudaStream_t st = as_stream(stream);
  if (!(wgrad_lin_use_tc() && wgrad_lin_tc_ok(P, Cin, Cout, ldy, ldx))) {
    int rc = segmif_wgrad(dy, ldy, coffy, x, ldx, coffx, 1, (int)((P + 15) / 16), 16, P, Cin, Cout, 1, 1, workspace, nchunk, grad, s_co,
                          1, s_ci, co_take, ci_take, stream);
    if (rc || !dbias) return rc;
    return segmif_colsum(dy, ldy, coffy, P, Cout, dbias, stream);
  }
  int want, bn;
  int64_t per;
  wgrad_lin_tc_plan(P, Cin, Cout, &want, &per, &bn);
  SEGMIF_REQUIRE(nchunk == want, "wgrad_lin: nchunk=%d, segmif_wgrad_lin_chunks says %d", nchunk, want);
  int rc = wgrad_lin_tc((const bf16*)dy + coffy, ldy, (const bf16*)x + coffx, ldx, P, Cin, Cout, workspace, nchunk, per, bn, dbias, st);
  if (rc) return rc;
  const int64_t n = (int64_t)Cout * Cin;
  if (nchunk >= 32)
    wgrad_reduce_kernel<<<grid_for(n, 32), 256, 0, st>>>(workspace, nchunk, Cout, 1, Cin, grad, s_co, 1, s_ci, co_take, ci_take);
  else
    wgrad_reduce_flat_kernel<<<grid_for(n, 256), 256, 0, st>>>(workspace, nchunk, Cout, 1, Cin, grad, s_co, 1, s_ci, co_take, ci_take);
  return check_launch("segmif_wgrad_lin(reduce)");
}

extern "C" int segmif_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                                 float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                                 segmif_stream_t stream) {
  SEGMIF_REQUIRE(param && grad && exp_avg && exp_avg_sq && n > 0 && step > 0, "adamw_step: bad arguments");
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  adamw_kernel<<<grid_for(n, 1024), 256, 0, as_stream(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                                  weight_decay, (float)bc1, (float)sqrt(bc2), grad_scale);
  return check_launch("segmif_adamw_step");
}

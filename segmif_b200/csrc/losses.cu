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

// ------------------------------------------------------------------------------------------------ streaming stencils
// SSIM, the Laplacian residuals and Sobel share one layout ("column streaming").  A block of 128 threads owns a strip of
// 128 image columns and walks down a range of rows; thread = column.  Rows arrive in GROUPS: cp.async copies the next
// group's rows (+ R halo columns each side, zero fill outside the image) straight into the other half of a
// double-buffered shared-memory block while the current group is consumed, so a block keeps ~10 KB in flight and
// synchronises twice per group instead of once per row.  For every row each thread applies the HORIZONTAL taps to its
// column, and the VERTICAL filter is a scatter into a ring of partial output accumulators held in REGISTERS: input row
// r adds g[r - o + R] * h(r) to every output row o in [r - R, r + R]; output r - R is then complete and is consumed on the
// spot.  The group length equals the ring length and the group loop is fully unrolled, so all ring indices are
// compile-time constants.  Each input pixel is read from HBM once (plus 2R/rows and 2R/128 halo overheads) and nothing but
// the loss partials is written; the first versions (32x32 tiles with halo in shared memory, one shared load per FMA)
// were bound by shared-memory bandwidth at 8-9 % of the HBM roofline.
constexpr int kStripW = 128;

__device__ __forceinline__ void cp_async4(void* smem_dst, const float* gsrc, bool ok) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(ok ? 4 : 0));
}

struct StreamGeom {
  int x0, y0, y1;          // first column, output rows [y0, y1)
  int col;                 // this thread's column
  bool col_ok;
};
__device__ __forceinline__ StreamGeom stream_geom(int H, int W, int rows_per_chunk) {
  StreamGeom g;
  g.x0 = blockIdx.x * kStripW;
  g.y0 = blockIdx.y * rows_per_chunk;
  g.y1 = min(H, g.y0 + rows_per_chunk);
  g.col = g.x0 + threadIdx.x;
  g.col_ok = g.col < W;
  return g;
}

// rows [r0, r0 + ROWS) of NIMG planes -> buf[row][image][R + 128 + R]; rows / columns outside the image (or >= r_end) read 0
template <int NIMG, int R, int ROWS>
__device__ __forceinline__ void issue_rows(const float* const (&src)[NIMG], float (*buf)[NIMG][kStripW + 2 * R], int r0, int r_end,
                                           int H, int W, const StreamGeom& g) {
  const int t = threadIdx.x;
  const bool has_halo = t < 2 * R;
  const int hslot = t < R ? t : kStripW + t;                       // left halo slots [0, R), right [R + 128, 2R + 128)
  const int hcol = t < R ? g.x0 - R + t : g.x0 + kStripW + (t - R);
  const bool hcol_ok = has_halo && (unsigned)hcol < (unsigned)W;
#pragma unroll
  for (int i = 0; i < ROWS; ++i) {
    const int r = r0 + i;
    const bool row_ok = (unsigned)r < (unsigned)H && r < r_end;
    const int64_t o = (int64_t)(row_ok ? r : 0) * W;
#pragma unroll
    for (int m = 0; m < NIMG; ++m) {
      cp_async4(&buf[i][m][R + t], src[m] + o + (g.col_ok ? g.col : 0), row_ok && g.col_ok);
      if (has_halo) cp_async4(&buf[i][m][hslot], src[m] + o + (hcol_ok ? hcol : 0), row_ok && hcol_ok);
    }
  }
  cp_async_commit();
}

// ------------------------------------------------------------------------------------------------ SSIM
struct Gauss11 { float g[11]; };

// packed fp32 pairs (FFMA2 / FMUL2): (mu1, mu2) and (E[x^2], E[y^2]) travel together -- the kernel is issue bound
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                     rc = *reinterpret_cast<unsigned long long*>(&c), rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}

__global__ void __launch_bounds__(128) ssim_kernel(const float* __restrict__ a, const float* __restrict__ b, int H, int W,
                                                   int rows_per_chunk, Gauss11 win, float* __restrict__ partials) {
  constexpr int R = 5, N = 11;
  __shared__ float lines[2][N][2][kStripW + 2 * R];
  __shared__ float sred[8];
  const StreamGeom g = stream_geom(H, W, rows_per_chunk);
  const float* const src[2] = {a + (int64_t)blockIdx.z * H * W, b + (int64_t)blockIdx.z * H * W};
  const int t = threadIdx.x;
  float2 am[N], as[N];                                     // ring of (mu1, mu2) and (E[x^2], E[y^2]) partial sums
  float ax[N];                                             // ... and E[xy]
#pragma unroll
  for (int j = 0; j < N; ++j) { am[j] = make_float2(0.f, 0.f); as[j] = make_float2(0.f, 0.f); ax[j] = 0.f; }
  float local = 0.f;
  const int r_begin = g.y0 - R, r_end = g.y1 + R;         // input rows that touch this chunk's outputs
  issue_rows<2, R, N>(src, lines[0], r_begin, r_end, H, W, g);
  int pb = 0;
  for (int rb = r_begin; rb < r_end; rb += N, pb ^= 1) {
    issue_rows<2, R, N>(src, lines[pb ^ 1], rb + N, r_end, H, W, g);
    cp_async_wait<1>();
    __syncthreads();
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const int r = rb + i;
      if (r < r_end) {                                    // block-uniform
        if ((unsigned)r < (unsigned)H) {
          const float* la = lines[pb][i][0] + t;
          const float* lb = lines[pb][i][1] + t;
          float2 hm = make_float2(0.f, 0.f), hs = make_float2(0.f, 0.f);
          float hx = 0.f;
#pragma unroll
          for (int k = 0; k < N; ++k) {
            const float2 v = make_float2(la[k], lb[k]);
            const float2 w = make_float2(win.g[k], win.g[k]);
            hm = ffma2(w, v, hm);
            hs = ffma2(w, fmul2(v, v), hs);
            hx = fmaf(win.g[k], v.x * v.y, hx);
          }
#pragma unroll
          for (int j = 0; j < N; ++j) {                   // output row o = r - R + j gets weight g[N - 1 - j]
            const float2 w = make_float2(win.g[N - 1 - j], win.g[N - 1 - j]);
            const int sl = (i + j) % N;
            am[sl] = ffma2(w, hm, am[sl]);
            as[sl] = ffma2(w, hs, as[sl]);
            ax[sl] = fmaf(w.x, hx, ax[sl]);
          }
        }
        const int o = r - R;                              // complete now: slot i
        if (o >= g.y0 && o < g.y1 && g.col_ok) {
          const float m1 = am[i].x, m2 = am[i].y;
          const float mu1_sq = m1 * m1, mu2_sq = m2 * m2, mu12 = m1 * m2;
          const float sg1 = as[i].x - mu1_sq, sg2 = as[i].y - mu2_sq, sg12 = ax[i] - mu12;
          const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
          local += ((2.f * mu12 + C1) * (2.f * sg12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sg1 + sg2 + C2));
        }
        am[i] = make_float2(0.f, 0.f);
        as[i] = make_float2(0.f, 0.f);
        ax[i] = 0.f;
      }
    }
    __syncthreads();                                      // this half is refilled by the next iteration's cp.async
  }
  cp_async_wait<0>();
  const float tot = block_sum_256(local, sred);
  if (threadIdx.x == 0) partials[((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = tot;
}

// ------------------------------------------------------------------------------------------------ Laplacian
// lap_loss.py:39-71: the (k,k) Gaussian exp(-(dx^2+dy^2)/(2 sigma^2)) normalised to sum 1 is the outer product of the
// normalised 1-D Gaussians, so each blur is a horizontal and a vertical k-tap pass.  residual = img - blur is built in the
// accumulator ring itself (+centre when the centre row passes, -g_v * h_k for every row).
struct LapTaps { float g3[3], g5[5], g7[7]; };

// NIMG = 3: LapLoss2 (input, ir, vis -> target = max(res ir, res vis));  NIMG = 2: LapLoss (input, target)
template <int NIMG>
__global__ void __launch_bounds__(128) laploss_kernel(const float* __restrict__ inp, const float* __restrict__ p1,
                                                      const float* __restrict__ p2, int H, int W, int rows_per_chunk,
                                                      LapTaps tp, float* __restrict__ partials) {
  constexpr int R = 3, N = 7;
  __shared__ float lines[2][N][NIMG][kStripW + 2 * R];
  __shared__ float sred[8];
  const StreamGeom g = stream_geom(H, W, rows_per_chunk);
  const int64_t off = (int64_t)blockIdx.z * H * W;
  const float* src[NIMG];
  src[0] = inp + off;
  src[1] = p1 + off;
  if (NIMG == 3) src[NIMG - 1] = p2 + off;
  const int t = threadIdx.x;
  float acc[N][NIMG][3];                                   // [ring slot][image][scale 3/5/7]
#pragma unroll
  for (int j = 0; j < N; ++j)
#pragma unroll
    for (int m = 0; m < NIMG; ++m)
#pragma unroll
      for (int k = 0; k < 3; ++k) acc[j][m][k] = 0.f;
  float l3 = 0.f, l5 = 0.f, l7 = 0.f;
  const int r_begin = g.y0 - R, r_end = g.y1 + R;
  issue_rows<NIMG, R, N>(src, lines[0], r_begin, r_end, H, W, g);
  int pb = 0;
  for (int rb = r_begin; rb < r_end; rb += N, pb ^= 1) {
    issue_rows<NIMG, R, N>(src, lines[pb ^ 1], rb + N, r_end, H, W, g);
    cp_async_wait<1>();
    __syncthreads();
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const int r = rb + i;
      if (r < r_end) {
        if ((unsigned)r < (unsigned)H) {
#pragma unroll
          for (int m = 0; m < NIMG; ++m) {
            float v[N];
#pragma unroll
            for (int k = 0; k < N; ++k) v[k] = lines[pb][i][m][t + k];
            float h3 = 0.f, h5 = 0.f, h7 = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) h3 = fmaf(tp.g3[k], v[2 + k], h3);
#pragma unroll
            for (int k = 0; k < 5; ++k) h5 = fmaf(tp.g5[k], v[1 + k], h5);
#pragma unroll
            for (int k = 0; k < 7; ++k) h7 = fmaf(tp.g7[k], v[k], h7);
#pragma unroll
            for (int j = 0; j < N; ++j) {                 // output row o = r - R + j; d = r - o
              const int sl = (i + j) % N;
              const int d = R - j;
              if (d >= -1 && d <= 1) acc[sl][m][0] = fmaf(-tp.g3[d + 1], h3, acc[sl][m][0]);
              if (d >= -2 && d <= 2) acc[sl][m][1] = fmaf(-tp.g5[d + 2], h5, acc[sl][m][1]);
              acc[sl][m][2] = fmaf(-tp.g7[d + 3], h7, acc[sl][m][2]);
            }
            const int sc = (i + R) % N;                   // o == r: the centre pixel of the residual
            acc[sc][m][0] += v[R]; acc[sc][m][1] += v[R]; acc[sc][m][2] += v[R];
          }
        }
        const int o = r - R;
        if (o >= g.y0 && o < g.y1 && g.col_ok) {
          float tgt[3];
#pragma unroll
          for (int k = 0; k < 3; ++k) tgt[k] = NIMG == 3 ? fmaxf(acc[i][1][k], acc[i][NIMG - 1][k]) : acc[i][1][k];
          l3 += fabsf(acc[i][0][0] - tgt[0]);
          l5 += fabsf(acc[i][0][1] - tgt[1]);
          l7 += fabsf(acc[i][0][2] - tgt[2]);
        }
#pragma unroll
        for (int m = 0; m < NIMG; ++m)
#pragma unroll
          for (int k = 0; k < 3; ++k) acc[i][m][k] = 0.f;
      }
    }
    __syncthreads();
  }
  cp_async_wait<0>();
  const int64_t blk = ((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  const float t3 = block_sum_256(l3, sred);
  const float t5 = block_sum_256(l5, sred);
  const float t7 = block_sum_256(l7, sred);
  if (threadIdx.x == 0) { partials[blk * 3 + 0] = t3; partials[blk * 3 + 1] = t5; partials[blk * 3 + 2] = t7; }
}

// ------------------------------------------------------------------------------------------------ Entropy
// A warp walks P-row x 32-column segments (32 / P patches; the P lanes [q P, (q+1) P) hold patch q).  core/Entropy.py
// evaluates all 32 Gaussian kernels (sigma = 0.01, bins 1/31 = 3.2 sigma apart) for every pixel; a bin more than 3 bins
// away from the pixel's nearest bin is >= 11 sigma off and weighs < exp(-60), below fp32 resolution of the patch sums, so
// each lane evaluates only the 7 bins around each of its P pixels (7 exp per pixel instead of 32: the first version
// was MUFU-issue bound) and adds them into its PRIVATE column of a per-warp shared-memory histogram hist[bin][lane] --
// plain read-modify-write, no atomics (a shared-memory float atomicAdd is a compare-and-swap loop; ncu showed the
// atomic version stalled on it with 70 M bank conflicts).  Then lane = bin: the P columns of a patch are summed,
// normalised (one warp sum per patch) and -p log p accumulated per lane.  Rows are 33 floats apart: the scatter hits
// bank (bin + lane) % 32 (neighbouring pixels fall into neighbouring bins: rarely equal), the gather bank (lane + column).
struct Bins32 { float b[32]; };

__device__ __forceinline__ float ex2_ftz(float x) {        // one MUFU, no denormal fix-up code (weights < 2^-126 are 0 anyway)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int P>
__global__ void __launch_bounds__(256) entropy_kernel(const float* __restrict__ img, int B, int H, int W, Bins32 bins,
                                                      float* __restrict__ partials) {
  constexpr int NP = 32 / P;                               // patches per segment
  __shared__ float hist[8][32][33];
  __shared__ float sbin[32];
  __shared__ float sred[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  if (threadIdx.x < 32) sbin[threadIdx.x] = bins.b[threadIdx.x];
  __syncthreads();
  const int segs_x = (W + 31) / 32;
  const int prow = H / P;
  const int64_t nseg = (int64_t)B * prow * segs_x;
  const float inv_sigma = 1.0f / 0.01f;
  const float nhl2e = -0.5f * 1.4426950408889634f;
  float (*hw)[33] = hist[warp];
  float total = 0.f;
  for (int64_t sidx = (int64_t)blockIdx.x * nwarp + warp; sidx < nseg; sidx += (int64_t)gridDim.x * nwarp) {
    const int sx = (int)(sidx % segs_x);
    const int py = (int)((sidx / segs_x) % prow);
    const int64_t b = sidx / ((int64_t)segs_x * prow);
    const int x = sx * 32 + lane;
    float val[P];
#pragma unroll
    for (int dy = 0; dy < P; ++dy) val[dy] = x < W ? img[(b * H + (int64_t)py * P + dy) * W + x] : 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) hw[j][lane] = 0.f;        // own column
    if (x < W) {
#pragma unroll
      for (int dy = 0; dy < P; ++dy) {
        const int j0 = min(31, max(0, __float2int_rn(val[dy] * 31.0f)));
#pragma unroll
        for (int dj = -3; dj <= 3; ++dj) {
          const int j = j0 + dj;
          if ((unsigned)j < 32u) {
            const float r = (val[dy] - sbin[j]) * inv_sigma;
            hw[j][lane] += ex2_ftz(nhl2e * (r * r));
          }
        }
      }
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      if (sx * 32 + q * P >= W) break;                     // warp-uniform (W % P == 0)
      float acc = 0.f;
#pragma unroll
      for (int p = 0; p < P; ++p) acc += hw[lane][q * P + p];
      float pdf = acc * (1.0f / (float)(P * P));
      const float norm = warp_sum(pdf) + 1e-40f;
      pdf = pdf * __frcp_rn(norm) + 1e-40f;
      // p log p of an (almost) empty bin: the reference's 1e-40 * log(1e-40) ~ -1e-38; dropped (and __logf would flush
      // the denormal to -inf)
      if (pdf > 1e-30f) total -= pdf * __logf(pdf);
    }
    __syncwarp();
  }
  total = warp_sum(total);
  if (lane != 0) total = 0.f;
  const float tsum = block_sum_256(total, sred);
  if (threadIdx.x == 0) partials[blockIdx.x] = tsum;
}

// Thread-per-patch form (P = 4, 8, 16; W % 4 == 0): a thread owns a whole P x P patch and a PRIVATE 32-bin column
// hist[bin][thread] of shared memory -- with the thread index innermost every access of a warp hits 32 different banks
// whatever the (data dependent) bin is, so the read-modify-writes are conflict free (the warp-per-segment kernel above
// measured 72 M bank conflicts and 89 % shared-pipe utilisation at configs[4]: 0.76 ms, 5 % of the HBM roofline).
// Per pixel only the FOUR bins jl-1 .. jl+2 around jl = floor(31 v) are evaluated: any other bin is >= 2 bin widths
// = 6.45 sigma away and weighs < 1e-9 of the nearest one, far below the fp32 resolution of the normalised histogram.
// The finishing pass (32 bins -> registers, column zeroed for the next patch, normalise, -p log p) is per thread too:
// no shuffles, no barriers inside the loop.  Loads are 16-byte, consecutive threads = consecutive patches of a row.
template <int P>
__global__ void __launch_bounds__(128) entropy_patch_kernel(const float* __restrict__ img, int B, int H, int W, Bins32 bins,
                                                            float* __restrict__ partials) {
  __shared__ float hist[32 * 128];
  __shared__ float sbin[32];
  __shared__ float sred[8];
  const int tid = threadIdx.x;
  if (tid < 32) sbin[tid] = bins.b[tid];
#pragma unroll
  for (int k = 0; k < 32; ++k) hist[k * 128 + tid] = 0.f;
  __syncthreads();
  const int pw = W / P, ph = H / P;
  const int64_t npatch = (int64_t)B * ph * pw;
  const float inv_sigma = 1.0f / 0.01f;
  const float nhl2e = -0.5f * 1.4426950408889634f;
  float* col = hist + tid;
  float total = 0.f;
  for (int64_t p = (int64_t)blockIdx.x * 128 + tid; p < npatch; p += (int64_t)gridDim.x * 128) {
    const int px = (int)(p % pw);
    const int py = (int)((p / pw) % ph);
    const int64_t b = p / ((int64_t)pw * ph);
    const float* base = img + (b * H + (int64_t)py * P) * W + (int64_t)px * P;
#pragma unroll 1
    for (int dy = 0; dy < P; ++dy) {
#pragma unroll
      for (int x4 = 0; x4 < P / 4; ++x4) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(base + (int64_t)dy * W) + x4);
        const float vv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float v = vv[e];
          int jl = __float2int_rd(v * 31.0f);
          jl = jl < 0 ? 0 : (jl > 31 ? 31 : jl);
#pragma unroll
          for (int dj = -1; dj <= 2; ++dj) {
            const int j = jl + dj;
            if ((unsigned)j < 32u) {
              const float r = (v - sbin[j]) * inv_sigma;
              col[j * 128] += ex2_ftz(nhl2e * (r * r));
            }
          }
        }
      }
    }
    float h[32], norm = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      h[k] = col[k * 128] * (1.0f / (float)(P * P));
      col[k * 128] = 0.f;
      norm += h[k];
    }
    const float inv = __frcp_rn(norm + 1e-40f);
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float pdf = h[k] * inv + 1e-40f;
      if (pdf > 1e-30f) total -= pdf * __logf(pdf);       // an (almost) empty bin contributes ~1e-38 in the reference: dropped
    }
  }
  const float tsum = block_sum_256(total, sred);
  if (tid == 0) partials[blockIdx.x] = tsum;
}

// ------------------------------------------------------------------------------------------------ Sobel + L1
// column streaming with a three-row register window per image: per row d = right - left and s = left + 2 c + right;
// gx(o) = d(o-1) + 2 d(o) + d(o+1), gy(o) = s(o-1) - s(o+1)   (core/loss.py:634-650, zero padding)
__global__ void __launch_bounds__(128) sobel_l1_kernel(const float* __restrict__ x, const float* __restrict__ y, int H, int W,
                                                       int rows_per_chunk, float* __restrict__ partials) {
  constexpr int R = 1, N = 8;
  __shared__ float lines[2][N][2][kStripW + 2 * R];
  __shared__ float sred[8];
  const StreamGeom g = stream_geom(H, W, rows_per_chunk);
  const float* const src[2] = {x + (int64_t)blockIdx.z * H * W, y + (int64_t)blockIdx.z * H * W};
  const int t = threadIdx.x;
  const int r_begin = g.y0 - R, r_end = g.y1 + R;
  float2 d0 = make_float2(0.f, 0.f), d1 = d0, s0 = d0, s1 = d0;      // rows r-2 (0) and r-1 (1), (x image, y image)
  float l1 = 0.f, lg = 0.f;
  issue_rows<2, R, N>(src, lines[0], r_begin, r_end, H, W, g);
  int pb = 0;
  for (int rb = r_begin; rb < r_end; rb += N, pb ^= 1) {
    issue_rows<2, R, N>(src, lines[pb ^ 1], rb + N, r_end, H, W, g);
    cp_async_wait<1>();
    __syncthreads();
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const int r = rb + i;
      if (r < r_end) {
        const float* lx = lines[pb][i][0] + t;
        const float* ly = lines[pb][i][1] + t;
        const float2 lf = make_float2(lx[0], ly[0]), ce = make_float2(lx[1], ly[1]), rt = make_float2(lx[2], ly[2]);
        const float2 d2 = make_float2(rt.x - lf.x, rt.y - lf.y);
        const float2 s2 = make_float2(lf.x + 2.f * ce.x + rt.x, lf.y + 2.f * ce.y + rt.y);
        if (r >= g.y0 && r < g.y1 && g.col_ok) l1 += fabsf(ce.x - ce.y);
        const int o = r - R;
        if (o >= g.y0 && o < g.y1 && g.col_ok) {
          const float gxx = d0.x + 2.f * d1.x + d2.x, gyx = s0.x - s2.x;
          const float gxy = d0.y + 2.f * d1.y + d2.y, gyy = s0.y - s2.y;
          lg += fabsf((fabsf(gxx) + fabsf(gyx)) - (fabsf(gxy) + fabsf(gyy)));
        }
        d0 = d1; d1 = d2; s0 = s1; s1 = s2;
      }
    }
    __syncthreads();
  }
  cp_async_wait<0>();
  const int64_t blk = ((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  const float a = block_sum_256(l1, sred);
  const float gsum = block_sum_256(lg, sred);
  if (threadIdx.x == 0) { partials[blk * 2] = a; partials[blk * 2 + 1] = gsum; }
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

// column-streaming launch geometry: strips of 128 columns x row chunks of <= 128 rows (halo rows re-read: 2R / rows)
struct StreamGrid { dim3 grid; int rows_per_chunk; int per_image; };
static StreamGrid stream_grid(int B, int H, int W) {
  StreamGrid g;
  const int strips = (W + kStripW - 1) / kStripW;
  int chunks = (H + 127) / 128;
  // small images: split rows further so that at least ~2 blocks per SM exist (never below 32 rows per chunk)
  while ((int64_t)strips * chunks * B < 2 * 148 && (H + chunks) / (chunks + 1) >= 32) ++chunks;
  g.rows_per_chunk = (H + chunks - 1) / chunks;
  chunks = (H + g.rows_per_chunk - 1) / g.rows_per_chunk;
  g.grid = dim3(strips, chunks, B);
  g.per_image = strips * chunks;
  return g;
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
  const StreamGrid sg = stream_grid(B, H, W);
  cudaStream_t st = as_stream(stream);
  ssim_kernel<<<sg.grid, 128, 0, st>>>(img1, img2, H, W, sg.rows_per_chunk, win, partial_area(workspace));
  int rc = check_launch("segmif_ssim_fwd");
  if (rc) return rc;
  const int per = sg.per_image;
  if (per_image) return finish(workspace, B, per, 1, 0, 1.0 / ((double)H * W), out, st, "segmif_ssim_fwd");
  return finish(workspace, 1, per * B, 1, 0, 1.0 / ((double)B * H * W), out, st, "segmif_ssim_fwd");
}

static LapTaps make_lap_taps() {
  // lap_loss.py:39-60 builds exp(-(dx^2+dy^2)/(2 sigma^2)) / (2 pi sigma^2) on a (k,k) grid and divides by its sum; that is
  // the outer product of e_i / sum(e) with e_i = exp(-(i - mean)^2 / (2 sigma^2)), sigma = 2 (the constant cancels)
  LapTaps k;
  const int sizes[3] = {3, 5, 7};
  float* dst[3] = {k.g3, k.g5, k.g7};
  for (int s = 0; s < 3; ++s) {
    const int n = sizes[s];
    const double mean = (n - 1) / 2.0;
    double e[7], sum = 0.0;
    for (int i = 0; i < n; ++i) { e[i] = exp(-(i - mean) * (i - mean) / 8.0); sum += e[i]; }
    for (int i = 0; i < n; ++i) dst[s][i] = (float)(e[i] / sum);
  }
  return k;
}

extern "C" int segmif_laploss2_fwd(const float* inp, const float* ir, const float* vis, int B, int H, int W,
                                   float* workspace, float* out, segmif_stream_t stream) {
  SEGMIF_REQUIRE(inp && ir && vis && workspace && out, "laploss2: null pointer");
  SEGMIF_REQUIRE(B > 0 && H > 0 && W > 0, "laploss2: empty input");
  static const LapTaps ker = make_lap_taps();
  const StreamGrid sg = stream_grid(B, H, W);
  cudaStream_t st = as_stream(stream);
  laploss_kernel<3><<<sg.grid, 128, 0, st>>>(inp, ir, vis, H, W, sg.rows_per_chunk, ker, partial_area(workspace));
  int rc = check_launch("segmif_laploss2_fwd");
  if (rc) return rc;
  return finish(workspace, 1, sg.per_image * B, 3, 1, 1.0 / ((double)B * H * W), out, st, "segmif_laploss2_fwd");
}

extern "C" int segmif_laploss_fwd(const float* inp, const float* target, int B, int H, int W, float* workspace,
                                  float* out, segmif_stream_t stream) {
  SEGMIF_REQUIRE(inp && target && workspace && out, "laploss: null pointer");
  SEGMIF_REQUIRE(B > 0 && H > 0 && W > 0, "laploss: empty input");
  static const LapTaps ker = make_lap_taps();
  const StreamGrid sg = stream_grid(B, H, W);
  cudaStream_t st = as_stream(stream);
  laploss_kernel<2><<<sg.grid, 128, 0, st>>>(inp, target, nullptr, H, W, sg.rows_per_chunk, ker, partial_area(workspace));
  int rc = check_launch("segmif_laploss_fwd");
  if (rc) return rc;
  return finish(workspace, 1, sg.per_image * B, 3, 1, 1.0 / ((double)B * H * W), out, st, "segmif_laploss_fwd");
}

extern "C" int segmif_entropy_fwd(const float* img, int B, int H, int W, int patch, float* workspace, float* out,
                                  segmif_stream_t stream) {
  SEGMIF_REQUIRE(img && workspace && out, "entropy: null pointer");
  SEGMIF_REQUIRE(patch == 2 || patch == 4 || patch == 8 || patch == 16, "entropy: patch size %d unsupported (2,4,8,16)", patch);
  SEGMIF_REQUIRE(H % patch == 0 && W % patch == 0 && B > 0, "entropy: H and W must be multiples of the patch size");
  Bins32 bins;   // torch.linspace(0, 1, 32) in fp32: symmetric two-sided formula
  const float step = 1.0f / 31.0f;
  for (int i = 0; i < 32; ++i) bins.b[i] = i < 16 ? 0.0f + step * (float)i : 1.0f - step * (float)(31 - i);
  cudaStream_t st = as_stream(stream);
  float* part = partial_area(workspace);
  if (patch >= 4 && W % 4 == 0 && ((uintptr_t)img & 15) == 0) {
    const int64_t npatch = (int64_t)B * (H / patch) * (W / patch);
    const int nb = (int)std::min<int64_t>(148 * 12, (npatch + 127) / 128);
    switch (patch) {
      case 4: entropy_patch_kernel<4><<<nb, 128, 0, st>>>(img, B, H, W, bins, part); break;
      case 8: entropy_patch_kernel<8><<<nb, 128, 0, st>>>(img, B, H, W, bins, part); break;
      default: entropy_patch_kernel<16><<<nb, 128, 0, st>>>(img, B, H, W, bins, part); break;
    }
    int rc2 = check_launch("segmif_entropy_fwd");
    if (rc2) return rc2;
    return finish(workspace, 1, nb, 1, 2, 1.0, out, st, "segmif_entropy_fwd");
  }
  const int nblocks = 148 * 8;
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
  const StreamGrid sg = stream_grid(B, H, W);
  cudaStream_t st = as_stream(stream);
  sobel_l1_kernel<<<sg.grid, 128, 0, st>>>(x, y, H, W, sg.rows_per_chunk, partial_area(workspace));
  int rc = check_launch("segmif_sobel_l1_fwd");
  if (rc) return rc;
  return finish(workspace, 1, sg.per_image * B, 2, 3, 1.0 / (double)n, out, st, "segmif_sobel_l1_fwd");
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

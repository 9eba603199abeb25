// Training data path on the device: what the reference's DataLoader workers do per sample on the host
// with Pillow, OpenCV (through mmcv) and numpy -- datasets/voc_fusion3.py:169-209, datasets/imutils.py:34-49,69-91,121-129,
// 199-249,295-391 -- restated so that every output BYTE is the one the reference produces:
//   resize_tables   Pillow Resample.c precompute_coeffs + normalize_coeffs_8bpc (triangle filter, 22-bit fixed point) and the
//                   accumulated-coordinate index walk of Geometry.c (NEAREST), in IEEE double like the C code
//   resize_h / _v   ImagingResampleHorizontal_8bpc / Vertical_8bpc: uint8 result of the horizontal pass feeds the vertical one
//   label_pad       NEAREST-resized, flipped label placed at (H_pad, W_pad) inside the ignore_index canvas (random_crop2)
//   window_stats    np.unique(...) of each of the ten candidate crop windows: #classes, largest class, non-ignored pixels
//   finish          flip -> PhotoMetricDistortion (convert / OpenCV BGR<->HSV, uint8 and float32 flavours, including OpenCV's
//                   vector-body / scalar-tail rounding difference by column) -> mean_rgb canvas -> crop -> / 255.0 -> CHW
// Only the random draws (Python `random`, `np.random`, in the reference's order) and the accept/reject decision of the crop
// window stay on the host (segmif_b200/datasets/imutils.py).  All kernels are batched over samples (grid.z) and HBM-bound
// byte work: one thread per output pixel, planar uint8 intermediates so that warps read and write whole 32-byte sectors.
#include <algorithm>

#include "common.cuh"

namespace segmif {

constexpr int kPrecisionBits = 32 - 8 - 2;   // Resample.c PRECISION_BITS

// The 352-byte descriptor is staged once per block in shared memory: a per-thread copy would live in local memory (the op arrays
// are indexed dynamically) and cost more traffic than the pixels.
__device__ __forceinline__ const segmif_dp_sample& stage_sample(const segmif_dp_sample* __restrict__ samples, int idx, segmif_dp_sample* sh) {
  static_assert(sizeof(segmif_dp_sample) % 4 == 0, "descriptor is copied as 32-bit words");
  const uint32_t* src = reinterpret_cast<const uint32_t*>(samples + idx);
  uint32_t* dst = reinterpret_cast<uint32_t*>(sh);
  for (int i = threadIdx.x; i < (int)(sizeof(segmif_dp_sample) / 4); i += blockDim.x) dst[i] = src[i];
  __syncthreads();
  return *sh;
}

__device__ __forceinline__ const int32_t* tab(const int32_t* arena, const segmif_dp_sample& s, int which) {
  // [xmin_x nw][xcnt_x nw][k_x nw*ks_x][xmin_y nh][xcnt_y nh][k_y nh*ks_y][near_x nw][near_y nh]
  const int32_t* p = arena + s.tab_off;
  if (which == 0) return p;
  p += s.nw;
  if (which == 1) return p;
  p += s.nw;
  if (which == 2) return p;
  p += (int64_t)s.nw * s.ks_x;
  if (which == 3) return p;
  p += s.nh;
  if (which == 4) return p;
  p += s.nh;
  if (which == 5) return p;
  p += (int64_t)s.nh * s.ks_y;
  if (which == 6) return p;
  return p + s.nw;
}

// ---------------------------------------------------------------------------------------------------------------- tables
// grid (ceil(max(nh, nw) / 128), 2, n): axis 0 = x (W -> nw), 1 = y (H -> nh).  Thread (axis, xx) writes one coefficient row;
// thread 0 of block 0 of each axis walks the NEAREST index table sequentially (the reference accumulates `xo += a0`).
__global__ void __launch_bounds__(128) resize_tables_kernel(const segmif_dp_sample* __restrict__ samples, int32_t* __restrict__ arena) {
  __shared__ segmif_dp_sample sh_sample;
  const segmif_dp_sample& s = stage_sample(samples, blockIdx.z, &sh_sample);
  if (!s.resized) return;
  const int axis = blockIdx.y;
  const int in_size = axis == 0 ? s.W : s.H, out_size = axis == 0 ? s.nw : s.nh, ks = axis == 0 ? s.ks_x : s.ks_y;
  int32_t* xmin_t = const_cast<int32_t*>(tab(arena, s, axis == 0 ? 0 : 3));
  int32_t* xcnt_t = const_cast<int32_t*>(tab(arena, s, axis == 0 ? 1 : 4));
  int32_t* k_t = const_cast<int32_t*>(tab(arena, s, axis == 0 ? 2 : 5));
  int32_t* near_t = const_cast<int32_t*>(tab(arena, s, axis == 0 ? 6 : 7));
  const double scale = __ddiv_rn((double)in_size, (double)out_size);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double xo = __dmul_rn(scale, 0.5);
    for (int x = 0; x < out_size; ++x) {
      int xin = (int)xo;                         // COORD(): truncation
      near_t[x] = min(max(xin, 0), in_size - 1);
      xo = __dadd_rn(xo, scale);
    }
  }
  const int xx = blockIdx.x * 128 + threadIdx.x;
  if (xx >= out_size) return;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = filterscale;            // bilinear support 1.0 * filterscale
  const double ss = __ddiv_rn(1.0, filterscale);
  const double center = __dmul_rn(__dadd_rn((double)xx, 0.5), scale);
  int lo = (int)__dadd_rn(__dsub_rn(center, support), 0.5);
  if (lo < 0) lo = 0;
  int hi = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
  if (hi > in_size) hi = in_size;
  const int n = hi - lo;
  double ww = 0.0;
  for (int x = 0; x < n; ++x) {
    double a = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + lo), center), 0.5), ss);
    if (a < 0.0) a = -a;
    const double w = a < 1.0 ? __dsub_rn(1.0, a) : 0.0;
    ww = __dadd_rn(ww, w);
  }
  for (int x = 0; x < ks; ++x) {
    int32_t kq = 0;
    if (x < n) {
      double a = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + lo), center), 0.5), ss);
      if (a < 0.0) a = -a;
      double w = a < 1.0 ? __dsub_rn(1.0, a) : 0.0;
      if (ww != 0.0) w = __ddiv_rn(w, ww);
      const double q = __dmul_rn(w, (double)(1 << kPrecisionBits));
      kq = w < 0.0 ? (int32_t)__dadd_rn(-0.5, q) : (int32_t)__dadd_rn(0.5, q);
    }
    k_t[(int64_t)xx * ks + x] = kq;
  }
  xmin_t[xx] = lo;
  xcnt_t[xx] = n;
}

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= kPrecisionBits;
  return (uint8_t)min(max(v, 0), 255);
}

// ---------------------------------------------------------------------------------------------------------------- resize
// Row pitch of every uint8 intermediate (label canvas, horizontal-pass output, resized region): the width rounded up to 16, so
// that a thread can move four neighbouring pixels with one aligned 32-bit access.
__host__ __device__ __forceinline__ int pitch16(int w) { return (w + 15) & ~15; }

// Horizontal pass over the source rows the vertical pass will need [src_y0, src_y1) and the output columns [roi_x0, roi_x1):
// tmp is PLANAR uint8 [4 + mask_c][rows][pitch] (ir, vis c0, vis c1, vis c2, mask c0 [, c1, c2]).  Four output columns per thread.
// grid (ceil(pitch/4/128), rows, n).
__global__ void __launch_bounds__(128) resize_h_kernel(const segmif_dp_sample* __restrict__ samples, const int32_t* __restrict__ arena,
                                                       uint8_t* __restrict__ tmp_arena) {
  __shared__ segmif_dp_sample sh_sample;
  const segmif_dp_sample& s = stage_sample(samples, blockIdx.z, &sh_sample);
  if (!s.resized) return;
  const int rows = s.src_y1 - s.src_y0, roi_w = s.roi_x1 - s.roi_x0, pitch = pitch16(roi_w);
  const int r = blockIdx.y, x4 = (blockIdx.x * 128 + threadIdx.x) * 4;
  if (r >= rows || x4 >= roi_w) return;
  const int y = s.src_y0 + r;
  const int32_t* xmin_t = tab(arena, s, 0);
  const int32_t* xcnt_t = tab(arena, s, 1);
  const int32_t* k_t = tab(arena, s, 2);
  const int mc = s.mask_c;                 // 1: plane (voc_fusion3.py), 3: H x W x 3 image (voc_fusion2.py)
  uint32_t pk[7] = {0u, 0u, 0u, 0u, 0u, 0u, 0u};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (x4 + q >= roi_w) break;
    const int xx = s.roi_x0 + x4 + q;
    const int lo = xmin_t[xx], n = xcnt_t[xx];
    const int32_t* k = k_t + (int64_t)xx * s.ks_x;
    const uint8_t* ir = s.ir + (int64_t)y * s.W + lo;
    const uint8_t* vis = s.vis + ((int64_t)y * s.W + lo) * 3;
    const uint8_t* mk = s.mask + ((int64_t)y * s.W + lo) * mc;
    int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0, a3 = a0, a4 = a0, a5 = a0, a6 = a0;
    for (int j = 0; j < n; ++j) {
      const int kj = k[j];
      a0 += kj * ir[j];
      a1 += kj * vis[3 * j];
      a2 += kj * vis[3 * j + 1];
      a3 += kj * vis[3 * j + 2];
      a4 += kj * mk[mc * j];
      if (mc == 3) {
        a5 += kj * mk[3 * j + 1];
        a6 += kj * mk[3 * j + 2];
      }
    }
    pk[0] |= (uint32_t)clip8(a0) << (8 * q);
    pk[1] |= (uint32_t)clip8(a1) << (8 * q);
    pk[2] |= (uint32_t)clip8(a2) << (8 * q);
    pk[3] |= (uint32_t)clip8(a3) << (8 * q);
    pk[4] |= (uint32_t)clip8(a4) << (8 * q);
    pk[5] |= (uint32_t)clip8(a5) << (8 * q);
    pk[6] |= (uint32_t)clip8(a6) << (8 * q);
  }
  uint8_t* t = tmp_arena + s.tmp_off + (int64_t)r * pitch + x4;
  const int64_t plane = (int64_t)rows * pitch;
#pragma unroll
  for (int c = 0; c < 7; ++c)
    if (c < 4 + mc) *reinterpret_cast<uint32_t*>(t + c * plane) = pk[c];
}

// Vertical pass: resized region, planar uint8 [4 + mask_c][roi_h][pitch], four columns per thread.  grid (ceil(pitch/4/128), roi_h, n).
__global__ void __launch_bounds__(128) resize_v_kernel(const segmif_dp_sample* __restrict__ samples, const int32_t* __restrict__ arena,
                                                       const uint8_t* __restrict__ tmp_arena, uint8_t* __restrict__ rs_arena) {
  __shared__ segmif_dp_sample sh_sample;
  const segmif_dp_sample& s = stage_sample(samples, blockIdx.z, &sh_sample);
  if (!s.resized) return;
  const int rows = s.src_y1 - s.src_y0, roi_w = s.roi_x1 - s.roi_x0, roi_h = s.roi_y1 - s.roi_y0, pitch = pitch16(roi_w);
  const int yo = blockIdx.y, x4 = (blockIdx.x * 128 + threadIdx.x) * 4;
  if (yo >= roi_h || x4 >= roi_w) return;
  const int yy = s.roi_y0 + yo;
  const int lo = tab(arena, s, 3)[yy], n = tab(arena, s, 4)[yy];
  const int32_t* k = tab(arena, s, 5) + (int64_t)yy * s.ks_y;
  const int64_t tplane = (int64_t)rows * pitch, oplane = (int64_t)roi_h * pitch;
  const uint8_t* t = tmp_arena + s.tmp_off + (int64_t)(lo - s.src_y0) * pitch + x4;
  uint8_t* o = rs_arena + s.rs_off + (int64_t)yo * pitch + x4;
#pragma unroll
  for (int c = 0; c < 7; ++c) {
    if (c >= 4 + s.mask_c) break;
    int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0, a3 = a0;
    for (int j = 0; j < n; ++j) {
      const uint32_t v = *reinterpret_cast<const uint32_t*>(t + c * tplane + (int64_t)j * pitch);
      const int kj = k[j];
      a0 += kj * (int)(v & 255u);
      a1 += kj * (int)((v >> 8) & 255u);
      a2 += kj * (int)((v >> 16) & 255u);
      a3 += kj * (int)(v >> 24);
    }
    *reinterpret_cast<uint32_t*>(o + c * oplane) =
        (uint32_t)clip8(a0) | ((uint32_t)clip8(a1) << 8) | ((uint32_t)clip8(a2) << 16) | ((uint32_t)clip8(a3) << 24);
  }
}

// ---------------------------------------------------------------------------------------------------------------- labels
// pad_label of random_crop2 as uint8 [PH][pitch16(PW)] (the reference holds it as float32; the values are the same integers).
// Four pixels per thread, one 32-bit store.  grid (ceil(max pitch / 4 / 128), max PH, n).
__global__ void __launch_bounds__(128) label_pad_kernel(const segmif_dp_sample* __restrict__ samples, const int32_t* __restrict__ arena,
                                                        uint8_t* __restrict__ lab_arena, int ignore_index) {
  __shared__ segmif_dp_sample sh_sample;
  const segmif_dp_sample& s = stage_sample(samples, blockIdx.z, &sh_sample);
  const int pitch = pitch16(s.PW);
  const int y = blockIdx.y, x4 = (blockIdx.x * 128 + threadIdx.x) * 4;
  if (y >= s.PH || x4 >= pitch) return;
  const int ys = y - s.pad_h;
  const bool row_in = ys >= 0 && ys < s.nh;
  const uint8_t* src_row = row_in ? s.label + (int64_t)(s.resized ? tab(arena, s, 7)[ys] : ys) * s.W : nullptr;
  const int32_t* near_x = tab(arena, s, 6);
  uint32_t pk = 0u;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int xs = x4 + q - s.pad_w;
    uint32_t v = (uint32_t)ignore_index;
    if (row_in && xs >= 0 && xs < s.nw) {
      const int xr = s.flip ? s.nw - 1 - xs : xs;
      v = src_row[s.resized ? near_x[xr] : xr];
    }
    pk |= v << (8 * q);
  }
  *reinterpret_cast<uint32_t*>(lab_arena + s.lab_off + (int64_t)y * pitch + x4) = pk;
}

// Histogram of each candidate window: grid (10 candidates, kHistSlices row slices, n).  Labels are piecewise constant, so each
// thread walks 32 consecutive pixels and issues one shared-memory atomic per RUN of equal values; non-empty bins are then added
// to hist_ws [n][10][256] (zeroed by the entry point).
constexpr int kHistSlices = 8;
__global__ void __launch_bounds__(256) window_hist_kernel(const segmif_dp_sample* __restrict__ samples, const uint8_t* __restrict__ lab_arena,
                                                          int crop, unsigned int* __restrict__ hist_ws) {
  __shared__ segmif_dp_sample sh_sample;
  const segmif_dp_sample& s = stage_sample(samples, blockIdx.z, &sh_sample);
  const int cand = blockIdx.x;
  __shared__ unsigned int hist[256];
  hist[threadIdx.x] = 0u;
  __syncthreads();
  const int pitch = pitch16(s.PW);
  const int rps = (crop + kHistSlices - 1) / kHistSlices, r0 = blockIdx.y * rps, r1 = min(crop, r0 + rps);
  const int segs = (crop + 31) / 32, total = max(r1 - r0, 0) * segs;
  const uint8_t* lab = lab_arena + s.lab_off + (int64_t)s.cand_hs[cand] * pitch + s.cand_ws[cand];
  for (int t = threadIdx.x; t < total; t += 256) {
    const int row = r0 + t / segs, x0 = (t % segs) * 32, x1 = min(crop, x0 + 32);
    const uint8_t* p = lab + (int64_t)row * pitch;
    unsigned cur = p[x0], cnt = 1;
    for (int x = x0 + 1; x < x1; ++x) {
      const unsigned v = p[x];
      if (v == cur) {
        ++cnt;
      } else {
        atomicAdd(&hist[cur], cnt);
        cur = v;
        cnt = 1;
      }
    }
    atomicAdd(&hist[cur], cnt);
  }
  __syncthreads();
  if (hist[threadIdx.x]) atomicAdd(hist_ws + ((int64_t)blockIdx.z * 10 + cand) * 256 + threadIdx.x, hist[threadIdx.x]);
}

// One block per (candidate, sample): {#values != ignore present, count of the most frequent one, non-ignored pixels}.
__global__ void __launch_bounds__(256) window_stats_kernel(const unsigned int* __restrict__ hist_ws, int ignore_index, int32_t* __restrict__ stats) {
  const int cand = blockIdx.x;
  unsigned int c = threadIdx.x == (unsigned)ignore_index ? 0u : hist_ws[((int64_t)blockIdx.y * 10 + cand) * 256 + threadIdx.x];
  unsigned int nz = c ? 1u : 0u, mx = c, sm = c;
  for (int o = 16; o; o >>= 1) {
    nz += __shfl_xor_sync(0xffffffffu, nz, o);
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    sm += __shfl_xor_sync(0xffffffffu, sm, o);
  }
  __shared__ unsigned int red[3][8];
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = nz;
    red[1][threadIdx.x >> 5] = mx;
    red[2][threadIdx.x >> 5] = sm;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int a = 0, b = 0, d = 0;
    for (int w = 0; w < 8; ++w) {
      a += red[0][w];
      b = max(b, red[1][w]);
      d += red[2][w];
    }
    int32_t* o = stats + ((int64_t)blockIdx.y * 10 + cand) * 3;
    o[0] = (int32_t)a;
    o[1] = (int32_t)b;
    o[2] = (int32_t)d;
  }
}

// ---------------------------------------------------------------------------------------------------------------- colour
// OpenCV color_hsv.simd.hpp restated (see oracle/data_oracle.py for the pinning): `tail8` / `tail32` say whether the pixel's
// COLUMN falls into the scalar tail of OpenCV's row loop (x >= (W / lanes) * lanes), which rounds differently from the body.
struct Px {
  float b, g, r;          // channel 0, 1, 2 of the array handed to mmcv (the reference passes RGB; the maths is positional)
};

__constant__ int kHsvSector[6][3] = {{1, 3, 0}, {1, 0, 2}, {3, 0, 1}, {0, 2, 1}, {0, 1, 3}, {2, 1, 0}};

__device__ __forceinline__ int hsv_div_table(int num, double den) {   // cvRound((num << 12) / den), ties to even like cvRound
  return (int)rint(__ddiv_rn((double)(num << 12), den));
}

// div_tab: [0,256) sdiv, [256,512) hdiv (OpenCV builds the same tables once; here once per block that needs them)
__device__ __forceinline__ void bgr2hsv_u8(int b, int g, int r, const int* __restrict__ div_tab, int& h, int& s, int& v) {
  v = max(max(b, g), r);
  const int vmin = min(min(b, g), r);
  const int diff = v - vmin;
  const int vr = v == r ? -1 : 0, vg = v == g ? -1 : 0;
  const int sdiv = div_tab[v];
  const int hdiv = div_tab[256 + diff];
  s = (diff * sdiv + (1 << 11)) >> 12;
  h = (vr & (g - b)) + (~vr & ((vg & (b - r + 2 * diff)) + ((~vg) & (r - g + 4 * diff))));
  h = (h * hdiv + (1 << 11)) >> 12;
  h += h < 0 ? 180 : 0;
}

__device__ __forceinline__ Px hsv2bgr_core(float hh, float s, float v) {
  const float pre = floorf(hh);
  const float f = __fsub_rn(hh, pre);
  int sector = (int)pre % 6;
  float t[4];
  t[0] = v;
  t[1] = __fmul_rn(v, __fsub_rn(1.f, s));
  t[2] = __fmul_rn(v, __fmaf_rn(-s, f, 1.f));
  t[3] = __fmul_rn(v, __fmaf_rn(-s, __fsub_rn(1.f, f), 1.f));
  Px o;
  o.b = t[kHsvSector[sector][0]];
  o.g = t[kHsvSector[sector][1]];
  o.r = t[kHsvSector[sector][2]];
  return o;
}

__device__ __forceinline__ void hsv2bgr_u8(int h, int s, int v, bool tail32, int& b, int& g, int& r) {
  const Px o = hsv2bgr_core(__fmul_rn((float)h, (float)(6.0 / 180.0)), __fmul_rn((float)s, (float)(1 / 255.0)),
                            __fmul_rn((float)v, (float)(1 / 255.0)));
  const float fb = __fmul_rn(o.b, 255.f), fg = __fmul_rn(o.g, 255.f), fr = __fmul_rn(o.r, 255.f);
  b = min(max((int)(tail32 ? rintf(fb) : floorf(fb)), 0), 255);
  g = min(max((int)(tail32 ? rintf(fg) : floorf(fg)), 0), 255);
  r = min(max((int)(tail32 ? rintf(fr) : floorf(fr)), 0), 255);
}

__device__ __forceinline__ void bgr2hsv_f32(Px p, bool tail8, float& h, float& s, float& v) {
  v = fmaxf(fmaxf(p.b, p.g), p.r);
  const float vmin = fminf(fminf(p.b, p.g), p.r);
  const float diff = __fsub_rn(v, vmin);
  const float eps = 1.1920928955078125e-07f;
  s = __fdiv_rn(diff, __fadd_rn(fabsf(v), eps));
  const float d = __fdiv_rn(60.f, __fadd_rn(diff, eps));
  if (v == p.r) {
    h = __fmul_rn(__fsub_rn(p.g, p.b), d);
    if (h < 0.f) h = tail8 ? __fadd_rn(h, 360.f) : __fmaf_rn(__fsub_rn(p.g, p.b), d, 360.f);
  } else {
    h = v == p.g ? __fmaf_rn(__fsub_rn(p.b, p.r), d, 120.f) : __fmaf_rn(__fsub_rn(p.r, p.g), d, 240.f);
    if (h < 0.f) h = __fadd_rn(h, 360.f);
  }
}

__device__ __forceinline__ Px hsv2bgr_f32(float h, float s, float v) {
  if (s == 0.f) return Px{v, v, v};
  return hsv2bgr_core(__fmul_rn(h, (float)(6.0 / 360.0)), s, v);
}

__device__ __forceinline__ float convert_u8(float x, float alpha, float beta) {   // imutils.py:308-312
  const float y = __fadd_rn(__fmul_rn(x, alpha), beta);
  return (float)(int)fminf(fmaxf(y, 0.f), 255.f);
}

__device__ __forceinline__ int py_mod180(int v) {
  v %= 180;
  return v < 0 ? v + 180 : v;
}

// PhotoMetricDistortion on one pixel.  The dtype of the image (uint8 or float32) at each op is resolved by the host
// (op_u8[i]); values of a uint8 image are held as exact small integers in the float registers.
__device__ __forceinline__ Px photometric(Px p, const segmif_dp_sample& s, const int* __restrict__ div_tab, bool tail8, bool tail32) {
  for (int i = 0; i < s.n_ops; ++i) {
    const int kind = s.op_kind[i];
    if (kind == SEGMIF_DP_OP_CONVERT) {
      p.b = convert_u8(p.b, s.op_alpha[i], s.op_beta[i]);
      p.g = convert_u8(p.g, s.op_alpha[i], s.op_beta[i]);
      p.r = convert_u8(p.r, s.op_alpha[i], s.op_beta[i]);
    } else if (s.op_u8[i]) {
      int h, sa, v, b, g, r;
      bgr2hsv_u8((int)p.b, (int)p.g, (int)p.r, div_tab, h, sa, v);
      if (kind == SEGMIF_DP_OP_SATURATION) sa = (int)convert_u8((float)sa, s.op_alpha[i], 0.f);
      else h = py_mod180(h + s.op_delta[i]);
      hsv2bgr_u8(h, sa, v, tail32, b, g, r);
      p = Px{(float)b, (float)g, (float)r};
    } else {
      float h, sa, v;
      bgr2hsv_f32(p, tail8, h, sa, v);
      if (kind == SEGMIF_DP_OP_SATURATION) sa = convert_u8(sa, s.op_alpha[i], 0.f);     // uint8 0 or 1 written into the float plane
      else h = (float)py_mod180((int)h + s.op_delta[i]);
      p = hsv2bgr_f32(h, sa, v);
    }
  }
  return p;
}

// ---------------------------------------------------------------------------------------------------------------- finish
// One thread per pixel of the crop window: out_* fp32 [n][3][crop][crop] (CHW, / 255.0), label fp32 [n][crop][crop] and, when
// label_i64 is given, the same labels as int64.  grid (ceil(crop*crop/256), 1, n).
__global__ void __launch_bounds__(256) finish_kernel(const segmif_dp_sample* __restrict__ samples, const uint8_t* __restrict__ rs_arena,
                                                     const uint8_t* __restrict__ lab_arena, int crop, float mean0, float mean1, float mean2,
                                                     float* __restrict__ out_ir, float* __restrict__ out_vis, float* __restrict__ out_mask,
                                                     float* __restrict__ out_label, int64_t* __restrict__ label_i64) {
  __shared__ segmif_dp_sample sh_sample;
  const segmif_dp_sample& s = stage_sample(samples, blockIdx.z, &sh_sample);
  __shared__ int div_tab[512];
  bool need_tab = false;
  for (int k = 0; k < s.n_ops; ++k) need_tab |= s.op_kind[k] != SEGMIF_DP_OP_CONVERT && s.op_u8[k];
  if (need_tab) {                                               // uniform per block
    for (int k = threadIdx.x; k < 512; k += 256) {
      const int d = k & 255;
      div_tab[k] = d == 0 ? 0 : (k < 256 ? hsv_div_table(255, (double)d) : hsv_div_table(180, 6.0 * d));
    }
    __syncthreads();
  }
  const int i = blockIdx.x * 256 + threadIdx.x;
  const int n = crop * crop;
  if (i >= n) return;
  const int y = i / crop, x = i - y * crop;
  const int yp = s.hs + y, xp = s.ws + x;                       // canvas coordinates
  const int ys = yp - s.pad_h, xs = xp - s.pad_w;               // coordinates in the (flipped) image PhotoMetricDistortion saw
  const int64_t o = (int64_t)blockIdx.z * 3 * n + i;
  const uint8_t lab = lab_arena[s.lab_off + (int64_t)yp * pitch16(s.PW) + xp];
  out_label[(int64_t)blockIdx.z * n + i] = (float)lab;
  if (label_i64) label_i64[(int64_t)blockIdx.z * n + i] = (int64_t)lab;
  float ir, mk[3];
  Px p;
  if (ys >= 0 && ys < s.nh && xs >= 0 && xs < s.nw) {
    const int xr = s.flip ? s.nw - 1 - xs : xs;
    if (s.resized) {
      const int pitch = pitch16(s.roi_x1 - s.roi_x0), roi_h = s.roi_y1 - s.roi_y0;
      const int64_t plane = (int64_t)roi_h * pitch;
      const uint8_t* q = rs_arena + s.rs_off + (int64_t)(ys - s.roi_y0) * pitch + (xr - s.roi_x0);
      ir = (float)q[0];
      p = Px{(float)q[plane], (float)q[2 * plane], (float)q[3 * plane]};
      mk[0] = (float)q[4 * plane];
      mk[1] = s.mask_c == 3 ? (float)q[5 * plane] : mk[0];
      mk[2] = s.mask_c == 3 ? (float)q[6 * plane] : mk[0];
    } else {
      const int64_t q = (int64_t)ys * s.W + xr;
      ir = (float)s.ir[q];
      p = Px{(float)s.vis[3 * q], (float)s.vis[3 * q + 1], (float)s.vis[3 * q + 2]};
      mk[0] = (float)s.mask[q * s.mask_c];
      mk[1] = s.mask_c == 3 ? (float)s.mask[q * 3 + 1] : mk[0];
      mk[2] = s.mask_c == 3 ? (float)s.mask[q * 3 + 2] : mk[0];
    }
    p = photometric(p, s, div_tab, xs >= (s.nw / 8) * 8, xs >= (s.nw / 32) * 32);
    out_ir[o] = __fdiv_rn(ir, 255.f);
    out_ir[o + n] = __fdiv_rn(ir, 255.f);
    out_ir[o + 2 * n] = __fdiv_rn(ir, 255.f);
    out_mask[o] = __fdiv_rn(mk[0], 255.f);
    out_mask[o + n] = __fdiv_rn(mk[1], 255.f);
    out_mask[o + 2 * n] = __fdiv_rn(mk[2], 255.f);
  } else {
    p = Px{mean0, mean1, mean2};
    const float m0 = __fdiv_rn(mean0, 255.f), m1 = __fdiv_rn(mean1, 255.f), m2 = __fdiv_rn(mean2, 255.f);
    out_ir[o] = m0;
    out_ir[o + n] = m1;
    out_ir[o + 2 * n] = m2;
    out_mask[o] = m0;
    out_mask[o + n] = m1;
    out_mask[o + 2 * n] = m2;
  }
  out_vis[o] = __fdiv_rn(p.b, 255.f);
  out_vis[o + n] = __fdiv_rn(p.g, 255.f);
  out_vis[o + 2 * n] = __fdiv_rn(p.r, 255.f);
}

// aug=False (validation): image / 255.0 on uint8 arrays is a float64 division in numpy.  HWC uint8 -> CHW float64.
__global__ void __launch_bounds__(256) u8_to_chw_f64_kernel(const uint8_t* __restrict__ src, int64_t hw, int C, int rep, double* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= hw) return;
  for (int c = 0; c < 3; ++c) dst[c * hw + i] = __ddiv_rn((double)src[rep ? i : i * C + c], 255.0);
}

}  // namespace segmif

using namespace segmif;

static int max_of(const segmif_dp_sample* host, int n, int (*f)(const segmif_dp_sample&)) {
  int m = 0;
  for (int i = 0; i < n; ++i) m = std::max(m, f(host[i]));
  return m;
}

extern "C" int segmif_dp_label_stage(const segmif_dp_sample* samples_dev, const segmif_dp_sample* samples_host, int n, int crop,
                                     int ignore_index, int32_t* table_arena, uint8_t* label_arena, int32_t* hist_ws, int32_t* stats,
                                     segmif_stream_t stream) {
  SEGMIF_REQUIRE(samples_dev && samples_host && table_arena && label_arena && hist_ws && stats && n > 0 && crop > 0, "dp_label_stage: bad arguments");
  SEGMIF_REQUIRE(ignore_index >= 0 && ignore_index <= 255, "dp_label_stage: ignore_index=%d must fit the uint8 label", ignore_index);
  for (int i = 0; i < n; ++i) {
    const segmif_dp_sample& s = samples_host[i];
    SEGMIF_REQUIRE(s.PH >= crop && s.PW >= crop && s.PH >= s.nh && s.PW >= s.nw, "dp_label_stage: sample %d canvas %dx%d too small", i, s.PH, s.PW);
    SEGMIF_REQUIRE(s.ks_x <= 64 && s.ks_y <= 64, "dp_label_stage: sample %d scale out of range", i);
    for (int c = 0; c < 10; ++c)
      SEGMIF_REQUIRE(s.cand_hs[c] >= 0 && s.cand_hs[c] + crop <= s.PH && s.cand_ws[c] >= 0 && s.cand_ws[c] + crop <= s.PW,
                     "dp_label_stage: sample %d candidate %d outside the canvas", i, c);
  }
  cudaStream_t st = as_stream(stream);
  const int mo = max_of(samples_host, n, [](const segmif_dp_sample& s) { return s.resized ? std::max(s.nh, s.nw) : 0; });
  if (mo > 0) resize_tables_kernel<<<dim3(ceil_div(mo, 128), 2, n), 128, 0, st>>>(samples_dev, table_arena);
  const int mph = max_of(samples_host, n, [](const segmif_dp_sample& s) { return (int)s.PH; });
  const int mpw = max_of(samples_host, n, [](const segmif_dp_sample& s) { return (int)s.PW; });
  label_pad_kernel<<<dim3(ceil_div(pitch16(mpw) / 4, 128), mph, n), 128, 0, st>>>(samples_dev, table_arena, label_arena, ignore_index);
  {
    cudaError_t e = cudaMemsetAsync(hist_ws, 0, (size_t)n * 10 * 256 * sizeof(int32_t), st);
    if (e != cudaSuccess) { set_error("dp_label_stage: cudaMemsetAsync failed: %s", cudaGetErrorString(e)); return SEGMIF_ERR_CUDA; }
  }
  window_hist_kernel<<<dim3(10, kHistSlices, n), 256, 0, st>>>(samples_dev, label_arena, crop, reinterpret_cast<unsigned int*>(hist_ws));
  window_stats_kernel<<<dim3(10, n), 256, 0, st>>>(reinterpret_cast<const unsigned int*>(hist_ws), ignore_index, stats);
  return check_launch("segmif_dp_label_stage");
}

extern "C" int segmif_dp_image_stage(const segmif_dp_sample* samples_dev, const segmif_dp_sample* samples_host, int n, int crop,
                                     const float* mean_rgb, const int32_t* table_arena, uint8_t* tmp_arena, uint8_t* resized_arena,
                                     const uint8_t* label_arena, float* out_ir, float* out_vis, float* out_mask, float* out_label,
                                     int64_t* out_label_i64, segmif_stream_t stream) {
  SEGMIF_REQUIRE(samples_dev && samples_host && mean_rgb && table_arena && label_arena && out_ir && out_vis && out_mask && out_label && n > 0 && crop > 0,
                 "dp_image_stage: bad arguments");
  int rows = 0, roi_w = 0, roi_h = 0;
  for (int i = 0; i < n; ++i) {
    const segmif_dp_sample& s = samples_host[i];
    SEGMIF_REQUIRE(s.hs >= 0 && s.hs + crop <= s.PH && s.ws >= 0 && s.ws + crop <= s.PW, "dp_image_stage: sample %d window outside the canvas", i);
    SEGMIF_REQUIRE(s.n_ops >= 0 && s.n_ops <= SEGMIF_DP_MAX_OPS, "dp_image_stage: sample %d has %d ops", i, s.n_ops);
    SEGMIF_REQUIRE(s.mask_c == 1 || s.mask_c == 3, "dp_image_stage: sample %d mask_c=%d must be 1 or 3", i, s.mask_c);
    if (!s.resized) continue;
    SEGMIF_REQUIRE(tmp_arena && resized_arena, "dp_image_stage: resize workspaces missing");
    SEGMIF_REQUIRE(0 <= s.roi_x0 && s.roi_x0 < s.roi_x1 && s.roi_x1 <= s.nw && 0 <= s.roi_y0 && s.roi_y0 < s.roi_y1 && s.roi_y1 <= s.nh &&
                       0 <= s.src_y0 && s.src_y0 < s.src_y1 && s.src_y1 <= s.H,
                   "dp_image_stage: sample %d has an empty or out-of-range region", i);
    rows = std::max(rows, s.src_y1 - s.src_y0);
    roi_w = std::max(roi_w, s.roi_x1 - s.roi_x0);
    roi_h = std::max(roi_h, s.roi_y1 - s.roi_y0);
  }
  cudaStream_t st = as_stream(stream);
  if (rows > 0) {
    resize_h_kernel<<<dim3(ceil_div(pitch16(roi_w) / 4, 128), rows, n), 128, 0, st>>>(samples_dev, table_arena, tmp_arena);
    resize_v_kernel<<<dim3(ceil_div(pitch16(roi_w) / 4, 128), roi_h, n), 128, 0, st>>>(samples_dev, table_arena, tmp_arena, resized_arena);
  }
  finish_kernel<<<dim3(ceil_div(crop * crop, 256), 1, n), 256, 0, st>>>(samples_dev, resized_arena, label_arena, crop, mean_rgb[0], mean_rgb[1],
                                                                     mean_rgb[2], out_ir, out_vis, out_mask, out_label, out_label_i64);
  return check_launch("segmif_dp_image_stage");
}

extern "C" int segmif_dp_u8_to_chw_f64(const unsigned char* src, int H, int W, int C, double* dst, segmif_stream_t stream) {
  SEGMIF_REQUIRE(src && dst && H > 0 && W > 0 && (C == 1 || C == 3), "dp_u8_to_chw_f64: bad arguments (C must be 1 or 3)");
  const int64_t hw = (int64_t)H * W;
  u8_to_chw_f64_kernel<<<(unsigned)ceil_div(hw, (int64_t)256), 256, 0, as_stream(stream)>>>(src, hw, C, C == 1, dst);
  return check_launch("segmif_dp_u8_to_chw_f64");
}

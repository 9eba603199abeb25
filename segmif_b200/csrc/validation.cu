// Validation / post-processing on the device: what the reference does per image on the host with
// numpy + sklearn after a device->host copy.
//   confusion_matrix   test_segmentation.py:173-176 -> sklearn.metrics.confusion_matrix(label, prediction, labels=[0..nc-1]),
//                      accumulated over the test set (conf_total += conf); rows = ground truth, columns = prediction;
//                      pairs with either value outside [0, nc) are dropped, as sklearn does for values not in `labels`
//   fused_to_uint8     val_performance.py:447-460: clamp to [0,1] -> uint8(255 x) -> NHWC -> (u - min) / (max - min) over the
//                      whole batch -> uint8(255 y).  The reference's numpy arithmetic is reproduced exactly: truncating
//                      casts, the renormalisation in double precision.
#include <algorithm>

#include "common.cuh"

namespace segmif {

// block-private histogram of nc*nc bins in shared memory, one 64-bit global atomic per non-empty bin and block
__global__ void __launch_bounds__(256) confusion_kernel(const int64_t* __restrict__ truth, const int64_t* __restrict__ pred,
                                                        int64_t n, int nc, unsigned long long* __restrict__ conf) {
  extern __shared__ unsigned int hist[];
  const int bins = nc * nc;
  for (int i = threadIdx.x; i < bins; i += 256) hist[i] = 0u;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const int64_t t = truth[i], p = pred[i];
    if ((uint64_t)t < (uint64_t)nc && (uint64_t)p < (uint64_t)nc) atomicAdd(&hist[(int)t * nc + (int)p], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < bins; i += 256)
    if (hist[i]) atomicAdd(conf + i, (unsigned long long)hist[i]);
}

__global__ void minmax_init_kernel(unsigned int* mm) { mm[0] = 255u; mm[1] = 0u; }

// pass 1: u = uint8(255 * clamp(x, 0, 1)) for the NCHW fp32 image; batch-wide min / max of u
__global__ void __launch_bounds__(256) to_uint8_minmax_kernel(const float* __restrict__ rgb, int64_t n, unsigned int* __restrict__ mm) {
  unsigned int lo = 255u, hi = 0u;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const float v = fminf(fmaxf(rgb[i], 0.f), 1.f);
    const unsigned int u = (unsigned int)(255.0f * v);          // np.uint8(255.0 * float32 array): float32 product, truncation
    lo = min(lo, u);
    hi = max(hi, u);
  }
  for (int o = 16; o; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(mm, lo);
    atomicMax(mm + 1, hi);
  }
}

// pass 2: out[b][y][x][c] = uint8(255.0 * ((u - min) / (max - min)))   (NCHW fp32 -> NHWC uint8)
__global__ void __launch_bounds__(256) to_uint8_renorm_kernel(const float* __restrict__ rgb, const unsigned int* __restrict__ mm,
                                                              unsigned char* __restrict__ out, int B, int64_t HW) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= B * HW) return;
  const int64_t b = i / HW, p = i - b * HW;
  const unsigned int lo = mm[0], hi = mm[1];
  const double den = (double)(hi - lo);                         // 0 -> inf / nan in numpy; a constant image maps to 0 here
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = fminf(fmaxf(rgb[(b * 3 + c) * HW + p], 0.f), 1.f);
    const unsigned int u = (unsigned int)(255.0f * v);
    const double y = den > 0.0 ? (double)(u - lo) / den : 0.0;
    out[i * 3 + c] = (unsigned char)(255.0 * y);
  }
}

}  // namespace segmif

using namespace segmif;

extern "C" int segmif_confusion_matrix(const int64_t* truth, const int64_t* pred, int64_t n, int num_classes, int64_t* conf,
                                       segmif_stream_t stream) {
  SEGMIF_REQUIRE(truth && pred && conf && n >= 0, "confusion_matrix: bad arguments");
  SEGMIF_REQUIRE(num_classes > 0 && num_classes <= 64, "confusion_matrix: num_classes=%d must be in [1, 64]", num_classes);
  if (n == 0) return SEGMIF_OK;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, (int64_t)256 * 16), 148 * 4));
  confusion_kernel<<<grid, 256, (size_t)num_classes * num_classes * sizeof(unsigned int), as_stream(stream)>>>(
      truth, pred, n, num_classes, reinterpret_cast<unsigned long long*>(conf));
  return check_launch("segmif_confusion_matrix");
}

/* minmax: two uint32 of workspace; written here (initialised to {255, 0} by this call) */
extern "C" int segmif_fused_to_uint8(const float* rgb, unsigned char* out_nhwc, unsigned int* minmax, int B, int64_t HW,
                                     segmif_stream_t stream) {
  SEGMIF_REQUIRE(rgb && out_nhwc && minmax && B > 0 && HW > 0, "fused_to_uint8: bad arguments");
  cudaStream_t st = as_stream(stream);
  minmax_init_kernel<<<1, 1, 0, st>>>(minmax);
  const int64_t n = (int64_t)B * 3 * HW;
  to_uint8_minmax_kernel<<<(int)std::min<int64_t>(ceil_div(n, (int64_t)256 * 8), 148 * 8), 256, 0, st>>>(rgb, n, minmax);
  to_uint8_renorm_kernel<<<(unsigned)ceil_div((int64_t)B * HW, (int64_t)256), 256, 0, st>>>(rgb, minmax, out_nhwc, B, HW);
  return check_launch("segmif_fused_to_uint8");
}

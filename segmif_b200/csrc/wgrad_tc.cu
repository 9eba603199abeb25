// Weight gradient of the fusion network's 3x3 (dilated) convolutions on tcgen05 tensor cores:
//     dW[co][tap][ci] = sum_p dY[p][co] * X[p + tap][ci]                 (core/model_fusion.py:118-157,1047-1067 backward)
// -- a GEMM whose contraction index is the PIXEL, so with pixel-major activations both operands are "MN-major" (the
// non-contracted index is the contiguous one).  The mma.sync version (train_ops.cu: wgrad_kernel<9>) went through
// ldmatrix.trans and reached ~200 TFLOP/s; it was 25 % of the train_fusion step.
// Formulation: D[ci][co] per tap, M = 128 input channels (two 64-channel SW128 blocks, LBO apart), N = 32 output channels,
// K = 16 pixels per MMA = one row of the 8 x 16-pixel tile.
//   * A operand: the (8+2d) x (16+2d) halo tile of X as TMA delivers it, [halo pixel][64 ch] rows of 128 B with the 128-byte
//     swizzle.  For tap (ky, kx) and tile row kr the 16 pixels are 16 CONSECUTIVE rows of that tile starting at row
//     (kr + ky d) * HXP + kx d, so each of the nine taps is just a different start address of the same tile (the swizzle XOR
//     follows absolute address bits -- tc_common.cuh -- so any 128-byte row start is legal).
//   * B operand: the 8 x 16 dY tile, [pixel][32 ch] rows of 64 B, 64-byte swizzle.
//   * nine fp32 accumulators [128 x 32] stay resident in TMEM (288 of 512 columns) over ALL tiles of the block; they are
//     read once at the end and written as per-block partials, reduced in fixed order by wgrad_reduce_kernel.
// Warp roles: 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2..5 = final epilogue.
#include <algorithm>

#include "tc_common.cuh"

namespace segmif {

namespace {

constexpr int kThreads = 192;
constexpr int TH = 8, TW = 16;
constexpr int Y_BYTES = TH * TW * 64;               // 32 channels x 2 B per pixel

struct WgTcArgs {
  float* partials;                                  // [gridDim.x][Cout][NT][Cin]
  int B, H, W, Cin, Cout, tiles_x, tiles_y, nstages;
};

template <int DIL>
struct Cfg {
  static constexpr int NT = DIL == 0 ? 1 : 9;
  static constexpr int HXP = TW + 2 * DIL, HROWS = TH + 2 * DIL;
  static constexpr int XB_BYTES = HROWS * HXP * 128;                 // one 64-channel block of the halo tile
  static constexpr int XB_STRIDE = (XB_BYTES + 1023) & ~1023;
  static constexpr int STAGE = 2 * XB_STRIDE + Y_BYTES;              // multiple of 1024
  static constexpr uint32_t TMEM_COLS = NT * 32 <= 32 ? 32 : 512;
};

// smem matrix descriptor, MN-major operand: [0,14) addr>>4 | [16,30) LBO>>4 (stride between 64-element groups along M/N) |
// [32,46) SBO>>4 (stride between 8-row K groups) | bit 46 version | [61,64) swizzle (2 = 128B, 4 = 64B)
__device__ __forceinline__ uint64_t mn_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

template <int DIL>
__global__ void __launch_bounds__(kThreads, 1) wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmX,
                                                               const __grid_constant__ CUtensorMap tmY, const WgTcArgs a) {
  using C = Cfg<DIL>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[4], empty[4], done;
  __shared__ uint32_t tmem_slot;
  const int NS = a.nstages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ci0 = blockIdx.y * 128, co0 = blockIdx.z * 32;
  const bool two_blocks = ci0 + 64 < a.Cin;                 // second 64-channel block holds real channels
  const int tiles_per_img = a.tiles_x * a.tiles_y;
  const int num_tiles = tiles_per_img * a.B;

  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&tmX);
    tc::prefetch_tmap(&tmY);
    for (int s = 0; s < 4; ++s) {
      tc::mbar_init(full + s, 1);
      tc::mbar_init(empty + s, 1);
    }
    tc::mbar_init(&done, 1);
    tc::fence_barrier_init();
  }
  if (!two_blocks) {                                        // the unused half of A must read as zeros
    for (int s = 0; s < NS; ++s) {
      uint4* z = reinterpret_cast<uint4*>(smem + s * C::STAGE + C::XB_STRIDE);
      for (int i = threadIdx.x; i < C::XB_STRIDE / 16; i += kThreads) z[i] = make_uint4(0, 0, 0, 0);
    }
    tc::fence_proxy_async();
  }
  if (warp == 1) tc::tmem_alloc(&tmem_slot, C::TMEM_COLS);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    if (tc::elect_one()) {
      const uint32_t bytes = (uint32_t)(C::XB_BYTES * (two_blocks ? 2 : 1) + Y_BYTES);
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int b = tile / tiles_per_img, rem = tile - b * tiles_per_img;
        const int y0 = (rem / a.tiles_x) * TH, x0 = (rem % a.tiles_x) * TW;
        const int s = it % NS;
        tc::mbar_wait(empty + s, ((it / NS) & 1) ^ 1);
        tc::mbar_expect_tx(full + s, bytes);
        uint8_t* st = smem + s * C::STAGE;
        tc::tma_load_4d(st, &tmX, full + s, ci0, x0 - DIL, y0 - DIL, b);
        if (two_blocks) tc::tma_load_4d(st + C::XB_STRIDE, &tmX, full + s, ci0 + 64, x0 - DIL, y0 - DIL, b);
        tc::tma_load_4d(st + 2 * C::XB_STRIDE, &tmY, full + s, co0, x0, y0, b);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    constexpr uint32_t idesc = tc::make_idesc_bf16(128, 32) | (1u << 15) | (1u << 16);      // A and B MN-major
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int s = it % NS;
      tc::mbar_wait(full + s, (it / NS) & 1);
      tc::tc_fence_after();
      if (tc::elect_one()) {
        const uint32_t sx = smem_u32(smem + s * C::STAGE);
        const uint32_t sy = sx + 2 * C::XB_STRIDE;
#pragma unroll 1
        for (int kr = 0; kr < TH; ++kr) {
          const uint64_t bd = mn_desc(sy + kr * TW * 64, 0, 512, 4);
#pragma unroll
          for (int t = 0; t < C::NT; ++t) {
            const int ky = C::NT == 9 ? t / 3 : 0, kx = C::NT == 9 ? t % 3 : 0;
            const uint32_t row = (uint32_t)((kr + ky * DIL) * C::HXP + kx * DIL);
            const uint64_t ad = mn_desc(sx + row * 128, C::XB_STRIDE, 1024, 2);
            tc::umma_bf16(tmem_base + t * 32, ad, bd, idesc, (it | kr) != 0 ? 1u : 0u);
          }
        }
        tc::umma_commit(empty + s);
      }
      __syncwarp();
    }
    if (tc::elect_one()) tc::umma_commit(&done);
    __syncwarp();
  } else {
    const int quad = warp & 3;
    tc::mbar_wait(&done, 0);
    tc::tc_fence_after();
    const int ci = ci0 + quad * 32 + lane;
    float* out = a.partials + (size_t)blockIdx.x * a.Cout * C::NT * a.Cin;
#pragma unroll 1
    for (int t = 0; t < C::NT; ++t) {
      float v[32];
      tc::tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(t * 32), v);
      if (ci < a.Cin) {
#pragma unroll
        for (int j = 0; j < 32; ++j) out[((size_t)(co0 + j) * C::NT + t) * a.Cin + ci] = v[j];
      }
    }
    tc::tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, C::TMEM_COLS);
}

int encode(CUtensorMap* out, const void* base, int C, int ld, int B, int H, int W, uint32_t bc, uint32_t bw, uint32_t bh,
           CUtensorMapSwizzle swz, const char* what) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return SEGMIF_ERR_CUDA;
  cuuint64_t gd[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t gs[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
  cuuint32_t bx[4] = {bc, bw, bh, 1}, es[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                   C == ld ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_64B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("%s: cuTensorMapEncodeTiled failed (CUresult %d)", what, (int)r); return SEGMIF_ERR_CUDA; }
  return SEGMIF_OK;
}

template <int DIL>
int launch(const CUtensorMap& tmX, const CUtensorMap& tmY, const WgTcArgs& a0, int nchunk, cudaStream_t st) {
  using C = Cfg<DIL>;
  WgTcArgs a = a0;
  a.nstages = std::max(2, std::min(4, (227 * 1024 - 2048) / C::STAGE));
  const size_t smem = (size_t)a.nstages * C::STAGE + 1024;
  static bool cfg = false;
  if (!cfg) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel<DIL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("wgrad_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return SEGMIF_ERR_CUDA; }
    cfg = true;
  }
  dim3 grid(nchunk, (a.Cin + 127) / 128, a.Cout / 32);
  wgrad_tc_kernel<DIL><<<grid, kThreads, smem, st>>>(tmX, tmY, a);
  return check_launch("segmif_wgrad (tcgen05)");
}

}  // namespace

bool wgrad_tc_ok(int B, int H, int W, int Cin, int Cout, int taps, int dil, int ldy, int ldx) {
  if (taps != 9 || (dil != 1 && dil != 2)) return false;
  return Cout % 32 == 0 && Cin % 8 == 0 && ldy % 8 == 0 && ldx % 8 == 0 && B > 0 && H > 0 && W > 0;
}

int wgrad_tc_chunks(int B, int H, int W, int Cin, int Cout) {
  const int tiles = B * ((H + TH - 1) / TH) * ((W + TW - 1) / TW);
  const int blocks = ((Cin + 127) / 128) * (Cout / 32);
  return std::max(1, std::min(tiles, std::max(1, 148 / blocks)));
}

// dy, x: channel offsets already applied; partials [nchunk][Cout][9][Cin]
int wgrad_tc(const void* dy, int ldy, const void* x, int ldx, int B, int H, int W, int Cin, int Cout, int dil, float* partials,
             int nchunk, cudaStream_t st) {
  CUtensorMap tmX, tmY;
  if (int rc = encode(&tmX, x, Cin, ldx, B, H, W, 64, TW + 2 * dil, TH + 2 * dil, CU_TENSOR_MAP_SWIZZLE_128B, "wgrad_tc(X)")) return rc;
  if (int rc = encode(&tmY, dy, Cout, ldy, B, H, W, 32, TW, TH, CU_TENSOR_MAP_SWIZZLE_64B, "wgrad_tc(dY)")) return rc;
  WgTcArgs a;
  a.partials = partials; a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout;
  a.tiles_x = (W + TW - 1) / TW; a.tiles_y = (H + TH - 1) / TH; a.nstages = 2;
  if (nchunk > a.tiles_x * a.tiles_y * B) { set_error("wgrad_tc: more chunks (%d) than tiles", nchunk); return SEGMIF_ERR_INVALID; }
  return dil == 2 ? launch<2>(tmX, tmY, a, nchunk, st) : launch<1>(tmX, tmY, a, nchunk, st);
}

}  // namespace segmif

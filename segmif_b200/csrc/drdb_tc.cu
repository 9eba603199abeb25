// DRDB growth layers in "push" form on tcgen05 tensor cores (segmif_drdb_push_tc_fwd).
//
// The reference computes layer j as a 3x3 dilated conv over ALL earlier slabs (N = 32 output channels, K = 9*Cin_j):
// with N = 32 every MMA is fed 4 KB of A from shared memory for 16 cycles of math, so the pull form tops out at
// ~40 % of the tensor peak (measured 609 TFLOP/s, profiles/r1_bench_v3_*).  Because convolution is linear in its input
// channels, the same numbers can be produced slab by slab instead: when slab s (x0, then each new g_j) is available,
// one pass multiplies its halo tile with the weights of EVERY later layer restricted to that slab's input channels
// (N = 32 * #later layers, up to 128), finishes the next layer (g = relu(P + acc + bias)) and adds the rest into bf16
// partial pre-activations P kept in HBM.  Six launches per DRDB:
//   x0 -> {g1, P2, P3} (N=96) | x0 -> {P4, P5} (N=64) | g1 -> {g2, P3+, P4+, P5+} (N=128) | g2 -> {g3, P4+, P5+} (N=96)
//   | g3 -> {g4, P5+} (N=64) | g4 -> {g5} (N=32)
// Same machinery as conv_tc.cu: persistent CTAs, the step's weights resident in shared memory (TMA), one 4-D TMA halo
// box per tile whose nine taps are shifted descriptor views, double-buffered TMEM accumulators.  32-channel slabs use
// only the first two K=16 sub-steps of each 128-byte row (the upper half of the box is out of range -> zero-filled,
// never multiplied) and their weights are packed two taps per 128-byte row.
#include <algorithm>

#include "dataflow.cuh"
#include "tc_common.cuh"

namespace segmif {

struct PushGroup {          // one 32-channel output group of the step
  const float* bias;        // added when non-null (the group that completes a layer)
  const bf16* pin;          // partial pre-activation to add (pixel-major), or null
  bf16* dst;                // g slice of the growth buffer, or P slice
  int ld_pin, coff_pin, ld_dst, coff_dst, relu;
};

struct PushArgs {
  PushGroup g[4];
  int B, H, W, tiles_x, tiles_y;
  unsigned* signal;        // dataflow.cuh: finished-tile counters [B * tiles_y] published for concurrently running consumers
  unsigned long long* timing;
};

constexpr int kPushThreads = 192;

template <int NOUT, int NSUB, int KSLAB>
struct PushCfg {
  static constexpr int TH = 16, TW = 8 * NSUB, DIL = 2;
  static constexpr int HROWS = TH + 2 * DIL;
  static constexpr int HXP = TW + 2 * DIL;                             // exact: swizzle follows absolute address bits
  static constexpr int A_BYTES = HROWS * HXP * 128;
  static constexpr int A_STRIDE = (A_BYTES + 1023) & ~1023;
  static constexpr int NWT = KSLAB == 64 ? 9 : 5;                    // weight tiles: one per tap, or one per tap pair
  static constexpr int W_TILE_BYTES = NOUT * 128;
  static constexpr int W_BYTES = NWT * W_TILE_BYTES;
  static constexpr int ACC_COLS = NSUB * NOUT;
  static constexpr uint32_t TMEM_COLS = (2 * ACC_COLS) <= 64 ? 64 : (2 * ACC_COLS) <= 128 ? 128 : (2 * ACC_COLS) <= 256 ? 256 : 512;
  static constexpr int NSTAGES = ((227 * 1024 - 2048 - W_BYTES) / A_STRIDE) >= 4 ? 4 : ((227 * 1024 - 2048 - W_BYTES) / A_STRIDE) >= 3 ? 3 : 2;
  static constexpr size_t SMEM = (size_t)W_BYTES + (size_t)NSTAGES * A_STRIDE + 13 * 8 + 16;
};

template <int NOUT, int NSUB, int KSLAB>
__global__ void __launch_bounds__(kPushThreads, 1) drdb_push_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                       const __grid_constant__ CUtensorMap tmW,
                                                                       const PushArgs a) {
  using Cfg = PushCfg<NOUT, NSUB, KSLAB>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sW = smem;
  uint8_t* sA = smem + Cfg::W_BYTES;
  constexpr int NS = Cfg::NSTAGES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::W_BYTES + NS * Cfg::A_STRIDE);
  uint64_t* empty = full + 4;
  uint64_t* wfull = empty + 4;
  uint64_t* tmem_full = wfull + 1;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_img = a.tiles_x * a.tiles_y;
  const int num_tiles = tiles_per_img * a.B;

  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmW);
    for (int s = 0; s < 4; ++s) {
      tc::mbar_init(full + s, 1);
      tc::mbar_init(empty + s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(tmem_full + s, 1);
      tc::mbar_init(tmem_empty + s, 4);
    }
    tc::mbar_init(wfull, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) df_mark_begin(a.timing);

  if (warp == 0) {
    if (tc::elect_one()) {
      tc::mbar_expect_tx(wfull, (uint32_t)Cfg::W_BYTES);
      for (int t = 0; t < Cfg::NWT; ++t) tc::tma_load_2d(sW + t * Cfg::W_TILE_BYTES, &tmW, wfull, t * 64, 0);
      int lt = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
        const int b = tile / tiles_per_img, rem = tile - b * tiles_per_img;
        const int y0 = (rem / a.tiles_x) * Cfg::TH, x0 = (rem % a.tiles_x) * Cfg::TW;
        const int s = lt % NS;
        tc::mbar_wait(empty + s, ((lt / NS) & 1) ^ 1);
        tc::mbar_expect_tx(full + s, Cfg::A_BYTES);
        tc::tma_load_4d(sA + s * Cfg::A_STRIDE, &tmA, full + s, 0, x0 - Cfg::DIL, y0 - Cfg::DIL, b);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    constexpr uint32_t idesc = tc::make_idesc_bf16(128, NOUT);
    constexpr uint64_t A_HI = (uint64_t)tc::desc_hi_sw128(Cfg::HXP * 128) << 32, B_HI = (uint64_t)tc::desc_hi_sw128(1024) << 32;
    const bool leader = tc::elect_one();
    tc::mbar_wait(wfull, 0);
    uint64_t w_d = B_HI | (uint64_t)(smem_u32(sW) >> 4);         // 64-bit descriptors: one UIADD3.64 per operand per MMA
    asm volatile("" : "+l"(w_d));
    int lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      const int s = lt % NS, buf = lt & 1;     // halo-tile ring slot and TMEM accumulator buffer of this tile
      tc::mbar_wait(tmem_empty + buf, ((lt >> 1) & 1) ^ 1);
      tc::mbar_wait(full + s, (lt / NS) & 1);
      tc::tc_fence_after();
      if (leader) {
        const uint32_t acc = tmem_base + (uint32_t)(buf * Cfg::ACC_COLS);
        uint64_t a_d0 = A_HI | (uint64_t)(smem_u32(sA + s * Cfg::A_STRIDE) >> 4);
        asm volatile("" : "+l"(a_d0));                  // opaque base: per-MMA offsets stay immediates of one UIADD3.64
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int ky = t / 3, kx = t % 3;
#pragma unroll
          for (int sub = 0; sub < NSUB; ++sub) {
#pragma unroll
            for (int k = 0; k < KSLAB / 16; ++k) {
              const uint32_t a_off = (uint32_t)(((ky * Cfg::DIL * Cfg::HXP + kx * Cfg::DIL + sub * 8) * 128 + k * 32) >> 4);
              const uint32_t w_off = (uint32_t)(((KSLAB == 64 ? t * Cfg::W_TILE_BYTES : (t >> 1) * Cfg::W_TILE_BYTES + (t & 1) * 64) + k * 32) >> 4);
              tc::umma_bf16(acc + (uint32_t)(sub * NOUT), a_d0 + a_off, w_d + w_off, idesc,
                                 (t == 0 && k == 0) ? 0u : 1u);
            }
          }
        }
        tc::umma_commit(empty + s);
        tc::umma_commit(tmem_full + buf);
      }
      __syncwarp();
    }
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + lane, ty = r >> 3, tx = r & 7;
    int lt = 0;
    int prev_b = -1, prev_t = 0;                   // dataflow: the tile whose stores are not yet published
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      const int buf = lt & 1;
      const int b = tile / tiles_per_img, rem = tile - b * tiles_per_img;
      const int y = (rem / a.tiles_x) * Cfg::TH + ty;
      const int x0 = (rem % a.tiles_x) * Cfg::TW + tx;
      tc::mbar_wait(tmem_full + buf, (lt >> 1) & 1);
      tc::tc_fence_after();
      if (a.signal && prev_b >= 0) {               // publish the previous tile: its stores were issued a whole tile ago
        __threadfence();
        tc::named_bar_sync(2, 128);
        if (warp == 2 && lane == 0) df_signal(a.signal, prev_b, a.tiles_y, prev_t);
      }
#pragma unroll 1
      for (int sg = 0; sg < NSUB * (NOUT / 32); ++sg) {
        const int sub = sg / (NOUT / 32), gi = sg % (NOUT / 32);
        const int x = x0 + sub * 8;
        const PushGroup& g = a.g[gi];
        float v[32];
        tc::tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * Cfg::ACC_COLS + sub * NOUT + gi * 32), v);
        if (y < a.H && x < a.W) {
          const int64_t pix = ((int64_t)b * a.H + y) * a.W + x;
          if (g.pin) {
            const uint4* pp = reinterpret_cast<const uint4*>(g.pin + pix * g.ld_pin + g.coff_pin);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 t = pp[j];
              const float2 p0 = unpack_bf16x2(t.x), p1 = unpack_bf16x2(t.y), p2 = unpack_bf16x2(t.z), p3 = unpack_bf16x2(t.w);
              v[8 * j] += p0.x; v[8 * j + 1] += p0.y; v[8 * j + 2] += p1.x; v[8 * j + 3] += p1.y;
              v[8 * j + 4] += p2.x; v[8 * j + 5] += p2.y; v[8 * j + 6] += p3.x; v[8 * j + 7] += p3.y;
            }
          }
          if (g.bias) {
            const float4* bp = reinterpret_cast<const float4*>(g.bias);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 bv = __ldg(bp + j);
              v[4 * j] += bv.x; v[4 * j + 1] += bv.y; v[4 * j + 2] += bv.z; v[4 * j + 3] += bv.w;
            }
          }
          if (g.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          uint4* d = reinterpret_cast<uint4*>(g.dst + pix * g.ld_dst + g.coff_dst);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            d[j] = make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                              pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(tmem_empty + buf);
      prev_b = b; prev_t = rem / a.tiles_x;
    }
    if (a.signal && prev_b >= 0) {
      __threadfence();
      tc::named_bar_sync(2, 128);
      if (warp == 2 && lane == 0) df_signal(a.signal, prev_b, a.tiles_y, prev_t);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  if (threadIdx.x == 0) df_mark_end(a.timing);
}

template <int NOUT, int NSUB, int KSLAB>
static int launch_push(const segmif_drdb_push_params* p, cudaStream_t st, const ConvDfExtra* x = nullptr) {
  using Cfg = PushCfg<NOUT, NSUB, KSLAB>;
  auto kern = drdb_push_tc_kernel<NOUT, NSUB, KSLAB>;
  static bool configured = false;
  static int sms = 148;
  if (!configured) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    if (err != cudaSuccess) { set_error("drdb_push: %zu bytes of shared memory refused: %s", Cfg::SMEM, cudaGetErrorString(err)); return SEGMIF_ERR_CUDA; }
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    configured = true;
  }
  CUtensorMap tmA, tmW;
  {
    // only the slab's own channels are in range: the upper half of a 32-channel slab's 64-channel box is zero-filled
    const uint64_t dims[4] = {(uint64_t)p->slab_width, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->B};
    const uint64_t strides[3] = {(uint64_t)p->ld_src * 2, (uint64_t)p->W * p->ld_src * 2, (uint64_t)p->H * p->W * p->ld_src * 2};
    const uint32_t box[4] = {64, (uint32_t)Cfg::HXP, (uint32_t)Cfg::HROWS, 1};
    int rc = make_tmap_bf16(&tmA, reinterpret_cast<const bf16*>(p->src) + p->slab_offset, 4, dims, strides, box, true, "drdb_push(A)",
                            p->slab_width == p->ld_src ? 256 : 64);   // a channel slab of wider rows: do not widen L2 misses
    if (rc) return rc;
  }
  {
    const uint64_t wcols = (uint64_t)Cfg::NWT * 64;
    const uint64_t dims[2] = {wcols, (uint64_t)NOUT};
    const uint64_t strides[1] = {wcols * 2};
    const uint32_t box[2] = {64, (uint32_t)NOUT};
    int rc = make_tmap_bf16(&tmW, p->weight, 2, dims, strides, box, true, "drdb_push(W)");
    if (rc) return rc;
  }
  PushArgs a;
  for (int i = 0; i < 4; ++i) {
    const segmif_drdb_push_group& s = p->groups[i];
    a.g[i].bias = s.bias; a.g[i].pin = reinterpret_cast<const bf16*>(s.partial_in); a.g[i].dst = reinterpret_cast<bf16*>(s.dst);
    a.g[i].ld_pin = s.ld_partial_in; a.g[i].coff_pin = s.coff_partial_in; a.g[i].ld_dst = s.ld_dst; a.g[i].coff_dst = s.coff_dst;
    a.g[i].relu = s.relu;
  }
  a.B = p->B; a.H = p->H; a.W = p->W;
  a.tiles_x = (p->W + Cfg::TW - 1) / Cfg::TW; a.tiles_y = (p->H + Cfg::TH - 1) / Cfg::TH;
  a.signal = x ? x->df.signal : nullptr;
  a.timing = x ? x->df.timing : nullptr;
  const int num_tiles = a.tiles_x * a.tiles_y * a.B;
  const int ctas = (x && x->max_ctas > 0) ? std::min(x->max_ctas, sms) : sms;
  kern<<<std::min(num_tiles, ctas), kPushThreads, Cfg::SMEM, st>>>(tmA, tmW, a);
  return check_launch("segmif_drdb_push_tc_fwd");
}

}  // namespace segmif

using namespace segmif;

static int drdb_push_dispatch(const segmif_drdb_push_params* p, cudaStream_t st, const ConvDfExtra* x) {
  SEGMIF_REQUIRE(p && p->src && p->weight, "drdb_push: null pointer");
  SEGMIF_REQUIRE(p->slab_width == 32 || p->slab_width == 64, "drdb_push: slab_width=%d must be 32 or 64", p->slab_width);
  SEGMIF_REQUIRE(p->n_out == 32 || p->n_out == 64 || p->n_out == 96 || p->n_out == 128, "drdb_push: n_out=%d must be 32/64/96/128", p->n_out);
  SEGMIF_REQUIRE(p->ld_src % 8 == 0 && p->slab_offset % 8 == 0 && p->slab_offset + p->slab_width <= p->ld_src, "drdb_push: bad slab");
  SEGMIF_REQUIRE(((uintptr_t)p->src & 15) == 0 && ((uintptr_t)p->weight & 15) == 0, "drdb_push: pointers must be 16-byte aligned");
  for (int i = 0; i < p->n_out / 32; ++i) {
    const segmif_drdb_push_group& g = p->groups[i];
    SEGMIF_REQUIRE(g.dst && ((uintptr_t)g.dst & 15) == 0 && g.ld_dst % 8 == 0 && g.coff_dst % 8 == 0, "drdb_push: group %d dst misaligned", i);
    SEGMIF_REQUIRE(!g.partial_in || (((uintptr_t)g.partial_in & 15) == 0 && g.ld_partial_in % 8 == 0 && g.coff_partial_in % 8 == 0), "drdb_push: group %d partial_in misaligned", i);
    SEGMIF_REQUIRE(!g.bias || ((uintptr_t)g.bias & 15) == 0, "drdb_push: group %d bias misaligned", i);
  }
  if (p->slab_width == 64) {
    if (p->n_out == 96) return launch_push<96, 1, 64>(p, st, x);      // 108 KB of weights: one sub-tile per box
    if (p->n_out == 64) return launch_push<64, 2, 64>(p, st, x);
    if (p->n_out == 32) return launch_push<32, 2, 64>(p, st, x);
    set_error("drdb_push: n_out=128 with a 64-channel slab does not fit shared memory");
    return SEGMIF_ERR_INVALID;
  }
  if (p->n_out == 128) return launch_push<128, 2, 32>(p, st, x);
  if (p->n_out == 96) return launch_push<96, 2, 32>(p, st, x);
  if (p->n_out == 64) return launch_push<64, 2, 32>(p, st, x);
  return launch_push<32, 2, 32>(p, st, x);
}

namespace segmif {
int drdb_push_df(const segmif_drdb_push_params* p, const ConvDfExtra& x, cudaStream_t st) { return drdb_push_dispatch(p, st, &x); }
// tile width of the instance the dispatch selects (64-channel slab: N = 96 uses 8-pixel-wide tiles, the others 16)
int drdb_push_tile_w(int slab_width, int n_out) { return (slab_width == 64 && n_out == 96) ? 8 : 16; }
}  // namespace segmif

extern "C" int segmif_drdb_push_tc_fwd(const segmif_drdb_push_params* p, segmif_stream_t stream) {
  return drdb_push_dispatch(p, as_stream(stream), nullptr);
}

// 3x3 (optionally dilated) stride-1 convolution as an implicit GEMM on tcgen05 tensor cores (segmif_conv3x3_tc_fwd):
// the DRDB growth convolutions (65 % of the pipeline's FLOPs) and conv2 / conv21 of Fusion_Network3_ac.
//
//  * persistent kernel, one CTA per SM; ALL packed weights of the layer ([Cout][9][Cin] bf16, <= 144 KB) are
//    loaded once per CTA by TMA and stay resident in shared memory;
//  * per output tile of 16 x (8*NSUB) pixels and per 64-channel slab, ONE 4-D TMA box load brings the tile plus its
//    dilation halo (rows x cols x 64 ch, 128-byte swizzle, out-of-image pixels and channels >= Cin zero-filled by
//    the TMA unit = the conv's zero padding for free).  The nine taps are nine shifted VIEWS of that box:
//    MMA row r = (ty, tx) = (r / 8, r % 8) sits at line (ty + ky*dil) * HXP + tx + kx*dil, i.e. a K-major SW128
//    operand with 8-row groups HXP*128 bytes apart whose start is shifted by (ky*dil*HXP + kx*dil) lines --
//    L2->SM traffic drops ~5x versus gathering every tap separately (2..4 boxes in flight, sized to shared memory);
//  * one elected thread issues tcgen05.mma (M=128, N=Cout, K=16) with precomputed descriptor words; accumulators live
//    in TMEM, double buffered so the epilogue of tile i overlaps the MMAs of tile i+1;
//  * epilogue: tcgen05.ld -> (+ partial pre-activation tile, TMA-loaded) -> bias -> ReLU/PReLU -> bf16 -> shared-memory
//    staging tile -> ONE TMA store per sub-tile into the channel slice of the destination (image borders clipped by
//    the TMA unit).  No LSU global access in the loop: per-lane 16-byte stores of a row-per-thread layout cost 32
//    L1tex wavefronts per instruction and capped the tensor pipe at 13-23 % (profiles/r1_ncu_conv3x3_tc_v3_*).
//    The staging tiles cost 16..48 KB; when they would take a slot from the halo ring (wide layers: weights 72 KB) the
//    host selects the DIRECT epilogue instead: partial rows prefetched into registers before the accumulator wait,
//    16-byte global stores.  The ring depth matters more: with two 51 KB slots a 2-slab layer holds ONE tile in
//    flight and every tile pays the ~1.8 us box-load latency (profiles/r1_ncu_conv3x3_tc_v7_*: MMA warp 50 % in wait).
// Warp roles: 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2..5 = epilogue.
#include <algorithm>
#include <type_traits>

#include "dataflow.cuh"
#include "tc_common.cuh"

namespace segmif {

struct ConvTcArgs {
  const float* bias;
  const float* alpha;
  int B, H, W, nchunks, act, has_pre;
  int tiles_x, tiles_y, cin, ksteps_last, nstages;
  int staged;              // 1: partial / output tiles go through shared-memory staging + TMA; 0: per-thread global access
  const bf16* pre;         // direct mode: partial pre-activations (channel offset applied), pitch ld_pre
  bf16* dst;               // direct mode: destination slice (channel offset applied), pitch ld_dst
  int ld_pre, ld_dst;
  int y_shift;             // the tile grid starts y_shift rows above the image (dataflow.cuh); 0 otherwise
  Dataflow df;             // cross-kernel dependencies (df.enabled == 0: plain stream-ordered kernel)
};

constexpr int kConvTcThreads = 192;

template <int COUT, int DIL, int NSUB>
struct ConvTcCfg {
  static constexpr int TH = 16, TW = 8 * NSUB;
  static constexpr int HROWS = TH + 2 * DIL;
  static constexpr int HXP = TW + 2 * DIL;     // halo columns; any pitch works: the swizzle follows absolute smem address bits
  static constexpr int A_BYTES = HROWS * HXP * 128;                    // bytes one TMA box delivers
  static constexpr int A_STRIDE = (A_BYTES + 1023) & ~1023;            // ring slots start on 1024-byte boundaries
  static constexpr int W_TILE_BYTES = COUT * 128;                      // one (slab, tap) weight tile
  static constexpr int ACC_COLS = NSUB * COUT;                         // TMEM columns per accumulator buffer
  static constexpr uint32_t TMEM_COLS = (2 * ACC_COLS) <= 32 ? 32 : (2 * ACC_COLS) <= 64 ? 64 : (2 * ACC_COLS) <= 128 ? 128 : 256;
  static constexpr int ROW_BYTES = COUT * 2;                           // one pixel of the output / partial tile
  static constexpr int SUB_BYTES = 128 * ROW_BYTES;                    // one 16x8 sub-tile
  static constexpr int OUT_BYTES = NSUB * SUB_BYTES;                   // staging for the TMA stores of one tile
};

template <int COUT, int DIL, int NSUB>
__global__ void __launch_bounds__(kConvTcThreads, 1) conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                       const __grid_constant__ CUtensorMap tmW,
                                                                       const __grid_constant__ CUtensorMap tmO,
                                                                       const __grid_constant__ CUtensorMap tmP,
                                                                       const ConvTcArgs a) {
  using Cfg = ConvTcCfg<COUT, DIL, NSUB>;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int w_bytes = a.nchunks * 9 * Cfg::W_TILE_BYTES;
  const int NS = a.nstages;                      // halo-tile ring depth (2..4), chosen by the host to fill shared memory
  uint8_t* sW = smem;
  uint8_t* sA = smem + w_bytes;
  uint8_t* sOut = sA + NS * Cfg::A_STRIDE;                             // [NSUB][128 px][COUT] bf16
  uint8_t* sPre = sOut + (a.staged ? Cfg::OUT_BYTES : 0);              // [2][NSUB][128 px][COUT] bf16 (only with pre_add)
  uint64_t* full = reinterpret_cast<uint64_t*>(sPre + ((a.staged && a.has_pre) ? 2 * Cfg::OUT_BYTES : 0));
  uint64_t* empty = full + 4;
  uint64_t* wfull = empty + 4;
  uint64_t* tmem_full = wfull + 1;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* pfull = tmem_empty + 2;
  uint64_t* pempty = pfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_img = a.tiles_x * a.tiles_y;
  const int num_tiles = tiles_per_img * a.B;

  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmW);
    tc::prefetch_tmap(&tmO);
    if (a.has_pre) tc::prefetch_tmap(&tmP);
    for (int s = 0; s < 4; ++s) {
      tc::mbar_init(full + s, 1);
      tc::mbar_init(empty + s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(tmem_full + s, 1);
      tc::mbar_init(tmem_empty + s, 4);        // one arrive per epilogue warp
      tc::mbar_init(pfull + s, 1);
      tc::mbar_init(pempty + s, 4);
    }
    tc::mbar_init(wfull, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0 && a.df.enabled) df_mark_begin(a.df.timing);

  if (warp == 0) {
    if (tc::elect_one()) {
      // resident weights: nchunks * 9 tiles of [COUT x 64]
      tc::mbar_expect_tx(wfull, (uint32_t)w_bytes);
      for (int c = 0; c < a.nchunks; ++c)
        for (int t = 0; t < 9; ++t)
          tc::tma_load_2d(sW + (c * 9 + t) * Cfg::W_TILE_BYTES, &tmW, wfull, t * a.cin + c * 64, 0);
      int it = 0, lt = 0;
      DfSeen seen0 = {-1, -1, 0}, seen1 = {-1, -1, 0};
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
        const int b = tile / tiles_per_img, rem = tile - b * tiles_per_img;
        const int y0 = (rem / a.tiles_x) * Cfg::TH - a.y_shift, x0 = (rem % a.tiles_x) * Cfg::TW;
        if (a.df.enabled) {                     // producer stages must have finished the rows this tile reads
          df_wait(a.df.dep[1], a.df.error, b, y0, y0 + Cfg::TH, a.H, seen1);
          df_wait(a.df.dep[0], a.df.error, b, y0, y0 + Cfg::TH, a.H, seen0);
        }
        if (a.has_pre && a.staged) {            // partial pre-activation tile of this output tile
          const int pb = lt & 1;
          tc::mbar_wait(pempty + pb, ((lt >> 1) & 1) ^ 1);
          tc::mbar_expect_tx(pfull + pb, Cfg::OUT_BYTES);
#pragma unroll
          for (int sub = 0; sub < NSUB; ++sub)
            tc::tma_load_4d(sPre + pb * Cfg::OUT_BYTES + sub * Cfg::SUB_BYTES, &tmP, pfull + pb, 0, x0 + sub * 8, y0, b);
        }
        for (int c = 0; c < a.nchunks; ++c, ++it) {
          const int s = it % NS;
          const uint32_t ph = (it / NS) & 1;
          tc::mbar_wait(empty + s, ph ^ 1);
          tc::mbar_expect_tx(full + s, Cfg::A_BYTES);
          tc::tma_load_4d(sA + s * Cfg::A_STRIDE, &tmA, full + s, c * 64, x0 - DIL, y0 - DIL, b);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // Whole warp walks the loop (uniform control flow -> descriptors live in uniform registers); the elected lane issues.
    constexpr uint32_t idesc = tc::make_idesc_bf16(128, COUT);
    // 64-bit descriptors advanced with ONE uniform 64-bit add each (UIADD3.64): ~2 SASS instructions per MMA.  Building them
    // from separate lo / hi words cost 8 per MMA, which made every N=32 launch issue-bound (55 cycles per 17-cycle MMA).
    constexpr uint64_t A_HI = (uint64_t)tc::desc_hi_sw128(Cfg::HXP * 128) << 32, B_HI = (uint64_t)tc::desc_hi_sw128(1024) << 32;
    const bool leader = tc::elect_one();
    tc::mbar_wait(wfull, 0);
    const uint64_t w_d0 = B_HI | (uint64_t)(smem_u32(sW) >> 4);
    int it = 0, lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      const int buf = lt & 1;
      tc::mbar_wait(tmem_empty + buf, ((lt >> 1) & 1) ^ 1);
      tc::tc_fence_after();
      const uint32_t acc = tmem_base + (uint32_t)(buf * Cfg::ACC_COLS);
      for (int c = 0; c < a.nchunks; ++c, ++it) {
        const int s = it % NS;
        tc::mbar_wait(full + s, (it / NS) & 1);
        tc::tc_fence_after();
        if (leader) {
          uint64_t a_d0 = A_HI | (uint64_t)(smem_u32(sA + s * Cfg::A_STRIDE) >> 4);
          uint64_t w_d = w_d0 + (uint64_t)(c * 9 * (Cfg::W_TILE_BYTES >> 4));
          asm volatile("" : "+l"(a_d0), "+l"(w_d));        // opaque bases: the per-MMA offsets stay 32-bit immediates of one UIADD3.64
          const uint32_t first = c != 0 ? 1u : 0u;
          const int klim = (c == a.nchunks - 1) ? a.ksteps_last : 4;
          // trailing channels of a partly filled last slab are zeros: only the first klim k-steps are issued.  One uniform
          // branch per slab selects a fully unrolled, predicate-free instance.
          auto issue = [&](auto klim_c) {
            constexpr int KL = decltype(klim_c)::value;
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              const int ky = t / 3, kx = t % 3;
#pragma unroll
              for (int sub = 0; sub < NSUB; ++sub) {
#pragma unroll
                for (int k = 0; k < KL; ++k) {
                  const uint32_t a_off = (uint32_t)(((ky * DIL * Cfg::HXP + kx * DIL + sub * 8) * 128 + k * 32) >> 4);
                  const uint32_t w_off = (uint32_t)((t * Cfg::W_TILE_BYTES + k * 32) >> 4);
                  tc::umma_bf16(acc + (uint32_t)(sub * COUT), a_d0 + a_off, w_d + w_off, idesc, (t == 0 && k == 0) ? first : 1u);
                }
              }
            }
          };
          if (klim == 4) issue(std::integral_constant<int, 4>{});
          else if (klim == 2) issue(std::integral_constant<int, 2>{});
          else if (klim == 1) issue(std::integral_constant<int, 1>{});
          else issue(std::integral_constant<int, 3>{});
          tc::umma_commit(empty + s);
        }
        __syncwarp();
      }
      if (leader) tc::umma_commit(tmem_full + buf);
      __syncwarp();
    }
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + lane;              // MMA row == TMEM lane == pixel (r / 8, r % 8) of the sub-tile
    // act(v) = v >= 0 ? v : slope * v covers none (slope 1), ReLU (0) and PReLU (alpha): branch-free, tiny code --
    // the epilogue must stay resident in the instruction cache next to the unrolled MMA issue loop.
    const float slope = a.act == SEGMIF_ACT_PRELU ? *a.alpha : (a.act == SEGMIF_ACT_RELU ? 0.f : 1.f);
    float bias_r[COUT];
#pragma unroll
    for (int j = 0; j < COUT / 4; ++j) {
      const float4 bv = __ldg(reinterpret_cast<const float4*>(a.bias) + j);
      bias_r[4 * j] = bv.x; bias_r[4 * j + 1] = bv.y; bias_r[4 * j + 2] = bv.z; bias_r[4 * j + 3] = bv.w;
    }
    const bool store_leader = (warp == 2 && lane == 0);
    int lt = 0;
    if (a.staged) {
      int prev_b = -1, prev_t = 0;               // dataflow: tile whose TMA stores are still in flight
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
        const int buf = lt & 1;
        const int b = tile / tiles_per_img, rem = tile - b * tiles_per_img;
        const int y0 = (rem / a.tiles_x) * Cfg::TH - a.y_shift, x0 = (rem % a.tiles_x) * Cfg::TW;
        if (a.has_pre) tc::mbar_wait(pfull + buf, (lt >> 1) & 1);
        tc::mbar_wait(tmem_full + buf, (lt >> 1) & 1);
        tc::tc_fence_after();
        if (store_leader) tc::bulk_wait_read0();               // the previous tile's TMA stores have drained the staging tile
        tc::named_bar_sync(1, 128);
#pragma unroll
        for (int sc = 0; sc < NSUB * (COUT / 32); ++sc) {
          const int sub = sc / (COUT / 32), c = (sc % (COUT / 32)) * 32;
          float v[32];
          tc::tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * Cfg::ACC_COLS + sub * COUT + c), v);
          if (a.has_pre) {
            const uint4* pp = reinterpret_cast<const uint4*>(sPre + buf * Cfg::OUT_BYTES + sub * Cfg::SUB_BYTES + r * Cfg::ROW_BYTES + c * 2);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 t = pp[j];
              const float2 p0 = unpack_bf16x2(t.x), p1 = unpack_bf16x2(t.y), p2 = unpack_bf16x2(t.z), p3 = unpack_bf16x2(t.w);
              v[8 * j] += p0.x; v[8 * j + 1] += p0.y; v[8 * j + 2] += p1.x; v[8 * j + 3] += p1.y;
              v[8 * j + 4] += p2.x; v[8 * j + 5] += p2.y; v[8 * j + 6] += p3.x; v[8 * j + 7] += p3.y;
            }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float t = v[j] + bias_r[c + j];
            v[j] = t >= 0.f ? t : slope * t;
          }
          uint4* d = reinterpret_cast<uint4*>(sOut + sub * Cfg::SUB_BYTES + r * Cfg::ROW_BYTES + c * 2);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            d[j] = make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                              pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
        }
        tc::tc_fence_before();
        tc::fence_proxy_async();                   // staging writes -> visible to the TMA (async proxy)
        __syncwarp();
        if (lane == 0) {
          tc::mbar_arrive(tmem_empty + buf);       // this warp has drained the accumulator buffer
          if (a.has_pre) tc::mbar_arrive(pempty + buf);
        }
        tc::named_bar_sync(1, 128);
        if (store_leader) {
#pragma unroll
          for (int sub = 0; sub < NSUB; ++sub) tc::tma_store_4d(&tmO, sOut + sub * Cfg::SUB_BYTES, 0, x0 + sub * 8, y0, b);
          tc::bulk_commit();
          if (a.df.enabled && a.df.signal) {
            // publish the PREVIOUS tile: all bulk groups but the one just committed have completed their global writes
            asm volatile("cp.async.bulk.wait_group 1;\n" ::: "memory");
            if (prev_b >= 0) df_signal(a.df.signal, prev_b, a.tiles_y, prev_t);
            prev_b = b; prev_t = rem / a.tiles_x;
          }
        }
      }
      if (store_leader && a.df.enabled && a.df.signal && prev_b >= 0) {
        tc::bulk_wait_all0();
        df_signal(a.df.signal, prev_b, a.tiles_y, prev_t);
      }
    } else {
      // direct epilogue: this thread owns pixel (r / 8, r % 8) of every sub-tile
      const int ty = r >> 3, tx = r & 7;
      DfSeen seen_e = {-1, -1, 0};
      int prev_b = -1, prev_t = 0;                 // dataflow: the tile whose stores are not yet published
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
        const int buf = lt & 1;
        const int b = tile / tiles_per_img, rem = tile - b * tiles_per_img;
        const int yt0 = (rem / a.tiles_x) * Cfg::TH - a.y_shift;
        const int y = yt0 + ty, x0 = (rem % a.tiles_x) * Cfg::TW + tx;
        const bool y_ok = (unsigned)y < (unsigned)a.H;
        const size_t rowpix = ((size_t)b * a.H + (y_ok ? y : 0)) * a.W;
        uint4 pv[NSUB][COUT / 8];
        if (a.has_pre) {                           // partial rows in flight while the MMAs of this tile still run
          if (a.df.enabled) {                      // ... once their producer stage has published them
            if (lane == 0) df_wait(a.df.dep[1], a.df.error, b, yt0, yt0 + Cfg::TH, a.H, seen_e);
            __syncwarp();
          }
#pragma unroll
          for (int sub = 0; sub < NSUB; ++sub) {
            const int x = x0 + sub * 8;
            const bool ok = y_ok && x < a.W;
            const uint4* pp = reinterpret_cast<const uint4*>(a.pre + (rowpix + x) * a.ld_pre);
#pragma unroll
            for (int j = 0; j < COUT / 8; ++j)      // written by a concurrently running kernel in dataflow mode: no .nc path
              pv[sub][j] = ok ? (a.df.enabled ? __ldcg(pp + j) : __ldg(pp + j)) : make_uint4(0, 0, 0, 0);
          }
        }
        tc::mbar_wait(tmem_full + buf, (lt >> 1) & 1);
        tc::tc_fence_after();
        if (a.df.enabled && a.df.signal && prev_b >= 0) {
          // publish the PREVIOUS tile now: its stores were issued a whole tile ago, so the fence no longer waits on them
          __threadfence();
          tc::named_bar_sync(2, 128);
          if (store_leader) df_signal(a.df.signal, prev_b, a.tiles_y, prev_t);
        }
#pragma unroll
        for (int sc = 0; sc < NSUB * (COUT / 32); ++sc) {
          const int sub = sc / (COUT / 32), c = (sc % (COUT / 32)) * 32;
          float v[32];
          tc::tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * Cfg::ACC_COLS + sub * COUT + c), v);
          if (a.has_pre) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 t = pv[sub][c / 8 + j];
              const float2 p0 = unpack_bf16x2(t.x), p1 = unpack_bf16x2(t.y), p2 = unpack_bf16x2(t.z), p3 = unpack_bf16x2(t.w);
              v[8 * j] += p0.x; v[8 * j + 1] += p0.y; v[8 * j + 2] += p1.x; v[8 * j + 3] += p1.y;
              v[8 * j + 4] += p2.x; v[8 * j + 5] += p2.y; v[8 * j + 6] += p3.x; v[8 * j + 7] += p3.y;
            }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float t = v[j] + bias_r[c + j];
            v[j] = t >= 0.f ? t : slope * t;
          }
          const int x = x0 + sub * 8;
          if (y_ok && x < a.W) {
            uint4* d = reinterpret_cast<uint4*>(a.dst + (rowpix + x) * a.ld_dst + c);
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
      if (a.df.enabled && a.df.signal && prev_b >= 0) {   // every epilogue thread's stores are ordered before the published count
        __threadfence();
        tc::named_bar_sync(2, 128);
        if (store_leader) df_signal(a.df.signal, prev_b, a.tiles_y, prev_t);
      }
    }
    if (store_leader) tc::bulk_wait_all0();
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  if (threadIdx.x == 0 && a.df.enabled) df_mark_end(a.df.timing);
}

template <int COUT, int DIL, int NSUB>
static size_t conv_tc_fixed_smem(int nchunks, bool has_pre, bool staged) {
  using Cfg = ConvTcCfg<COUT, DIL, NSUB>;
  return (size_t)nchunks * 9 * Cfg::W_TILE_BYTES + (staged ? Cfg::OUT_BYTES + (has_pre ? 2 * Cfg::OUT_BYTES : 0) : 0) + 17 * 8 + 16;
}

template <int COUT, int DIL, int NSUB>
static int launch_conv_tc(const segmif_conv_params* p, cudaStream_t st, const ConvDfExtra* x = nullptr) {
  using Cfg = ConvTcCfg<COUT, DIL, NSUB>;
  const int nchunks = (p->Cin + 63) / 64;
  const bool has_pre = p->pre_add != nullptr;
  const size_t limit = 227 * 1024 - 1024;
  auto depth = [&](bool staged_) {
    return (int)std::min<size_t>(4, (limit - conv_tc_fixed_smem<COUT, DIL, NSUB>(nchunks, has_pre, staged_)) / Cfg::A_STRIDE);
  };
  const bool staged = depth(true) >= depth(false);       // staging tiles only when they do not cost a ring slot
  const size_t fixed = conv_tc_fixed_smem<COUT, DIL, NSUB>(nchunks, has_pre, staged);
  const int nstages = std::max(2, depth(staged));
  const size_t smem = fixed + (size_t)nstages * Cfg::A_STRIDE;
  auto kern = conv3x3_tc_kernel<COUT, DIL, NSUB>;
  static bool configured = false;          // opt in once to the full 227 KB (the size varies with Cin; never during graph capture)
  static int sms = 148;
  if (!configured) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (err != cudaSuccess) { set_error("conv3x3_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(err)); return SEGMIF_ERR_CUDA; }
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    configured = true;
  }
  CUtensorMap tmA, tmW, tmO, tmP;
  {
    const uint64_t dims[4] = {(uint64_t)p->Cin, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->B};
    const uint64_t strides[3] = {(uint64_t)p->ld_src * 2, (uint64_t)p->W * p->ld_src * 2, (uint64_t)p->H * p->W * p->ld_src * 2};
    const uint32_t box[4] = {64, (uint32_t)Cfg::HXP, (uint32_t)Cfg::HROWS, 1};
    int rc = make_tmap_bf16(&tmA, reinterpret_cast<const bf16*>(p->src) + p->src_coff, 4, dims, strides, box, true, "conv3x3_tc(A)",
                            p->Cin == p->ld_src ? 256 : 64);       // a channel slab of wider rows: do not widen L2 misses
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)9 * p->Cin, (uint64_t)COUT};
    const uint64_t strides[1] = {(uint64_t)9 * p->Cin * 2};
    const uint32_t box[2] = {64, (uint32_t)COUT};
    int rc = make_tmap_bf16(&tmW, p->weight, 2, dims, strides, box, true, "conv3x3_tc(W)");
    if (rc) return rc;
  }
  {
    // output: the COUT-channel slice of the pixel-major destination; boxes of 16 x 8 pixels, clipped at the borders
    const uint64_t dims[4] = {(uint64_t)COUT, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->B};
    const uint64_t strides[3] = {(uint64_t)p->ld_dst * 2, (uint64_t)p->W * p->ld_dst * 2, (uint64_t)p->H * p->W * p->ld_dst * 2};
    const uint32_t box[4] = {(uint32_t)COUT, 8, 16, 1};
    int rc = make_tmap_bf16(&tmO, reinterpret_cast<bf16*>(p->dst) + p->dst_coff, 4, dims, strides, box, false, "conv3x3_tc(O)");
    if (rc) return rc;
    tmP = tmO;
    if (has_pre) {
      const uint64_t ps[3] = {(uint64_t)p->ld_pre * 2, (uint64_t)p->W * p->ld_pre * 2, (uint64_t)p->H * p->W * p->ld_pre * 2};
      rc = make_tmap_bf16(&tmP, reinterpret_cast<const bf16*>(p->pre_add) + p->pre_coff, 4, dims, ps, box, false, "conv3x3_tc(P)", COUT == p->ld_pre ? 256 : 64);
      if (rc) return rc;
    }
  }
  ConvTcArgs a;
  a.bias = p->bias; a.alpha = p->prelu_alpha;
  a.B = p->B; a.H = p->H; a.W = p->W; a.nchunks = nchunks; a.act = p->act; a.has_pre = has_pre ? 1 : 0;
  a.y_shift = x ? x->y_shift : 0;
  if (x) a.df = x->df; else { a.df = Dataflow(); a.df.enabled = 0; a.df.signal = nullptr; a.df.error = nullptr; a.df.timing = nullptr; a.df.dep[0].flags = a.df.dep[1].flags = nullptr; }
  a.tiles_x = (p->W + Cfg::TW - 1) / Cfg::TW; a.tiles_y = (p->H + a.y_shift + Cfg::TH - 1) / Cfg::TH;
  a.cin = p->Cin;
  a.ksteps_last = ((p->Cin - 1) % 64) / 16 + 1;
  a.nstages = nstages;
  a.staged = staged ? 1 : 0;
  a.pre = has_pre ? reinterpret_cast<const bf16*>(p->pre_add) + p->pre_coff : nullptr;
  a.dst = reinterpret_cast<bf16*>(p->dst) + p->dst_coff;
  a.ld_pre = p->ld_pre; a.ld_dst = p->ld_dst;
  const int num_tiles = a.tiles_x * a.tiles_y * a.B;
  const int ctas = (x && x->max_ctas > 0) ? std::min(x->max_ctas, sms) : sms;
  kern<<<std::min(num_tiles, ctas), kConvTcThreads, smem, st>>>(tmA, tmW, tmO, tmP, a);
  return check_launch("segmif_conv3x3_tc_fwd");
}

template <int COUT, int DIL, int NSUB>
static bool conv_tc_fits(int nchunks, bool has_pre) {
  return conv_tc_fixed_smem<COUT, DIL, NSUB>(nchunks, has_pre, false) + 2 * (size_t)ConvTcCfg<COUT, DIL, NSUB>::A_STRIDE <= 227 * 1024 - 1024;
}

}  // namespace segmif

using namespace segmif;

static int conv3x3_tc_dispatch(const segmif_conv_params* p, cudaStream_t st, const ConvDfExtra* x) {
  SEGMIF_REQUIRE(p && p->src && p->weight && p->dst && p->bias, "conv3x3_tc: null pointer (bias is required)");
  SEGMIF_REQUIRE(p->KH == 3 && p->KW == 3 && p->stride == 1 && (p->dil == 1 || p->dil == 2) && p->pad == p->dil,
                 "conv3x3_tc: only 3x3, stride 1, dilation 1 or 2 with 'same' padding");
  SEGMIF_REQUIRE(p->Cout == 32 || p->Cout == 64, "conv3x3_tc: Cout=%d must be 32 or 64", p->Cout);
  SEGMIF_REQUIRE(p->Cin % 16 == 0 && p->Cin > 0 && p->ld_src % 8 == 0 && p->src_coff % 8 == 0, "conv3x3_tc: Cin must be a multiple of 16, pitch/offset multiples of 8");
  SEGMIF_REQUIRE(p->dst_dtype == SEGMIF_BF16 && p->ld_dst % 8 == 0 && p->dst_coff % 8 == 0, "conv3x3_tc: dst must be bf16 with 16-byte aligned slices");
  SEGMIF_REQUIRE(p->residual == nullptr, "conv3x3_tc: residual is not supported");
  SEGMIF_REQUIRE(p->pre_add == nullptr || (((uintptr_t)p->pre_add & 15) == 0 && p->ld_pre % 8 == 0 && p->pre_coff % 8 == 0), "conv3x3_tc: pre_add must be 16-byte aligned");
  SEGMIF_REQUIRE(p->act != SEGMIF_ACT_GELU, "conv3x3_tc: GELU is not supported");
  SEGMIF_REQUIRE(p->act != SEGMIF_ACT_PRELU || p->prelu_alpha, "conv3x3_tc: PReLU needs prelu_alpha");
  SEGMIF_REQUIRE(p->src_coff + p->Cin <= p->ld_src && p->dst_coff + p->Cout <= p->ld_dst, "conv3x3_tc: channel slice exceeds pitch");
  SEGMIF_REQUIRE(((uintptr_t)p->src & 15) == 0 && ((uintptr_t)p->weight & 15) == 0 && ((uintptr_t)p->dst & 15) == 0 && ((uintptr_t)p->bias & 15) == 0,
                 "conv3x3_tc: pointers must be 16-byte aligned");
  const int nchunks = (p->Cin + 63) / 64;
  const bool pre = p->pre_add != nullptr;
  if (p->Cout == 32 && p->dil == 2) {
    if (conv_tc_fits<32, 2, 2>(nchunks, pre)) return launch_conv_tc<32, 2, 2>(p, st, x);
    SEGMIF_REQUIRE((conv_tc_fits<32, 2, 1>(nchunks, pre)), "conv3x3_tc: Cin=%d too large for resident weights", p->Cin);
    return launch_conv_tc<32, 2, 1>(p, st, x);
  }
  if (p->Cout == 32 && p->dil == 1) {
    if (conv_tc_fits<32, 1, 2>(nchunks, pre)) return launch_conv_tc<32, 1, 2>(p, st, x);
    SEGMIF_REQUIRE((conv_tc_fits<32, 1, 1>(nchunks, pre)), "conv3x3_tc: Cin=%d too large for resident weights", p->Cin);
    return launch_conv_tc<32, 1, 1>(p, st, x);
  }
  if (p->Cout == 64 && p->dil == 1) {
    SEGMIF_REQUIRE((conv_tc_fits<64, 1, 1>(nchunks, pre)), "conv3x3_tc: Cin=%d too large for resident weights", p->Cin);
    return launch_conv_tc<64, 1, 1>(p, st, x);
  }
  SEGMIF_REQUIRE((conv_tc_fits<64, 2, 1>(nchunks, pre)), "conv3x3_tc: Cin=%d too large for resident weights", p->Cin);
  return launch_conv_tc<64, 2, 1>(p, st, x);
}

namespace segmif {
int conv3x3_tc_df(const segmif_conv_params* p, const ConvDfExtra& x, cudaStream_t st) { return conv3x3_tc_dispatch(p, st, &x); }
// tile width (pixels) the dispatch above selects: the consumer of this stage's counters needs tiles per tile row
int conv3x3_tc_tile_w(int Cin, int Cout, int dil, bool has_pre) {
  const int nchunks = (Cin + 63) / 64;
  if (Cout == 32 && dil == 2) return conv_tc_fits<32, 2, 2>(nchunks, has_pre) ? 16 : 8;
  if (Cout == 32 && dil == 1) return conv_tc_fits<32, 1, 2>(nchunks, has_pre) ? 16 : 8;
  return 8;
}
}  // namespace segmif

extern "C" int segmif_conv3x3_tc_fwd(const segmif_conv_params* p, segmif_stream_t stream) {
  return conv3x3_tc_dispatch(p, as_stream(stream), nullptr);
}

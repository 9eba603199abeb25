// Pass 3 of the hierarchical interactive attention on tcgen05 tensor cores (segmif_ffm_apply_lr_fwd):
//   out_s = LayerNorm( x_s + y3 Mz_s + relu(x_s Wu_s^T + bu_s) Mv_s + b_end_s ),  s = 1, 2,   y3 = relu(upsample(Q)[0:64])
// per 128-pixel tile, as two chained GEMM stages whose intermediate never leaves the SM:
//   stage 1  U_s = X_s Wu_s^T              (TMA tiles of x1 / x2 -> tcgen05.mma -> TMEM)
//   epi 1    u_s = relu(U_s + bu_s) -> bf16 -> shared memory, written in the 128-byte-swizzled K-major layout the
//            tensor core reads; y3 is interpolated from the L2-resident low-resolution Q into the same layout
//   stage 2  O_s = y3 Mz_s^T + u_s Mv_s^T  (folded per-image matrices resident in shared memory)
//   epi 2    + b_end + residual (the x_s tile still in shared memory) -> LayerNorm with ONE PIXEL ROW PER THREAD (the
//            TMEM lane layout), no shuffles -> bf16 -> swizzled staging tile -> TMA store into the channel slice.
// 24 MMAs per tile; HBM traffic = 256 B/px in + 256 B/px out.  The mma.sync version it replaces (ffm.cu) needed
// 0.76 ms per call against a 0.2 ms HBM floor: 16-pixel-per-warp register GEMM chains, 237 registers, 8 warps/SM.
// Warp roles: 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2..5 = epilogue of stream 1, 6..9 = stream 2
// (both groups interpolate y3); grid = (CTAs per image, B), each CTA walks tiles of its image.
#include <algorithm>

#include "tc_common.cuh"

namespace segmif {

struct FfmTcArgs {
  const bf16* q;           // [B, qh, qw, 128] low-resolution pre-activation (channels 0..63 = y3 half)
  const float* bproj;      // [b1u 64][b2u 64]
  const float* bend;       // [2][64]
  const float* ln_g;       // [2][64]
  const float* ln_b;       // [2][64]
  float eps, sy, sx;
  int qh, qw, H, W;
  int64_t HW;
};

constexpr int kFfmTcThreads = 320;
constexpr int kTileBytes = 128 * 128;        // [128 px][64 ch] bf16

__device__ __forceinline__ void ffm_lr_src(int dst, float scale, int in_size, int& i0, int& i1, float& l0, float& l1) {
  float s = scale * ((float)dst + 0.5f) - 0.5f;
  s = s < 0.f ? 0.f : s;
  i0 = (int)s;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  l1 = s - (float)i0;
  l0 = 1.f - l1;
}

// byte offset of 16-byte chunk `chunk` of row `row` inside a [rows][128 B] tile with the 128-byte swizzle
__device__ __forceinline__ uint32_t sw128_off(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

__global__ void __launch_bounds__(kFfmTcThreads, 1) ffm_apply_tc_kernel(const __grid_constant__ CUtensorMap tmX1,
                                                                        const __grid_constant__ CUtensorMap tmX2,
                                                                        const __grid_constant__ CUtensorMap tmW,
                                                                        const __grid_constant__ CUtensorMap tmM,
                                                                        const __grid_constant__ CUtensorMap tmO1,
                                                                        const __grid_constant__ CUtensorMap tmO2,
                                                                        const FfmTcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sW = smem;                              // [2][64][128 B]      Wu1, Wu2
  uint8_t* sM = sW + 2 * 8192;                     // [4][64][128 B]      Mz1, Mv1, Mz2, Mv2 of this image
  uint8_t* sX = sM + 4 * 8192;                     // [2 stages][2 streams][128][128 B]
  uint8_t* sY = sX + 4 * kTileBytes;               // [2] y3 tiles: tile i+1's is interpolated while the MMAs of tile i run
  uint8_t* sU = sY + 2 * kTileBytes;               // [2 streams] u tiles
  uint8_t* sO = sU + 2 * kTileBytes;               // [2 streams] output staging
  uint64_t* wfull = reinterpret_cast<uint64_t*>(sO + 2 * kTileBytes);
  uint64_t* xfull = wfull + 1;                     // [2]
  uint64_t* xempty = xfull + 2;                    // [2]  8 arrivals (epilogue warps, after the residual read)
  uint64_t* g1_full = xempty + 2;                  // stage-1 accumulators ready
  uint64_t* a2_ready = g1_full + 1;                // y3 / u tiles written (8 arrivals)
  uint64_t* g2_full = a2_ready + 1;                // stage-2 accumulators ready
  uint64_t* tile_done = g2_full + 1;               // epilogue 2 has drained TMEM (8 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tile_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int ntiles = (int)((a.HW + 127) / 128);

  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&tmX1); tc::prefetch_tmap(&tmX2); tc::prefetch_tmap(&tmW); tc::prefetch_tmap(&tmM);
    tc::prefetch_tmap(&tmO1); tc::prefetch_tmap(&tmO2);
    tc::mbar_init(wfull, 1);
    for (int s = 0; s < 2; ++s) { tc::mbar_init(xfull + s, 1); tc::mbar_init(xempty + s, 8); }
    tc::mbar_init(g1_full, 1);
    tc::mbar_init(a2_ready, 8);
    tc::mbar_init(g2_full, 1);
    tc::mbar_init(tile_done, 8);
    tc::fence_barrier_init();
  }
  // TMEM columns: [0,128) stage-1 accumulators of even tiles, [128,256) stage-2 accumulators, [256,384) stage-1 of odd tiles.
  // Round 2: the stage-1 MMAs of tile i+1 are issued right behind the stage-2 MMAs of tile i (own accumulator buffer), and the
  // epilogue warps interpolate y3 of tile i+1 while they wait for stage 2 of tile i -- per tile the chain was
  // y3 | MMA1 -> epi 1 -> MMA2 -> epi 2 with the tensor pipe idle during both epilogues and the warps idle during both MMAs.
  if (warp == 1) tc::tmem_alloc(tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (tc::elect_one()) {
      tc::mbar_expect_tx(wfull, 6 * 8192);
      tc::tma_load_2d(sW, &tmW, wfull, 0, 0);
      tc::tma_load_2d(sW + 8192, &tmW, wfull, 0, 64);
      for (int m = 0; m < 4; ++m) tc::tma_load_2d(sM + m * 8192, &tmM, wfull, 0, (b * 4 + m) * 64);
      int it = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const int s = it & 1;
        tc::mbar_wait(xempty + s, ((it >> 1) & 1) ^ 1);
        tc::mbar_expect_tx(xfull + s, 2 * kTileBytes);
        tc::tma_load_3d(sX + (s * 2 + 0) * kTileBytes, &tmX1, xfull + s, 0, t * 128, b);
        tc::tma_load_3d(sX + (s * 2 + 1) * kTileBytes, &tmX2, xfull + s, 0, t * 128, b);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    constexpr uint32_t idesc = tc::make_idesc_bf16(128, 64);
    constexpr uint64_t HI = (uint64_t)tc::desc_hi_sw128(1024) << 32;
    const bool leader = tc::elect_one();
    tc::mbar_wait(wfull, 0);
    uint64_t w_d = HI | (smem_u32(sW) >> 4), m_d = HI | (smem_u32(sM) >> 4), y_d = HI | (smem_u32(sY) >> 4), u_d = HI | (smem_u32(sU) >> 4);
    asm volatile("" : "+l"(w_d), "+l"(m_d), "+l"(y_d), "+l"(u_d));   // opaque bases: offsets stay immediates of one UIADD3.64
    auto stage1 = [&](int it) {                                     // U_s = X_s Wu_s^T of tile `it` into its accumulator buffer
      const int s = it & 1;
      tc::mbar_wait(xfull + s, (it >> 1) & 1);
      tc::tc_fence_after();
      if (leader) {
        uint64_t x_d = HI | (smem_u32(sX + s * 2 * kTileBytes) >> 4);
        asm volatile("" : "+l"(x_d));
        const uint32_t acc1 = tmem_base + (uint32_t)(s * 256);
#pragma unroll
        for (int st = 0; st < 2; ++st)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc::umma_bf16(acc1 + st * 64, x_d + (uint64_t)(st * (kTileBytes >> 4) + k * 2), w_d + (uint64_t)(st * 512 + k * 2), idesc, k > 0 ? 1u : 0u);
        tc::umma_commit(g1_full);
      }
      __syncwarp();
    };
    int it = 0;
    if ((int)blockIdx.x < ntiles) stage1(0);
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const int s = it & 1;
      tc::mbar_wait(a2_ready, it & 1);                             // u tiles (and, transitively, this tile's y3) written
      if (it > 0) tc::mbar_wait(tile_done, (it - 1) & 1);          // stage-2 accumulators of the previous tile have been drained
      tc::tc_fence_after();
      if (leader) {
        const uint64_t y_ds = y_d + (uint64_t)(s * (kTileBytes >> 4));
#pragma unroll
        for (int st = 0; st < 2; ++st) {
#pragma unroll
          for (int k = 0; k < 4; ++k)            // y3 Mz_s^T
            tc::umma_bf16(tmem_base + 128 + st * 64, y_ds + (uint64_t)(k * 2), m_d + (uint64_t)((2 * st) * 512 + k * 2), idesc, k > 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 4; ++k)            // + u_s Mv_s^T
            tc::umma_bf16(tmem_base + 128 + st * 64, u_d + (uint64_t)(st * (kTileBytes >> 4) + k * 2), m_d + (uint64_t)((2 * st + 1) * 512 + k * 2), idesc, 1u);
        }
        tc::umma_commit(g2_full);
      }
      __syncwarp();
      if (t + (int)gridDim.x < ntiles) stage1(it + 1);             // next tile's projection runs behind this tile's stage 2
    }
  } else {
    const int ew = warp - 2;                       // 0..7
    const int st = ew >> 2;                        // stream handled by this warp group
    const int quad = warp & 3;
    const int r = quad * 32 + lane;                // tile row == TMEM lane of this thread
    const int et = ew * 32 + lane;                 // 0..255 for the y3 interpolation
    const int bar_id = 1 + st;
    const bool store_leader = ((ew & 3) == 0 && lane == 0);
    const CUtensorMap* tmO = st == 0 ? &tmO1 : &tmO2;
    uint8_t* sUs = sU + st * kTileBytes;
    uint8_t* sOs = sO + st * kTileBytes;
    // ---- y3 = relu(bilerp(Q[..., 0:64])) of tile `t` -> sY[buf].  Safe without a barrier: the last reader of sY[buf] was stage 2
    // of the tile two iterations back, whose completion (g2_full) these warps have already waited for.
    auto write_y3 = [&](int t, int buf) {
      const int64_t p0 = (int64_t)t * 128;
      uint8_t* sYb = sY + buf * kTileBytes;
      {
        const int pixel = et >> 1, half = et & 1;
        const int64_t p = p0 + pixel;
        uint4 outv[4];
        if (p < a.HW) {
          const unsigned pu = (unsigned)p;                     // HW < 2^31 (checked by the host)
          const int Y = (int)(pu / (unsigned)a.W), X = (int)(pu - (unsigned)Y * (unsigned)a.W);
          int y0, y1, x0, x1;
          float hy0, hy1, wx0, wx1;
          ffm_lr_src(Y, a.sy, a.qh, y0, y1, hy0, hy1);
          ffm_lr_src(X, a.sx, a.qw, x0, x1, wx0, wx1);
          const bf16* base = a.q + (int64_t)b * a.qh * a.qw * 128 + half * 32;
          const bf16* p00 = base + ((int64_t)y0 * a.qw + x0) * 128;
          const bf16* p01 = base + ((int64_t)y0 * a.qw + x1) * 128;
          const bf16* p10 = base + ((int64_t)y1 * a.qw + x0) * 128;
          const bf16* p11 = base + ((int64_t)y1 * a.qw + x1) * 128;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float v0[8], v1[8], v2[8], v3[8], o[8];
            load8(p00 + c * 8, v0); load8(p01 + c * 8, v1); load8(p10 + c * 8, v2); load8(p11 + c * 8, v3);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = fmaxf(hy0 * (wx0 * v0[j] + wx1 * v1[j]) + hy1 * (wx0 * v2[j] + wx1 * v3[j]), 0.f);
            outv[c] = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
          }
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c) outv[c] = make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(sYb + sw128_off(pixel, half * 4 + c)) = outv[c];
      }
    };
    int it = 0;
    if ((int)blockIdx.x < ntiles) write_y3(blockIdx.x, 0);
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const int s = it & 1;
      const int64_t p0 = (int64_t)t * 128;
      // ---- epilogue 1: u_s = relu(U_s + bu_s) -> sU_s
      tc::mbar_wait(g1_full, it & 1);
      tc::tc_fence_after();
#pragma unroll
      for (int hc = 0; hc < 2; ++hc) {
        float v[32];
        tc::tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(s * 256 + st * 64 + hc * 32), v);
        const float4* bp = reinterpret_cast<const float4*>(a.bproj + st * 64 + hc * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bv = __ldg(bp + j);
          v[4 * j] = fmaxf(v[4 * j] + bv.x, 0.f); v[4 * j + 1] = fmaxf(v[4 * j + 1] + bv.y, 0.f);
          v[4 * j + 2] = fmaxf(v[4 * j + 2] + bv.z, 0.f); v[4 * j + 3] = fmaxf(v[4 * j + 3] + bv.w, 0.f);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<uint4*>(sUs + sw128_off(r, hc * 4 + j)) =
              make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                         pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
      }
      tc::tc_fence_before();
      tc::fence_proxy_async();                     // sY / sU writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(a2_ready);
      if (t + (int)gridDim.x < ntiles) write_y3(t + (int)gridDim.x, s ^ 1);      // overlaps stage 2 of this tile
      // ---- epilogue 2: + b_end + residual, LayerNorm over the 64 channels of this thread's pixel, TMA store
      tc::mbar_wait(g2_full, it & 1);
      tc::tc_fence_after();
      float v[64];
      {
        float lo[32], hi[32];
        tc::tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(128 + st * 64), lo);
        tc::tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(128 + st * 64 + 32), hi);
#pragma unroll
        for (int j = 0; j < 32; ++j) { v[j] = lo[j]; v[32 + j] = hi[j]; }
      }
      const uint8_t* sXs = sX + (s * 2 + st) * kTileBytes;
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 xr = *reinterpret_cast<const uint4*>(sXs + sw128_off(r, c));
        const float2 x0 = unpack_bf16x2(xr.x), x1 = unpack_bf16x2(xr.y), x2 = unpack_bf16x2(xr.z), x3 = unpack_bf16x2(xr.w);
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.bend + st * 64 + c * 8));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(a.bend + st * 64 + c * 8 + 4));
        v[8 * c] += x0.x + b0.x; v[8 * c + 1] += x0.y + b0.y; v[8 * c + 2] += x1.x + b0.z; v[8 * c + 3] += x1.y + b0.w;
        v[8 * c + 4] += x2.x + b1.x; v[8 * c + 5] += x2.y + b1.y; v[8 * c + 6] += x3.x + b1.z; v[8 * c + 7] += x3.y + b1.w;
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) { tc::mbar_arrive(tile_done); tc::mbar_arrive(xempty + s); }   // TMEM and the x tiles are free again
#pragma unroll
      for (int j = 0; j < 64; ++j) sum += v[j];
      const float mean = sum * (1.f / 64.f);
      float sq = 0.f;
#pragma unroll
      for (int j = 0; j < 64; ++j) { const float d = v[j] - mean; sq = fmaf(d, d, sq); }
      const float rstd = rsqrtf(sq * (1.f / 64.f) + a.eps);
      if (store_leader) tc::bulk_wait_read0();     // the previous tile's store has drained this group's staging tile
      tc::named_bar_sync(bar_id, 128);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(a.ln_g + st * 64 + c * 8));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(a.ln_g + st * 64 + c * 8 + 4));
        const float4 e0 = __ldg(reinterpret_cast<const float4*>(a.ln_b + st * 64 + c * 8));
        const float4 e1 = __ldg(reinterpret_cast<const float4*>(a.ln_b + st * 64 + c * 8 + 4));
        const float o0 = (v[8 * c] - mean) * rstd * g0.x + e0.x, o1 = (v[8 * c + 1] - mean) * rstd * g0.y + e0.y;
        const float o2 = (v[8 * c + 2] - mean) * rstd * g0.z + e0.z, o3 = (v[8 * c + 3] - mean) * rstd * g0.w + e0.w;
        const float o4 = (v[8 * c + 4] - mean) * rstd * g1.x + e1.x, o5 = (v[8 * c + 5] - mean) * rstd * g1.y + e1.y;
        const float o6 = (v[8 * c + 6] - mean) * rstd * g1.z + e1.z, o7 = (v[8 * c + 7] - mean) * rstd * g1.w + e1.w;
        *reinterpret_cast<uint4*>(sOs + sw128_off(r, c)) = make_uint4(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3), pack_bf16x2(o4, o5), pack_bf16x2(o6, o7));
      }
      tc::fence_proxy_async();
      tc::named_bar_sync(bar_id, 128);
      if (store_leader) {
        tc::tma_store_3d(tmO, sOs, 0, (int)p0, b);
        tc::bulk_commit();
      }
    }
    if (store_leader) tc::bulk_wait_all0();
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------ pass 1 (Gram)
// G_s = u_s^T u_s over the pixels of this CTA's tiles, s = 1, 2, 3 (u_s = relu(x_s Wu_s^T + bu_s); u_3 = relu(upsample(Q)
// [64:128])).  The projection is the same tensor-core stage as in the apply pass; its epilogue writes u_s TRANSPOSED
// ([channel][pixel], 128-byte swizzle, two 64-pixel K blocks) so that the Gram products are ordinary K-major MMAs with
// K = pixels: streams 1 and 2 stacked as one M = N = 128 product (the two diagonal 64x64 blocks are G_1, G_2; the
// off-diagonal blocks are free riders on an idle tensor pipe), stream 3 as M = 128 (rows 64..127 read whatever follows
// in shared memory and are never looked at), N = 64.  Accumulators stay in TMEM for the whole kernel and are written
// once per CTA as fp32 partial sums (same [B, nchunk, 3, 64, 64] contract as the mma.sync kernel it replaces, which
// spent 620 us per call at 12 % occupancy / 237 registers).
constexpr int kGramKb = 192 * 128;            // one 64-pixel K block: rows 0..63 u1^T, 64..127 u2^T, 128..191 u3^T
constexpr int kGramBuf = 2 * kGramKb;         // 128 pixels

__global__ void __launch_bounds__(kFfmTcThreads, 1) ffm_gram_tc_kernel(const __grid_constant__ CUtensorMap tmX1,
                                                                       const __grid_constant__ CUtensorMap tmX2,
                                                                       const __grid_constant__ CUtensorMap tmW,
                                                                       const FfmTcArgs a, float* __restrict__ partials) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sW = smem;                              // [2][64][128 B]
  uint8_t* sX = sW + 2 * 8192;                     // [2 stages][2 streams][128][128 B]
  uint8_t* sT = sX + 4 * kTileBytes;               // [2 buffers][2 K blocks][192 rows][128 B] + 8 KB of zeros
  uint64_t* wfull = reinterpret_cast<uint64_t*>(sT + 2 * kGramBuf + 8192);
  uint64_t* xfull = wfull + 1;                     // [2]
  uint64_t* xempty = xfull + 2;                    // [2]  projection MMAs have read the x tiles
  uint64_t* g1_full = xempty + 2;                  // [2]  projection accumulators ready
  uint64_t* u_free = g1_full + 2;                  // [2]  epilogue has drained them (8 arrivals)
  uint64_t* t_ready = u_free + 2;                  // [2]  transposed tiles written (8 arrivals)
  uint64_t* t_free = t_ready + 2;                  // [2]  Gram MMAs have read them
  uint64_t* g_done = t_free + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(g_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int ntiles = (int)((a.HW + 127) / 128);
  const int my_tiles = ((int)blockIdx.x < ntiles) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&tmX1); tc::prefetch_tmap(&tmX2); tc::prefetch_tmap(&tmW);
    tc::mbar_init(wfull, 1);
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(xfull + s, 1); tc::mbar_init(xempty + s, 1); tc::mbar_init(g1_full + s, 1);
      tc::mbar_init(u_free + s, 8); tc::mbar_init(t_ready + s, 8); tc::mbar_init(t_free + s, 1);
    }
    tc::mbar_init(g_done, 1);
    tc::fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 8192 / 16; i += kFfmTcThreads) reinterpret_cast<uint4*>(sT + 2 * kGramBuf)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 1) tc::tmem_alloc(tmem_slot, 512);
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;          // cols [0,256): U double buffer; [256,384): D12; [384,448): D3

  if (warp == 0) {
    if (tc::elect_one()) {
      tc::mbar_expect_tx(wfull, 2 * 8192);
      tc::tma_load_2d(sW, &tmW, wfull, 0, 0);
      tc::tma_load_2d(sW + 8192, &tmW, wfull, 0, 64);
      for (int it = 0; it < my_tiles; ++it) {
        const int t = blockIdx.x + it * gridDim.x, s = it & 1;
        tc::mbar_wait(xempty + s, ((it >> 1) & 1) ^ 1);
        tc::mbar_expect_tx(xfull + s, 2 * kTileBytes);
        tc::tma_load_3d(sX + (s * 2 + 0) * kTileBytes, &tmX1, xfull + s, 0, t * 128, b);
        tc::tma_load_3d(sX + (s * 2 + 1) * kTileBytes, &tmX2, xfull + s, 0, t * 128, b);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    constexpr uint32_t idesc_p = tc::make_idesc_bf16(128, 64), idesc_g12 = tc::make_idesc_bf16(128, 128), idesc_g3 = tc::make_idesc_bf16(128, 64);
    constexpr uint64_t HI = (uint64_t)tc::desc_hi_sw128(1024) << 32;
    const bool leader = tc::elect_one();
    tc::mbar_wait(wfull, 0);
    uint64_t w_d = HI | (smem_u32(sW) >> 4), x_d0 = HI | (smem_u32(sX) >> 4), t_d0 = HI | (smem_u32(sT) >> 4);
    asm volatile("" : "+l"(w_d), "+l"(x_d0), "+l"(t_d0));
    auto gram = [&](int it) {                      // Gram MMAs of tile `it` (whole warp calls, the leader issues)
      const int tb = it & 1;
      tc::mbar_wait(t_ready + tb, (it >> 1) & 1);
      tc::tc_fence_after();
      if (leader) {
        const uint64_t t_d = t_d0 + (uint64_t)(tb * (kGramBuf >> 4));
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t d12 = t_d + (uint64_t)((kb * kGramKb + k * 32) >> 4);
            const uint64_t d3 = d12 + (uint64_t)((128 * 128) >> 4);
            const uint32_t acc = (it | kb | k) != 0 ? 1u : 0u;
            tc::umma_bf16(tmem_base + 256, d12, d12, idesc_g12, acc);
            tc::umma_bf16(tmem_base + 384, d3, d3, idesc_g3, acc);
          }
        tc::umma_commit(t_free + tb);
      }
      __syncwarp();
    };
    for (int it = 0; it < my_tiles; ++it) {
      const int s = it & 1;
      tc::mbar_wait(u_free + s, ((it >> 1) & 1) ^ 1);
      tc::mbar_wait(xfull + s, (it >> 1) & 1);
      tc::tc_fence_after();
      if (leader) {
        const uint64_t x_d = x_d0 + (uint64_t)(s * 2 * (kTileBytes >> 4));
#pragma unroll
        for (int st = 0; st < 2; ++st)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc::umma_bf16(tmem_base + s * 128 + st * 64, x_d + (uint64_t)(st * (kTileBytes >> 4) + k * 2), w_d + (uint64_t)(st * 512 + k * 2), idesc_p, k > 0 ? 1u : 0u);
        tc::umma_commit(g1_full + s);
        tc::umma_commit(xempty + s);
      }
      __syncwarp();
      if (it > 0) gram(it - 1);                    // the projection of tile `it` is already queued behind it
    }
    if (my_tiles > 0) gram(my_tiles - 1);
    if (leader) tc::umma_commit(g_done);
    __syncwarp();
  } else {
    const int ew = warp - 2;                       // 0..7
    const int st = ew >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;                // pixel of the tile == TMEM lane
    const int et = ew * 32 + lane;
    for (int it = 0; it < my_tiles; ++it) {
      const int t = blockIdx.x + it * gridDim.x, s = it & 1;
      const int64_t p0 = (int64_t)t * 128;
      uint8_t* sTb = sT + s * kGramBuf;
      // ---- u3 = relu(bilerp(Q[..., 64:128])) for (pixel, 32-channel half) = (et >> 1, et & 1); loads first, then the
      //      wait for the buffer (Gram MMAs of tile it - 2)
      const int pixel = et >> 1, half = et & 1;
      uint4 u3v[4];
      {
        const int64_t p = p0 + pixel;
        if (p < a.HW) {
          const unsigned pu = (unsigned)p;
          const int Y = (int)(pu / (unsigned)a.W), X = (int)(pu - (unsigned)Y * (unsigned)a.W);
          int y0, y1, x0, x1;
          float hy0, hy1, wx0, wx1;
          ffm_lr_src(Y, a.sy, a.qh, y0, y1, hy0, hy1);
          ffm_lr_src(X, a.sx, a.qw, x0, x1, wx0, wx1);
          const bf16* base = a.q + (int64_t)b * a.qh * a.qw * 128 + 64 + half * 32;
          const bf16* p00 = base + ((int64_t)y0 * a.qw + x0) * 128;
          const bf16* p01 = base + ((int64_t)y0 * a.qw + x1) * 128;
          const bf16* p10 = base + ((int64_t)y1 * a.qw + x0) * 128;
          const bf16* p11 = base + ((int64_t)y1 * a.qw + x1) * 128;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float v0[8], v1[8], v2[8], v3[8], o[8];
            load8(p00 + c * 8, v0); load8(p01 + c * 8, v1); load8(p10 + c * 8, v2); load8(p11 + c * 8, v3);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = fmaxf(hy0 * (wx0 * v0[j] + wx1 * v1[j]) + hy1 * (wx0 * v2[j] + wx1 * v3[j]), 0.f);
            u3v[c] = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
          }
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c) u3v[c] = make_uint4(0, 0, 0, 0);
        }
      }
      tc::mbar_wait(t_free + s, ((it >> 1) & 1) ^ 1);
      {
        uint8_t* kbp = sTb + (pixel >> 6) * kGramKb;
        const int pin = pixel & 63, pch = pin >> 3, pb = (pin & 7) * 2;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t w4[4] = {u3v[c].x, u3v[c].y, u3v[c].z, u3v[c].w};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int row = 128 + half * 32 + c * 8 + j;
            const uint16_t hv = (uint16_t)((j & 1) ? (w4[j >> 1] >> 16) : (w4[j >> 1] & 0xffffu));
            *reinterpret_cast<uint16_t*>(kbp + row * 128 + ((pch ^ (row & 7)) << 4) + pb) = hv;
          }
        }
      }
      // ---- u_s = relu(U_s + bu_s) (zero past the image), transposed into rows st*64 .. st*64+63
      tc::mbar_wait(g1_full + s, (it >> 1) & 1);
      tc::tc_fence_after();
      {
        const bool live = (p0 + r) < a.HW;
        uint8_t* kbp = sTb + (r >> 6) * kGramKb;
        const int pin = r & 63, pch = pin >> 3, pb = (pin & 7) * 2;
#pragma unroll
        for (int hc = 0; hc < 2; ++hc) {
          float v[32];
          tc::tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(s * 128 + st * 64 + hc * 32), v);
          const float4* bp = reinterpret_cast<const float4*>(a.bproj + st * 64 + hc * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bv = __ldg(bp + j);
            v[4 * j] = fmaxf(v[4 * j] + bv.x, 0.f); v[4 * j + 1] = fmaxf(v[4 * j + 1] + bv.y, 0.f);
            v[4 * j + 2] = fmaxf(v[4 * j + 2] + bv.z, 0.f); v[4 * j + 3] = fmaxf(v[4 * j + 3] + bv.w, 0.f);
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int row = st * 64 + hc * 32 + j;
            const bf16 hv = __float2bfloat16_rn(live ? v[j] : 0.f);
            *reinterpret_cast<bf16*>(kbp + row * 128 + ((pch ^ (row & 7)) << 4) + pb) = hv;
          }
        }
      }
      tc::tc_fence_before();
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) { tc::mbar_arrive(u_free + s); tc::mbar_arrive(t_ready + s); }
    }
    // ---- partial sums of this CTA: group 0 writes G1 (rows 0..63) / G2 (rows 64..127) of D12, group 1 writes G3
    tc::mbar_wait(g_done, 0);
    tc::tc_fence_after();
    float* out = partials + ((int64_t)b * gridDim.x + blockIdx.x) * 3 * 4096;
    const int which = st == 0 ? (r >> 6) : 2;
    const int row = r & 63;
    const uint32_t col0 = st == 0 ? (uint32_t)(256 + (r >> 6) * 64) : 384u;
#pragma unroll
    for (int hc = 0; hc < 2; ++hc) {
      float v[32];
      tc::tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + col0 + (uint32_t)(hc * 32), v);    // warp-collective
      if (my_tiles == 0) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      if (st == 0 || r < 64) {
        float4* d = reinterpret_cast<float4*>(out + which * 4096 + row * 64 + hc * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, 512);
}

static int make_px_map(CUtensorMap* m, const void* base, int coff, int ld, int64_t HW, int B, const char* what) {
  const uint64_t dims[3] = {64, (uint64_t)HW, (uint64_t)B};
  const uint64_t strides[2] = {(uint64_t)ld * 2, (uint64_t)HW * ld * 2};
  const uint32_t box[3] = {64, 128, 1};
  return make_tmap_bf16(m, reinterpret_cast<const bf16*>(base) + coff, 3, dims, strides, box, true, what);
}

}  // namespace segmif

using namespace segmif;

extern "C" int segmif_ffm_apply_lr_fwd(const void* x1, int ld1, int coff1, const void* x2, int ld2, int coff2,
                                       const void* q3, int qh, int qw, int H, int W, const void* wproj,
                                       const float* bproj, const void* folded, const float* bend, const float* ln_gamma,
                                       const float* ln_beta, float eps, void* out1, int ldo1, int coffo1, void* out2,
                                       int ldo2, int coffo2, int B, segmif_stream_t stream) {
  SEGMIF_REQUIRE(x1 && x2 && q3 && wproj && bproj && folded && bend && ln_gamma && ln_beta && out1 && out2, "ffm_apply_lr: null pointer");
  SEGMIF_REQUIRE(ld1 % 8 == 0 && ld2 % 8 == 0 && coff1 % 8 == 0 && coff2 % 8 == 0 && ldo1 % 8 == 0 && ldo2 % 8 == 0 &&
                 coffo1 % 8 == 0 && coffo2 % 8 == 0, "ffm_apply_lr: pitches/offsets must be multiples of 8");
  SEGMIF_REQUIRE(qh > 0 && qw > 0 && H > 0 && W > 0 && B > 0, "ffm_apply_lr: bad sizes");
  SEGMIF_REQUIRE((((uintptr_t)x1 | (uintptr_t)x2 | (uintptr_t)q3 | (uintptr_t)wproj | (uintptr_t)folded | (uintptr_t)out1 | (uintptr_t)out2 |
                   (uintptr_t)bproj | (uintptr_t)bend | (uintptr_t)ln_gamma | (uintptr_t)ln_beta) & 15) == 0, "ffm_apply_lr: pointers must be 16-byte aligned");
  const int64_t HW = (int64_t)H * W;
  SEGMIF_REQUIRE(HW < (1ll << 31), "ffm_apply_lr: H*W must be below 2^31");
  CUtensorMap tmX1, tmX2, tmW, tmM, tmO1, tmO2;
  int rc;
  if ((rc = make_px_map(&tmX1, x1, coff1, ld1, HW, B, "ffm_apply(x1)"))) return rc;
  if ((rc = make_px_map(&tmX2, x2, coff2, ld2, HW, B, "ffm_apply(x2)"))) return rc;
  if ((rc = make_px_map(&tmO1, out1, coffo1, ldo1, HW, B, "ffm_apply(out1)"))) return rc;
  if ((rc = make_px_map(&tmO2, out2, coffo2, ldo2, HW, B, "ffm_apply(out2)"))) return rc;
  {
    const uint64_t dims[2] = {64, 128};
    const uint64_t strides[1] = {128};
    const uint32_t box[2] = {64, 64};
    if ((rc = make_tmap_bf16(&tmW, wproj, 2, dims, strides, box, true, "ffm_apply(W)"))) return rc;
    const uint64_t mdims[2] = {64, (uint64_t)B * 4 * 64};
    if ((rc = make_tmap_bf16(&tmM, folded, 2, mdims, strides, box, true, "ffm_apply(M)"))) return rc;
  }
  FfmTcArgs a;
  a.q = reinterpret_cast<const bf16*>(q3); a.bproj = bproj; a.bend = bend; a.ln_g = ln_gamma; a.ln_b = ln_beta;
  a.eps = eps; a.sy = (float)qh / (float)H; a.sx = (float)qw / (float)W; a.qh = qh; a.qw = qw; a.H = H; a.W = W; a.HW = HW;
  const size_t smem = 2 * 8192 + 4 * 8192 + 4 * kTileBytes + 4 * kTileBytes + 2 * kTileBytes + 10 * 8 + 16;
  static bool configured = false;
  static int sms = 148;
  if (!configured) {
    cudaError_t err = cudaFuncSetAttribute(ffm_apply_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) { set_error("ffm_apply_lr: cudaFuncSetAttribute: %s", cudaGetErrorString(err)); return SEGMIF_ERR_CUDA; }
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    configured = true;
  }
  const int ntiles = (int)((HW + 127) / 128);
  const int per_image = std::min(ntiles, std::max(1, sms / B));
  dim3 grid(per_image, B);
  ffm_apply_tc_kernel<<<grid, kFfmTcThreads, smem, as_stream(stream)>>>(tmX1, tmX2, tmW, tmM, tmO1, tmO2, a);
  return check_launch("segmif_ffm_apply_lr_fwd");
}

extern "C" int segmif_ffm_gram_lr_fwd(const void* x1, int ld1, int coff1, const void* x2, int ld2, int coff2,
                                      const void* q3, int qh, int qw, int H, int W, const void* wproj,
                                      const float* bproj, float* partials, int nchunk, int B, segmif_stream_t stream) {
  SEGMIF_REQUIRE(x1 && x2 && q3 && wproj && bproj && partials, "ffm_gram_lr: null pointer");
  SEGMIF_REQUIRE(ld1 % 8 == 0 && ld2 % 8 == 0 && coff1 % 8 == 0 && coff2 % 8 == 0, "ffm_gram_lr: pitches/offsets must be multiples of 8");
  SEGMIF_REQUIRE(nchunk > 0 && B > 0 && qh > 0 && qw > 0 && H > 0 && W > 0, "ffm_gram_lr: bad sizes");
  SEGMIF_REQUIRE((((uintptr_t)x1 | (uintptr_t)x2 | (uintptr_t)q3 | (uintptr_t)wproj | (uintptr_t)bproj | (uintptr_t)partials) & 15) == 0,
                 "ffm_gram_lr: pointers must be 16-byte aligned");
  const int64_t HW = (int64_t)H * W;
  SEGMIF_REQUIRE(HW < (1ll << 31), "ffm_gram_lr: H*W must be below 2^31");
  CUtensorMap tmX1, tmX2, tmW;
  int rc;
  if ((rc = make_px_map(&tmX1, x1, coff1, ld1, HW, B, "ffm_gram(x1)"))) return rc;
  if ((rc = make_px_map(&tmX2, x2, coff2, ld2, HW, B, "ffm_gram(x2)"))) return rc;
  {
    const uint64_t dims[2] = {64, 128};
    const uint64_t strides[1] = {128};
    const uint32_t box[2] = {64, 64};
    if ((rc = make_tmap_bf16(&tmW, wproj, 2, dims, strides, box, true, "ffm_gram(W)"))) return rc;
  }
  FfmTcArgs a;
  a.q = reinterpret_cast<const bf16*>(q3); a.bproj = bproj; a.bend = nullptr; a.ln_g = nullptr; a.ln_b = nullptr;
  a.eps = 0.f; a.sy = (float)qh / (float)H; a.sx = (float)qw / (float)W; a.qh = qh; a.qw = qw; a.H = H; a.W = W; a.HW = HW;
  const size_t smem = 2 * 8192 + 4 * kTileBytes + 2 * kGramBuf + 8192 + 16 * 8 + 16;
  static bool configured = false;
  if (!configured) {
    cudaError_t err = cudaFuncSetAttribute(ffm_gram_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) { set_error("ffm_gram_lr: cudaFuncSetAttribute: %s", cudaGetErrorString(err)); return SEGMIF_ERR_CUDA; }
    configured = true;
  }
  dim3 grid(nchunk, B);
  ffm_gram_tc_kernel<<<grid, kFfmTcThreads, smem, as_stream(stream)>>>(tmX1, tmX2, tmW, a, partials);
  return check_launch("segmif_ffm_gram_lr_fwd");
}

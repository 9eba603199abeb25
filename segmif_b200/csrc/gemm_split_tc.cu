// Strict-precision contraction on the 5th-generation tensor cores (segmif_split_gemm_fwd): the fp32-parity mode of
// every dense layer on the path (nn.Linear, 1x1 conv, the 3x3 / dilated 3x3 convolutions of the fusion network).
//
// north_star asks for fp32 parity (<= 1e-3 end to end, bit-exact labels); tcgen05 has no fp32 operand kind, so an
// fp32 value is carried as THREE bf16 planes  a = a0 + a1 + a2  (a0 = bf16(a), a1 = bf16(a - a0), a2 = bf16(a - a0 - a1):
// 3 x 8 mantissa bits = the 24 bits of fp32) and a product of two such values is evaluated as the six partial products
// of order <= 2^-16:   a.w ~= a0 w0 + (a0 w1 + a1 w0 + a0 w2 + a2 w0 + a1 w1),   dropped terms <= 2^-24 |a||w|.
// Every partial product is an ordinary kind::f16 MMA (bf16 x bf16 is exact in fp32); the leading term accumulates in
// one TMEM accumulator and the five corrections in a SECOND one, so the rounding of the running fp32 sum happens
// K/16 times per output as in a plain GEMM instead of 6 K/16 times, and the corrections (2^-8 smaller) round at
// 2^-8 of that.  The epilogue adds the two accumulators.
//
// Structure = gemm_tc.cu (TMA producer warp, one MMA-issuing thread, four epilogue warps, 4-stage operand ring,
// double-buffered accumulators, persistent tile loop) with a K loop that runs over  term x tap x k-block:
//   * the A operand has one tensor map per plane; in ROW mode a tile is 128 consecutive rows of [M, K]; in PATCH mode
//     (convolutions) a tile is a 16 x 8 pixel patch of a pixel-major [B, H, W, C] tensor fetched with a 4-D box at
//     (x0 + dx[tap], y0 + dy[tap]) -- out-of-image pixels are zero-filled by the TMA unit = the conv's zero padding;
//   * the W operand is [N][tap][plane][Kp] (Kp = K rounded up to 64, zero padded), so a k-block of any (tap, plane)
//     is one 2-D box.
// nterms = 1 gives the plain bf16 product, 3 a 2-plane (16-bit) product, 6 the full fp32-grade product.
// Epilogue: bias -> act (none | ReLU | PReLU) -> + fp32 residual -> fp32 store and / or a 3-plane bf16 split store
// (the operand format of the next strict layer); any N (columns beyond N masked), any pitch (scalar path when a slice is
// not 16-byte aligned).
#include <algorithm>

#include "tc_common.cuh"

namespace segmif {

struct SplitArgs {
  const float* bias;
  const float* alpha;
  const float* res;
  float* dst;
  bf16* dsp;
  int64_t dsp_plane;
  int M, N, act;
  int ld_res, res_coff, ld_dst, dst_coff, ld_dsp, dsp_coff;
  int vec;                       // all slices 16-byte aligned: vector loads / stores
  int patch, H, W, tiles_x, tiles_per_img;
  int ntaps, nterms, nkb;
  int n_tiles, num_tiles;
  int dx[9], dy[9];
};

constexpr int kSpStages = 4;
constexpr int kSpThreads = 192;

__device__ __forceinline__ void split3_bf16(float v, bf16& h, bf16& m, bf16& l) {
  h = __float2bfloat16_rn(v);
  const float r1 = v - __bfloat162float(h);
  m = __float2bfloat16_rn(r1);
  const float r2 = r1 - __bfloat162float(m);
  l = __float2bfloat16_rn(r2);
}

__device__ __forceinline__ void split_epilogue_row32(const SplitArgs& e, float (&v)[32], int64_t m, int n, float slope) {
  const int nvalid = e.N - n < 32 ? e.N - n : 32;
  if (e.vec && nvalid == 32) {
    if (e.bias) {
      const float4* bp = reinterpret_cast<const float4*>(e.bias + n);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 bv = __ldg(bp + j);
        v[4 * j] += bv.x; v[4 * j + 1] += bv.y; v[4 * j + 2] += bv.z; v[4 * j + 3] += bv.w;
      }
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = v[j] >= 0.f ? v[j] : slope * v[j];
    if (e.res) {
      const float4* rp = reinterpret_cast<const float4*>(e.res + m * e.ld_res + e.res_coff + n);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 t = rp[j];
        v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
      }
    }
    if (e.dst) {
      float4* d = reinterpret_cast<float4*>(e.dst + m * e.ld_dst + e.dst_coff + n);
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
    if (e.dsp) {
      uint32_t ph[16], pm[16], pl[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        bf16 h0, m0, l0, h1, m1, l1;
        split3_bf16(v[2 * j], h0, m0, l0);
        split3_bf16(v[2 * j + 1], h1, m1, l1);
        ph[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
        pm[j] = (uint32_t)__bfloat16_as_ushort(m0) | ((uint32_t)__bfloat16_as_ushort(m1) << 16);
        pl[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
      }
      bf16* base = e.dsp + m * e.ld_dsp + e.dsp_coff + n;
      uint4* d0 = reinterpret_cast<uint4*>(base);
      uint4* d1 = reinterpret_cast<uint4*>(base + e.dsp_plane);
      uint4* d2 = reinterpret_cast<uint4*>(base + 2 * e.dsp_plane);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        d0[j] = make_uint4(ph[4 * j], ph[4 * j + 1], ph[4 * j + 2], ph[4 * j + 3]);
        d1[j] = make_uint4(pm[4 * j], pm[4 * j + 1], pm[4 * j + 2], pm[4 * j + 3]);
        d2[j] = make_uint4(pl[4 * j], pl[4 * j + 1], pl[4 * j + 2], pl[4 * j + 3]);
      }
    }
    return;
  }
  // scalar path: column tail (N % 32 != 0, e.g. the 9 class logits) or unaligned slices
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (j < nvalid) {
      float x = v[j];
      if (e.bias) x += __ldg(e.bias + n + j);
      x = x >= 0.f ? x : slope * x;
      if (e.res) x += e.res[m * e.ld_res + e.res_coff + n + j];
      if (e.dst) e.dst[m * e.ld_dst + e.dst_coff + n + j] = x;
      if (e.dsp) {
        bf16 h, mm, l;
        split3_bf16(x, h, mm, l);
        bf16* base = e.dsp + m * e.ld_dsp + e.dsp_coff + n + j;
        base[0] = h; base[e.dsp_plane] = mm; base[2 * e.dsp_plane] = l;
      }
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(kSpThreads, 1) gemm_split_tc_kernel(const __grid_constant__ CUtensorMap tmA0,
                                                                      const __grid_constant__ CUtensorMap tmA1,
                                                                      const __grid_constant__ CUtensorMap tmA2,
                                                                      const __grid_constant__ CUtensorMap tmB,
                                                                      const SplitArgs e) {
  constexpr int A_BYTES = 128 * 128, B_BYTES = BN * 128;
  constexpr uint32_t TMEM_COLS = 4 * BN;                   // {main, correction} x double buffer
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + kSpStages * A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + kSpStages * B_BYTES);
  uint64_t* empty = full + kSpStages;
  uint64_t* tmem_full = empty + kSpStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&tmA0);
    tc::prefetch_tmap(&tmB);
    for (int s = 0; s < kSpStages; ++s) { tc::mbar_init(full + s, 1); tc::mbar_init(empty + s, 1); }
    for (int s = 0; s < 2; ++s) { tc::mbar_init(tmem_full + s, 1); tc::mbar_init(tmem_empty + s, 4); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, TMEM_COLS);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int per_term = e.ntaps * e.nkb;
  const int total_k = e.nterms * per_term;

  if (warp == 0) {
    if (tc::elect_one()) {
      int it = 0;
      for (int tile = blockIdx.x; tile < e.num_tiles; tile += gridDim.x) {
        const int mt = tile / e.n_tiles, n0 = (tile % e.n_tiles) * BN;
        int m0 = mt * 128, b = 0, y0 = 0, x0 = 0;
        if (e.patch) {
          b = mt / e.tiles_per_img;
          const int rem = mt - b * e.tiles_per_img;
          y0 = (rem / e.tiles_x) * 16;
          x0 = (rem % e.tiles_x) * 8;
        }
        for (int term = 0; term < e.nterms; ++term) {
          const int pa = (0x120100 >> (4 * term)) & 0xF;       // A plane of term: 0 0 1 0 2 1
          const int pw = (0x102010 >> (4 * term)) & 0xF;       // W plane of term: 0 1 0 2 0 1
          const CUtensorMap* tm = pa == 0 ? &tmA0 : (pa == 1 ? &tmA1 : &tmA2);
          for (int tap = 0; tap < e.ntaps; ++tap) {
            for (int kk = 0; kk < e.nkb; ++kk, ++it) {
              const int s = it % kSpStages;
              const uint32_t ph = (it / kSpStages) & 1;
              tc::mbar_wait(empty + s, ph ^ 1);
              tc::mbar_expect_tx(full + s, A_BYTES + B_BYTES);
              if (e.patch)
                tc::tma_load_4d(sA + s * A_BYTES, tm, full + s, kk * 64, x0 + e.dx[tap], y0 + e.dy[tap], b);
              else
                tc::tma_load_2d(sA + s * A_BYTES, tm, full + s, kk * 64, m0);
              tc::tma_load_2d(sB + s * B_BYTES, &tmB, full + s, ((tap * 3 + pw) * e.nkb + kk) * 64, n0);
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    constexpr uint32_t idesc = tc::make_idesc_bf16(128, BN);
    constexpr uint64_t HI = (uint64_t)tc::desc_hi_sw128(1024) << 32;
    const bool leader = tc::elect_one();
    int it = 0, lt = 0;
    for (int tile = blockIdx.x; tile < e.num_tiles; tile += gridDim.x, ++lt) {
      const int buf = lt & 1;
      tc::mbar_wait(tmem_empty + buf, ((lt >> 1) & 1) ^ 1);
      tc::tc_fence_after();
      const uint32_t acc_main = tmem_base + (uint32_t)(buf * 2 * BN);
      for (int i = 0; i < total_k; ++i, ++it) {
        const int s = it % kSpStages;
        tc::mbar_wait(full + s, (it / kSpStages) & 1);
        tc::tc_fence_after();
        if (leader) {
          const bool corr = i >= per_term;
          const uint32_t acc = acc_main + (corr ? (uint32_t)BN : 0u);
          const bool first = (i == 0) || (i == per_term);
          uint64_t a_d = HI | (uint64_t)(smem_u32(sA + s * A_BYTES) >> 4), b_d = HI | (uint64_t)(smem_u32(sB + s * B_BYTES) >> 4);
          asm volatile("" : "+l"(a_d), "+l"(b_d));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc::umma_bf16(acc, a_d + (uint64_t)(k * 2), b_d + (uint64_t)(k * 2), idesc, (first && k == 0) ? 0u : 1u);
          tc::umma_commit(empty + s);
        }
        __syncwarp();
      }
      if (leader) tc::umma_commit(tmem_full + buf);
      __syncwarp();
    }
  } else {
    const int quad = warp & 3;
    const float slope = e.act == SEGMIF_ACT_PRELU ? *e.alpha : (e.act == SEGMIF_ACT_RELU ? 0.f : 1.f);
    int lt = 0;
    for (int tile = blockIdx.x; tile < e.num_tiles; tile += gridDim.x, ++lt) {
      const int buf = lt & 1;
      const int mt = tile / e.n_tiles, n0 = (tile % e.n_tiles) * BN;
      const int r = quad * 32 + lane;
      int64_t m;
      bool valid;
      if (e.patch) {
        const int b = mt / e.tiles_per_img;
        const int rem = mt - b * e.tiles_per_img;
        const int y = (rem / e.tiles_x) * 16 + (r >> 3), x = (rem % e.tiles_x) * 8 + (r & 7);
        valid = y < e.H && x < e.W;
        m = ((int64_t)b * e.H + y) * e.W + x;
      } else {
        m = (int64_t)mt * 128 + r;
        valid = m < e.M;
      }
      tc::mbar_wait(tmem_full + buf, (lt >> 1) & 1);
      tc::tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        float v[32];
        const uint32_t ta = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * 2 * BN + c);
        tc::tmem_ld32(ta, v);
        if (e.nterms > 1) {
          float w[32];
          tc::tmem_ld32(ta + (uint32_t)BN, w);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += w[j];
        }
        if (valid && (n0 + c) < e.N) split_epilogue_row32(e, v, m, n0 + c, slope);
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(tmem_empty + buf);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int BN>
static int launch_split(const CUtensorMap* tmA, const CUtensorMap& tmB, const SplitArgs& e, int64_t m_tiles, cudaStream_t st) {
  constexpr size_t smem = (size_t)kSpStages * (128 * 128 + BN * 128) + (2 * kSpStages + 4) * 8 + 16;
  auto kern = gemm_split_tc_kernel<BN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) { set_error("split_gemm: cudaFuncSetAttribute: %s", cudaGetErrorString(err)); return SEGMIF_ERR_CUDA; }
    configured = true;
  }
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  SplitArgs a = e;
  a.n_tiles = (int)ceil_div(e.N, BN);
  const int64_t num_tiles = (int64_t)a.n_tiles * m_tiles;
  SEGMIF_REQUIRE(num_tiles < (1ll << 31), "split_gemm: too many tiles");
  a.num_tiles = (int)num_tiles;
  const int ctas_per_sm = BN <= 64 ? 2 : 1;      // TMEM: 4*BN columns per CTA, 512 per SM
  const int grid = (int)std::min<int64_t>(num_tiles, (int64_t)sms * ctas_per_sm);
  kern<<<grid, kSpThreads, smem, st>>>(tmA[0], tmA[1], tmA[2], tmB, a);
  return check_launch("segmif_split_gemm_fwd");
}

static int split_gemm_impl(const segmif_split_gemm_params* p, cudaStream_t st) {
  SEGMIF_REQUIRE(p && p->a_planes && p->w_planes && (p->dst || p->dst_planes), "split_gemm: null pointer");
  SEGMIF_REQUIRE(p->N > 0 && p->K > 0, "split_gemm: bad sizes");
  SEGMIF_REQUIRE(p->nterms == 1 || p->nterms == 3 || p->nterms == 6, "split_gemm: nterms must be 1, 3 or 6");
  SEGMIF_REQUIRE(p->K % 8 == 0 && p->ld_a % 8 == 0 && p->a_coff % 8 == 0 && p->a_plane_stride % 8 == 0,
                 "split_gemm: K, ld_a, a_coff and the plane stride must be multiples of 8");
  SEGMIF_REQUIRE(p->a_coff + p->K <= p->ld_a, "split_gemm: channel slice exceeds pitch");
  SEGMIF_REQUIRE(((uintptr_t)p->a_planes & 15) == 0 && ((uintptr_t)p->w_planes & 15) == 0, "split_gemm: operands must be 16-byte aligned");
  SEGMIF_REQUIRE(p->act == SEGMIF_ACT_NONE || p->act == SEGMIF_ACT_RELU || (p->act == SEGMIF_ACT_PRELU && p->prelu_alpha),
                 "split_gemm: act must be none, ReLU or PReLU (with prelu_alpha)");
  const bool patch = p->ntaps > 0;
  const int ntaps = patch ? p->ntaps : 1;
  SEGMIF_REQUIRE(ntaps <= 9, "split_gemm: at most 9 taps");
  int64_t M, m_tiles;
  SplitArgs e;
  if (patch) {
    SEGMIF_REQUIRE(p->B > 0 && p->H > 0 && p->W > 0, "split_gemm: patch mode needs B, H, W");
    M = (int64_t)p->B * p->H * p->W;
    e.tiles_x = (p->W + 7) / 8;
    e.tiles_per_img = e.tiles_x * ((p->H + 15) / 16);
    m_tiles = (int64_t)e.tiles_per_img * p->B;
  } else {
    SEGMIF_REQUIRE(p->M > 0, "split_gemm: bad M");
    M = p->M;
    e.tiles_x = e.tiles_per_img = 1;
    m_tiles = ceil_div(M, 128);
  }
  const int nkb = (int)ceil_div(p->K, 64);
  const int Kp = nkb * 64;
  CUtensorMap tmA[3], tmB;
  const int nplanes = p->nterms == 1 ? 1 : (p->nterms == 3 ? 2 : 3);
  for (int pl = 0; pl < 3; ++pl) {
    const bf16* base = reinterpret_cast<const bf16*>(p->a_planes) + (pl < nplanes ? pl : 0) * p->a_plane_stride + p->a_coff;
    int rc;
    if (patch) {
      const uint64_t dims[4] = {(uint64_t)p->K, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->B};
      const uint64_t strides[3] = {(uint64_t)p->ld_a * 2, (uint64_t)p->W * p->ld_a * 2, (uint64_t)p->H * p->W * p->ld_a * 2};
      const uint32_t box[4] = {64, 8, 16, 1};
      rc = make_tmap_bf16(&tmA[pl], base, 4, dims, strides, box, true, "split_gemm(A, patch)", p->K == p->ld_a ? 256 : 64);
    } else {
      const uint64_t dims[2] = {(uint64_t)p->K, (uint64_t)M};
      const uint64_t strides[1] = {(uint64_t)p->ld_a * 2};
      const uint32_t box[2] = {64, 128};
      rc = make_tmap_bf16(&tmA[pl], base, 2, dims, strides, box, true, "split_gemm(A)", p->K == p->ld_a ? 256 : 64);
    }
    if (rc) return rc;
  }
  const int BN = p->N > 64 ? 128 : (p->N > 32 ? 64 : 32);
  {
    const uint64_t dims[2] = {(uint64_t)ntaps * 3 * Kp, (uint64_t)p->N};
    const uint64_t strides[1] = {(uint64_t)ntaps * 3 * Kp * 2};
    const uint32_t box[2] = {64, (uint32_t)BN};
    int rc = make_tmap_bf16(&tmB, p->w_planes, 2, dims, strides, box, true, "split_gemm(W)");
    if (rc) return rc;
  }
  e.bias = p->bias; e.alpha = p->prelu_alpha; e.res = p->residual; e.dst = p->dst;
  e.dsp = reinterpret_cast<bf16*>(p->dst_planes); e.dsp_plane = p->dst_plane_stride;
  e.M = (int)std::min<int64_t>(M, 0x7fffffff); e.N = p->N; e.act = p->act;
  SEGMIF_REQUIRE(M < (1ll << 31), "split_gemm: M must be below 2^31");
  e.ld_res = p->ld_res; e.res_coff = p->res_coff; e.ld_dst = p->ld_dst; e.dst_coff = p->dst_coff;
  e.ld_dsp = p->ld_dp; e.dsp_coff = p->dp_coff;
  SEGMIF_REQUIRE(!p->dst || p->dst_coff + p->N <= p->ld_dst, "split_gemm: dst slice exceeds pitch");
  SEGMIF_REQUIRE(!p->dst_planes || p->dp_coff + p->N <= p->ld_dp, "split_gemm: dst_planes slice exceeds pitch");
  SEGMIF_REQUIRE(!p->residual || p->res_coff + p->N <= p->ld_res, "split_gemm: residual slice exceeds pitch");
  bool vec = (!p->bias || ((uintptr_t)p->bias & 15) == 0);
  if (p->dst) vec = vec && ((uintptr_t)p->dst & 15) == 0 && p->ld_dst % 4 == 0 && p->dst_coff % 4 == 0;
  if (p->residual) vec = vec && ((uintptr_t)p->residual & 15) == 0 && p->ld_res % 4 == 0 && p->res_coff % 4 == 0;
  if (p->dst_planes) vec = vec && ((uintptr_t)p->dst_planes & 15) == 0 && p->ld_dp % 8 == 0 && p->dp_coff % 8 == 0 && p->dst_plane_stride % 8 == 0;
  e.vec = vec ? 1 : 0;
  e.patch = patch ? 1 : 0; e.H = p->H; e.W = p->W;
  e.ntaps = ntaps; e.nterms = p->nterms; e.nkb = nkb;
  for (int t = 0; t < 9; ++t) { e.dx[t] = (patch && t < ntaps) ? p->tap_dx[t] : 0; e.dy[t] = (patch && t < ntaps) ? p->tap_dy[t] : 0; }
  e.n_tiles = e.num_tiles = 0;
  if (BN == 128) return launch_split<128>(tmA, tmB, e, m_tiles, st);
  if (BN == 64) return launch_split<64>(tmA, tmB, e, m_tiles, st);
  return launch_split<32>(tmA, tmB, e, m_tiles, st);
}

}  // namespace segmif

using namespace segmif;

extern "C" int segmif_split_gemm_fwd(const segmif_split_gemm_params* p, segmif_stream_t stream) {
  return split_gemm_impl(p, as_stream(stream));
}

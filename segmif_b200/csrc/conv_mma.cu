// Implicit-GEMM convolution / linear layer, bf16 operands, fp32 accumulation (segmif_conv_fwd).
//
// GEMM view: M = B*Ho*Wo output pixels, N = Cout, K = KH*KW*Cin.  K is walked channel-chunk outer /
// tap inner so that the KH*KW shifted reads of one 32-channel slab hit L1 instead of L2 (cp.async.ca).
// A rows are gathered per pixel with zero fill outside the image (cp.async src-size 0), so padding,
// stride and dilation cost nothing and no im2col buffer exists.  4-stage cp.async ring, XOR-swizzled
// 64-byte smem rows, ldmatrix + mma.sync m16n8k16.  The epilogue fuses bias, ReLU/PReLU/GELU, the
// residual add and the channel-offset store that implements DRDB's dense concatenation in place.
#include <algorithm>

#include "common.cuh"

namespace segmif {

struct ConvArgs {
  const bf16* src;
  const bf16* wgt;
  const float* bias;
  const float* alpha;
  const void* res;
  void* dst;
  int B, H, W, Cin, ld_src, src_coff;
  int KH, KW, stride, pad, dil, Ho, Wo, Cout;
  int act, res_dtype, ld_res, res_coff, dst_dtype, ld_dst, dst_coff;
  int M;
  int ksplit;      // > 1: blockIdx.z owns a slice of the k-blocks and accumulates into a pre-zeroed fp32 dst with atomics
};

constexpr int kBK = 32;      // channels per k-block (64 bytes per smem row)
constexpr int kStages = 4;

__device__ __forceinline__ int swz64(int row, int chunk) { return chunk ^ ((row >> 1) & 3); }

template <int BM, int BN, int WARPS_M, int WARPS_N>
__global__ void __launch_bounds__(WARPS_M* WARPS_N * 32) conv_mma_kernel(const ConvArgs a) {
  constexpr int T = WARPS_M * WARPS_N * 32;
  constexpr int WM = BM / WARPS_M, WN = BN / WARPS_N;
  constexpr int MT = WM / 16, NT = WN / 8;
  static_assert(WN == 32, "warp N tile fixed at 32");
  constexpr int A_ITERS = (BM * 4) / T;
  constexpr int B_ITERS = (BN * 4 + T - 1) / T;
  static_assert((BM * 4) % T == 0, "A tile must divide evenly");

  extern __shared__ __align__(128) uint8_t smem_raw[];
  bf16* sA = reinterpret_cast<bf16*>(smem_raw);                       // [kStages][BM][32]
  bf16* sB = sA + kStages * BM * kBK;                                 // [kStages][BN][32]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int warp_m = warp / WARPS_N, warp_n = warp % WARPS_N;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int taps = a.KH * a.KW;
  const int KB_all = (a.Cin / kBK) * taps;
  const int kb_per = (KB_all + a.ksplit - 1) / a.ksplit;
  const int kb_first = blockIdx.z * kb_per;
  const int KB = max(0, min(KB_all, kb_first + kb_per) - kb_first);      // k-blocks of this split

  // ---- per-thread gather coordinates of the A rows this thread copies (fixed for the whole K loop)
  int a_iy0[A_ITERS], a_ix0[A_ITERS];
  int64_t a_boff[A_ITERS];
  bool a_ok[A_ITERS];
#pragma unroll
  for (int i = 0; i < A_ITERS; ++i) {
    const int row = (tid + i * T) >> 2;
    const int m = m0 + row;
    a_ok[i] = m < a.M;
    const int mm = a_ok[i] ? m : 0;
    const int ox = mm % a.Wo;
    const int t2 = mm / a.Wo;
    const int oy = t2 % a.Ho;
    const int b = t2 / a.Ho;
    a_iy0[i] = oy * a.stride - a.pad;
    a_ix0[i] = ox * a.stride - a.pad;
    a_boff[i] = (int64_t)b * a.H * a.W;
  }

  auto load_stage = [&](int stage, int kb_local) {
    const int kb = kb_first + kb_local;
    const int cidx = kb / taps, tap = kb - cidx * taps;
    const int ky = tap / a.KW, kx = tap - ky * a.KW;
    bf16* dA = sA + stage * BM * kBK;
#pragma unroll
    for (int i = 0; i < A_ITERS; ++i) {
      const int idx = tid + i * T;
      const int row = idx >> 2, chunk = idx & 3;
      const int iy = a_iy0[i] + ky * a.dil, ix = a_ix0[i] + kx * a.dil;
      const bool ok = a_ok[i] && (unsigned)iy < (unsigned)a.H && (unsigned)ix < (unsigned)a.W;
      const bf16* g = a.src;
      if (ok) g += (a_boff[i] + (int64_t)iy * a.W + ix) * a.ld_src + a.src_coff + cidx * kBK + chunk * 8;
      cp_async16_ca(smem_u32(dA + row * kBK + swz64(row, chunk) * 8), g, ok ? 16 : 0);
    }
    bf16* dB = sB + stage * BN * kBK;
    const int64_t wk = (int64_t)tap * a.Cin + cidx * kBK;
    const int64_t wrow = (int64_t)taps * a.Cin;
#pragma unroll
    for (int i = 0; i < B_ITERS; ++i) {
      const int idx = tid + i * T;
      if (idx < BN * 4) {
        const int row = idx >> 2, chunk = idx & 3;
        const bool ok = (n0 + row) < a.Cout;
        const bf16* g = ok ? a.wgt + (int64_t)(n0 + row) * wrow + wk + chunk * 8 : a.wgt;
        cp_async16_cg(smem_u32(dB + row * kBK + swz64(row, chunk) * 8), g, ok ? 16 : 0);
      }
    }
  };

  float acc[MT][NT][4];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.0f;

#pragma unroll
  for (int s = 0; s < kStages - 1; ++s) {
    if (s < KB) load_stage(s, s);
    cp_async_commit();
  }

  for (int kb = 0; kb < KB; ++kb) {
    cp_async_wait<kStages - 2>();
    __syncthreads();
    const int nk = kb + kStages - 1;
    if (nk < KB) load_stage(nk % kStages, nk);
    cp_async_commit();

    const bf16* tA = sA + (kb % kStages) * BM * kBK;
    const bf16* tB = sB + (kb % kStages) * BN * kBK;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {     // two k16 steps per 32-channel block
      uint32_t af[MT][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int row = warp_m * WM + mt * 16 + (lane & 15);
        const int chunk = ks * 2 + (lane >> 4);
        ldmatrix_x4(af[mt], smem_u32(tA + row * kBK + swz64(row, chunk) * 8));
      }
      uint32_t bfr[NT / 2][4];
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {
        const int row = warp_n * WN + np * 16 + (lane & 7) + ((lane >> 4) << 3);
        const int chunk = ks * 2 + ((lane >> 3) & 1);
        ldmatrix_x4(bfr[np], smem_u32(tB + row * kBK + swz64(row, chunk) * 8));
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
          mma_bf16_16816(acc[mt][nt], af[mt], bfr[nt >> 1][(nt & 1) * 2], bfr[nt >> 1][(nt & 1) * 2 + 1]);
    }
  }
  cp_async_wait<0>();

  // ---- epilogue -------------------------------------------------------------------------------
  const float alpha = (a.act == SEGMIF_ACT_PRELU) ? *a.alpha : 0.0f;
  const int g = lane >> 2, tq = lane & 3;
  const bool dst_vec = ((a.ld_dst | a.dst_coff) & 1) == 0;
  const bool res_vec = ((a.ld_res | a.res_coff) & 1) == 0;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int m = m0 + warp_m * WM + mt * 16 + g + half * 8;
      if (m >= a.M) continue;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int n = n0 + warp_n * WN + nt * 8 + tq * 2;
        if (n >= a.Cout) continue;
        const bool pair = (n + 1) < a.Cout;
        float v0 = acc[mt][nt][half * 2 + 0], v1 = acc[mt][nt][half * 2 + 1];
        if (a.bias && blockIdx.z == 0) {
          v0 += a.bias[n];
          if (pair) v1 += a.bias[n + 1];
        }
        if (a.ksplit > 1) {                 // split-K: fp32 atomics into the zeroed destination, nothing else fused
          float* d = reinterpret_cast<float*>(a.dst) + (int64_t)m * a.ld_dst + a.dst_coff + n;
          atomicAdd(d, v0);
          if (pair) atomicAdd(d + 1, v1);
          continue;
        }
        v0 = apply_act(v0, a.act, alpha);
        v1 = apply_act(v1, a.act, alpha);
        if (a.res) {
          const int64_t ro = (int64_t)m * a.ld_res + a.res_coff + n;
          if (a.res_dtype == SEGMIF_F32) {
            const float* r = reinterpret_cast<const float*>(a.res) + ro;
            if (pair && res_vec) { float2 t = *reinterpret_cast<const float2*>(r); v0 += t.x; v1 += t.y; }
            else { v0 += r[0]; if (pair) v1 += r[1]; }
          } else {
            const bf16* r = reinterpret_cast<const bf16*>(a.res) + ro;
            if (pair && res_vec) { float2 t = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(r)); v0 += t.x; v1 += t.y; }
            else { v0 += __bfloat162float(r[0]); if (pair) v1 += __bfloat162float(r[1]); }
          }
        }
        const int64_t d_o = (int64_t)m * a.ld_dst + a.dst_coff + n;
        if (a.dst_dtype == SEGMIF_F32) {
          float* d = reinterpret_cast<float*>(a.dst) + d_o;
          if (pair && dst_vec) *reinterpret_cast<float2*>(d) = make_float2(v0, v1);
          else { d[0] = v0; if (pair) d[1] = v1; }
        } else {
          bf16* d = reinterpret_cast<bf16*>(a.dst) + d_o;
          if (pair && dst_vec) *reinterpret_cast<uint32_t*>(d) = pack_bf16x2(v0, v1);
          else { d[0] = __float2bfloat16_rn(v0); if (pair) d[1] = __float2bfloat16_rn(v1); }
        }
      }
    }
  }
}

template <int BM, int BN, int WARPS_M, int WARPS_N>
static int launch_conv(const ConvArgs& a, cudaStream_t st) {
  constexpr int smem = kStages * (BM + BN) * kBK * (int)sizeof(bf16);
  auto kern = conv_mma_kernel<BM, BN, WARPS_M, WARPS_N>;
  static bool configured = false;     // one-time opt-in to > 48 KB dynamic smem (idempotent, race-benign)
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { set_error("conv: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return SEGMIF_ERR_CUDA; }
    configured = true;
  }
  ConvArgs b = a;
  const int64_t tiles = ceil_div(a.M, BM) * ceil_div(a.Cout, BN);
  const int KB = (a.Cin / kBK) * a.KH * a.KW;
  b.ksplit = 1;
  // Few output tiles but a long reduction (Attention.sr: M = B*Nk = 2400 rows, K up to 4096): spread K over CTAs.
  if (tiles < 100 && KB >= 16 && a.dst_dtype == SEGMIF_F32 && a.act == SEGMIF_ACT_NONE && a.res == nullptr &&
      a.ld_dst == a.Cout && a.dst_coff == 0) {
    b.ksplit = (int)std::min<int64_t>(KB / 4, std::max<int64_t>(1, (148 * 2) / tiles));
    if (b.ksplit > 1) {
      cudaError_t e = cudaMemsetAsync(a.dst, 0, (size_t)a.M * a.Cout * sizeof(float), st);
      if (e != cudaSuccess) { set_error("conv: cudaMemsetAsync failed: %s", cudaGetErrorString(e)); return SEGMIF_ERR_CUDA; }
    }
  }
  dim3 grid((unsigned)ceil_div(a.M, BM), (unsigned)ceil_div(a.Cout, BN), (unsigned)b.ksplit);
  kern<<<grid, WARPS_M * WARPS_N * 32, smem, st>>>(b);
  return check_launch("segmif_conv_fwd");
}

}  // namespace segmif

using namespace segmif;

extern "C" int segmif_conv_fwd(const segmif_conv_params* p, segmif_stream_t stream) {
  SEGMIF_REQUIRE(p && p->src && p->weight && p->dst, "conv: null pointer");
  SEGMIF_REQUIRE(p->Cin > 0 && p->Cin % 32 == 0, "conv: Cin=%d must be a positive multiple of 32", p->Cin);
  SEGMIF_REQUIRE(p->KH >= 1 && p->KW >= 1 && p->stride >= 1 && p->dil >= 1 && p->pad >= 0, "conv: bad geometry");
  SEGMIF_REQUIRE(p->B > 0 && p->H > 0 && p->W > 0 && p->Ho > 0 && p->Wo > 0 && p->Cout > 0, "conv: bad sizes");
  SEGMIF_REQUIRE((p->ld_src % 8) == 0 && (p->src_coff % 8) == 0, "conv: src pitch/offset must be multiples of 8 elements");
  SEGMIF_REQUIRE(p->src_coff + p->Cin <= p->ld_src, "conv: src channels exceed pitch");
  SEGMIF_REQUIRE(p->dst_coff + p->Cout <= p->ld_dst, "conv: dst channels exceed pitch");
  SEGMIF_REQUIRE(p->act != SEGMIF_ACT_PRELU || p->prelu_alpha, "conv: PReLU needs prelu_alpha");
  SEGMIF_REQUIRE(p->pre_add == nullptr, "conv: pre_add is only supported by segmif_conv3x3_tc_fwd");
  SEGMIF_REQUIRE((reinterpret_cast<uintptr_t>(p->src) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->weight) & 15) == 0,
                 "conv: src/weight must be 16-byte aligned");
  const int64_t M = (int64_t)p->B * p->Ho * p->Wo;
  SEGMIF_REQUIRE(M < (1ll << 31), "conv: too many output pixels");
  ConvArgs a;
  a.src = reinterpret_cast<const bf16*>(p->src);
  a.wgt = reinterpret_cast<const bf16*>(p->weight);
  a.bias = p->bias; a.alpha = p->prelu_alpha; a.res = p->residual; a.dst = p->dst;
  a.B = p->B; a.H = p->H; a.W = p->W; a.Cin = p->Cin; a.ld_src = p->ld_src; a.src_coff = p->src_coff;
  a.KH = p->KH; a.KW = p->KW; a.stride = p->stride; a.pad = p->pad; a.dil = p->dil; a.Ho = p->Ho; a.Wo = p->Wo;
  a.Cout = p->Cout; a.act = p->act; a.res_dtype = p->res_dtype; a.ld_res = p->ld_res; a.res_coff = p->res_coff;
  a.dst_dtype = p->dst_dtype; a.ld_dst = p->ld_dst; a.dst_coff = p->dst_coff; a.M = (int)M;
  cudaStream_t st = as_stream(stream);
  if (p->Cout <= 32) return launch_conv<256, 32, 8, 1>(a, st);
  if (p->Cout <= 64 || (p->Cout % 128) != 0 && (p->Cout % 64) == 0) return launch_conv<128, 64, 4, 2>(a, st);
  return launch_conv<128, 128, 2, 4>(a, st);
}

// Weight (and bias) gradient of a linear layer on tcgen05 tensor cores:
//     dW[co][ci] = sum_t dY[t][co] * X[t][ci]          db[co] = sum_t dY[t][co]
// (core/mix_transformer.py Mlp / Attention / OverlapPatchEmbed projections, core/segformer_head.py MLP + linear_fuse,
// core/model_fusion.py CrossPath projections -- backward of train.py:216-226, :361-381).
// The contraction index is the TOKEN, so with token-major activations both operands are "MN-major" (the non-contracted index
// is the contiguous one) -- the operand form of wgrad_tc.cu, without taps or halos:
//   * D[ci][co]: M = 128 input channels (two 64-channel SW128 atoms, LBO apart), N = up to 256 output channels (up to four
//     atoms), K = 16 tokens per MMA = 16 consecutive 128-byte rows of the TMA tiles; fp32 accumulator in TMEM for the whole
//     token range of the block;
//   * a block walks its token range in 64-token stages through a 4-deep TMA ring (48 KB per stage at N = 256);
//   * the bias gradient rides along: a second accumulator takes  ONES[128 x 16] x dY  (every row = the column sums), so
//     dY is read from HBM once for both gradients (the separate column-sum kernel re-read it: 1.2 ms of a 10.8 ms step);
//   * grid = (token chunks, ceil(Cin / 128), ceil(Cout / 256)); per-block partials [chunk][Cout][Cin] are reduced in fixed
//     order by the existing wgrad_reduce kernels (deterministic); the bias partials are added atomically, as the column-sum
//     kernel did.
// The mma.sync version (train_ops.cu wgrad_lin_kernel, ldmatrix.trans on both operands) ran at ~120 TFLOP/s and needed 2 x 148 / tiles
// chunks to fill the machine.  Warp roles: 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2..5 = final epilogue.
#include <algorithm>

#include "tc_common.cuh"

namespace segmif {

namespace {

constexpr int kThreads = 192;
constexpr int KT = 64;                        // tokens per stage
constexpr int SUB = KT * 128;                 // one 64-channel atom of a stage: 64 rows of 128 B
constexpr int kMaxBN = 256;
constexpr int kOnesBytes = 2 * 2048;          // two atoms of 16 rows x 128 B of bf16 1.0

struct WlArgs {
  float* partials;                            // [nchunk][Cout][Cin]
  float* dbias;                               // += column sums of dY, or nullptr
  int64_t P, tok_per_chunk;
  int Cin, Cout, BN, nstages;
};

// MN-major shared-memory operand descriptor (see wgrad_tc.cu): LBO = distance between 64-element atoms along M / N,
// SBO = distance between 8-row groups along K, 128-byte swizzle.
__device__ __forceinline__ uint64_t mn_desc128(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(kThreads, 1) wgrad_lin_tc_kernel(const __grid_constant__ CUtensorMap tmX,
                                                                   const __grid_constant__ CUtensorMap tmY, const WlArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[4], empty[4], done;
  __shared__ uint32_t tmem_slot;
  const int NS = a.nstages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ci0 = blockIdx.y * 128, co0 = blockIdx.z * a.BN;
  const bool two_atoms = ci0 + 64 < a.Cin;
  const int ny = min(a.BN, ((a.Cout - co0) + 63) & ~63) >> 6;          // 64-channel atoms of dY this block holds
  const int N = ny * 64;
  const int stage_bytes = (2 + (a.BN >> 6)) * SUB;
  uint8_t* ones = smem + NS * stage_bytes;
  const int64_t t0 = (int64_t)blockIdx.x * a.tok_per_chunk;
  const int64_t t1 = min(a.P, t0 + a.tok_per_chunk);
  const int nst = (int)((t1 - t0 + KT - 1) / KT);                      // >= 1 (host)
  const bool do_bias = a.dbias != nullptr && blockIdx.y == 0;

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
  if (!two_atoms) {                                                   // the unused half of A must read as zeros
    for (int s = 0; s < NS; ++s) {
      uint4* z = reinterpret_cast<uint4*>(smem + s * stage_bytes + SUB);
      for (int i = threadIdx.x; i < SUB / 16; i += kThreads) z[i] = make_uint4(0, 0, 0, 0);
    }
  }
  for (int i = threadIdx.x; i < kOnesBytes / 4; i += kThreads) reinterpret_cast<uint32_t*>(ones)[i] = 0x3F803F80u;   // bf16 1.0 x 2
  tc::fence_proxy_async();
  if (warp == 1) tc::tmem_alloc(&tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    if (tc::elect_one()) {
      const uint32_t bytes = (uint32_t)((1 + (two_atoms ? 1 : 0) + ny) * SUB);
      for (int it = 0; it < nst; ++it) {
        const int s = it % NS;
        tc::mbar_wait(empty + s, ((it / NS) & 1) ^ 1);
        tc::mbar_expect_tx(full + s, bytes);
        uint8_t* st = smem + s * stage_bytes;
        const int row = (int)(t0 + (int64_t)it * KT);
        tc::tma_load_2d(st, &tmX, full + s, ci0, row);
        if (two_atoms) tc::tma_load_2d(st + SUB, &tmX, full + s, ci0 + 64, row);
        for (int j = 0; j < ny; ++j) tc::tma_load_2d(st + (2 + j) * SUB, &tmY, full + s, co0 + j * 64, row);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    const uint32_t idesc = tc::make_idesc_bf16(128, N) | (1u << 15) | (1u << 16);      // A and B MN-major
    const uint32_t so = smem_u32(ones);
    for (int it = 0; it < nst; ++it) {
      const int s = it % NS;
      tc::mbar_wait(full + s, (it / NS) & 1);
      tc::tc_fence_after();
      if (tc::elect_one()) {
        const uint32_t sx = smem_u32(smem + s * stage_bytes);
        const uint32_t sy = sx + 2 * SUB;
#pragma unroll
        for (int k = 0; k < KT / 16; ++k) {
          const uint64_t ad = mn_desc128(sx + k * 2048, SUB, 1024);
          const uint64_t bd = mn_desc128(sy + k * 2048, SUB, 1024);
          const uint32_t acc = (it | k) != 0 ? 1u : 0u;
          tc::umma_bf16(tmem_base, ad, bd, idesc, acc);
          if (do_bias) tc::umma_bf16(tmem_base + kMaxBN, mn_desc128(so, 2048, 1024), bd, idesc, acc);
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
    float* out = a.partials + (size_t)blockIdx.x * a.Cout * a.Cin;
#pragma unroll 1
    for (int c = 0; c < N; c += 32) {
      float v[32];
      tc::tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c, v);
      if (ci < a.Cin) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (co0 + c + j < a.Cout) out[(size_t)(co0 + c + j) * a.Cin + ci] = v[j];
      }
      if (do_bias && quad == 0) {                       // every row of the second accumulator holds the column sums: take row 0
        tc::tmem_ld32(tmem_base + (uint32_t)(kMaxBN + c), v);
        if (lane == 0) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (co0 + c + j < a.Cout) atomicAdd(a.dbias + co0 + c + j, v[j]);
        }
      }
    }
    tc::tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, 512);
}

}  // namespace

bool wgrad_lin_tc_ok(int64_t P, int Cin, int Cout, int ldy, int ldx) {
  return P > 0 && P < ((int64_t)1 << 31) && Cin % 8 == 0 && Cout % 8 == 0 && ldy % 8 == 0 && ldx % 8 == 0 && Cin > 0 && Cout > 0;
}

// token chunks (= partial buffers) and tokens per chunk (a multiple of the 64-token stage; every chunk is non-empty)
void wgrad_lin_tc_plan(int64_t P, int Cin, int Cout, int* nchunk, int64_t* tok_per_chunk, int* bn) {
  const int BN = std::min(kMaxBN, (Cout + 63) & ~63);
  const int64_t tiles = (int64_t)((Cin + 127) / 128) * ((Cout + BN - 1) / BN);
  int64_t want = std::max<int64_t>(1, std::min<int64_t>((P + KT - 1) / KT, 148 / std::max<int64_t>(1, tiles)));
  const int64_t per = (((P + want - 1) / want) + KT - 1) / KT * KT;
  *tok_per_chunk = per;
  *nchunk = (int)((P + per - 1) / per);
  *bn = BN;
}

// dy, x: channel offsets already applied.  partials [nchunk][Cout][Cin]; dbias (+=) may be null.
int wgrad_lin_tc(const void* dy, int ldy, const void* x, int ldx, int64_t P, int Cin, int Cout, float* partials, int nchunk,
                 int64_t tok_per_chunk, int bn, float* dbias, cudaStream_t st) {
  CUtensorMap tmX, tmY;
  const uint32_t box[2] = {64, (uint32_t)KT};
  {
    const uint64_t dims[2] = {(uint64_t)Cin, (uint64_t)P}, strides[1] = {(uint64_t)ldx * 2};
    if (int rc = make_tmap_bf16(&tmX, x, 2, dims, strides, box, true, "wgrad_lin_tc(X)", Cin == ldx ? 256 : 128)) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)Cout, (uint64_t)P}, strides[1] = {(uint64_t)ldy * 2};
    if (int rc = make_tmap_bf16(&tmY, dy, 2, dims, strides, box, true, "wgrad_lin_tc(dY)", Cout == ldy ? 256 : 128)) return rc;
  }
  WlArgs a;
  a.partials = partials; a.dbias = dbias; a.P = P; a.tok_per_chunk = tok_per_chunk; a.Cin = Cin; a.Cout = Cout; a.BN = bn;
  const int stage_bytes = (2 + bn / 64) * SUB;
  a.nstages = std::max(2, std::min(4, (227 * 1024 - 2048 - kOnesBytes) / stage_bytes));
  const size_t smem = (size_t)a.nstages * stage_bytes + kOnesBytes + 1024;
  static bool cfg = false;
  if (!cfg) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_lin_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * (2 + kMaxBN / 64) * SUB + kOnesBytes + 1024);
    if (e != cudaSuccess) { set_error("wgrad_lin_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return SEGMIF_ERR_CUDA; }
    cfg = true;
  }
  dim3 grid(nchunk, (Cin + 127) / 128, (Cout + bn - 1) / bn);
  wgrad_lin_tc_kernel<<<grid, kThreads, smem, st>>>(tmX, tmY, a);
  return check_launch("segmif_wgrad_lin (tcgen05)");
}

}  // namespace segmif

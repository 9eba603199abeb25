// Linear layer / 1x1 conv on the 5th-generation tensor cores (segmif_linear_tc_fwd):
//   dst[m, n] = residual[m, n] + act(bias[n] + sum_k A[m, k] W[n, k])
// A: bf16 pixel-major rows (pitch lda), W: bf16 [N, K] K-major.  Both operands are staged by TMA
// (cp.async.bulk.tensor.2d, 128-byte swizzle, out-of-range rows/columns zero-filled) into a 4-stage
// shared-memory ring; ONE thread issues tcgen05.mma (M=128, N=BN, K=16 per instruction) accumulating in
// TMEM; four epilogue warps read the accumulator with tcgen05.ld (one output row per thread) and fuse
// bias / ReLU / PReLU / GELU / residual and the channel-slice store.
// Warp roles: 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2..5 = epilogue (TMEM lane quadrant = warp % 4).
#include <algorithm>

#include "dataflow.cuh"
#include "tc_common.cuh"

#ifndef SEGMIF_TC_STAGES_SMALL
#define SEGMIF_TC_STAGES_SMALL 3
#endif

namespace segmif {

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  if (!fn) set_error("cuTensorMapEncodeTiled is not available from this driver");
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, bool swizzle128, const char* what, int l2_promotion_bytes) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return SEGMIF_ERR_CUDA;
  cuuint64_t gd[5], gs[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   l2_promotion_bytes >= 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                   : l2_promotion_bytes >= 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                   : l2_promotion_bytes >= 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("%s: cuTensorMapEncodeTiled failed (CUresult %d)", what, (int)r);
    return SEGMIF_ERR_CUDA;
  }
  return SEGMIF_OK;
}

struct TcEpilogue {
  const float* bias;
  const float* alpha;
  const void* res;
  void* dst;
  int M, N, act, res_dtype, ld_res, res_coff, dst_dtype, ld_dst, dst_coff;
  const float* row_scale;      // optional: act(.) is multiplied by row_scale[row / rows_per_scale] before the residual add
  int rows_per_scale;          // (timm DropPath in train mode: one 0 or 1/keep factor per sample)
};

// one output row (this thread) x 32 columns starting at n.  `slope`: act(v) = v >= 0 ? v : slope*v covers none (1),
// ReLU (0) and PReLU (alpha) without branches; GELU is a separate template flag so that its 32 inlined erf bodies do
// not bloat the common instantiation (the epilogue has to stay in the instruction cache).
// The residual row segment (32 values) is fetched into registers ahead of the accumulator it is added to: issued before
// the wait on the MMA barrier / while the previous 32 columns are processed, so its L2 / HBM latency is off the critical
// path (ncu before: long-scoreboard 10.6 stalls per issue in the DRDB 1x1 launch, the epilogue waiting on these loads).
struct ResRow32 { uint4 q[8]; };
__device__ __forceinline__ void load_res_row32(const TcEpilogue& e, ResRow32& r, int64_t m, int n) {
  const int64_t ro = m * e.ld_res + e.res_coff + n;
  if (e.res_dtype == SEGMIF_F32) {
    const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(e.res) + ro);
#pragma unroll
    for (int j = 0; j < 8; ++j) r.q[j] = p[j];
  } else {
    const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(e.res) + ro);
#pragma unroll
    for (int j = 0; j < 4; ++j) r.q[j] = p[j];
  }
}

template <bool GELU>
__device__ __forceinline__ void epilogue_row32(const TcEpilogue& e, float (&v)[32], const ResRow32& rr, int64_t m, int n, float slope,
                                               float rscale) {
  if (e.bias) {
    const float4* bp = reinterpret_cast<const float4*>(e.bias + n);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 bv = __ldg(bp + j);
      v[4 * j] += bv.x; v[4 * j + 1] += bv.y; v[4 * j + 2] += bv.z; v[4 * j + 3] += bv.w;
    }
  }
  if (GELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = v[j] >= 0.f ? v[j] : slope * v[j];
  }
  if (e.row_scale) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= rscale;
  }
  if (e.res) {
    if (e.res_dtype == SEGMIF_F32) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint4 t = rr.q[j];
        v[4 * j] += __uint_as_float(t.x); v[4 * j + 1] += __uint_as_float(t.y); v[4 * j + 2] += __uint_as_float(t.z); v[4 * j + 3] += __uint_as_float(t.w);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 t = rr.q[j];
        const float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y), c = unpack_bf16x2(t.z), d = unpack_bf16x2(t.w);
        v[8 * j] += a.x; v[8 * j + 1] += a.y; v[8 * j + 2] += b.x; v[8 * j + 3] += b.y;
        v[8 * j + 4] += c.x; v[8 * j + 5] += c.y; v[8 * j + 6] += d.x; v[8 * j + 7] += d.y;
      }
    }
  }
  const int64_t d_o = m * e.ld_dst + e.dst_coff + n;
  if (e.dst_dtype == SEGMIF_F32) {
    float4* d = reinterpret_cast<float4*>(reinterpret_cast<float*>(e.dst) + d_o);
#pragma unroll
    for (int j = 0; j < 8; ++j) d[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  } else {
    uint4* d = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(e.dst) + d_o);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      d[j] = make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                        pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
  }
}

constexpr int kTcStagesDefault = 4;
// BN <= 64: 3 stages (72 KB / 60 KB of shared memory) so that THREE CTAs share an SM -- the N <= 64 launches are HBM bound (the DRDB 1x1:
// 1.4 GB per launch) and their row-per-thread epilogue waits on residual loads; a third CTA per SM keeps more loads in flight.
template <int BN> struct TcStages { static constexpr int value = BN <= 64 ? (SEGMIF_TC_STAGES_SMALL) : 4; };
constexpr int kTcThreads = 192;

// dataflow.cuh: the A rows are pixels of [B, H, W] images written by a concurrently running producer stage
struct GemmDfDev {
  DfDep dep;
  unsigned* error;
  unsigned long long* timing;
  int H, W, HW;
};

// Persistent: each CTA walks output tiles (m-tile major, n-tile minor) with a stride of gridDim.x.  The 4-stage
// operand ring runs across tile boundaries and the TMEM accumulator is double buffered, so TMA, MMA and the
// epilogue of consecutive tiles overlap.
// WKN: the weight is stored [K, N] (N contiguous) and used as an MN-major B operand: per stage 64 contraction rows of
// 128-byte SW128 lines, one 8 KB box per 64 output columns (LBO apart), 8-row groups 1024 bytes apart (SBO).
template <int BN, bool GELU, bool WKN>
__global__ void __launch_bounds__(kTcThreads, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmB,
                                                                const TcEpilogue e, const int num_k_blocks,
                                                                const int n_tiles, const int num_tiles, const GemmDfDev d) {
  constexpr int A_BYTES = 128 * 128, B_BYTES = BN * 128;
  constexpr int kTcStages = TcStages<BN>::value;
  constexpr uint32_t TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + kTcStages * A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + kTcStages * B_BYTES);
  uint64_t* empty = full + kTcStages;
  uint64_t* tmem_full = empty + kTcStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmB);
    for (int s = 0; s < kTcStages; ++s) { tc::mbar_init(full + s, 1); tc::mbar_init(empty + s, 1); }
    for (int s = 0; s < 2; ++s) { tc::mbar_init(tmem_full + s, 1); tc::mbar_init(tmem_empty + s, 4); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, TMEM_COLS);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) df_mark_begin(d.timing);

  if (warp == 0) {
    if (tc::elect_one()) {
      int it = 0;
      DfSeen seen = {-1, -1, 0};
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * 128, n0 = (tile % n_tiles) * BN;
        if (d.dep.flags) {                       // rows m0 .. m0+127 = pixels of one or two images
          const int m1 = (m0 + 127 < e.M ? m0 + 127 : e.M - 1);
          const int b0 = m0 / d.HW, b1 = m1 / d.HW;
          const int ya = (m0 - b0 * d.HW) / d.W, yb = (m1 - b1 * d.HW) / d.W + 1;
          if (b0 == b1) {
            df_wait(d.dep, d.error, b0, ya, yb, d.H, seen);
          } else {
            df_wait(d.dep, d.error, b0, ya, d.H, d.H, seen);
            df_wait(d.dep, d.error, b1, 0, yb, d.H, seen);
          }
        }
        for (int kb = 0; kb < num_k_blocks; ++kb, ++it) {
          const int s = it % kTcStages;
          const uint32_t ph = (it / kTcStages) & 1;
          tc::mbar_wait(empty + s, ph ^ 1);
          tc::mbar_expect_tx(full + s, A_BYTES + B_BYTES);
          tc::tma_load_2d(sA + s * A_BYTES, &tmA, full + s, kb * 64, m0);
          if (WKN) {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tc::tma_load_2d(sB + s * B_BYTES + j * 8192, &tmB, full + s, n0 + j * 64, kb * 64);
          } else {
            tc::tma_load_2d(sB + s * B_BYTES, &tmB, full + s, kb * 64, n0);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    constexpr uint32_t idesc = tc::make_idesc_bf16(128, BN) | (WKN ? (1u << 16) : 0u);
    constexpr uint64_t HI = (uint64_t)tc::desc_hi_sw128(1024) << 32;
    constexpr uint64_t HI_B = HI | (WKN ? ((uint64_t)(8192 >> 4) << 16) : 0);         // LBO between 64-column groups
    constexpr uint64_t B_KSTEP = WKN ? (16 * 128) >> 4 : 2;                           // descriptor advance per K = 16
    const bool leader = tc::elect_one();
    int it = 0, lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      const int buf = lt & 1;
      tc::mbar_wait(tmem_empty + buf, ((lt >> 1) & 1) ^ 1);
      tc::tc_fence_after();
      const uint32_t acc = tmem_base + (uint32_t)(buf * BN);
      for (int kb = 0; kb < num_k_blocks; ++kb, ++it) {
        const int s = it % kTcStages;
        tc::mbar_wait(full + s, (it / kTcStages) & 1);
        tc::tc_fence_after();
        if (leader) {
          uint64_t a_d = HI | (uint64_t)(smem_u32(sA + s * A_BYTES) >> 4), b_d = HI_B | (uint64_t)(smem_u32(sB + s * B_BYTES) >> 4);
          asm volatile("" : "+l"(a_d), "+l"(b_d));     // opaque bases: k offsets stay immediates of one UIADD3.64 each
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc::umma_bf16(acc, a_d + (uint64_t)(k * 2), b_d + (uint64_t)k * B_KSTEP, idesc, (kb | k) != 0 ? 1u : 0u);
          tc::umma_commit(empty + s);          // frees the smem stage when these MMAs have read it
        }
        __syncwarp();
      }
      if (leader) tc::umma_commit(tmem_full + buf);   // accumulator of this tile complete
      __syncwarp();
    }
  } else {
    const int quad = warp & 3;                 // TMEM lanes [32*quad, 32*quad+32) are visible to this warp
    const float slope = e.act == SEGMIF_ACT_PRELU ? *e.alpha : (e.act == SEGMIF_ACT_RELU ? 0.f : 1.f);
    int lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      const int buf = lt & 1;
      const int m0 = (tile / n_tiles) * 128, n0 = (tile % n_tiles) * BN;
      const int64_t m = (int64_t)m0 + quad * 32 + lane;
      const bool has_res = e.res != nullptr && m < e.M;
      const float rscale = (e.row_scale != nullptr && m < e.M) ? e.row_scale[m / e.rows_per_scale] : 1.f;
      ResRow32 cur, nxt;
      if (has_res && n0 < e.N) load_res_row32(e, cur, m, n0);          // in flight while the MMAs of this tile finish
      tc::mbar_wait(tmem_full + buf, (lt >> 1) & 1);
      tc::tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        if (has_res && c + 32 < BN && (n0 + c + 32) < e.N) load_res_row32(e, nxt, m, n0 + c + 32);
        float v[32];
        tc::tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * BN + c), v);   // warp-collective
        if (m < e.M && (n0 + c) < e.N) epilogue_row32<GELU>(e, v, cur, m, n0 + c, slope, rscale);
        cur = nxt;
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(tmem_empty + buf);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, TMEM_COLS);
  if (threadIdx.x == 0) df_mark_end(d.timing);
}

template <int BN, bool GELU, bool WKN = false>
static int launch_gemm_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcEpilogue& e, int K, cudaStream_t st,
                          const GemmDfExtra* x = nullptr) {
  constexpr int kTcStages = TcStages<BN>::value;
  constexpr size_t smem = (size_t)kTcStages * (128 * 128 + BN * 128) + (2 * kTcStages + 4) * 8 + 16;
  auto kern = gemm_tc_kernel<BN, GELU, WKN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) { set_error("linear_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(err)); return SEGMIF_ERR_CUDA; }
    configured = true;
  }
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  const int n_tiles = (int)ceil_div(e.N, BN), m_tiles = (int)ceil_div(e.M, 128);
  const int64_t num_tiles = (int64_t)n_tiles * m_tiles;
  const int ctas_per_sm = BN <= 64 ? (kTcStages <= 3 ? 3 : 2) : 1;          // 72 KB (BN=64) / 60 KB (BN=32) of smem with 3 stages: three CTAs fit
  int grid = (int)std::min<int64_t>(num_tiles, (int64_t)sms * ctas_per_sm);
  GemmDfDev d;
  d.dep.flags = nullptr; d.dep.target = 0; d.dep.tiles_y = d.dep.shift = d.dep.halo = 0; d.error = nullptr; d.timing = nullptr; d.H = d.W = d.HW = 1;
  if (x) {
    d.dep = x->dep; d.error = x->error; d.timing = x->timing; d.H = x->H; d.W = x->W; d.HW = x->H * x->W;
    if (x->max_ctas > 0) grid = std::min(grid, x->max_ctas);
  }
  kern<<<grid, kTcThreads, smem, st>>>(tmA, tmB, e, (int)ceil_div(K, 64), n_tiles, (int)num_tiles, d);
  return check_launch("segmif_linear_tc_fwd");
}

static int linear_tc_impl(const segmif_linear_params* p, cudaStream_t st, const GemmDfExtra* x = nullptr) {
  SEGMIF_REQUIRE(p && p->src && p->weight && p->dst, "linear_tc: null pointer");
  SEGMIF_REQUIRE(p->M > 0 && p->N > 0 && p->K > 0, "linear_tc: bad sizes");
  SEGMIF_REQUIRE(p->K % 8 == 0 && p->ld_src % 8 == 0 && p->src_coff % 8 == 0, "linear_tc: K, ld_src, src_coff must be multiples of 8");
  SEGMIF_REQUIRE(p->N % 32 == 0, "linear_tc: N=%d must be a multiple of 32", p->N);
  SEGMIF_REQUIRE(p->bias == nullptr || ((uintptr_t)p->bias & 15) == 0, "linear_tc: bias must be 16-byte aligned");
  SEGMIF_REQUIRE(p->src_coff + p->K <= p->ld_src && p->dst_coff + p->N <= p->ld_dst, "linear_tc: channel slice exceeds pitch");
  const int dalign = p->dst_dtype == SEGMIF_F32 ? 4 : 8;
  SEGMIF_REQUIRE(p->ld_dst % dalign == 0 && p->dst_coff % dalign == 0, "linear_tc: dst pitch/offset must be 16-byte multiples");
  if (p->residual) {
    const int ralign = p->res_dtype == SEGMIF_F32 ? 4 : 8;
    SEGMIF_REQUIRE(p->ld_res % ralign == 0 && p->res_coff % ralign == 0, "linear_tc: residual pitch/offset must be 16-byte multiples");
  }
  SEGMIF_REQUIRE(p->act != SEGMIF_ACT_PRELU || p->prelu_alpha, "linear_tc: PReLU needs prelu_alpha");
  SEGMIF_REQUIRE(((uintptr_t)p->src & 15) == 0 && ((uintptr_t)p->weight & 15) == 0 && ((uintptr_t)p->dst & 15) == 0,
                 "linear_tc: pointers must be 16-byte aligned");
  const int BN = (p->N % 128 == 0 || p->N > 192) ? 128 : (p->N % 64 == 0 ? 64 : 32);
  SEGMIF_REQUIRE(!p->weight_kn || (p->N % 64 == 0 && p->act != SEGMIF_ACT_GELU), "linear_tc: weight_kn needs N %% 64 == 0 and no GELU");
  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[2] = {(uint64_t)p->K, (uint64_t)p->M};
    const uint64_t strides[1] = {(uint64_t)p->ld_src * 2};
    const uint32_t box[2] = {64, 128};
    int rc = make_tmap_bf16(&tmA, reinterpret_cast<const bf16*>(p->src) + p->src_coff, 2, dims, strides, box, true, "linear_tc(A)");
    if (rc) return rc;
  }
  if (p->weight_kn) {
    const uint64_t dims[2] = {(uint64_t)p->N, (uint64_t)p->K};
    const uint64_t strides[1] = {(uint64_t)p->N * 2};
    const uint32_t box[2] = {64, 64};
    int rc = make_tmap_bf16(&tmB, p->weight, 2, dims, strides, box, true, "linear_tc(W, [K,N])");
    if (rc) return rc;
  } else {
    const uint64_t dims[2] = {(uint64_t)p->K, (uint64_t)p->N};
    const uint64_t strides[1] = {(uint64_t)p->K * 2};
    const uint32_t box[2] = {64, (uint32_t)BN};
    int rc = make_tmap_bf16(&tmB, p->weight, 2, dims, strides, box, true, "linear_tc(W)");
    if (rc) return rc;
  }
  TcEpilogue e;
  e.bias = p->bias; e.alpha = p->prelu_alpha; e.res = p->residual; e.dst = p->dst;
  e.M = p->M; e.N = p->N; e.act = p->act; e.res_dtype = p->res_dtype; e.ld_res = p->ld_res; e.res_coff = p->res_coff;
  e.dst_dtype = p->dst_dtype; e.ld_dst = p->ld_dst; e.dst_coff = p->dst_coff;
  e.row_scale = p->row_scale; e.rows_per_scale = p->rows_per_scale > 0 ? p->rows_per_scale : 1;
  SEGMIF_REQUIRE(!p->row_scale || p->rows_per_scale > 0, "linear_tc: row_scale needs rows_per_scale > 0");
  if (p->act == SEGMIF_ACT_GELU) {
    if (BN == 128) return launch_gemm_tc<128, true>(tmA, tmB, e, p->K, st, x);
    if (BN == 64) return launch_gemm_tc<64, true>(tmA, tmB, e, p->K, st, x);
    return launch_gemm_tc<32, true>(tmA, tmB, e, p->K, st, x);
  }
  if (p->weight_kn) {
    if (BN == 128) return launch_gemm_tc<128, false, true>(tmA, tmB, e, p->K, st, x);
    return launch_gemm_tc<64, false, true>(tmA, tmB, e, p->K, st, x);
  }
  if (BN == 128) return launch_gemm_tc<128, false>(tmA, tmB, e, p->K, st, x);
  if (BN == 64) return launch_gemm_tc<64, false>(tmA, tmB, e, p->K, st, x);
  return launch_gemm_tc<32, false>(tmA, tmB, e, p->K, st, x);
}

int linear_tc_df(const segmif_linear_params* p, const GemmDfExtra& x, cudaStream_t st) { return linear_tc_impl(p, st, &x); }

}  // namespace segmif

using namespace segmif;

extern "C" int segmif_linear_tc_fwd(const segmif_linear_params* p, segmif_stream_t stream) {
  return linear_tc_impl(p, as_stream(stream));
}

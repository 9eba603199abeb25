// Spatial-reduction attention core on tcgen05 (core/mix_transformer.py:103-110): out = softmax(q k^T * scale) v per
// (batch, head) for the common case of the SegMiF encoder -- head dim 64 and at most 320 keys after the spatial reduction
// (cfg 2: Nk = 300 in every stage), where a whole score row fits in tensor memory and no online rescaling is needed:
//   S[128 x Nk] = Q K^T        tcgen05.mma, Q and K K-major SW128 tiles from TMA, accumulator in TMEM (<= 320 columns)
//   softmax                    one thread per query row: two passes over its TMEM row (max, then exp2 / sum), P written
//                              to shared memory as bf16 in the K-major SW128 layout the next MMA reads
//   O[128 x 64] = P V          tcgen05.mma, V read as an MN-major operand (the [key][d] tile as TMA delivers it)
// K and V of one (batch, head) stay resident in shared memory while the CTA walks that head's query tiles; Q tiles are
// double buffered.  The mma.sync flash kernel (attention.cu) remains for head dim 32 and Nk > 320.
// Warp roles: 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2..5 = softmax + epilogue (thread = query row).
#include <algorithm>
#include <cstdlib>

#include "tc_common.cuh"

namespace segmif {

namespace {

constexpr int kThreads = 192;
constexpr int MAXKB = 5;                                  // 64-key blocks -> Nk <= 320
constexpr int KV_BLOCK = 64 * 128;                        // [64 keys][64 d] bf16
constexpr int P_BLOCK = 128 * 128;                        // [128 queries][64 keys] bf16
constexpr int Q_TILE = 128 * 128;                         // [128 queries][64 d] bf16

struct AttnArgs {
  bf16* out;
  float* lse;
  int ldo, heads, N, Nk, nkb, q_tiles;
  float scale_log2e;
};

__device__ __forceinline__ uint64_t kmajor_desc(uint32_t saddr) {          // K-major SW128, 8-row groups 1024 B apart
  return ((uint64_t)tc::desc_hi_sw128(1024) << 32) | (uint64_t)((saddr >> 4) & 0x3FFF);
}
__device__ __forceinline__ uint64_t mnmajor_desc(uint32_t saddr) {         // MN-major SW128, one 64-element atom wide
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kThreads, 1) sr_attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                      const __grid_constant__ CUtensorMap tmK,
                                                                      const __grid_constant__ CUtensorMap tmV, const AttnArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sK = smem;
  uint8_t* sV = sK + MAXKB * KV_BLOCK;
  uint8_t* sP = sV + MAXKB * KV_BLOCK;
  uint8_t* sQ = sP + MAXKB * P_BLOCK;
  __shared__ uint64_t kv_full, q_full[2], q_empty[2], s_full, p_full, o_full;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y, b = bh / a.heads, h = bh % a.heads;
  const int nkb = a.nkb, ncols = nkb * 64;

  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&tmQ);
    tc::prefetch_tmap(&tmK);
    tc::prefetch_tmap(&tmV);
    tc::mbar_init(&kv_full, 1);
    for (int s = 0; s < 2; ++s) { tc::mbar_init(q_full + s, 1); tc::mbar_init(q_empty + s, 1); }
    tc::mbar_init(&s_full, 1);
    tc::mbar_init(&p_full, 4);                            // one arrive per softmax warp
    tc::mbar_init(&o_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(&tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + 320;

  if (warp == 0) {
    if (tc::elect_one()) {
      tc::mbar_expect_tx(&kv_full, (uint32_t)(2 * nkb * KV_BLOCK));
      for (int j = 0; j < nkb; ++j) {
        tc::tma_load_3d(sK + j * KV_BLOCK, &tmK, &kv_full, h * 64, j * 64, b);
        tc::tma_load_3d(sV + j * KV_BLOCK, &tmV, &kv_full, h * 64, j * 64, b);
      }
      int it = 0;
      for (int t = blockIdx.x; t < a.q_tiles; t += gridDim.x, ++it) {
        const int s = it & 1;
        tc::mbar_wait(q_empty + s, ((it >> 1) & 1) ^ 1);
        tc::mbar_expect_tx(q_full + s, Q_TILE);
        tc::tma_load_3d(sQ + s * Q_TILE, &tmQ, q_full + s, h * 64, t * 128, b);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // S = Q K^T in one (N = ncols <= 256) or two (N = 160 + 160) instructions per k16 step
    const int n_first = ncols <= 256 ? ncols : 160;
    const uint32_t idesc_s1 = tc::make_idesc_bf16(128, n_first);
    const uint32_t idesc_s2 = tc::make_idesc_bf16(128, ncols - n_first > 0 ? ncols - n_first : 16);
    constexpr uint32_t idesc_o = tc::make_idesc_bf16(128, 64) | (1u << 16);        // B = V, MN-major
    tc::mbar_wait(&kv_full, 0);
    int it = 0;
    for (int t = blockIdx.x; t < a.q_tiles; t += gridDim.x, ++it) {
      const int s = it & 1;
      tc::mbar_wait(q_full + s, (it >> 1) & 1);
      tc::tc_fence_after();
      if (tc::elect_one()) {
        const uint64_t qd = kmajor_desc(smem_u32(sQ + s * Q_TILE));
        const uint64_t kd = kmajor_desc(smem_u32(sK));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          tc::umma_bf16(tmem_s, qd + (uint64_t)(k * 2), kd + (uint64_t)(k * 2), idesc_s1, k != 0 ? 1u : 0u);
          if (ncols > n_first)
            tc::umma_bf16(tmem_s + n_first, qd + (uint64_t)(k * 2), kd + (uint64_t)((n_first * 128) >> 4) + (uint64_t)(k * 2), idesc_s2,
                          k != 0 ? 1u : 0u);
        }
        tc::umma_commit(q_empty + s);
        tc::umma_commit(&s_full);
      }
      __syncwarp();
      tc::mbar_wait(&p_full, it & 1);                     // P of this tile is in shared memory, S is free again
      tc::tc_fence_after();
      if (tc::elect_one()) {
        for (int kb = 0; kb < nkb; ++kb) {
          const uint64_t pd = kmajor_desc(smem_u32(sP + kb * P_BLOCK));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t vd = mnmajor_desc(smem_u32(sV + kb * KV_BLOCK + k * 16 * 128));
            tc::umma_bf16(tmem_o, pd + (uint64_t)(k * 2), vd, idesc_o, (kb | k) != 0 ? 1u : 0u);
          }
        }
        tc::umma_commit(&o_full);
      }
      __syncwarp();
    }
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + lane;                       // query row within the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    int it = 0;
    for (int t = blockIdx.x; t < a.q_tiles; t += gridDim.x, ++it) {
      const int row = t * 128 + r;
      tc::mbar_wait(&s_full, it & 1);
      tc::tc_fence_after();
      float m = -INFINITY;
      for (int c = 0; c < ncols; c += 32) {
        float v[32];
        tc::tmem_ld32(tmem_s + lane_addr + (uint32_t)c, v);
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c + j < a.Nk) m = fmaxf(m, v[j]);
      }
      const float ms = m * a.scale_log2e;
      float l = 0.f;
      for (int c = 0; c < ncols; c += 32) {
        float v[32];
        tc::tmem_ld32(tmem_s + lane_addr + (uint32_t)c, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float p = c + j < a.Nk ? ex2f(fmaf(v[j], a.scale_log2e, -ms)) : 0.f;
          v[j] = p;
          l += p;
        }
        uint8_t* prow = sP + (c >> 6) * P_BLOCK + r * 128;
        const int cb = (c & 32) >> 3;                     // first 16-byte chunk of these 32 keys inside the 128-byte row
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 pk;
          pk.x = pack_bf16x2(v[i * 8 + 0], v[i * 8 + 1]);
          pk.y = pack_bf16x2(v[i * 8 + 2], v[i * 8 + 3]);
          pk.z = pack_bf16x2(v[i * 8 + 4], v[i * 8 + 5]);
          pk.w = pack_bf16x2(v[i * 8 + 6], v[i * 8 + 7]);
          *reinterpret_cast<uint4*>(prow + (((cb + i) ^ (r & 7)) << 4)) = pk;
        }
      }
      tc::tc_fence_before();
      tc::fence_proxy_async();                            // generic-proxy writes of P -> visible to the MMA (async proxy)
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&p_full);
      // epilogue of this tile
      tc::mbar_wait(&o_full, it & 1);
      tc::tc_fence_after();
      const float inv = 1.f / l;
      if (row < a.N && a.lse != nullptr) a.lse[(int64_t)bh * a.N + row] = ms + log2f(l);     // exp2 domain, as attention.cu
      bf16* op = a.out + ((int64_t)b * a.N + row) * a.ldo + h * 64;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float v[32];
        tc::tmem_ld32(tmem_o + lane_addr + (uint32_t)(half * 32), v);
        if (row < a.N) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 pk;
            pk.x = pack_bf16x2(v[i * 8 + 0] * inv, v[i * 8 + 1] * inv);
            pk.y = pack_bf16x2(v[i * 8 + 2] * inv, v[i * 8 + 3] * inv);
            pk.z = pack_bf16x2(v[i * 8 + 4] * inv, v[i * 8 + 5] * inv);
            pk.w = pack_bf16x2(v[i * 8 + 6] * inv, v[i * 8 + 7] * inv);
            *reinterpret_cast<uint4*>(op + half * 32 + i * 8) = pk;
          }
        }
      }
      tc::tc_fence_before();
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, 512);
}

int encode3(CUtensorMap* out, const void* base, int C, int rows, int ld, int B, uint32_t box_rows, const char* what) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return SEGMIF_ERR_CUDA;
  cuuint64_t gd[3] = {(cuuint64_t)C, (cuuint64_t)rows, (cuuint64_t)B};
  cuuint64_t gs[2] = {(cuuint64_t)ld * 2, (cuuint64_t)rows * ld * 2};
  cuuint32_t bx[3] = {64, box_rows, 1}, es[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("%s: cuTensorMapEncodeTiled failed (CUresult %d)", what, (int)r); return SEGMIF_ERR_CUDA; }
  return SEGMIF_OK;
}

}  // namespace

bool sr_attention_tc_supported(int B, int heads, int N, int Nk, int D, int ldq, int ldkv, int ldo, const void* q, const void* k,
                               const void* v, const void* out) {
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return D == 64 && Nk >= 1 && Nk <= MAXKB * 64 && N >= 1 && ldq % 8 == 0 && ldkv % 8 == 0 && ldo % 8 == 0 && al16(q) && al16(k) &&
         al16(v) && al16(out) && B * heads <= 65535 && B * heads >= 1;
}

bool sr_attention_tc_ok(int B, int heads, int N, int Nk, int D, int ldq, int ldkv, int ldo, const void* q, const void* k, const void* v,
                        const void* out) {
  static int enabled = -1;
  // opt-in (SEGMIF_ATTN_TC=1) until it beats the mma.sync kernel in the whole step: its softmax runs on four warps per SM
  // while the tensor pipe idles, and at cfg 2 the first measurement was 15.6 vs 15.2 ms/step
  if (enabled < 0) { const char* e = getenv("SEGMIF_ATTN_TC"); enabled = (e && e[0] == '1') ? 1 : 0; }
  if (!enabled) return false;
  return sr_attention_tc_supported(B, heads, N, Nk, D, ldq, ldkv, ldo, q, k, v, out);
}

int sr_attention_tc(const void* q, int ldq, const void* k, const void* v, int ldkv, void* out, int ldo, int B, int heads, int N, int Nk,
                    float scale, float* lse, cudaStream_t st) {
  CUtensorMap tmQ, tmK, tmV;
  const int C = heads * 64;
  if (int rc = encode3(&tmQ, q, C, N, ldq, B, 128, "sr_attention_tc(Q)")) return rc;
  if (int rc = encode3(&tmK, k, C, Nk, ldkv, B, 64, "sr_attention_tc(K)")) return rc;
  if (int rc = encode3(&tmV, v, C, Nk, ldkv, B, 64, "sr_attention_tc(V)")) return rc;
  AttnArgs a;
  a.out = (bf16*)out; a.lse = lse; a.ldo = ldo; a.heads = heads; a.N = N; a.Nk = Nk;
  a.nkb = (Nk + 63) / 64; a.q_tiles = (N + 127) / 128; a.scale_log2e = scale * 1.4426950408889634f;
  const size_t smem = (size_t)MAXKB * (2 * KV_BLOCK + P_BLOCK) + 2 * Q_TILE + 1024;
  static bool cfg = false;
  if (!cfg) {
    cudaError_t e = cudaFuncSetAttribute(sr_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("sr_attention_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return SEGMIF_ERR_CUDA; }
    cfg = true;
  }
  const int bhn = B * heads;
  // one CTA per SM (193 KB of shared memory, all of TMEM): never more CTAs than SMs, or the tail CTAs run as a second wave
  const int gx = std::max(1, std::min(a.q_tiles, 148 / bhn));
  sr_attention_tc_kernel<<<dim3(gx, bhn), kThreads, smem, st>>>(tmQ, tmK, tmV, a);
  return check_launch("segmif_sr_attention_fwd (tcgen05)");
}

}  // namespace segmif

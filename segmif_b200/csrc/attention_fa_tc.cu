// Spatial-reduction attention core on tcgen05 for ANY number of keys (core/mix_transformer.py:103-110, head dim 64):
// flash-style, 128-key blocks, online softmax with lazy rescaling -- cfg 2 (Nk = 300: three blocks) and the long-sequence
// stress of BASELINE configs[3] (MiT-B4 at 1024^2: Nk = 1024, scores [4, 1, 65 536, 1024] never materialised).
//
// Per CTA: one (batch, head) and TWO query tiles of 128 rows in flight (A, B); the key / value blocks stream through a 3-stage TMA
// ring and every block serves both tiles (K / V are re-read from L2 once per tile PAIR: 2 x 128 KB per head at Nk = 1024).
//   warp 0     TMA producer: the two Q tiles, K / V blocks ([128 keys][64 d] bf16, 128-byte swizzle)
//   warp 1     MMA issuer:   S_x(j) = Q_x K_j^T (M 128, N 128, K 64: 4 MMAs) into tile x's score buffer in TMEM;
//                            O_x += P_x(j) V_j (M 128, N 64, K 128: 8 MMAs, V as an MN-major operand); S_x(j+1) is issued right after
//                            PV_x(j), so it runs under the OTHER tile's softmax
//   warps 2-5  softmax + epilogue of tile A, warps 6-9 of tile B; thread = query row (the TMEM lane layout): block maximum, exp2,
//              P -> bf16 in the K-major SW128 layout the PV MMA reads, running sum.  While one tile's warps wait for their next
//              scores or for a tcgen05.ld, the other tile's warps keep the MUFU and the issue slots busy (one tile per CTA measured
//              239 us at MiT-B4 stage 1; two tiles 153 us; the mma.sync flash kernel 280 us).
// Online softmax: row statistics live in registers.  The reference maximum m_ref of a row is only moved when a block's maximum
// exceeds it by more than 8 (log2 domain): P then stays <= 2^8, well inside bf16 / fp32 range, and the accumulator row in TMEM is
// rescaled (tcgen05.ld -> multiply -> tcgen05.st, warp-collective, ordered after the previous block's PV by its completion
// barrier) only in those rare steps -- for softmax-normalised inputs typically in the first one or two blocks of a tile.
// The kernel is bound by the exponentials (128 x 128 per tile and block on the 16/clk MUFU of an SM ~ 1024 cycles, the MMAs of the
// same block ~ 512): the tensor pipe runs entirely under the softmax.
#include <algorithm>
#include <cstdlib>

#include "tc_common.cuh"

namespace segmif {

namespace {

constexpr int kFaThreads = 320;                          // TMA warp, MMA warp, 4 softmax warps per query tile x 2 tiles
constexpr int kKvStages = 3;
constexpr int K_BLOCK = 128 * 128;                        // [128 keys][64 d] bf16
constexpr int KV_STAGE = 2 * K_BLOCK;                     // K then V
constexpr int P_TILE = 2 * 128 * 128;                     // [2 x 64 keys][128 queries][128 B]
constexpr int Q_TILE = 128 * 128;

struct FaArgs {
  bf16* out;
  float* lse;
  int ldo, heads, N, Nk, nkb, q_pairs;
  int trim;            // 1: a partly filled last key block is processed up to its last valid key only (SEGMIF_FA_TRIM=0 disables)
  float scale_log2e;
};

__device__ __forceinline__ uint64_t fa_kmajor_desc(uint32_t saddr) {
  return ((uint64_t)tc::desc_hi_sw128(1024) << 32) | (uint64_t)((saddr >> 4) & 0x3FFF);
}
__device__ __forceinline__ uint64_t fa_mnmajor_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ float fa_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 32 lanes x 32 columns of fp32 back into tensor memory (the rescaled accumulator row)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};\n" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
      "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
      "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
      "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
      "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
}

// TWO query tiles (A, B) of the same (batch, head) are in flight per CTA: they share every K / V block of the ring, each has its
// own score buffer, accumulator (TMEM) and P tile (shared memory) and its own four softmax warps, so that while one tile's
// softmax waits for its next scores (or for a TMEM load) the other tile's softmax warps keep the MUFU / issue slots busy.
__global__ void __launch_bounds__(kFaThreads, 1) sr_attention_fa_tc_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                         const __grid_constant__ CUtensorMap tmK,
                                                                         const __grid_constant__ CUtensorMap tmV, const FaArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sKV = smem;                                     // [stages][K block | V block]
  uint8_t* sP = sKV + kKvStages * KV_STAGE;                // [2 tiles][P_TILE]
  uint8_t* sQ = sP + 2 * P_TILE;                           // [2 tiles][Q_TILE]
  __shared__ uint64_t q_full, q_empty, kv_full[kKvStages], kv_empty[kKvStages], s_full[2], p_full[2], p_empty[2], o_full[2], o_empty[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y, b = bh / a.heads, h = bh % a.heads;
  const int nkb = a.nkb;

  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&tmQ);
    tc::prefetch_tmap(&tmK);
    tc::prefetch_tmap(&tmV);
    tc::mbar_init(&q_full, 1);
    tc::mbar_init(&q_empty, 1);
    for (int s = 0; s < kKvStages; ++s) { tc::mbar_init(kv_full + s, 1); tc::mbar_init(kv_empty + s, 1); }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(s_full + s, 1);
      tc::mbar_init(p_full + s, 4);                       // one arrive per softmax warp of the tile
      tc::mbar_init(p_empty + s, 1);
      tc::mbar_init(o_full + s, 1);
      tc::mbar_init(o_empty + s, 4);
    }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(&tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;                    // S_A [0,128)  S_B [128,256)  O_A [256,320)  O_B [320,384)

  if (warp == 0) {
    if (tc::elect_one()) {
      int pl = 0, kit = 0;                                 // pairs done by this CTA, KV blocks issued
      for (int pr = blockIdx.x; pr < a.q_pairs; pr += gridDim.x, ++pl) {
        tc::mbar_wait(&q_empty, (pl & 1) ^ 1);
        tc::mbar_expect_tx(&q_full, 2 * Q_TILE);
        tc::tma_load_3d(sQ, &tmQ, &q_full, h * 64, pr * 256, b);
        tc::tma_load_3d(sQ + Q_TILE, &tmQ, &q_full, h * 64, pr * 256 + 128, b);
        for (int j = 0; j < nkb; ++j, ++kit) {
          const int s = kit % kKvStages;
          tc::mbar_wait(kv_empty + s, ((kit / kKvStages) & 1) ^ 1);
          tc::mbar_expect_tx(kv_full + s, KV_STAGE);
          tc::tma_load_3d(sKV + s * KV_STAGE, &tmK, kv_full + s, h * 64, j * 128, b);
          tc::tma_load_3d(sKV + s * KV_STAGE + K_BLOCK, &tmV, kv_full + s, h * 64, j * 128, b);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = tc::make_idesc_bf16(128, 128);
    constexpr uint32_t idesc_o = tc::make_idesc_bf16(128, 64) | (1u << 16);        // B = V, MN-major
    const bool leader = tc::elect_one();
    int pl = 0, kit = 0;
    for (int pr = blockIdx.x; pr < a.q_pairs; pr += gridDim.x, ++pl) {
      tc::mbar_wait(&q_full, pl & 1);
      tc::tc_fence_after();
      auto issue_s = [&](int x, int kblock_global) {       // S_x = Q_x K^T for the block in ring stage kblock_global % stages
        const int st = kblock_global % kKvStages;
        tc::mbar_wait(kv_full + st, (kblock_global / kKvStages) & 1);
        tc::tc_fence_after();
        if (leader) {
          const uint64_t qd = fa_kmajor_desc(smem_u32(sQ + x * Q_TILE));
          const uint64_t kd = fa_kmajor_desc(smem_u32(sKV + st * KV_STAGE));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc::umma_bf16(tmem_base + (uint32_t)(x * 128), qd + (uint64_t)(k * 2), kd + (uint64_t)(k * 2), idesc_s, k != 0 ? 1u : 0u);
          tc::umma_commit(s_full + x);
        }
        __syncwarp();
      };
      issue_s(0, kit);
      issue_s(1, kit);
      for (int j = 0; j < nkb; ++j) {
        const int gidx = pl * nkb + j;                     // per-tile block counter (phase of s_full / p_full / p_empty)
        for (int x = 0; x < 2; ++x) {
          tc::mbar_wait(p_full + x, gidx & 1);             // P_x(j) is in shared memory, S_x(j) has been consumed, O_x is rescaled
          if (j == 0) tc::mbar_wait(o_empty + x, (pl & 1) ^ 1);   // the previous pair's epilogue has drained accumulator x
          tc::tc_fence_after();
          if (leader) {
            const int st = (kit + j) % kKvStages;
            const uint32_t vbase = smem_u32(sKV + st * KV_STAGE + K_BLOCK);
            const uint32_t pbase = smem_u32(sP + x * P_TILE);
            const int ksteps = ((j + 1) * 128 <= a.Nk || !a.trim) ? 8 : (a.Nk - j * 128 + 15) >> 4;   // last block: valid keys only
#pragma unroll
            for (int k = 0; k < 8; ++k) {                   // 16 keys per step: P columns (k / 4) block, (k % 4) * 32 B; V rows 16 k
              if (k >= ksteps) break;
              const uint64_t pd = fa_kmajor_desc(pbase + (uint32_t)((k >> 2) * (128 * 128))) + (uint64_t)((k & 3) * 2);
              const uint64_t vd = fa_mnmajor_desc(vbase + (uint32_t)(k * 16 * 128));
              tc::umma_bf16(tmem_base + 256 + (uint32_t)(x * 64), pd, vd, idesc_o, (j | k) != 0 ? 1u : 0u);
            }
            tc::umma_commit(p_empty + x);                   // PV_x(j) complete: P_x free, O_x may be rescaled
            if (x == 1) tc::umma_commit(kv_empty + st);     // both tiles are done with block j
            if (j == nkb - 1) tc::umma_commit(o_full + x);
          }
          __syncwarp();
          if (j + 1 < nkb) issue_s(x, kit + j + 1);        // the next scores of this tile run under the OTHER tile's softmax
        }
      }
      if (leader) tc::umma_commit(&q_empty);               // every S of this pair has been issued
      __syncwarp();
      kit += nkb;
    }
  } else {
    const int x = (warp - 2) >> 2;                         // query tile of this warp (0 = A, 1 = B)
    const int quad = warp & 3;
    const int r = quad * 32 + lane;                       // query row within the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const uint32_t ts = tmem_base + lane_addr + (uint32_t)(x * 128);
    const uint32_t to = tmem_base + lane_addr + 256 + (uint32_t)(x * 64);
    uint8_t* pt = sP + x * P_TILE;
    int pl = 0;
    for (int pr = blockIdx.x; pr < a.q_pairs; pr += gridDim.x, ++pl) {
      const int row = pr * 256 + x * 128 + r;
      float m_ref = -INFINITY, l = 0.f;
      for (int j = 0; j < nkb; ++j) {
        const int gidx = pl * nkb + j;
        const int kbase = j * 128;
        const bool full = kbase + 128 <= a.Nk;             // block-uniform: only the last block masks keys
        // a partly filled last block is processed in 32-column steps up to the last valid key only (Nk = 300: 64 of 128
        // columns); the PV MMAs of that block stop at the same place, so the untouched P columns are never read
        const int cmax = (full || !a.trim) ? 128 : min(128, (a.Nk - kbase + 31) & ~31);
        tc::mbar_wait(s_full + x, gidx & 1);
        tc::tc_fence_after();
        // ---- block maximum (log2 domain)
        float bm = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < cmax; c += 32) {
          float v[32];
          tc::tmem_ld32(ts + (uint32_t)c, v);
          if (full) {
#pragma unroll
            for (int i = 0; i < 32; ++i) bm = fmaxf(bm, v[i]);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (kbase + c + i < a.Nk) bm = fmaxf(bm, v[i]);
          }
        }
        bm *= a.scale_log2e;
        // ---- lazy rescale: move the reference maximum only when this block exceeds it by more than 2^8
        const bool need = bm > m_ref + 8.f;
        float factor = 1.f;
        if (need) {
          factor = fa_ex2(m_ref - bm);                    // first block: exp2(-inf) = 0
          m_ref = bm;
          l *= factor;
        }
        if (j > 0) {
          tc::mbar_wait(p_empty + x, (gidx - 1) & 1);      // PV(j-1) has finished: P tile free, accumulator complete up to block j-1
          if (__any_sync(0xffffffffu, need)) {
            tc::tc_fence_after();
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {        // scale this warp's 32 accumulator rows (factor 1 for rows that keep m_ref)
              float o[32];
              tc::tmem_ld32(to + (uint32_t)(half * 32), o);
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] *= factor;
              tmem_st32(to + (uint32_t)(half * 32), o);
            }
          }
        }
        // ---- P_j = exp2(s - m_ref) -> bf16 -> shared memory (K-major SW128, two 64-key blocks), running sum
#pragma unroll 1
        for (int c = 0; c < cmax; c += 32) {
          float v[32];
          tc::tmem_ld32(ts + (uint32_t)c, v);
          if (full) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              v[i] = fa_ex2(fmaf(v[i], a.scale_log2e, -m_ref));
              l += v[i];
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              v[i] = kbase + c + i < a.Nk ? fa_ex2(fmaf(v[i], a.scale_log2e, -m_ref)) : 0.f;
              l += v[i];
            }
          }
          uint8_t* prow = pt + (c >> 6) * (128 * 128) + r * 128;
          const int cb = (c & 32) >> 3;
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
        tc::fence_proxy_async();                          // generic-proxy writes of P -> visible to the MMA (async proxy)
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(p_full + x);
      }
      // ---- epilogue of this tile
      tc::mbar_wait(o_full + x, pl & 1);
      tc::tc_fence_after();
      const float inv = 1.f / l;
      if (row < a.N && a.lse != nullptr) a.lse[(int64_t)bh * a.N + row] = m_ref + log2f(l);      // exp2 domain, as attention.cu
      bf16* op = a.out + ((int64_t)b * a.N + row) * a.ldo + h * 64;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float v[32];
        tc::tmem_ld32(to + (uint32_t)(half * 32), v);
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
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(o_empty + x);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, 512);
}

int fa_encode3(CUtensorMap* out, const void* base, int C, int rows, int ld, int B, uint32_t box_rows, const char* what) {
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

bool sr_attention_fa_tc_supported(int B, int heads, int N, int Nk, int D, int ldq, int ldkv, int ldo, const void* q, const void* k,
                                  const void* v, const void* out) {
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return D == 64 && Nk >= 1 && N >= 1 && ldq % 8 == 0 && ldkv % 8 == 0 && ldo % 8 == 0 && al16(q) && al16(k) && al16(v) && al16(out) &&
         B * heads <= 65535 && B * heads >= 1;
}

int sr_attention_fa_tc(const void* q, int ldq, const void* k, const void* v, int ldkv, void* out, int ldo, int B, int heads, int N,
                       int Nk, float scale, float* lse, cudaStream_t st) {
  CUtensorMap tmQ, tmK, tmV;
  const int C = heads * 64;
  if (int rc = fa_encode3(&tmQ, q, C, N, ldq, B, 128, "sr_attention_fa_tc(Q)")) return rc;
  if (int rc = fa_encode3(&tmK, k, C, Nk, ldkv, B, 128, "sr_attention_fa_tc(K)")) return rc;
  if (int rc = fa_encode3(&tmV, v, C, Nk, ldkv, B, 128, "sr_attention_fa_tc(V)")) return rc;
  FaArgs a;
  a.out = (bf16*)out; a.lse = lse; a.ldo = ldo; a.heads = heads; a.N = N; a.Nk = Nk;
  a.nkb = (Nk + 127) / 128; a.q_pairs = (N + 255) / 256; a.scale_log2e = scale * 1.4426950408889634f;
  static const int trim = []() { const char* e = getenv("SEGMIF_FA_TRIM"); return (e && e[0] == '0') ? 0 : 1; }();
  a.trim = trim;
  const size_t smem = (size_t)kKvStages * KV_STAGE + 2 * P_TILE + 2 * Q_TILE + 1024;
  static bool cfg = false;
  if (!cfg) {
    cudaError_t e = cudaFuncSetAttribute(sr_attention_fa_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("sr_attention_fa_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return SEGMIF_ERR_CUDA; }
    cfg = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int bhn = B * heads;
  // one CTA per SM (193 KB of shared memory, all of TMEM): never more CTAs than SMs, or the tail CTAs run as a second wave
  const int gx = std::max(1, std::min(a.q_pairs, std::max(1, sms / bhn)));
  sr_attention_fa_tc_kernel<<<dim3(gx, bhn), kFaThreads, smem, st>>>(tmQ, tmK, tmV, a);
  return check_launch("segmif_sr_attention_fwd (tcgen05, flash)");
}

}  // namespace segmif

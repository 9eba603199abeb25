// Backward of the spatial-reduction attention core (core/mix_transformer.py:107-111):
//   S = scale q k^T,  P = softmax(S),  O = P v        given dO:
//   Delta_i = sum_d dO_id O_id,  dV = P^T dO,  dP = dO V^T,  dS = P o (dP - Delta),  dQ = scale dS K,  dK = scale dS^T Q
// Same shape as the forward (attention.cu): 64 queries per CTA (4 warps x 16 rows), K/V streamed in 64-key tiles through
// a double-buffered cp.async ring, the probabilities recomputed from the log-sum-exp the training forward stored, no
// [N, Nk] matrix in memory.  dQ accumulates in registers over the key tiles; the per-tile dK / dV contributions
// (contractions over the CTA's 64 queries: P and dS go through shared memory so the transposed operand comes from
// ldmatrix.trans) are added to fp32 accumulators [B, Nk, 2C] with atomics -- Nk is small (spatial reduction), N is large.
#include "ffm_mma.cuh"

namespace segmif {

constexpr int kAbD = 64, kAbThreads = 128;

__global__ void __launch_bounds__(kAbThreads) sr_attention_bwd_kernel(
    const bf16* __restrict__ q, int ldq, const bf16* __restrict__ k, const bf16* __restrict__ v, int ldkv,
    const bf16* __restrict__ o, const bf16* __restrict__ dout, int ldo, const float* __restrict__ lse,
    bf16* __restrict__ dq, int lddq, float* __restrict__ dkv, int lddkv, int v_off, int heads, int N, int Nk,
    float scale, float scale_log2e) {
  constexpr int D = kAbD, BQ = 64, BK = 64, KS = D / 16;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  bf16* sQ = reinterpret_cast<bf16*>(smem_raw);   // [64][64] each, 128-byte swizzled rows
  bf16* sdO = sQ + 4096;
  bf16* sP = sdO + 4096;                          // first holds O (for Delta), then P of the current tile
  bf16* sDS = sP + 4096;
  bf16* sK = sDS + 4096;                          // [2][64][64]
  bf16* sV = sK + 2 * 4096;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int bh = blockIdx.y, b = bh / heads, h = bh % heads;
  const int q0 = blockIdx.x * BQ;
  const bf16* qb = q + ((int64_t)b * N) * ldq + h * D;
  const bf16* ob = o + ((int64_t)b * N) * ldo + h * D;
  const bf16* dob = dout + ((int64_t)b * N) * ldo + h * D;
  const bf16* kb = k + ((int64_t)b * Nk) * ldkv + h * D;
  const bf16* vb = v + ((int64_t)b * Nk) * ldkv + h * D;

  for (int i = tid; i < BQ * 8; i += kAbThreads) {
    const int row = i >> 3, chunk = i & 7;
    const bool ok = (q0 + row) < N;
    const int sw = swz128(row, chunk) * 8;
    cp_async16_cg(smem_u32(sQ + row * D + sw), ok ? qb + (int64_t)(q0 + row) * ldq + chunk * 8 : qb, ok ? 16 : 0);
    cp_async16_cg(smem_u32(sdO + row * D + sw), ok ? dob + (int64_t)(q0 + row) * ldo + chunk * 8 : dob, ok ? 16 : 0);
    cp_async16_cg(smem_u32(sP + row * D + sw), ok ? ob + (int64_t)(q0 + row) * ldo + chunk * 8 : ob, ok ? 16 : 0);
  }
  auto load_kv = [&](int stage, int k0) {
    for (int i = tid; i < BK * 8; i += kAbThreads) {
      const int row = i >> 3, chunk = i & 7;
      const bool ok = (k0 + row) < Nk;
      const int64_t off = (int64_t)(k0 + row) * ldkv + chunk * 8;
      cp_async16_cg(smem_u32(sK + stage * 4096 + row * D + swz128(row, chunk) * 8), ok ? kb + off : kb, ok ? 16 : 0);
      cp_async16_cg(smem_u32(sV + stage * 4096 + row * D + swz128(row, chunk) * 8), ok ? vb + off : vb, ok ? 16 : 0);
    }
  };
  load_kv(0, 0);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  const int g = lane >> 2, tq = lane & 3;
  // Delta and log-sum-exp of this thread's two rows
  float delta[2], lrow[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = warp * 16 + g + r * 8;
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int chunk = tq * 2 + c;
      float a[8], d8[8];
      load8(sP + row * D + swz128(row, chunk) * 8, a);
      load8(sdO + row * D + swz128(row, chunk) * 8, d8);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc = fmaf(a[e], d8[e], acc);
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    delta[r] = acc;
    lrow[r] = (q0 + row) < N ? lse[(int64_t)bh * N + q0 + row] : 0.f;
  }
  uint32_t qf[KS][4], dof[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    const int row = warp * 16 + (lane & 15), chunk = ks * 2 + (lane >> 4);
    ldmatrix_x4(qf[ks], smem_u32(sQ + row * D + swz128(row, chunk) * 8));
    ldmatrix_x4(dof[ks], smem_u32(sdO + row * D + swz128(row, chunk) * 8));
  }
  float dqa[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) dqa[i][j] = 0.f;
  __syncthreads();                                // sP (holding O) may now be overwritten

  const int ntiles = (Nk + BK - 1) / BK;
  for (int t = 0; t < ntiles; ++t) {
    if (t + 1 < ntiles) load_kv((t + 1) & 1, (t + 1) * BK);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const bf16* tK = sK + (t & 1) * 4096;
    const bf16* tV = sV + (t & 1) * 4096;
    float s[8][4], dp[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { s[i][j] = 0.f; dp[i][j] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t bfr[4];
        const int row = np * 16 + (lane & 7) + ((lane >> 4) << 3), chunk = ks * 2 + ((lane >> 3) & 1);
        ldmatrix_x4(bfr, smem_u32(tK + row * D + swz128(row, chunk) * 8));
        mma_bf16_16816(s[np * 2], qf[ks], bfr[0], bfr[1]);
        mma_bf16_16816(s[np * 2 + 1], qf[ks], bfr[2], bfr[3]);
        ldmatrix_x4(bfr, smem_u32(tV + row * D + swz128(row, chunk) * 8));
        mma_bf16_16816(dp[np * 2], dof[ks], bfr[0], bfr[1]);
        mma_bf16_16816(dp[np * 2 + 1], dof[ks], bfr[2], bfr[3]);
      }
    }
    // P = exp2(S * scale*log2e - lse);  dS = P * (dP - Delta) * scale;  both to shared memory as bf16 (rows = queries)
    const int kbase = t * BK + tq * 2;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int row = warp * 16 + g + half * 8;
        const bool rlive = (q0 + row) < N;
        float p2[2], d2[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = half * 2 + e;
          const bool live = rlive && (kbase + nt * 8 + e) < Nk;
          const float p = live ? ex2_approx(s[nt][j] * scale_log2e - lrow[half]) : 0.f;
          p2[e] = p;
          d2[e] = p * (dp[nt][j] - delta[half]) * scale;
          s[nt][j] = d2[e];                       // keep dS (fp32) for dQ
        }
        *reinterpret_cast<uint32_t*>(sP + row * D + swz128(row, nt) * 8 + tq * 2) = pack_bf16x2(p2[0], p2[1]);
        *reinterpret_cast<uint32_t*>(sDS + row * D + swz128(row, nt) * 8 + tq * 2) = pack_bf16x2(d2[0], d2[1]);
      }
    }
    // dQ += dS K   (A = dS from registers, B = K[key][d] through ldmatrix.trans)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t pf[4];
      pf[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
      pf[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
      pf[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pf[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t bfr[4];
        const int row = kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3), chunk = np * 2 + (lane >> 4);
        ldmatrix_x4_trans(bfr, smem_u32(tK + row * D + swz128(row, chunk) * 8));
        mma_bf16_16816(dqa[np * 2], pf, bfr[0], bfr[1]);
        mma_bf16_16816(dqa[np * 2 + 1], pf, bfr[2], bfr[3]);
      }
    }
    __syncthreads();                              // P and dS of all 64 queries are in shared memory
    // dV[keys 16w..16w+15][d] += P^T dO,  dK[...] += dS^T Q   over the CTA's 64 queries
    {
      float av[8][4], ak[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { av[i][j] = 0.f; ak[i][j] = 0.f; }
      gram16x64_acc(av, sP, sdO, warp, lane);
      gram16x64_acc(ak, sDS, sQ, warp, lane);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int key = t * BK + warp * 16 + g + half * 8;
        if (key >= Nk) continue;
        float* dst = dkv + ((int64_t)b * Nk + key) * lddkv + h * D + tq * 2;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          atomicAdd(dst + nt * 8, ak[nt][half * 2]);
          atomicAdd(dst + nt * 8 + 1, ak[nt][half * 2 + 1]);
          atomicAdd(dst + v_off + nt * 8, av[nt][half * 2]);
          atomicAdd(dst + v_off + nt * 8 + 1, av[nt][half * 2 + 1]);
        }
      }
    }
    __syncthreads();                              // before the next tile overwrites sP / sDS and refills the K/V stage
  }
  cp_async_wait<0>();
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int row = q0 + warp * 16 + g + half * 8;
    if (row >= N) continue;
    bf16* op = dq + ((int64_t)b * N + row) * lddq + h * D + tq * 2;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) *reinterpret_cast<uint32_t*>(op + nt * 8) = pack_bf16x2(dqa[nt][half * 2], dqa[nt][half * 2 + 1]);
  }
}

}  // namespace segmif

using namespace segmif;

extern "C" int segmif_sr_attention_bwd(const void* q, int ldq, const void* k, const void* v, int ldkv, const void* out,
                                       const void* dout, int ldo, const float* lse, void* dq, int lddq, float* dkv,
                                       int lddkv, int v_off, int B, int heads, int N, int Nk, int D, float scale,
                                       segmif_stream_t stream) {
  SEGMIF_REQUIRE(q && k && v && out && dout && lse && dq && dkv, "sr_attention_bwd: null pointer");
  SEGMIF_REQUIRE(D == 64, "sr_attention_bwd: head dim %d unsupported (64: every MiT-B1..B5 stage)", D);
  SEGMIF_REQUIRE(ldq % 8 == 0 && ldkv % 8 == 0 && ldo % 8 == 0 && lddq % 2 == 0, "sr_attention_bwd: pitches must be multiples of 8");
  SEGMIF_REQUIRE(Nk > 0, "sr_attention_bwd: Nk must be positive");
  if (B * heads == 0 || N == 0) return SEGMIF_OK;
  const size_t smem = 8 * 4096 * sizeof(bf16);
  static bool cfg = false;
  if (!cfg) {
    cfg = true;
    cudaError_t e = cudaFuncSetAttribute((const void*)sr_attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("sr_attention_bwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return SEGMIF_ERR_CUDA; }
  }
  dim3 grid((unsigned)ceil_div(N, 64), (unsigned)(B * heads));
  sr_attention_bwd_kernel<<<grid, kAbThreads, smem, as_stream(stream)>>>(
      (const bf16*)q, ldq, (const bf16*)k, (const bf16*)v, ldkv, (const bf16*)out, (const bf16*)dout, ldo, lse, (bf16*)dq, lddq,
      dkv, lddkv, v_off, heads, N, Nk, scale, scale * 1.4426950408889634f);
  return check_launch("segmif_sr_attention_bwd");
}

// Spatial-reduction attention core: out = softmax(q k^T * scale) v per (batch, head), never
// materialising the [N, Nk] score matrix (the reference writes 184 MB of scores per layer at cfg 2).
// Flash-style: 64 queries per CTA (4 warps x 16 rows), K/V streamed in 64-key tiles through a
// double-buffered cp.async ring, S and O accumulators in registers, online softmax in the exp2 domain.
#include <cstdlib>

#include "common.cuh"

namespace segmif {

template <int D>
__device__ __forceinline__ int swz_row(int row, int chunk) {
  // D=64: 128-byte rows, 8 chunks -> xor with row&7.  D=32: 64-byte rows, 4 chunks.
  return D == 64 ? (chunk ^ (row & 7)) : (chunk ^ ((row >> 1) & 3));
}

template <int D>
__global__ void __launch_bounds__(128) sr_attention_kernel(const bf16* __restrict__ q, int ldq,
                                                           const bf16* __restrict__ k, const bf16* __restrict__ v,
                                                           int ldkv, bf16* __restrict__ out, int ldo, int heads, int N,
                                                           int Nk, float scale_log2e, float* __restrict__ lse) {
  constexpr int BQ = 64, BK = 64, CH = D / 8;      // CH = 16-byte chunks per row
  constexpr int KS = D / 16;                       // k16 steps over head dim
  constexpr int NT_O = D / 8;                      // output n-tiles
  __shared__ __align__(128) bf16 sQ[BQ * D];
  __shared__ __align__(128) bf16 sK[2][BK * D];
  __shared__ __align__(128) bf16 sV[2][BK * D];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int bh = blockIdx.y, b = bh / heads, h = bh % heads;
  const int q0 = blockIdx.x * BQ;
  const bf16* qb = q + ((int64_t)b * N) * ldq + h * D;
  const bf16* kb = k + ((int64_t)b * Nk) * ldkv + h * D;
  const bf16* vb = v + ((int64_t)b * Nk) * ldkv + h * D;

  // Q tile
  for (int i = tid; i < BQ * CH; i += 128) {
    const int row = i / CH, chunk = i % CH;
    const bool ok = (q0 + row) < N;
    const bf16* g = ok ? qb + (int64_t)(q0 + row) * ldq + chunk * 8 : qb;
    cp_async16_cg(smem_u32(sQ + row * D + swz_row<D>(row, chunk) * 8), g, ok ? 16 : 0);
  }
  auto load_kv = [&](int stage, int k0) {
    for (int i = tid; i < BK * CH; i += 128) {
      const int row = i / CH, chunk = i % CH;
      const bool ok = (k0 + row) < Nk;
      const int64_t off = (int64_t)(k0 + row) * ldkv + chunk * 8;
      cp_async16_cg(smem_u32(sK[stage] + row * D + swz_row<D>(row, chunk) * 8), ok ? kb + off : kb, ok ? 16 : 0);
      cp_async16_cg(smem_u32(sV[stage] + row * D + swz_row<D>(row, chunk) * 8), ok ? vb + off : vb, ok ? 16 : 0);
    }
  };
  load_kv(0, 0);
  cp_async_commit();

  const int ntiles = (Nk + BK - 1) / BK;
  float o[NT_O][4];
#pragma unroll
  for (int i = 0; i < NT_O; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  uint32_t qf[KS][4];

  for (int t = 0; t < ntiles; ++t) {
    if (t + 1 < ntiles) load_kv((t + 1) & 1, (t + 1) * BK);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (t == 0) {
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const int row = warp * 16 + (lane & 15), chunk = ks * 2 + (lane >> 4);
        ldmatrix_x4(qf[ks], smem_u32(sQ + row * D + swz_row<D>(row, chunk) * 8));
      }
    }
    const bf16* tK = sK[t & 1];
    const bf16* tV = sV[t & 1];
    // S = Q K^T  (16 x 64 per warp)
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t bfr[4];
        const int row = np * 16 + (lane & 7) + ((lane >> 4) << 3), chunk = ks * 2 + ((lane >> 3) & 1);
        ldmatrix_x4(bfr, smem_u32(tK + row * D + swz_row<D>(row, chunk) * 8));
        mma_bf16_16816(s[np * 2], qf[ks], bfr[0], bfr[1]);
        mma_bf16_16816(s[np * 2 + 1], qf[ks], bfr[2], bfr[3]);
      }
    }
    // mask keys past Nk, scale into the exp2 domain, online softmax
    const int kbase = t * BK + (lane & 3) * 2;
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int key = kbase + nt * 8 + (j & 1);
        const float val = key < Nk ? s[nt][j] * scale_log2e : -INFINITY;
        s[nt][j] = val;
        mx[j >> 1] = fmaxf(mx[j >> 1], val);
      }
    }
    float corr[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_run[r], mx[r]);
      corr[r] = ex2_approx(m_run[r] - m_new);       // m_run = -inf on the first tile -> 0
      m_run[r] = m_new;
      l_run[r] *= corr[r];
    }
    float rs[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float p = ex2_approx(s[nt][j] - m_run[j >> 1]);
        s[nt][j] = p;
        rs[j >> 1] += p;
      }
    l_run[0] += rs[0];
    l_run[1] += rs[1];
#pragma unroll
    for (int i = 0; i < NT_O; ++i) { o[i][0] *= corr[0]; o[i][1] *= corr[0]; o[i][2] *= corr[1]; o[i][3] *= corr[1]; }
    // O += P V
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {          // 16 keys per step
      uint32_t pf[4];
      pf[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
      pf[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
      pf[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pf[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int np = 0; np < NT_O / 2; ++np) {
        uint32_t bfr[4];
        const int row = kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3), chunk = np * 2 + (lane >> 4);
        ldmatrix_x4_trans(bfr, smem_u32(tV + row * D + swz_row<D>(row, chunk) * 8));
        mma_bf16_16816(o[np * 2], pf, bfr[0], bfr[1]);
        mma_bf16_16816(o[np * 2 + 1], pf, bfr[2], bfr[3]);
      }
    }
    __syncthreads();     // everyone done with this K/V stage before it is refilled
  }

  // finalise: row sums across the quad, normalise, store bf16
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const int g = lane >> 2, tq = lane & 3;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = q0 + warp * 16 + g + r * 8;
    if (row >= N) continue;
    const float inv = 1.f / l_run[r];
    if (lse != nullptr && tq == 0) lse[(int64_t)bh * N + row] = m_run[r] + log2f(l_run[r]);   // exp2 domain, for the backward
    bf16* op = out + ((int64_t)b * N + row) * ldo + h * D + tq * 2;
#pragma unroll
    for (int nt = 0; nt < NT_O; ++nt)
      *reinterpret_cast<uint32_t*>(op + nt * 8) = pack_bf16x2(o[nt][r * 2] * inv, o[nt][r * 2 + 1] * inv);
  }
}

}  // namespace segmif

using namespace segmif;

static int sr_attention_impl(const void* q, int ldq, const void* k, const void* v, int ldkv, void* out, int ldo, int B,
                             int heads, int N, int Nk, int D, float scale, float* lse, segmif_stream_t stream) {
  SEGMIF_REQUIRE(q && k && v && out, "sr_attention: null pointer");
  SEGMIF_REQUIRE(D == 64 || D == 32, "sr_attention: head dim %d unsupported (32 or 64)", D);
  SEGMIF_REQUIRE(ldq % 8 == 0 && ldkv % 8 == 0 && ldo % 2 == 0, "sr_attention: pitches must be multiples of 8");
  SEGMIF_REQUIRE(Nk > 0, "sr_attention: Nk must be positive");
  if (B * heads == 0 || N == 0) return SEGMIF_OK;
  // kernel selection (read once): SEGMIF_ATTN = fa (default: flash-style tcgen05 kernel, head dim 64, any Nk) | mma (the mma.sync
  // flash kernel, also used for head dim 32) | SEGMIF_ATTN_TC=1 (the first tcgen05 kernel: whole score row in TMEM, Nk <= 320)
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("SEGMIF_ATTN");
    mode = (e && e[0] == 'm') ? 1 : 0;
  }
  if (sr_attention_tc_ok(B, heads, N, Nk, D, ldq, ldkv, ldo, q, k, v, out))
    return sr_attention_tc(q, ldq, k, v, ldkv, out, ldo, B, heads, N, Nk, scale, lse, as_stream(stream));
  if (mode == 0 && ldo % 8 == 0 && sr_attention_fa_tc_supported(B, heads, N, Nk, D, ldq, ldkv, ldo, q, k, v, out))
    return sr_attention_fa_tc(q, ldq, k, v, ldkv, out, ldo, B, heads, N, Nk, scale, lse, as_stream(stream));
  dim3 grid((unsigned)ceil_div(N, 64), (unsigned)(B * heads));
  const float sl2 = scale * 1.4426950408889634f;
  if (D == 64)
    sr_attention_kernel<64><<<grid, 128, 0, as_stream(stream)>>>((const bf16*)q, ldq, (const bf16*)k, (const bf16*)v, ldkv, (bf16*)out, ldo, heads, N, Nk, sl2, lse);
  else
    sr_attention_kernel<32><<<grid, 128, 0, as_stream(stream)>>>((const bf16*)q, ldq, (const bf16*)k, (const bf16*)v, ldkv, (bf16*)out, ldo, heads, N, Nk, sl2, lse);
  return check_launch("segmif_sr_attention_fwd");
}

extern "C" int segmif_sr_attention_fwd(const void* q, int ldq, const void* k, const void* v, int ldkv, void* out,
                                       int ldo, int B, int heads, int N, int Nk, int D, float scale,
                                       segmif_stream_t stream) {
  return sr_attention_impl(q, ldq, k, v, ldkv, out, ldo, B, heads, N, Nk, D, scale, nullptr, stream);
}

/* the flash-style tcgen05 kernel explicitly (head dim 64, any Nk); lse may be NULL */
extern "C" int segmif_sr_attention_fa_fwd(const void* q, int ldq, const void* k, const void* v, int ldkv, void* out, int ldo,
                                          int B, int heads, int N, int Nk, int D, float scale, float* lse, segmif_stream_t stream) {
  SEGMIF_REQUIRE(q && k && v && out, "sr_attention_fa: null pointer");
  SEGMIF_REQUIRE(sr_attention_fa_tc_supported(B, heads, N, Nk, D, ldq, ldkv, ldo, q, k, v, out),
                 "sr_attention_fa: needs head dim 64, 16-byte aligned pointers and pitches that are multiples of 8");
  return sr_attention_fa_tc(q, ldq, k, v, ldkv, out, ldo, B, heads, N, Nk, scale, lse, as_stream(stream));
}

/* the tcgen05 kernel explicitly (head dim 64, Nk <= 320); lse may be NULL */
extern "C" int segmif_sr_attention_tc_fwd(const void* q, int ldq, const void* k, const void* v, int ldkv, void* out, int ldo,
                                          int B, int heads, int N, int Nk, int D, float scale, float* lse, segmif_stream_t stream) {
  SEGMIF_REQUIRE(q && k && v && out, "sr_attention_tc: null pointer");
  SEGMIF_REQUIRE(sr_attention_tc_supported(B, heads, N, Nk, D, ldq, ldkv, ldo, q, k, v, out),
                 "sr_attention_tc: needs head dim 64, 1 <= Nk <= 320, 16-byte aligned pointers and pitches that are multiples of 8");
  return sr_attention_tc(q, ldq, k, v, ldkv, out, ldo, B, heads, N, Nk, scale, lse, as_stream(stream));
}

extern "C" int segmif_sr_attention_train_fwd(const void* q, int ldq, const void* k, const void* v, int ldkv, void* out,
                                             int ldo, int B, int heads, int N, int Nk, int D, float scale, float* lse,
                                             segmif_stream_t stream) {
  SEGMIF_REQUIRE(lse, "sr_attention_train: the log-sum-exp output [B*heads, N] is required");
  return sr_attention_impl(q, ldq, k, v, ldkv, out, ldo, B, heads, N, Nk, D, scale, lse, stream);
}

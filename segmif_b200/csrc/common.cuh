// Shared device/host helpers for the segmif_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/segmif_b200.h"

namespace segmif {

// ---- error plumbing (host) -------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);      // cudaGetLastError -> SEGMIF_ERR_CUDA + message

#define SEGMIF_REQUIRE(cond, ...)                      \
  do {                                                 \
    if (!(cond)) {                                     \
      segmif::set_error(__VA_ARGS__);                  \
      return SEGMIF_ERR_INVALID;                       \
    }                                                  \
  } while (0)

static inline cudaStream_t as_stream(segmif_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- device helpers ----------------------------------------------------------------------------
typedef __nv_bfloat16 bf16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// 16-byte async copy global->shared; src_bytes == 0 zero-fills the destination.
__device__ __forceinline__ void cp_async16_ca(uint32_t saddr, const void* g, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;\n" ::"r"(saddr), "l"(g), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async16_cg(uint32_t saddr, const void* g, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(saddr), "l"(g), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr));
}

// D(16x8,f32) += A(16x16,bf16,row) * B(16x8,bf16,col)
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// 2^x through one MUFU.EX2 with flush-to-zero: exp2f() wraps the same instruction in denormal range fix-ups (3-4 extra
// instructions per call) that softmax weights in (0, 1] never need; -inf -> +0
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

__device__ __forceinline__ float apply_act(float v, int act, float alpha) {
  switch (act) {
    case SEGMIF_ACT_RELU: return fmaxf(v, 0.0f);
    case SEGMIF_ACT_PRELU: return v >= 0.0f ? v : alpha * v;
    case SEGMIF_ACT_GELU: return gelu_erf(v);
    default: return v;
  }
}

// load/store one element or small vectors in either storage dtype
template <typename T> __device__ __forceinline__ float ld_as_float(const T* p);
template <> __device__ __forceinline__ float ld_as_float<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld_as_float<bf16>(const bf16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void st_from_float(T* p, float v);
template <> __device__ __forceinline__ void st_from_float<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_from_float<bf16>(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

// 8 consecutive channels (16 B in bf16, 32 B in f32)
__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8(bf16* p, const float (&v)[8]) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

// dwconv_tma.cu: TMA-staged depthwise 3x3 kernels (used whenever the channel count is a multiple of 64)
bool dwconv_tma_ok(int B, int H, int W, int C);
int dwconv_tma_fwd(const void* x, const float* w9c, const float* bias, void* y, int B, int H, int W, int C, int flip, int gelu,
                   cudaStream_t st);
int64_t dwconv_tma_bwd_workspace(int B, int H, int W, int C);
int dwconv_tma_gelu_bwd(const void* x, const float* w9c, const float* bias, const void* dy, void* dz, int B, int H, int W, int C,
                        float* dw9c, float* dbias, float* workspace, cudaStream_t st);

// attention_tc.cu: spatial-reduction attention on tcgen05 (head dim 64, Nk <= 320)
bool sr_attention_tc_supported(int B, int heads, int N, int Nk, int D, int ldq, int ldkv, int ldo, const void* q, const void* k,
                               const void* v, const void* out);
bool sr_attention_tc_ok(int B, int heads, int N, int Nk, int D, int ldq, int ldkv, int ldo, const void* q, const void* k, const void* v,
                        const void* out);
int sr_attention_tc(const void* q, int ldq, const void* k, const void* v, int ldkv, void* out, int ldo, int B, int heads, int N, int Nk,
                    float scale, float* lse, cudaStream_t st);

// attention_fa_tc.cu: flash-style spatial-reduction attention on tcgen05 (head dim 64, any Nk)
bool sr_attention_fa_tc_supported(int B, int heads, int N, int Nk, int D, int ldq, int ldkv, int ldo, const void* q, const void* k,
                                  const void* v, const void* out);
int sr_attention_fa_tc(const void* q, int ldq, const void* k, const void* v, int ldkv, void* out, int ldo, int B, int heads, int N,
                       int Nk, float scale, float* lse, cudaStream_t st);

// wgrad_tc.cu: 3x3 weight gradients on tcgen05 (MN-major operands, accumulators resident in TMEM)
bool wgrad_lin_tc_ok(int64_t P, int Cin, int Cout, int ldy, int ldx);
void wgrad_lin_tc_plan(int64_t P, int Cin, int Cout, int* nchunk, int64_t* tok_per_chunk, int* bn);
int wgrad_lin_tc(const void* dy, int ldy, const void* x, int ldx, int64_t P, int Cin, int Cout, float* partials, int nchunk,
                 int64_t tok_per_chunk, int bn, float* dbias, cudaStream_t st);
bool wgrad_tc_ok(int B, int H, int W, int Cin, int Cout, int taps, int dil, int ldy, int ldx);
int wgrad_tc_chunks(int B, int H, int W, int Cin, int Cout);
int wgrad_tc(const void* dy, int ldy, const void* x, int ldx, int B, int H, int W, int Cin, int Cout, int dil, float* partials,
             int nchunk, cudaStream_t st);

}  // namespace segmif

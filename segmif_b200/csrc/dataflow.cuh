// Device-side dependency tracking between CONCURRENTLY running persistent kernels (segmif_drdb_dataflow_fwd):
// the seven stages of one DRDB (x0 push a / b, the four pull layers, the 1x1) run at the same time on disjoint groups of
// SMs and hand rows to each other through L2 instead of through HBM.  A producer kernel counts finished tiles per
// (image, tile row) in global memory; a consumer tile waits until the producer tile rows covering the pixel rows it is
// about to read are complete.
//
// Tile-row geometry: every stage uses 16-row tiles, but stage l shifts its tile grid UP by shift_l rows (tile row t covers
// image rows [16 t - shift_l, 16 t - shift_l + 16)).  With shift_l = 2 (l - 1) the 2-row dilation halo of layer l's tile
// row t lies inside layer l-1's tile rows t-1 and t, so no stage ever waits for a tile row BELOW its own: all stages
// sweep the images top to bottom a tile row or two apart and the live working set is a few tile rows (tens of MB).
//
// Memory ordering: producers publish with  [stores] -> (bulk-store completion | st.global) -> fence.proxy.async ->
// __threadfence -> red.release.gpu;  consumers  ld.acquire.gpu (poll) -> fence.proxy.async -> TMA loads.
// A poll that does not succeed within ~0.5 s sets *error and returns (wrong numbers, but never a hung GPU).
#pragma once
#include "common.cuh"

namespace segmif {

struct DfDep {
  const unsigned* flags;   // [B * tiles_y] finished-tile counters of the producer stage (nullptr: no dependency)
  unsigned target;         // tiles per tile row of the producer
  int tiles_y, shift, halo;
};

struct Dataflow {
  DfDep dep[2];
  unsigned* signal;        // this stage's counters [B * tiles_y] (nullptr: do not publish)
  unsigned* error;
  int enabled;
  unsigned long long* timing;   // 2 words (diagnostics) or nullptr
};

__device__ __forceinline__ unsigned df_load_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void df_fence_proxy_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// What a polling thread already knows: producer tile rows t_lo .. t_done of image b have been seen complete.  A CTA walks
// its tiles in row order, so consecutive tiles mostly need rows that were verified for the previous tile: without this
// memo every tile paid two or three dependent L2 round trips (ld.acquire) in the TMA-producer warp, directly on the path
// that keeps the halo-tile ring full.
struct DfSeen { int b, t_done, t_lo; };

// waits until the producer rows covering image rows [ya - halo, yb + halo) of image b are complete
static __device__ __noinline__ void df_wait(const DfDep& e, unsigned* error, int b, int ya, int yb, int H, DfSeen& seen) {
  if (e.flags == nullptr) return;
  ya = ya - e.halo < 0 ? 0 : ya - e.halo;
  yb = yb + e.halo > H ? H : yb + e.halo;
  if (ya >= yb) return;
  int t0 = (ya + e.shift) >> 4;
  int t1 = (yb - 1 + e.shift) >> 4;
  if (t1 > e.tiles_y - 1) t1 = e.tiles_y - 1;
  if (seen.b != b || t0 > seen.t_done + 1 || t0 < seen.t_lo) { seen.b = b; seen.t_lo = t0; seen.t_done = t0 - 1; }   // new interval
  if (t0 <= seen.t_done) t0 = seen.t_done + 1;
  if (t0 > t1) return;
  for (int t = t0; t <= t1; ++t) {
    const unsigned* f = e.flags + (size_t)b * e.tiles_y + t;
    if (df_load_acquire(f) >= e.target) continue;
    const long long start = clock64();
    unsigned ns = 32;
    while (df_load_acquire(f) < e.target) {
      __nanosleep(ns);
      if (ns < 1024) ns <<= 1;
      if (clock64() - start > 1000000000ll) {      // ~0.5 s at 2 GHz: a bug, not a slow producer
        if (error) atomicExch(error, 1u);
        break;
      }
    }
  }
  seen.t_done = t1;
  df_fence_proxy_all();
}

// one thread, after the tile's global writes are complete from its point of view
__device__ __forceinline__ void df_signal(unsigned* counters, int b, int tiles_y, int t) {
  df_fence_proxy_all();
  __threadfence();
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counters + (size_t)b * tiles_y + t) : "memory");
}

// optional per-stage timing (diagnostics): t[0] = min over CTAs of the start time, t[1] = max of the end time (globaltimer, ns)
__device__ __forceinline__ unsigned long long df_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void df_mark_begin(unsigned long long* t) { if (t) atomicMin(t, df_now()); }
__device__ __forceinline__ void df_mark_end(unsigned long long* t) { if (t) atomicMax(t + 1, df_now()); }

// ---- stage launchers with dataflow arguments (defined next to their kernels) -----------------------------------------
struct ConvDfExtra {
  Dataflow df;
  int y_shift;
  int max_ctas;        // 0: one CTA per SM
};
int conv3x3_tc_df(const segmif_conv_params* p, const ConvDfExtra& x, cudaStream_t st);
int conv3x3_tc_tile_w(int Cin, int Cout, int dil, bool has_pre);
int drdb_push_df(const segmif_drdb_push_params* p, const ConvDfExtra& x, cudaStream_t st);
int drdb_push_tile_w(int slab_width, int n_out);
struct GemmDfExtra {
  DfDep dep;           // producer of the A rows (tile rows of an image of H x W pixels)
  unsigned* error;
  unsigned long long* timing;
  int H, W;
  int max_ctas;
};
int linear_tc_df(const segmif_linear_params* p, const GemmDfExtra& x, cudaStream_t st);

}  // namespace segmif

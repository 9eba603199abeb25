// One DRDB (core/model_fusion.py:134-157) as SEVEN CONCURRENT persistent kernels on disjoint groups of SMs
// (segmif_drdb_dataflow_fwd): x0 push a (-> g1, P2, P3), x0 push b (-> P4, P5), the four pull layers over the g-slabs and
// the 1x1 + ReLU + residual, chained by the per-tile-row counters of dataflow.cuh instead of by kernel boundaries.
//
// Why: run one after the other, the stages move ~6.8 GB per DRDB through HBM (every layer re-reads the growing slabs and
// the partial pre-activations the previous launch wrote: profiles/r1_ncu_conv3x3_tc_v10_summary.csv, 61 % of the HBM peak
// on the pull launches) against a fused floor of 0.63 GB.  Run together, a consumer reads a producer's rows a tile row or
// two after they were written, i.e. out of the 126 MB L2: the only compulsory HBM traffic left is x0 in, the result out
// and the eventual write-back of the slabs.  The kernels themselves are the sequential ones (conv_tc.cu, drdb_tc.cu,
// gemm_tc.cu) with a dependency wait in the TMA-producer warp and a counter update after each tile's stores.
//
// Launch structure: the caller's stream forks into six side streams (events), every stage is launched with a fixed
// number of CTAs (the sum never exceeds the SM count, and every CTA needs a whole SM's shared memory or most of it, so
// all seven grids are co-resident and a spinning consumer can never keep a producer off the machine), and the side
// streams join the caller's stream again.  The pattern is legal under stream capture (CUDA graphs).
#include <algorithm>

#include "dataflow.cuh"

namespace segmif {

struct DfStreams {
  cudaStream_t side[6];
  cudaEvent_t fork, join[6];
  bool ok = false;
};

static DfStreams* df_streams(int dev) {
  static DfStreams per_dev[16];
  if (dev < 0 || dev >= 16) return nullptr;
  DfStreams& s = per_dev[dev];
  if (!s.ok) {
    for (int i = 0; i < 6; ++i) {
      if (cudaStreamCreateWithFlags(&s.side[i], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
      if (cudaEventCreateWithFlags(&s.join[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    }
    if (cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    s.ok = true;
  }
  return &s;
}

static const int kShift[6] = {0, 0, 2, 4, 6, 8};     // push a, push b, layers 2..5 (dataflow.cuh: shift_l = 2 (l - 1))

}  // namespace segmif

using namespace segmif;

extern "C" size_t segmif_drdb_dataflow_workspace_bytes(int B, int H) {
  const size_t ty = (size_t)(H + 8 + 15) / 16;
  return (6 * (size_t)B * ty + 4 + 32) * sizeof(unsigned);     // counters, error word (+3 pad), 7 x {begin, end} timestamps (diagnostics)
}

extern "C" int segmif_drdb_dataflow_prepare(int device) {
  SEGMIF_REQUIRE(df_streams(device) != nullptr, "drdb_dataflow: could not create the side streams");
  return SEGMIF_OK;
}

extern "C" int segmif_drdb_dataflow_fwd(const segmif_drdb_dataflow_params* p, segmif_stream_t stream) {
  SEGMIF_REQUIRE(p && p->growth && p->partial && p->w_push_a && p->w_push_b && p->w_1x1 && p->bias_1x1 && p->out && p->flags,
                 "drdb_dataflow: null pointer");
  for (int i = 0; i < 4; ++i) SEGMIF_REQUIRE(p->w_pull[i], "drdb_dataflow: null pull weights");
  for (int i = 0; i < 5; ++i) SEGMIF_REQUIRE(p->bias[i], "drdb_dataflow: null bias");
  SEGMIF_REQUIRE(p->B > 0 && p->H > 0 && p->W > 0 && p->ld >= 224 && p->ld % 8 == 0 && p->ld_partial >= 128 && p->ld_partial % 8 == 0,
                 "drdb_dataflow: bad sizes (growth pitch >= 224, partial pitch >= 128)");
  cudaStream_t st = as_stream(stream);
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  DfStreams* ds = df_streams(dev);
  SEGMIF_REQUIRE(ds != nullptr, "drdb_dataflow: could not create the side streams");
  // SMs per stage: measured shares of the sequential launches (push a, push b, L2..L5, 1x1) unless the caller tunes them
  int ctas[7];
  int total = 0;
  for (int i = 0; i < 7; ++i) total += p->ctas[i];
  if (total == 0) {
    const int def[7] = {25, 17, 12, 16, 20, 26, 32};            // SMs of 148; the 1x1 stage runs two CTAs per SM
    int acc = 0;
    for (int i = 0; i < 7; ++i) { ctas[i] = std::max(1, def[i] * sms / 148); acc += ctas[i]; }
    ctas[6] += sms - acc;
    ctas[6] *= 2;
  } else {
    for (int i = 0; i < 7; ++i) ctas[i] = p->ctas[i];
    if (ctas[6] < 0) { total -= ctas[6]; ctas[6] = 0; }
    // the six conv stages need a whole SM per CTA; two CTAs of the 1x1 GEMM (96 KB of shared memory, 128 TMEM columns) share one
    const int sm_need = total - ctas[6] + (ctas[6] + 1) / 2;
    SEGMIF_REQUIRE(sm_need <= sms, "drdb_dataflow: the stages need %d SMs, the device has %d (all stages must be co-resident)", sm_need, sms);
    for (int i = 0; i < 6; ++i) SEGMIF_REQUIRE(ctas[i] > 0, "drdb_dataflow: stage %d has no CTAs", i);
  }
  const int B = p->B, H = p->H, W = p->W;
  const size_t ty_max = (size_t)(H + 8 + 15) / 16;
  unsigned* flags[6];
  int tiles_y[6];
  for (int s = 0; s < 6; ++s) { flags[s] = p->flags + (size_t)s * B * ty_max; tiles_y[s] = (H + kShift[s] + 15) / 16; }
  unsigned* error = p->flags + 6 * (size_t)B * ty_max;
  unsigned long long* timing = reinterpret_cast<unsigned long long*>(error + 4);     // [7][2]; begin words start at ~0 (atomicMin)
  cudaError_t ce = cudaMemsetAsync(p->flags, 0, segmif_drdb_dataflow_workspace_bytes(B, H), st);
  if (ce == cudaSuccess) ce = cudaMemsetAsync(timing, 0xff, 16 * sizeof(unsigned long long), st);
  if (ce != cudaSuccess) { set_error("drdb_dataflow: cudaMemsetAsync: %s", cudaGetErrorString(ce)); return SEGMIF_ERR_CUDA; }
  for (int s = 0; s < 7; ++s) cudaMemsetAsync(timing + 2 * s + 1, 0, sizeof(unsigned long long), st);
  // fork
  cudaEventRecord(ds->fork, st);
  for (int i = 0; i < 6; ++i) cudaStreamWaitEvent(ds->side[i], ds->fork, 0);
  auto stage_stream = [&](int s) { return s == 0 ? st : ds->side[s - 1]; };
  unsigned target[6];
  target[0] = (unsigned)((W + drdb_push_tile_w(64, 96) - 1) / drdb_push_tile_w(64, 96));
  target[1] = (unsigned)((W + drdb_push_tile_w(64, 64) - 1) / drdb_push_tile_w(64, 64));
  for (int j = 2; j <= 5; ++j) {
    const int tw = conv3x3_tc_tile_w(32 * (j - 1), 32, 2, true);
    target[j] = (unsigned)((W + tw - 1) / tw);
  }
  auto dep = [&](int s, int halo) {
    DfDep d;
    d.flags = flags[s]; d.target = target[s]; d.tiles_y = tiles_y[s]; d.shift = kShift[s]; d.halo = halo;
    return d;
  };
  DfDep none;
  none.flags = nullptr; none.target = 0; none.tiles_y = 0; none.shift = 0; none.halo = 0;
  int rc = SEGMIF_OK;
  // ---- stage 0 / 1: x0 slab pushed into all five layers
  for (int s = 0; s < 2 && rc == SEGMIF_OK; ++s) {
    segmif_drdb_push_params q;
    q.src = p->growth; q.weight = s == 0 ? p->w_push_a : p->w_push_b;
    q.B = B; q.H = H; q.W = W; q.ld_src = p->ld; q.slab_offset = 0; q.slab_width = 64; q.n_out = s == 0 ? 96 : 64;
    for (int i = 0; i < 4; ++i) { q.groups[i].bias = nullptr; q.groups[i].partial_in = nullptr; q.groups[i].dst = nullptr;
                                  q.groups[i].ld_partial_in = q.groups[i].coff_partial_in = q.groups[i].ld_dst = q.groups[i].coff_dst = q.groups[i].relu = 0; }
    if (s == 0) {
      q.groups[0].bias = p->bias[0]; q.groups[0].dst = p->growth; q.groups[0].ld_dst = p->ld; q.groups[0].coff_dst = 64; q.groups[0].relu = 1;
      for (int i = 1; i < 3; ++i) { q.groups[i].dst = p->partial; q.groups[i].ld_dst = p->ld_partial; q.groups[i].coff_dst = 32 * (i - 1); }
    } else {
      for (int i = 0; i < 2; ++i) { q.groups[i].dst = p->partial; q.groups[i].ld_dst = p->ld_partial; q.groups[i].coff_dst = 64 + 32 * i; }
    }
    ConvDfExtra x;
    x.df.enabled = 1; x.df.dep[0] = none; x.df.dep[1] = none; x.df.signal = flags[s]; x.df.error = error; x.df.timing = timing + 2 * s;
    x.y_shift = 0; x.max_ctas = ctas[s];
    rc = drdb_push_df(&q, x, stage_stream(s));
  }
  // ---- stages 2..5: layer j over the g-slabs, + its x0 share P_j, + bias, ReLU, written as slab g_j
  for (int j = 2; j <= 5 && rc == SEGMIF_OK; ++j) {
    segmif_conv_params q;
    q.src = p->growth; q.weight = p->w_pull[j - 2]; q.bias = p->bias[j - 1]; q.prelu_alpha = nullptr; q.residual = nullptr; q.dst = p->growth;
    q.B = B; q.H = H; q.W = W; q.Cin = 32 * (j - 1); q.ld_src = p->ld; q.src_coff = 64;
    q.KH = q.KW = 3; q.stride = 1; q.pad = 2; q.dil = 2; q.Ho = H; q.Wo = W; q.Cout = 32;
    q.act = SEGMIF_ACT_RELU; q.res_dtype = SEGMIF_F32; q.ld_res = 0; q.res_coff = 0;
    q.dst_dtype = SEGMIF_BF16; q.ld_dst = p->ld; q.dst_coff = 64 + 32 * (j - 1);
    q.pre_add = p->partial; q.ld_pre = p->ld_partial; q.pre_coff = 32 * (j - 2);
    ConvDfExtra x;
    x.df.enabled = 1; x.df.signal = flags[j]; x.df.error = error; x.df.timing = timing + 2 * j;
    x.df.dep[0] = dep(j == 2 ? 0 : j - 1, 2);                    // the slabs: previous layer (transitively all earlier ones)
    x.df.dep[1] = j == 2 ? none : dep(j <= 3 ? 0 : 1, 0);         // P_j: push a wrote P2, P3; push b wrote P4, P5
    x.y_shift = kShift[j]; x.max_ctas = ctas[j];
    rc = conv3x3_tc_df(&q, x, stage_stream(j));
  }
  // ---- stage 6: out = x0 + relu(conv1x1(all 224 channels)); ctas[6] < 0: run it AFTER the join, stream ordered, whole GPU
  const bool seq_1x1 = p->ctas[6] < 0;
  if (rc == SEGMIF_OK && !seq_1x1) {
    segmif_linear_params q;
    q.src = p->growth; q.weight = p->w_1x1; q.bias = p->bias_1x1; q.prelu_alpha = nullptr; q.residual = p->growth; q.dst = p->out;
    q.M = B * H * W; q.N = 64; q.K = 224; q.ld_src = p->ld; q.src_coff = 0; q.act = SEGMIF_ACT_RELU;
    q.res_dtype = SEGMIF_BF16; q.ld_res = p->ld; q.res_coff = 0; q.dst_dtype = SEGMIF_BF16; q.ld_dst = p->ld_out; q.dst_coff = p->out_coff;
    q.weight_kn = 0; q.row_scale = nullptr; q.rows_per_scale = 0;
    GemmDfExtra x;
    x.dep = dep(5, 0); x.error = error; x.timing = timing + 12; x.H = H; x.W = W; x.max_ctas = ctas[6];
    rc = linear_tc_df(&q, x, stage_stream(6));
  }
  // join (always, so that a capture in progress is left in a consistent state even after a failed launch)
  for (int i = 0; i < 6; ++i) {
    cudaEventRecord(ds->join[i], ds->side[i]);
    cudaStreamWaitEvent(st, ds->join[i], 0);
  }
  if (rc == SEGMIF_OK && seq_1x1) {
    segmif_linear_params q;
    q.src = p->growth; q.weight = p->w_1x1; q.bias = p->bias_1x1; q.prelu_alpha = nullptr; q.residual = p->growth; q.dst = p->out;
    q.M = B * H * W; q.N = 64; q.K = 224; q.ld_src = p->ld; q.src_coff = 0; q.act = SEGMIF_ACT_RELU;
    q.res_dtype = SEGMIF_BF16; q.ld_res = p->ld; q.res_coff = 0; q.dst_dtype = SEGMIF_BF16; q.ld_dst = p->ld_out; q.dst_coff = p->out_coff;
    q.weight_kn = 0; q.row_scale = nullptr; q.rows_per_scale = 0;
    rc = segmif_linear_tc_fwd(&q, stream);
  }
  return rc;
}

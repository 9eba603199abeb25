#!/bin/bash
# `ncu --set full` captures of the kernels added in round 2 (run under gpurun, ONE GPU); reports land in gpurun_out/ and are
# condensed by tools/ncu_summary.py into profiles/r2_ncu_round2_summary.csv.
set -u
mkdir -p gpurun_out
cap() {  # tag regex skip count command...
  local tag=$1 regex=$2 skip=$3 count=$4; shift 4
  timeout 400 ncu --set full --clock-control none --import-source on -k "regex:$regex" -s "$skip" -c "$count" -f -o "gpurun_out/ncu_r2_$tag" "$@" > "gpurun_out/ncu_r2_$tag.log" 2>&1
  echo "ncu $tag rc=$?"
}
cap split_gemm gemm_split_tc 40 3 python tools/strict_once.py 1
cap attention_fa sr_attention_fa 4 2 python tools/attn_bench.py
cap entropy_patch entropy_patch 1 1 python tools/loss_once.py
cap datapath "finish_kernel|resize_v_kernel|resize_h_kernel|window_hist" 4 4 python tools/dp_once.py
cap wgrad_lin_s3fc1 wgrad_lin_tc 583 1 python tools/wgrad_lin_bench.py
cap wgrad_lin_fuse wgrad_lin_tc 1219 1 python tools/wgrad_lin_bench.py
python tools/ncu_summary.py gpurun_out/ncu_r2_*.ncu-rep > gpurun_out/r2_ncu_round2_summary.csv
cat gpurun_out/r2_ncu_round2_summary.csv | cut -c1-400

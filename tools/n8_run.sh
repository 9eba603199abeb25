# One 8-GPU session: the default bench line at N=8 (inference replicas + secondary training steps with their all-reduce) and
# the train_seg workload with the single all-reduce (default) and with the overlapped two-bucket schedule.
set -x
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}.json 2> gpurun_out/bench_n${N}.err
tail -c 400 gpurun_out/bench_n${N}.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --workload train_seg --steps 20 --warmup 3 > gpurun_out/train_seg_n${N}_single.json 2> gpurun_out/train_seg_n${N}_single.err
SEGMIF_OVERLAP_ALLREDUCE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $N --workload train_seg --steps 20 --warmup 3 > gpurun_out/train_seg_n${N}_overlap.json 2> gpurun_out/train_seg_n${N}_overlap.err
python - <<PY
import json
for f in ("bench_n$N", "train_seg_n${N}_single", "train_seg_n${N}_overlap"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["metric"], round(d["value"], 1), round(d["ms_per_step"], 3), d.get("replicas_in_sync"), d.get("clocks"))
        if "secondary" in d and d["secondary"]:
            for k in ("train_seg", "train_fusion_ce", "strict_mode"):
                print("   secondary", k, d["secondary"][k].get("value"), d["secondary"][k].get("ms_per_step"), d["secondary"][k].get("replicas_in_sync"), d["secondary"][k].get("error"))
    except Exception as e:
        print(f, "ERR", e)
PY

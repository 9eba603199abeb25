# train_seg / train_fusion_ce single-GPU step times (no secondary blocks, no CPU baseline)
for w in ${WORKLOADS:-train_seg train_fusion_ce}; do
  python bench.py --workload $w --steps 10 --warmup 3 --no-secondary --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['metric'], round(d['value'],1), d['unit'], round(d['ms_per_step'],3), 'ms/step', 'launches', d.get('gpu_launches'))"
done

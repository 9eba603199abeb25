# per-kernel durations of one batch of the device data path (ncu launch list; cold-cache, serialised)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/dp_launches.csv python tools/dp_once.py 2>&1 | tail -1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/dp_launches.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    try: agg.setdefault(r[ki][:40], []).append(float(r[vi].replace(",", "")))
    except Exception: pass
for k, v in agg.items(): print(f"{k:40s} n={len(v):3d} last={v[-1]/1e3:9.1f} us")
PY

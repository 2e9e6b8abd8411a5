#!/bin/bash
# usage: bash tools/gpu_sweep.sh <tag> "<variants>" "<carveouts>"   (scratch tuning sweep)
tag=$1; variants=$2; carves=$3
mkdir -p gpurun_out
for v in $variants; do
  so=libwavecu_${v}.so; [ "$v" = default ] && so=libwavecu.so
  for c in $carves; do
    WAVECU_CARVEOUT=$c WAVECU_SO=$so timeout 300 python bench.py --skip-cpu --steps 10 --warmup 3 > gpurun_out/${tag}_${v}_c${c}.json 2> gpurun_out/${tag}_${v}_c${c}.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_${v}_c${c}.json").read().strip().splitlines()[-1])
    print("%-10s carve %4s  value %.3e ms/step %.3f launch_us %.1f"%("${v}","${c}",d["value"],d["ms_per_step"],1e3*d["roofline"]["mean_launch_ms"]), flush=True)
except Exception as e: print("${v} ${c} failed", e)
PY
  done
done
if [ -n "$NCU_CARVES" ]; then
for c in $NCU_CARVES; do
  WAVECU_CARVEOUT=$c timeout 300 ncu --metrics launch__shared_mem_config_size,l1tex__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:correspond -s 18 -c 2 --csv python bench.py --skip-cpu --steps 1 --warmup 3 2>/dev/null | grep -E "config_size|hit_rate|duration" | awk -F'","' -v c=$c '{print "carve " c ": " $(NF-2) " " $NF}'
done
fi

#!/bin/bash
# A/B of the correspondence kernels under one gpurun call: parity suites with the tiled kernel, then the
# bench with the tiled kernel and with the round-1 LBVH walk (WAVECU_NN=walk).
# usage (under gpurun): bash tools/gpu_ab.sh <tag> [pytest args]
tag=${1:-ab}; shift
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q "$@" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${tag}_pytest.log
for nn in ${NN_VARIANTS:-walk walk1 tile}; do
  WAVECU_NN=$nn timeout 600 python bench.py --skip-cpu --steps 10 --warmup 3 > gpurun_out/${tag}_bench_${nn}.json 2> gpurun_out/${tag}_bench_${nn}.err; echo "bench $nn rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench_${nn}.json").read().strip().splitlines()[-1])
    print("${nn}", "value %.3e"%d["value"], "ms/step %.3f"%d["ms_per_step"], "e2e ms %.3f"%d["e2e"]["ms_per_step"], "launch_ms %.4f"%d["roofline"]["mean_launch_ms"], d["breakdown_ms_per_step"])
except Exception as e: print("${nn} failed", e); print(open("gpurun_out/${tag}_bench_${nn}.err").read()[-2000:])
PY
done

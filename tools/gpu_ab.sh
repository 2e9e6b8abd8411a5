#!/bin/bash
# A/B under one gpurun call: the GPU parity suites, then bench.py once per variant.
# usage (under gpurun): bash tools/gpu_ab.sh <tag> "<name>:<ENV=VAL,ENV=VAL>" ...   (a bare <name> = no env)
# SKIP_TESTS=1 skips pytest.
tag=${1:-ab}; shift
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${tag}_pytest.log
fi
for v in "$@"; do
  name=${v%%:*}; envs=""
  [ "$v" != "$name" ] && envs=$(echo "${v#*:}" | tr ',' ' ')
  env $envs timeout 600 python bench.py --skip-cpu --no-batch --steps 10 --warmup 3 > gpurun_out/${tag}_bench_${name}.json 2> gpurun_out/${tag}_bench_${name}.err; echo "bench $name rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench_${name}.json").read().strip().splitlines()[-1])
    print("${name}", "value %.3e"%d["value"], "ms/step %.3f"%d["ms_per_step"], "e2e ms %.3f"%d["e2e"]["ms_per_step"], "launch_ms %.4f"%d["roofline"]["mean_launch_ms"], d["breakdown_ms_per_step"])
except Exception as e: print("${name} failed", e); print(open("gpurun_out/${tag}_bench_${name}.err").read()[-2000:])
PY
done

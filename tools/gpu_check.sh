#!/bin/bash
# One gpurun call: sanitizer on the smoke match, the GPU parity suite, bench per tuning variant, ncu.
# usage (under gpurun): bash tools/gpu_check.sh <tag> [variant ...]
tag=${1:-x}; shift
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${tag}_gpu.txt 2>&1
SMOKE='import __graft_entry__ as g; g.smoke()'
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "$SMOKE" > gpurun_out/${tag}_san_mem.log 2>&1; echo "memcheck rc=$?"
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -c "$SMOKE" > gpurun_out/${tag}_san_race.log 2>&1; echo "racecheck rc=$?"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
for v in "$@"; do
  WAVECU_SO=libwavecu_${v}.so timeout 300 python bench.py --skip-cpu --steps 10 --warmup 3 > gpurun_out/${tag}_bench_${v}.json 2> gpurun_out/${tag}_bench_${v}.err; echo "bench $v rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench_${v}.json").read().strip().splitlines()[-1])
    print("${v}", "value %.3e"%d["value"], "ms/step %.3f"%d["ms_per_step"], "launch_ms %.4f"%d["roofline"]["mean_launch_ms"], d["breakdown_ms_per_step"])
except Exception as e: print("${v} failed", e)
PY
done
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; cat gpurun_out/${tag}_bench.json
[ -n "$SKIP_LAUNCHES" ] || timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --skip-cpu --steps 2 --warmup 3 > gpurun_out/${tag}_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:correspond -s ${NCU_SKIP:-18} -c ${NCU_COUNT:-4} -f -o gpurun_out/${tag}_prof_corr python bench.py --skip-cpu --steps 1 --warmup 3 > gpurun_out/${tag}_prof.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -20

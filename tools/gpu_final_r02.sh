#!/bin/bash
# Round-2 closing run under ONE gpurun call (1 GPU): GPU test suite, smoke, the default bench line, the reference arm,
# and the ncu launch list of the default bench command (scratch output in gpurun_out/r02z_*).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02z_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02z_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 > gpurun_out/r02z_ref.json 2> gpurun_out/r02z_ref.err; echo "reference rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02z_launches.csv python bench.py --skip-cpu --no-batch --steps 2 --warmup 3 > gpurun_out/r02z_launches.log 2>&1; echo "launch list rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02z_bench.json").read().strip().splitlines()[-1])
r = json.loads(open("gpurun_out/r02z_ref.json").read().strip().splitlines()[-1])
print("value %.4g  ms/step %.4f  e2e %.4g  single %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], {k: round(v, 4) if isinstance(v, float) else v for k, v in d["single_matcher"].items() if k != "what"}))
print("roofline frac %.4f  launches %d  clocks %s" % (d["roofline"]["frac"], d["gpu_launches"], d["clocks"]))
print("parity", {k: v for k, v in d["parity"].items() if isinstance(v, bool)})
print("cpu_baseline %.4g on %d cores; reference arm %.4g; e2e ratio %.1f" % (d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], r["value"], d["e2e"]["value"] / r["value"]))
for k in ("batch256", "gicp500k", "ndt1m5m"):
    s = d[k]
    print(k, "%.4g" % s["value"], "ms/step %.3f" % s["ms_per_step"], (s.get("several_matchers") or {}).get("ms_per_match"), s.get("scans_per_s"))
PY

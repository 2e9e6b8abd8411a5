#!/bin/bash
# Round-2 profile captures under ONE gpurun call (1 GPU): launch list of the default bench command, and
# `ncu --set full` of the dominant kernel of each workload plus the tiled correspondence kernel.
# Scratch output in gpurun_out/r02_*; tools/make_profiles_r02.py turns it into profiles/.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_cpp_shim.py -m gpu -x -q 2>&1 | tail -3
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum -c 700 --csv --log-file gpurun_out/r02_launches.csv python bench.py --skip-cpu --no-batch --steps 2 --warmup 3 > gpurun_out/r02_launches.log 2>&1; echo "launch list rc=$?"
timeout 600 $NCU --set full --import-source on -k regex:iterate_kernel -s 14 -c 4 -f -o gpurun_out/r02_iterate python bench.py --skip-cpu --no-batch --steps 1 --warmup 3 > /dev/null 2>&1; echo "iterate rc=$?"
WAVECU_NN=tile timeout 600 $NCU --set full --import-source on -k regex:correspond_tile -s 8 -c 4 -f -o gpurun_out/r02_tile python tools/tile_probe.py > /dev/null 2>&1; echo "tile rc=$?"
WAVECU_FUSED=0 timeout 600 $NCU --set full -k regex:"reduce_kernel|correspond_kernel" -s 20 -c 4 -f -o gpurun_out/r02_unfused python bench.py --skip-cpu --no-batch --steps 1 --warmup 3 > /dev/null 2>&1; echo "unfused rc=$?"
timeout 600 $NCU --set full -k regex:gicp_cost -s 300 -c 3 -f -o gpurun_out/r02_gicp python tools/gicp_probe.py 500000 > /dev/null 2>&1; echo "gicp rc=$?"
cat > /tmp/ndt_probe.py <<'PY'
import sys
sys.path.insert(0, ".")
import libwave_b200 as W
from bench import ndt_clouds
scan, big = ndt_clouds(0)
m = W.NDTMatcher(W.NDTMatcherParams(res=0.5))
m.setup(scan, big)
print(m.match(), m.iterations, m.stats())
PY
timeout 900 $NCU --set full -k regex:ndt_derivative -s 10 -c 3 -f -o gpurun_out/r02_ndt python /tmp/ndt_probe.py > gpurun_out/r02_ndt.log 2>&1; echo "ndt rc=$?"
ls -la gpurun_out/r02_*

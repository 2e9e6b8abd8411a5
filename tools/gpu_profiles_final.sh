#!/bin/bash
# Final round-2 captures of the GICP / NDT kernels (1 GPU): `ncu --set full` summaries for profiles/.
mkdir -p gpurun_out
NCU="ncu --clock-control none"
python tools/gicp_variant_timing.py libwavecu.so
python tools/ndt_variant_timing.py libwavecu.so
timeout 600 $NCU --set full -k regex:gicp_cost -s 300 -c 3 -f -o gpurun_out/r02f_gicp python tools/gicp_probe_one.py > /dev/null 2>&1; echo "gicp rc=$?"
timeout 900 $NCU --set full -k regex:"ndt_derivative|ndt_leaf" -s 0 -c 12 -f -o gpurun_out/r02f_ndt python tools/ndt_probe.py > /dev/null 2>&1; echo "ndt rc=$?"
timeout 600 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r02f_ndt_launches.csv python tools/ndt_probe.py > /dev/null 2>&1; echo "ndt launches rc=$?"

ncu --clock-control none --set full --import-source on -k regex:ndt_derivative -s 10 -c 2 -f -o gpurun_out/r02b_ndt python tools/ndt_probe.py > gpurun_out/r02b_ndt.log 2>&1; echo "ndt rc=$?"

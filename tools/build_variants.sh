#!/bin/bash
# Tuning builds of libwavecu.so (selected at run time with WAVECU_SO=<file name>); not part of the product.
# usage: tools/build_variants.sh name1:"-DFLAG=..." name2:"..."
set -e
cd "$(dirname "$0")/../libwave_b200/csrc"
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  make -j8 OUT=../libwavecu_${name}.so BUILD=../_build_${name} EXTRA="${flags}" > /dev/null
  echo "built libwavecu_${name}.so (${flags})"
  grep -A2 "correspond_kernel" ../_build_${name}/icp.ptxas.log | grep -E "registers|stack" || true
done

#!/bin/bash
# A/B of builds of the same library (selected with CHI_LIB), device-timed at configs[1] (Gaussian and
# pseudo-Voigt) by tools/quick_bench.py.  usage: tools/exp_variants.sh libchi.so libchi_x.so ...
for lib in "${@:-libchi.so}"; do
  [ -f caribou_hi_b200/$lib ] || continue
  echo "== $lib"
  CHI_LIB=$PWD/caribou_hi_b200/$lib QUICK=1 python tools/quick_bench.py 100000 2>&1 | grep -E "Gunits|rror"
done

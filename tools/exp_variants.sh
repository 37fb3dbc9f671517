#!/bin/bash
# A/B of the exp2 variants of the channel kernels (csrc/moments.cuh: CHI_EXP_REPL, CHI_EXP_U) and of the
# Gaussian instantiation without the VJP plumbing (CHI_EA_NOVJP=2): builds of the same library selected
# with CHI_LIB, device-timed at configs[1] (Gaussian and pseudo-Voigt) by tools/quick_bench.py.
for lib in libchi.so libchi_u.so libchi_repl.so libchi_replu.so; do
  [ -f caribou_hi_b200/$lib ] || continue
  for novjp in 1 2; do
    echo "== $lib CHI_EA_NOVJP=$novjp"
    CHI_LIB=$PWD/caribou_hi_b200/$lib CHI_EA_NOVJP=$novjp QUICK=1 python tools/quick_bench.py 100000 2>&1 | grep -E "Gunits|rror"
  done
done

#!/bin/bash
# ncu source-level capture of k_lvg_solve_v2 for prebuilt libraries: tools/prof.sh tag lib.so [tag lib.so ...]
# writes gpurun_out/prof_<tag>.ncu-rep (read here with ncu -i ... --page source/raw --csv)
while [ $# -ge 2 ]; do
  tag=$1; lib=$2; shift 2
  cp "$lib" radex_emcee_b200/libradex_b200.so
  ncu --set full --clock-control none --import-source on -k regex:k_lvg_solve_v2 -s 1 -c 1 -f -o gpurun_out/prof_$tag \
      python bench.py --log2n 13 --steps 1 --warmup 2 --no-cpu > gpurun_out/prof_$tag.log 2>&1
  tail -2 gpurun_out/prof_$tag.log
done

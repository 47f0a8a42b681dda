#!/bin/bash
# ncu source-level capture of k_lvg_solve_v2 with the in-tree library: tools/prof.sh tag [bench.py args...]
# writes gpurun_out/{prof_<tag>.ncu-rep, raw_<tag>.csv, lines_<tag>.txt}
tag=$1; shift
ncu --set full --clock-control none --import-source on -k regex:${KERNEL:-k_lvg_solve_v2} -s ${SKIP:-1} -c 1 -f -o gpurun_out/prof_$tag \
    python bench.py --log2n 14 --steps 1 --warmup 3 --no-cpu --no-extras --no-e2e "$@" > gpurun_out/prof_$tag.log 2>&1
ncu -i gpurun_out/prof_$tag.ncu-rep --page source --csv > gpurun_out/src_$tag.csv 2>/dev/null
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/raw_$tag.csv 2>/dev/null
d=$(mktemp -d); (cd $d && cuobjdump -xelf all $OLDPWD/radex_emcee_b200/libradex_b200.so > /dev/null && nvdisasm --print-line-info radex_b200.sm_100a.cubin > dis.txt)
python tools/ncu_lines.py gpurun_out/src_$tag.csv $d/dis.txt ${LINES_KERNEL:-${KERNEL:-k_lvg_solve_v2}} 40 > gpurun_out/lines_$tag.txt
rm -f gpurun_out/src_$tag.csv
tail -1 gpurun_out/prof_$tag.log | cut -c1-200

#!/bin/bash
# Everything profiles/r2_* is made from, on one B200: tools/final_pass.sh   (about 10 minutes; outputs under gpurun_out/r2g/)
out=gpurun_out/r2g; mkdir -p $out
python bench.py > $out/bench_n1.json 2> $out/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err
python bench.py --stop radex --no-cpu > $out/bench_n1_radex.json 2>> $out/bench_n1.err
tools/launch_split.sh r2g/final > $out/launch_split.txt 2>&1
python tools/make_traffic.py gpurun_out/r2g/final_launches.csv $out/traffic.json > /dev/null 2>&1
KERNEL=k_lvg_small LINES_KERNEL=k_lvg_smallILi3E SKIP=24 tools/prof.sh r2_S3 --keep small --park-max 7 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/raw_r2_S3.csv > $out/ncu_S3.md 2>&1; cat gpurun_out/lines_r2_S3.txt >> $out/ncu_S3.md
KERNEL=k_lvg_small LINES_KERNEL=k_lvg_smallILi5E SKIP=22 tools/prof.sh r2_S5 --keep k57 --park-max 7 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/raw_r2_S5.csv > $out/ncu_S5.md 2>&1; cat gpurun_out/lines_r2_S5.txt >> $out/ncu_S5.md
SKIP=13 tools/prof.sh r2_B --keep k8 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/raw_r2_B.csv > $out/ncu_B.md 2>&1; cat gpurun_out/lines_r2_B.txt >> $out/ncu_B.md
rm -f gpurun_out/prof_r2_S3.ncu-rep gpurun_out/prof_r2_S5.ncu-rep gpurun_out/prof_r2_B.ncu-rep
for a in "--ncomp 1 --walkers 100 --steps 200 --warmup 50" "--ncomp 1 --walkers 100 --steps 200 --warmup 50 --spec -1" \
         "--ncomp 1 --walkers 100 --steps 200 --warmup 50 --stop radex" \
         "--ncomp 1 --walkers 1600 --nsources 16 --steps 100 --warmup 30" "--ncomp 1 --walkers 1600 --nsources 16 --steps 100 --warmup 30 --spec -1" \
         "--ncomp 2 --walkers 16384 --steps 20 --warmup 10 --spread 0.1" \
         "--ncomp 2 --walkers 400 --steps 100 --warmup 30" "--ncomp 2 --walkers 400 --steps 100 --warmup 30 --spec -1" \
         "--ncomp 1 --walkers 16384 --steps 40 --warmup 10"; do
  python tools/bench_sampler.py $a >> $out/sampler.jsonl 2>> $out/sampler.err
done
bash tools/spec_sweep.sh > $out/spec_sweep.txt 2>&1
python tools/parity_report.py --n 2000 --n2 8192 --out $out/parity.json > $out/parity.log 2>&1
python tools/tol_sweep.py --log2n 18 > $out/tol_sweep.jsonl 2>> $out/bench_n1.err
python tools/timing.py run 19 > $out/sections.md 2>&1
tools/microbench2 > $out/fp64_issue.txt 2>&1; tools/microbench >> $out/fp64_issue.txt 2>&1
tail -c 400 $out/bench_n1.json; echo; cat $out/launch_split.txt | tail -14; cat $out/sampler.jsonl | cut -c1-200

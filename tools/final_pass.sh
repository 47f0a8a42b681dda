#!/bin/bash
# Everything profiles/r2_* is made from, on one B200: tools/final_pass.sh   (about 8 minutes; outputs under gpurun_out/r2f/)
out=gpurun_out/r2f; mkdir -p $out
python bench.py > $out/bench_n1.json 2> $out/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err
python bench.py --stop radex --no-cpu > $out/bench_n1_radex.json 2>> $out/bench_n1.err
tools/launch_split.sh r2f/final > $out/launch_split.txt 2>&1
KERNEL=k_lvg_small LINES_KERNEL=k_lvg_smallILi5E SKIP=22 tools/prof.sh r2_S5 --keep k57 --park-max 7 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/raw_r2_S5.csv > $out/ncu_S5.md 2>&1; cat gpurun_out/lines_r2_S5.txt >> $out/ncu_S5.md
SKIP=13 tools/prof.sh r2_B --keep k8 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/raw_r2_B.csv > $out/ncu_B.md 2>&1; cat gpurun_out/lines_r2_B.txt >> $out/ncu_B.md
rm -f gpurun_out/prof_r2_S5.ncu-rep gpurun_out/prof_r2_B.ncu-rep
for a in "--ncomp 1 --walkers 100 --steps 200 --warmup 50" "--ncomp 1 --walkers 100 --steps 200 --warmup 50 --stop radex" \
         "--ncomp 1 --walkers 1600 --nsources 16 --steps 100 --warmup 30" "--ncomp 2 --walkers 16384 --steps 20 --warmup 10 --spread 0.1" \
         "--ncomp 2 --walkers 400 --steps 100 --warmup 30" "--ncomp 1 --walkers 16384 --steps 40 --warmup 10"; do
  python tools/bench_sampler.py $a >> $out/sampler.jsonl 2>> $out/sampler.err
done
python tools/parity_report.py --n 2000 --n2 8192 --out $out/parity.json > $out/parity.log 2>&1
python tools/tol_sweep.py --log2n 18 > $out/tol_sweep.jsonl 2>> $out/bench_n1.err
tail -c 400 $out/bench_n1.json; echo; cat $out/launch_split.txt | tail -14; cat $out/sampler.jsonl | cut -c1-330

#!/bin/bash
# Per-launch durations and DRAM bytes of one 2^20 step (ncu, serialised, cold cache): tools/launch_split.sh tag [bench args]
# writes gpurun_out/<tag>_launches.csv and prints the share of every kernel in the last step + the DRAM bytes of the step
tag=$1; shift
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-extras --no-e2e "$@" > gpurun_out/${tag}_launches.log 2>&1
python tools/launch_times.py gpurun_out/${tag}_launches.csv

#!/bin/bash
# usage: scripts_prof.sh <tag>  -- ncu full capture of k_fused at 100 MB (run under gpurun)
python bench.py --bytes 100000000 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_fused -s 3 -c 1 -o gpurun_out/$1 python bench.py --bytes 100000000 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_$1.log 2>&1
tail -2 gpurun_out/ncu_$1.log | cut -c1-200

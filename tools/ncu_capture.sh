#!/bin/bash
# usage: scripts_prof.sh <tag> <kernel-regex> [bench args]  -- ncu full capture at 100 MB (run under gpurun)
TAG=$1; KRE=$2; shift 2
python bench.py --bytes ${NCU_BYTES:-400000000} --steps 2 --warmup 3 --no-e2e --no-cpu "$@" > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KRE -s 3 -c 1 -o gpurun_out/$TAG python bench.py --bytes ${NCU_BYTES:-400000000} --steps 2 --warmup 3 --no-e2e --no-cpu "$@" > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log | cut -c1-200

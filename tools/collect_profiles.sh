#!/bin/bash
# Round evidence, run under gpurun on one B200:  bash tools/collect_profiles.sh <tag>
# bench JSON lines (configs 2/3/4 + reference arm), the ncu launch list of the default bench command, and one
# `ncu --set full` capture per top kernel at the bench size.  Everything lands in gpurun_out/<tag>_*.
TAG=$1
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_config2.json 2> gpurun_out/${TAG}_c2.err
python bench.py --config 3 --steps 5 --warmup 3 --no-config5 > gpurun_out/${TAG}_bench_config3.json 2> gpurun_out/${TAG}_c3.err
python bench.py --config 4 --steps 3 --warmup 3 --no-config5 > gpurun_out/${TAG}_bench_config4.json 2> gpurun_out/${TAG}_c4.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_ref.err
python tools/latency_probe.py > gpurun_out/${TAG}_latency.json 2> gpurun_out/${TAG}_lat.err
python tools/pcie_probe.py > gpurun_out/${TAG}_pcie_1gpu.json 2> gpurun_out/${TAG}_pcie.err
# launch list of the bench command itself (per-launch times under ncu are cold-cache and serialised)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/${TAG}_ncu_launches_config2.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-config5 > gpurun_out/${TAG}_ncu_launches.log 2>&1
for K in k_route k_scan k_rank_scatter; do
  ncu --set full --clock-control none --import-source on -k regex:^${K}\$ -s 3 -c 1 -o gpurun_out/${TAG}_${K}_c2 \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-config5 > gpurun_out/${TAG}_ncu_${K}.log 2>&1
done
# k_emit launches twice per step (blocks, then the segments of long blocks -- none in configs 2 and 3): capture a pair
ncu --set full --clock-control none --import-source on -k regex:^k_emit\$ -s 6 -c 2 -o gpurun_out/${TAG}_k_emit_c2 \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-config5 > gpurun_out/${TAG}_ncu_k_emit.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:^k_emit\$ -s 6 -c 2 -o gpurun_out/${TAG}_k_emit_c3 \
  python bench.py --config 3 --steps 2 --warmup 3 --no-e2e --no-cpu --no-config5 > gpurun_out/${TAG}_ncu_k_emit_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:^k_route\$ -s 3 -c 1 -o gpurun_out/${TAG}_k_route_c4 \
  python bench.py --config 4 --steps 2 --warmup 3 --no-e2e --no-cpu --no-config5 > gpurun_out/${TAG}_ncu_k_route_c4.log 2>&1
# config 4: the launch list (k_emit, k_land, k_chain, k_emit over the segments, k_runs) and the segment kernels
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 200 --csv --log-file gpurun_out/${TAG}_ncu_launches_config4.csv \
  python bench.py --config 4 --steps 2 --warmup 1 --no-e2e --no-cpu --no-config5 > gpurun_out/${TAG}_ncu_launches_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:^k_runs\$ -s 3 -c 1 -o gpurun_out/${TAG}_k_runs_c4 \
  python bench.py --config 4 --steps 2 --warmup 3 --no-e2e --no-cpu --no-config5 > gpurun_out/${TAG}_ncu_k_runs_c4.log 2>&1
ls -la gpurun_out/${TAG}_* | head -40

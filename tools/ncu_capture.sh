#!/bin/bash
# ncu evidence (B200_PROFILING.md recipe): (1) launch list of a short headline bench, (2) --set full capture of the round
# 0-2 kernels of one 1080p image (second image of tools/one.py, host-driven schedule so that every phase is its own kernel).
mkdir -p gpurun_out
B="python bench.py --mode headline --steps 2 --warmup 3 --batch 8 --no-cpu-baseline"
$B > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
O="python tools/one.py 1920 1080 4 0 1 2"
$O > gpurun_out/plain_one.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:k_blur_tile|k_r0_graph|k_relabel|k_r0_edges|k_succ_scan|k_edges' -s 43 -c 10 -f -o gpurun_out/prof_r2 $O > gpurun_out/ncu_one.log 2>&1
echo "full capture rc=$?"
tail -n 3 gpurun_out/ncu_bench.log; tail -n 3 gpurun_out/ncu_one.log

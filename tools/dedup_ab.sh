#!/bin/bash
# A/B of the duplicate elimination between rounds: per-kernel tables and device timelines with and without it.
mkdir -p gpurun_out
for cfg in "1920 1080 4 0" "3840 2160 8 1" "16384 8192 4 0" "16384 8192 8 1"; do
  tag=$(echo $cfg | tr ' ' '_')
  GSEG_DEDUP=1 python tools/prof.py $cfg > /dev/null 2>&1; mv gpurun_out/prof_*.txt gpurun_out/dd1_$tag.txt 2>/dev/null
  GSEG_DEDUP=0 python tools/prof.py $cfg > /dev/null 2>&1; mv gpurun_out/prof_*.txt gpurun_out/dd0_$tag.txt 2>/dev/null
  echo "== $cfg"; head -2 gpurun_out/dd1_$tag.txt; head -2 gpurun_out/dd0_$tag.txt
done

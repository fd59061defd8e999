#!/bin/bash
# Final single-GPU measurements of the round: driver-style bench (both arms), per-kernel tables, ncu evidence.
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2>/dev/null; echo "reference rc=$?"
graph-algorithm-image-segmentation-gpgpu_b200/gseg_batch --synth 1920x1080 --n 256 --contexts 8 --steps 20 --warmup 5 > gpurun_out/final_cpp_batch.txt 2>&1; cat gpurun_out/final_cpp_batch.txt
for c in "1920 1080 4 0" "3840 2160 8 1" "1920 1080 4 2" "16384 8192 4 0"; do python tools/prof.py $c > /dev/null 2>&1; done
bash tools/ncu_capture.sh

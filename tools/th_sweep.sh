#!/bin/bash
# Blur tile height sweep (wave quantisation at 1080p: 64x32 tiles = 1020 blocks on 740 resident slots).
cd "$(dirname "$0")/.."
PK=graph-algorithm-image-segmentation-gpgpu_b200
mkdir -p gpurun_out/th
for th in 32 24 16 40; do
  so=gpurun_out/th/libgseg_th$th.so
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-ffp-contract=off -DTH=$th -shared -o $so $PK/csrc/gseg_api.cu $PK/csrc/gseg_pool.cu || exit 1
  echo "== TH $th"
  for cfg in "1920 1080 4 0" "3840 2160 8 1"; do
    GSEG_LIB=$PWD/$so GSEG_NOBUILD=1 python tools/prof.py $cfg > /dev/null 2>&1
    grep -hE "^k_blur_tile +0|persistent" gpurun_out/prof_*.txt; rm -f gpurun_out/prof_*.txt
  done
done
rm -rf gpurun_out/th

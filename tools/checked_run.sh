#!/bin/bash
# The memcheck that can be run on this pool (compute-sanitizer is closed on it): build libgseg.so with -DGSEG_CHECKED
# (index / capacity assertions wherever a device-computed id, list slot or arena offset becomes an address, gseg_device.cuh)
# and run the parity tests and a randomised sweep against it.  Any failed assertion ends the run with GSEG_E_INTERNAL
# "checked build: bounds check at site N failed", which fails the test that triggered it.
cd "$(dirname "$0")/.."
PK=graph-algorithm-image-segmentation-gpgpu_b200
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-ffp-contract=off -DGSEG_CHECKED -shared \
     -o gpurun_out/libgseg_checked.so $PK/csrc/gseg_api.cu $PK/csrc/gseg_pool.cu || exit 1
export GSEG_LIB=$PWD/gpurun_out/libgseg_checked.so
{
echo "# checked build (-DGSEG_CHECKED) on one B200: parity tests + randomised sweep, $(date -u +%Y-%m-%dT%H:%MZ)"
python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -q -x -k "not cli and not cpp_batch" 2>&1 | tail -4
GSEG_DEDUP=1 python -m pytest tests/test_gpu_parity.py -q -x -k "felz_partition or hierarchy_levels or schedules or randomised or full_size" 2>&1 | tail -3
python tools/fuzz.py 600 2>&1 | tail -3
} > gpurun_out/checked_build.txt 2>&1
cat gpurun_out/checked_build.txt

#!/bin/bash
# compute-sanitizer record: memcheck / racecheck / synccheck / initcheck on the 320x240 smoke
# configuration -- three variants, both schedules, tail forced on and off -- through the C++ CLI (no Python in the
# process).  Usage (on the GPU box): bash tools/sanitize.sh [out_dir]; writes one summary per tool.
set -u
OUT=${1:-gpurun_out/sanitize}
mkdir -p "$OUT"
BIN=graph-algorithm-image-segmentation-gpgpu_b200/gseg
export LD_LIBRARY_PATH=graph-algorithm-image-segmentation-gpgpu_b200:${LD_LIBRARY_PATH:-}
CS=/usr/local/cuda/bin/compute-sanitizer
run_case() { # tool, tag, args...
    local tool=$1 tag=$2; shift 2
    local log="$OUT/${tool}_${tag}.log"
    timeout 600 $CS --tool "$tool" --print-limit 20 --error-exitcode 9 "$BIN" "$@" 0.8 300 20 - /tmp/san_${tool}_${tag}.ppm > "$log" 2>&1
    local rc=$?
    local summ=$(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" "$log" | tail -1)
    echo "$tool $tag rc=$rc :: ${summ:-no summary line} :: $(grep -c 'got .* components' "$log") result line(s)" | tee -a "$OUT/summary.txt"
}
: > "$OUT/summary.txt"
for tool in memcheck racecheck synccheck initcheck; do
    for variant in felz hier superpix; do
        conn=8; [ "$variant" = superpix ] && conn=4
        run_case $tool ${variant}_device --synth 320x240:1 --variant $variant --conn $conn
        run_case $tool ${variant}_hostloop --synth 320x240:1 --variant $variant --conn $conn --host-loop
    done
    run_case $tool felz_tail_all --synth 320x240:1 --variant felz --conn 8 --tail 1073741824,1073741824
    run_case $tool felz_tail_off --synth 320x240:1 --variant felz --conn 8 --tail 0,0
    run_case $tool hier_1080p_tail_all --synth 640x360:3 --variant hier --conn 8 --tail 1073741824,1073741824
done
echo "done" >> "$OUT/summary.txt"

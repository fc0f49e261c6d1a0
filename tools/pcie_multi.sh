#!/bin/bash
# N concurrent copies of tools/_build/pcie_multi (one per GPU), for N = 1, 2, 4, 8 and each direction; optional CPU pinning.
# usage: tools/pcie_multi.sh [max gpus] > gpurun_out/r02_pcie_multi.jsonl
MAX=${1:-8}
P=tools/_build/pcie_multi
nproc_=$(nproc)
for pin in 0 1; do
for N in 1 2 4 8; do
  [ $N -gt $MAX ] && continue
  for mode in h2d d2h both; do
    start=$(python3 -c "import time; print(time.time() + 4)")
    echo "{\"procs\": $N, \"mode\": \"$mode\", \"cpu_pinning\": $pin}"
    for ((g = 0; g < N; g++)); do
      if [ $pin -eq 1 ]; then c0=$((g * nproc_ / N)); c1=$(((g + 1) * nproc_ / N - 1)); taskset -c $c0-$c1 $P $g $mode $start 3 & else $P $g $mode $start 3 & fi
    done
    wait
  done
done
done

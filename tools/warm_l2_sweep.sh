#!/bin/bash
# Run on the GPU box: per-kernel launch durations of a 4-layer C2 forward at several batch sizes, with ncu leaving the
# caches alone (--cache-control none: a consumer finds what its producer just wrote in L2) and with the default flush.
mkdir -p gpurun_out
for B in 4 8 16 32; do
  PROFILE_BATCH=$B PROFILE_ITERS=3 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv \
      --log-file gpurun_out/warm_B$B.csv python tools/profile_block.py > gpurun_out/warm_B$B.log 2>&1
  PROFILE_BATCH=$B PROFILE_ITERS=3 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/cold_B$B.csv python tools/profile_block.py > gpurun_out/cold_B$B.log 2>&1
done
for B in 4 8 16 32; do
  echo "== warm B=$B"; python tools/ncu_launch_summary.py gpurun_out/warm_B$B.csv
  echo "== cold B=$B"; python tools/ncu_launch_summary.py gpurun_out/cold_B$B.csv
done

#!/bin/bash
# Run on the GPU box (via gpurun): launch list of the bench command + full ncu captures of the three pipeline kernels.
set -x
mkdir -p gpurun_out
FFNO_B200_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
for k in ff_ts_kernel axis_pipe_kernel mix_pipe_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 2 -o gpurun_out/prof_$k \
      python tools/profile_block.py > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep

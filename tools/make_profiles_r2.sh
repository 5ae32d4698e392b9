#!/bin/bash
# Run on the GPU box (via gpurun): round-2 ncu evidence.  Launch list of the bench command, --set full captures of the
# three C2 pipeline kernels, launch lists + one --set full capture each for C4 and C5.
set -x
mkdir -p gpurun_out
FFNO_B200_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv \
    --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager-gpu > gpurun_out/r02_ncu_launch.log 2>&1
for k in ff_ts_kernel axis_pipe_kernel mix_pipe_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 2 -o gpurun_out/r02_prof_$k \
      python tools/profile_block.py > gpurun_out/r02_ncu_$k.log 2>&1
done
for c in c4 c5; do
  PROFILE_CONFIG=$c ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_$c.csv \
      python tools/profile_config.py > gpurun_out/r02_ncu_launch_$c.log 2>&1
done
PROFILE_CONFIG=c4 ncu --set full --clock-control none --import-source on -k regex:axis_pipe_kernel -s 10 -c 2 -o gpurun_out/r02_prof_axis_c4 \
    python tools/profile_config.py > gpurun_out/r02_ncu_axis_c4.log 2>&1
PROFILE_CONFIG=c5 ncu --set full --clock-control none --import-source on -k regex:axis_pipe_kernel -s 12 -c 2 -o gpurun_out/r02_prof_axis_c5 \
    python tools/profile_config.py > gpurun_out/r02_ncu_axis_c5.log 2>&1
PROFILE_CONFIG=c5 ncu --set full --clock-control none --import-source on -k regex:ff_ts_kernel -s 3 -c 1 -o gpurun_out/r02_prof_ff_c5 \
    python tools/profile_config.py > gpurun_out/r02_ncu_ff_c5.log 2>&1
ls -la gpurun_out/*.ncu-rep

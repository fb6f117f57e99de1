#!/bin/bash
# Run under gpurun (1 GPU).  Writes raw captures to gpurun_out/; summaries are distilled into profiles/ by profiles/summarize.py.
set -x
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu"
# (1) every launch with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/launches.log 2>&1
# (2) full captures of the hot kernels
ncu --set full --clock-control none --import-source on -k regex:ac_loss_grad -s 2 -c 1 -o gpurun_out/prof_loss -f $B > gpurun_out/prof_loss.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rollout_tc -s 1 -c 1 -o gpurun_out/prof_rollout -f $B > gpurun_out/prof_rollout.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:forward_ -s 1 -c 1 -o gpurun_out/prof_fwd -f $B > gpurun_out/prof_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:reduce_clip_adam -s 2 -c 1 -o gpurun_out/prof_adam -f $B > gpurun_out/prof_adam.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:env_step_kernel -s 1 -c 1 -o gpurun_out/prof_env -f $B > gpurun_out/prof_env.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_series_fastest -s 1 -c 1 -o gpurun_out/prof_gae -f $B > gpurun_out/prof_gae.log 2>&1
ls -la gpurun_out

#!/bin/bash
# Run under gpurun (1 GPU).  Writes raw captures to gpurun_out/; summaries are distilled into profiles/ by profiles/summarize.py <tag>.
# The bench replays a CUDA graph per iteration; the captures use eager launches (B200RL_GRAPH=0) so that every kernel is a plain launch.
set -x
mkdir -p gpurun_out
export B200RL_GRAPH=0
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu"
C5="python bench.py --config c5 --steps 1 --warmup 1"
# (1) every launch with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/launches.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 400 --csv --log-file gpurun_out/launches_c5.csv $C5 > gpurun_out/launches_c5.log 2>&1
# (2) full captures of the hot kernels
ncu --set full --clock-control none --import-source on -k regex:ac_loss_grad -s 2 -c 1 -o gpurun_out/prof_loss -f $B > gpurun_out/prof_loss.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rollout_tc -s 1 -c 1 -o gpurun_out/prof_rollout -f $B > gpurun_out/prof_rollout.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:reduce_clip_adam -s 2 -c 1 -o gpurun_out/prof_adam -f $B > gpurun_out/prof_adam.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pack_records -s 1 -c 1 -o gpurun_out/prof_pack -f $B > gpurun_out/prof_pack.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_series_fastest -s 1 -c 1 -o gpurun_out/prof_gae -f $B > gpurun_out/prof_gae.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sample_gather -s 300 -c 1 -o gpurun_out/prof_gather -f $C5 > gpurun_out/prof_gather.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dqn_loss_grad -s 300 -c 1 -o gpurun_out/prof_dqn -f $C5 > gpurun_out/prof_dqn.log 2>&1
ls -la gpurun_out

# usage: ab.sh name1=path1[,ENV=VAL...] name2=path2 ... : K7 / rollout / step time of each library variant (+ environment) on this box
for kv in "$@"; do
  name=${kv%%=*}; rest=${kv#*=}; lib=${rest%%,*}; envs=""
  if [ "$rest" != "$lib" ]; then envs=$(echo "${rest#*,}" | tr ',' ' '); fi
  env B200RL_LIB=$lib $envs timeout 250 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_$name.json").read().strip().splitlines()[-1])
    p=d["phases_per_rank"][0]
    print("$name", "Msteps/s", round(d["value"]/1e6,1), "step_ms", round(d["ms_per_step"],3), "K7_ms", round(d["roofline"]["ms_per_launch"],4), "rollout_ms", round(p["rollout_ms"],4), "k7sum", round(p["loss_backward_ms"],3), "k8sum", round(p["optimiser_exchange_ms"],3))
except Exception as e:
    print("$name", "FAILED", e, open("gpurun_out/ab_$name.err").read()[-400:])
PY
done

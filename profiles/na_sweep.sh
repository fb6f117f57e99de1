for na in 80 81 82 84; do
  B200RL_K7_ACTOR_CTAS=$na timeout 250 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/na_$na.json 2>gpurun_out/na_$na.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/na_$na.json").read().strip().splitlines()[-1])
print("c2 na=$na", round(d["value"]/1e6,1), round(d["ms_per_step"],3), round(d["roofline"]["ms_per_launch"],4))
PY
done
for na in 78 79 80 81; do
  B200RL_K7_ACTOR_CTAS=$na timeout 250 python bench.py --config c3 --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/na3_$na.json 2>gpurun_out/na3_$na.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/na3_$na.json").read().strip().splitlines()[-1])
print("c3 na=$na", round(d["value"]/1e6,1), round(d["ms_per_step"],3), round(d["roofline"]["ms_per_launch"],4))
PY
done

# usage: mgpu.sh N [extra env ...]: bench.py on N GPUs of this box (torchrun), prints the headline + per-rank phases
N=$1; shift
tag=$(echo "$*" | tr ' =' '__')
env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/mg_${N}_$tag.json 2> gpurun_out/mg_${N}_$tag.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/mg_${N}_$tag.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("N=$N $*", "Msteps/s", round(d["value"]/1e6,1), "step_ms", round(d["ms_per_step"],3), "weak", d.get("weak_scaling") and round(d["weak_scaling"]["value"]/1e6,1), "K7_ms", round(d["roofline"]["ms_per_launch"],4))
    for p in d["phases_per_rank"]: print("   ", {k:(round(v,3) if isinstance(v,float) else v) for k,v in p.items()})
    print("   grad_allreduce:", d["config"]["grad_allreduce"], "| replicas_bit_identical:", d.get("replicas_bit_identical"))
except Exception as e:
    print("N=$N $* FAILED", e); print(open("gpurun_out/mg_${N}_$tag.err").read()[-1500:])
PY

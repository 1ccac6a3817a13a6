# configs[3] on 4 GPUs / configs[2] on 2 GPUs (run under `gpurun --gpus N`): bash tools/ubench/run_scale4.sh N
N=${1:-4}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_scale_n$N.json 2> gpurun_out/r02_scale_n$N.err
python - <<PY
import json
f="gpurun_out/r02_scale_n$N.json"
try:
    d=json.loads([l for l in open(f) if l.startswith("{")][-1])
    print(f, {k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], "parity", d["parity"]["mismatches"], d["parity"]["findings_compared"], d["config"]["sharding"][:40])
    for r in d["ranks"]: print("  ", r["rank"], round(r["ms"],2), [round(x,1) for x in r["mission_ms"]])
except Exception as e:
    print(f, "FAILED", e, open(f.replace(".json",".err")).read()[-1500:])
PY

# configs[4] on 8 GPUs (run under `gpurun --gpus 8`)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_scale_n8.json 2> gpurun_out/r02_scale_n8.err
python - <<PY
import json
for f in ("gpurun_out/r02_scale_n8.json",):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, {k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], "parity", d["parity"]["mismatches"], d["parity"]["findings_compared"])
        for r in d["ranks"]: print("  ", r["rank"], round(r["ms"],2), [round(x,1) for x in r["mission_ms"]])
    except Exception as e:
        print(f, "FAILED", e, open(f.replace(".json",".err")).read()[-1500:])
PY

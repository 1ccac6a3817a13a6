# multi-GPU bench lines (run under `gpurun --gpus 8`): configs[4] on 8 GPUs, configs[3] on 4, configs[1] range-sharded over 8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 6 --warmup 3 > gpurun_out/r02_scale_n8.json 2> gpurun_out/r02_scale_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --config 1 --steps 20 --warmup 5 > gpurun_out/r02_single_mission_n8.json 2> gpurun_out/r02_single.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r02_scale_n4.json 2> gpurun_out/r02_scale_n4.err
python - <<PY
import json
for f in ("gpurun_out/r02_scale_n8.json","gpurun_out/r02_single_mission_n8.json","gpurun_out/r02_scale_n4.json"):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, {k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], "parity", d["parity"]["mismatches"], d["parity"]["findings_compared"])
        for r in d["ranks"]: print("  ", r["rank"], round(r["ms"],2), [round(x,1) for x in r["mission_ms"]])
    except Exception as e:
        print(f, "FAILED", e, open(f.replace(".json",".err").replace("r02_single_mission_n8","r02_single")).read()[-1500:])
PY

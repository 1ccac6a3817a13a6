# the missions of one rank side by side (one host thread + stream each) or one after the other
for cfg in "8 1" "8 0" "4 1" "4 0"; do set -- $cfg
  python bench.py --gpus $1 --as-rank 1 --mission-threads $2 --steps 3 --warmup 2 --no-cpu --no-e2e --no-parity > gpurun_out/bs.json 2> gpurun_out/bs.err; tail -c 200 gpurun_out/bs.err
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bs.json") if l.startswith("{")][-1])
r=d["roofline"]; print("N$1 threads $2", round(d["ms_per_step"],2), {k:round(v,1) for k,v in r["mission_ms_rank0"].items()})
PY
done

# GPU check of the current tree: -m gpu tests, headline bench, occupancy experiment for the members / declined kernels
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/b1.json 2> gpurun_out/b1.err; tail -c 300 gpurun_out/b1.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/b1.json") if l.startswith("{")][-1])
r=d["roofline"]; print("N1", d["ms_per_step"], d["parity"]["mismatches"], d["e2e"]["value"], d["e2e"]["d2h_bytes_per_step"], {k[:12]:round(v,3) for k,v in r["kernels_ms"].items()}, [round(x,3) for x in r["host_phase_ms"]])
PY
for cfg in "4 4" "6 6" "8 8"; do set -- $cfg; for m in 6 7; do
  SX_MEMBERS_MINB=$1 SX_DECLINED_MINB=$2 python bench.py --gpus 8 --as-rank 3 --only $m --steps 3 --warmup 2 --no-cpu --no-e2e --no-parity > gpurun_out/bm.json 2> gpurun_out/bm.err; tail -c 200 gpurun_out/bm.err
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bm.json") if l.startswith("{")][-1])
r=d["roofline"]; print("MINB $1 $2 mission $m", round(d["ms_per_step"],2), r["mission"], {k[:12]:round(v,2) for k,v in r["kernels_ms"].items()}, [round(x,1) for x in r["host_phase_ms"]])
PY
done; done

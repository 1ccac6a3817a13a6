# shared-memory carve-out preference of the exact-stage kernels: max shared (100) vs default (-1) vs max L1 (0)
for co in 100 -1 0 25; do
  SX_CARVEOUT=$co python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-parity > gpurun_out/b1.json 2> gpurun_out/b1.err; tail -c 300 gpurun_out/b1.err
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/b1.json") if l.startswith("{")][-1])
r=d["roofline"]; print("CARVEOUT $co N1", round(d["ms_per_step"],4), {k[:12]:round(v,3) for k,v in r["kernels_ms"].items()})
PY
  for cfg in "8 6" "4 3"; do set -- $cfg
  SX_CARVEOUT=$co python bench.py --gpus $1 --as-rank 1 --only $2 --steps 3 --warmup 2 --no-cpu --no-e2e --no-parity > gpurun_out/bm.json 2> gpurun_out/bm.err; tail -c 200 gpurun_out/bm.err
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bm.json") if l.startswith("{")][-1])
r=d["roofline"]; print("CARVEOUT $co N$1 mission $2", round(d["ms_per_step"],2), r["mission"], {k[:12]:round(v,2) for k,v in r["kernels_ms"].items()})
PY
  done
done

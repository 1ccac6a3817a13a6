# pieces of the exact stage behind ONE prefilter launch: sweep K and lanes on the headline workload
for cfg in "1 1" "2 2" "4 2" "4 4" "8 2" "8 4" "6 3"; do set -- $cfg
  SX_PIECES=$1 SX_LANES=$2 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-parity > gpurun_out/b1.json 2> gpurun_out/b1.err; tail -c 300 gpurun_out/b1.err
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/b1.json") if l.startswith("{")][-1])
r=d["roofline"]; print("K $1 LANES $2", round(d["ms_per_step"],4), {k[:12]:round(v,3) for k,v in r["kernels_ms"].items()}, [round(x,3) for x in r["host_phase_ms"]])
PY
done
python -m pytest tests -m gpu -x -q 2>&1 | tail -4

set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
for g in 444 412 380 348 296 222; do
  SX_PREF_GRID=$g python bench.py --steps 10 --warmup 3 --no-cpu --e2e-max-mib 64 > gpurun_out/sweep_g$g.json 2> gpurun_out/sweep_g$g.err
  python - <<PY
import json
d=json.load(open("gpurun_out/sweep_g$g.json"))
print("GRID $g ms_per_step", d["ms_per_step"], "pref_ms", d["roofline"]["kernels_ms"]["sx_prefilter_kernel"], "frac", d["roofline"]["frac"])
PY
done

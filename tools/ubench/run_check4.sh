# convergent byte-wise automaton (sx_fast_generic.cuh): the missions it serves, one at a time, then the parity tests
for cfg in "8 6" "8 7" "8 2" "4 3" "2 1"; do set -- $cfg
  python bench.py --gpus $1 --as-rank 1 --only $2 --steps 3 --warmup 2 --no-cpu --no-e2e --no-parity > gpurun_out/bm.json 2> gpurun_out/bm.err; tail -c 200 gpurun_out/bm.err
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bm.json") if l.startswith("{")][-1])
r=d["roofline"]; print("N$1 mission $2", round(d["ms_per_step"],2), r["mission"], {k[:12]:round(v,2) for k,v in r["kernels_ms"].items()})
PY
done
if [ -n "$SX_CHECK_TESTS" ]; then timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5; fi
if [ -n "$SX_CHECK_NCU" ]; then
ncu --set full --import-source on --clock-control none -k regex:"sx_sp_(declined|members)" -c 8 -f -o gpurun_out/prof_eucjpb \
  python bench.py --gpus 8 --as-rank 1 --only 6 --steps 1 --warmup 1 --no-cpu --no-e2e --no-parity > gpurun_out/ncu_eucjpb.log 2>&1
fi
if [ -n "$SX_CHECK_NCU2" ]; then
ncu --set full --import-source on --clock-control none -k regex:"sx_sp_(members|gather)" --launch-skip 4 -c 3 -f -o gpurun_out/prof_koi8 \
  python bench.py --gpus 8 --as-rank 1 --only 7 --steps 1 --warmup 1 --no-cpu --no-e2e --no-parity > gpurun_out/ncu_koi8.log 2>&1
fi
ls -la gpurun_out/

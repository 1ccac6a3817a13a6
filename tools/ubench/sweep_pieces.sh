# usage: bash tools/ubench/sweep_pieces.sh "<pieces> <grid>" ...  -- headline workload under different piece counts / prefilter grids
for cfg in "$@"; do
  set -- $cfg
  SX_PIECES=$1 SX_PREF_GRID=$2 python bench.py --steps 10 --warmup 3 --no-cpu --e2e-max-mib 64 > gpurun_out/sw_p$1_g$2.json 2> gpurun_out/sw_p$1_g$2.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/sw_p$1_g$2.json"))
    r=d["roofline"]
    print("PIECES $1 GRID $2 ms_per_step %.4f pref_ms %.4f pipeline %.4f host %s kernels %s" % (d["ms_per_step"], r["kernels_ms"]["sx_prefilter_kernel"], r["pipeline_ms"], [round(x,3) for x in r["host_phase_ms"]], {k[:14]:round(v,3) for k,v in r["kernels_ms"].items()}))
except Exception as e:
    print("PIECES $1 GRID $2 failed", e, open("gpurun_out/sw_p$1_g$2.err").read()[-800:])
PY
done

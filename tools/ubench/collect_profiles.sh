# Collect the round's profile artefacts on one B200 (run under gpurun); outputs under gpurun_out/ (keep below 64 MiB).
set -x
python bench.py > gpurun_out/r02_bench_line.json 2> gpurun_out/r02_bench.err
python bench.py --impl reference --steps 5 --warmup 1 --ref-seconds 40 > gpurun_out/r02_reference_line.json 2>> gpurun_out/r02_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-parity --no-e2e > gpurun_out/r02_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sx_prefilter|sx_sp_heads|sx_sp_gather|sx_sp_members" -s 6 -c 4 \
    -f -o gpurun_out/r02_prof python bench.py --steps 2 --warmup 1 --no-cpu --no-parity --no-e2e > gpurun_out/r02_ncu_full.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke.log 2>&1
tail -2 gpurun_out/r02_smoke.log
ls -la gpurun_out/ | tail -12

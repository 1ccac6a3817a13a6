# ncu --set full of the byte-wise exact-stage kernels (declined = heads without a mask engine, members) on the Big5 and EUC-JP missions
ncu --set full --import-source on --clock-control none -k regex:"sx_sp_(declined|members)" -c 2 -f -o gpurun_out/prof_big5b \
  python bench.py --gpus 4 --as-rank 1 --only 3 --steps 1 --warmup 1 --no-cpu --no-e2e --no-parity > gpurun_out/ncu_big5b.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"sx_sp_(declined|members)" -c 2 -f -o gpurun_out/prof_eucjpb \
  python bench.py --gpus 8 --as-rank 1 --only 6 --steps 1 --warmup 1 --no-cpu --no-e2e --no-parity > gpurun_out/ncu_eucjpb.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3

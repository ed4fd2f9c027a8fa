mkdir -p gpurun_out/r2g
ncu --set full --clock-control none --import-source on -k regex:k_advance_tile -s 6 -c 4 -o gpurun_out/r2g/prof_tile python bench.py --steps 6 --warmup 2 --no-cpu --no-e2e --sort-miss 0.0005 --sort-max 4 --sort-full 0 > gpurun_out/r2g/b2.log 2>&1
ls -la gpurun_out/r2g

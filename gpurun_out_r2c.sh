mkdir -p gpurun_out/r2c
timeout 600 python -m pytest tests/test_gpu_tile.py -x -q > gpurun_out/r2c/pytest_tile.log 2>&1; echo "tile rc=$?"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "fused or tiled or sort or deposit" > gpurun_out/r2c/pytest_parity.log 2>&1; echo "parity rc=$?"
tail -4 gpurun_out/r2c/pytest_tile.log; tail -4 gpurun_out/r2c/pytest_parity.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_advance|k_regroup|k_scan|k_copy_parked|k_after" -c 200 --csv --log-file gpurun_out/r2c/launches_tile.csv python bench.py --steps 12 --warmup 4 --no-cpu --no-e2e --sort-miss 0.002 --sort-max 4 --sort-full 0 > gpurun_out/r2c/b1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_advance_tile -s 10 -c 2 -o gpurun_out/r2c/prof_tile python bench.py --steps 6 --warmup 4 --no-cpu --no-e2e --sort-miss 0.002 --sort-max 64 --sort-full 0 > gpurun_out/r2c/b2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_advance_tiled -s 10 -c 2 -o gpurun_out/r2c/prof_legacy python bench.py --steps 6 --warmup 4 --no-cpu --no-e2e --advance-path 1 > gpurun_out/r2c/b3.log 2>&1
ls -la gpurun_out/r2c

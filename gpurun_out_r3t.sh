mkdir -p gpurun_out/r3t
timeout 1400 ncu --set full --clock-control none --import-source on -k regex:'k_advance_tile' --launch-skip 14 -c 12 -o gpurun_out/r3t/ncu_tile -f python bench.py --steps 10 --warmup 5 --no-cpu --no-e2e > gpurun_out/r3t/ncu_tile.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r3t

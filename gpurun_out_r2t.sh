mkdir -p gpurun_out/r2t
timeout 1400 ncu --set full --clock-control none --import-source on -k regex:'k_advance_tile' --launch-skip 12 -c 9 -o gpurun_out/r2t/ncu_tile -f python bench.py --steps 8 --warmup 5 --no-cpu --no-e2e > gpurun_out/r2t/ncu_tile.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r2t/ncu_tile.log
ls -la gpurun_out/r2t

mkdir -p gpurun_out/r2s
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_mcc_(test|collide|select_skip)' --launch-skip 30 -c 9 -o gpurun_out/r2s/ncu_mcc -f python bench.py --steps 6 --warmup 5 --no-cpu --no-e2e > gpurun_out/r2s/ncu_mcc.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r2s/ncu_mcc.log
ls -la gpurun_out/r2s

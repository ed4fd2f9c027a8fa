mkdir -p gpurun_out/r2d
timeout 600 python -m pytest tests/test_gpu_tile.py -x -q > gpurun_out/r2d/pytest_tile.log 2>&1; echo "tile rc=$?"
for cfg in "m002_x8:--sort-miss 0.002 --sort-max 8" "m002_x4:--sort-miss 0.002 --sort-max 4" "m005_x0:--sort-miss 0.005 --sort-max 0" "path1:--advance-path 1"; do
  name=${cfg%%:*}; args=${cfg#*:}
  python bench.py --steps 64 --warmup 8 --no-cpu --no-e2e --sort-full 0 $args > gpurun_out/r2d/bench_$name.json 2> gpurun_out/r2d/bench_$name.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_advance" -c 80 --csv --log-file gpurun_out/r2d/launches_tile.csv python bench.py --steps 12 --warmup 4 --no-cpu --no-e2e --sort-miss 0.002 --sort-max 8 --sort-full 0 > gpurun_out/r2d/b1.log 2>&1
tail -3 gpurun_out/r2d/pytest_tile.log
for f in gpurun_out/r2d/bench_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.load(open('$f')); r=d['roofline']
    print(' ms/step %.3f value %.3e kernel_frac %.3f avg_launch_ms %.3f share %.3f'%(d['ms_per_step'],d['value'],r['frac'],r['avg_launch_ms'],r['kernel_share_of_step']), r.get('window_stats(gather_miss,deposit_miss,moves,rounds)'))
except Exception as e: print(' failed',e)
"; done

mkdir -p gpurun_out/r2e
timeout 900 python -m pytest tests/test_gpu_tile.py tests/test_gpu_mcc.py -x -q > gpurun_out/r2e/pytest_a.log 2>&1; echo "tile+mcc rc=$?"
for cfg in "m0005_x5:--sort-miss 0.0005 --sort-max 5" "m0005_x4:--sort-miss 0.0005 --sort-max 4" "m0005_x6:--sort-miss 0.0005 --sort-max 6" "m0005_x3:--sort-miss 0.0005 --sort-max 3"; do
  name=${cfg%%:*}; args=${cfg#*:}
  python bench.py --steps 60 --warmup 8 --no-cpu --no-e2e --sort-full 0 $args > gpurun_out/r2e/bench_$name.json 2> gpurun_out/r2e/bench_$name.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/r2e/launches.csv python bench.py --steps 12 --warmup 6 --no-cpu --no-e2e --sort-miss 0.0005 --sort-max 4 --sort-full 0 > gpurun_out/r2e/b1.log 2>&1
tail -3 gpurun_out/r2e/pytest_a.log
for f in gpurun_out/r2e/bench_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.load(open('$f')); r=d['roofline']
    print(' ms/step %.3f value %.3e kernel_frac %.3f avg_launch_ms %.3f share %.3f'%(d['ms_per_step'],d['value'],r['frac'],r['avg_launch_ms'],r['kernel_share_of_step']), r.get('window_stats(gather_miss,deposit_miss,moves,rounds)'))
except Exception as e: print(' failed',e)
"; done

mkdir -p gpurun_out/r2m
for cfg in "x4:--sort-miss 0.0005 --sort-max 4" "x5:--sort-miss 0.0005 --sort-max 5" "x6:--sort-miss 0.0005 --sort-max 6" "x8m002:--sort-miss 0.002 --sort-max 8" "x3:--sort-miss 0.0005 --sort-max 3"; do
  name=${cfg%%:*}; args=${cfg#*:}
  python bench.py --steps 60 --warmup 8 --no-cpu --no-e2e --sort-full 0 $args > gpurun_out/r2m/bench_$name.json 2> gpurun_out/r2m/bench_$name.err
done
for f in gpurun_out/r2m/bench_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.load(open('$f')); r=d['roofline']
    print(' ms/step %.3f value %.3e kernel_frac %.3f avg_launch_ms %.3f share %.3f'%(d['ms_per_step'],d['value'],r['frac'],r['avg_launch_ms'],r['kernel_share_of_step']), r.get('window_stats(gather_miss,deposit_miss,moves,rounds)'))
except Exception as e: print(' failed',e)
"; done

mkdir -p gpurun_out/r3n
for cfg in "0.005 5" "0.005 6" "0.01 6" "0.02 6" "0.02 7" "0.02 8" "0.01 5"; do
  set -- $cfg
  tag=m$1_x$2
  python bench.py --steps 48 --warmup 12 --no-cpu --no-e2e --sort-miss $1 --sort-max $2 > gpurun_out/r3n/bench_$tag.json 2> gpurun_out/r3n/bench_$tag.err
  python -c "
import json
d=json.load(open('gpurun_out/r3n/bench_$tag.json')); r=d['roofline']
print('$tag ms/step %.3f measured %.3f kernel_frac %.3f step_frac %.3f'%(d['ms_per_step'],r['ms_per_step_measured'],r['frac'],r['step_frac']), r['reorder_in_timed_region']['e-'])"
done

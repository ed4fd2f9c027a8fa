mkdir -p gpurun_out/r2z
for cfg in "0.0005 4" "0.002 4" "0.005 4" "0.005 5" "0.005 6" "0.02 6" "0.02 8" "0.05 10"; do
  set -- $cfg
  tag=m$1_x$2
  python bench.py --steps 40 --warmup 12 --no-cpu --no-e2e --sort-miss $1 --sort-max $2 > gpurun_out/r2z/bench_$tag.json 2> gpurun_out/r2z/bench_$tag.err
  python -c "
import json
d=json.load(open('gpurun_out/r2z/bench_$tag.json')); r=d['roofline']
print('$tag ms/step %.3f measured %.3f kernel_frac %.3f avg_launch_ms %.3f share %.3f'%(d['ms_per_step'],r['ms_per_step_measured'],r['frac'],r['avg_launch_ms'],r['kernel_share_of_step']), r['reorder_in_timed_region']['e-'], list(r.values())[9]['e-'][:2] if False else '')"
done

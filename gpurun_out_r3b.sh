mkdir -p gpurun_out/r3b
for v in pf0 pf1 pf2; do
  cp exp/lib_$v.so iskra_b200/libiskra_b200.so
  python bench.py --steps 40 --warmup 12 --no-cpu --no-e2e > gpurun_out/r3b/bench_$v.json 2> gpurun_out/r3b/bench_$v.err
  python -c "
import json
d=json.load(open('gpurun_out/r3b/bench_$v.json')); r=d['roofline']
print('$v ms/step %.3f measured %.3f kernel_frac %.3f avg_launch_ms %.3f share %.3f'%(d['ms_per_step'],r['ms_per_step_measured'],r['frac'],r['avg_launch_ms'],r['kernel_share_of_step']))"
done

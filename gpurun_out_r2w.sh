mkdir -p gpurun_out/r2w
for v in base l2g32 l2g64; do
  cp exp/lib_$v.so iskra_b200/libiskra_b200.so
  python bench.py --steps 40 --warmup 8 --no-cpu --no-e2e > gpurun_out/r2w/bench_$v.json 2> gpurun_out/r2w/bench_$v.err
  grep "L2 fetch" gpurun_out/r2w/bench_$v.err | head -1
  python -c "
import json
d=json.load(open('gpurun_out/r2w/bench_$v.json')); r=d['roofline']
print('$v ms/step %.3f measured %.3f kernel_frac %.3f avg_launch_ms %.3f share %.3f'%(d['ms_per_step'],r['ms_per_step_measured'],r['frac'],r['avg_launch_ms'],r['kernel_share_of_step']))"
done
cp exp/lib_l2g32.so iskra_b200/libiskra_b200.so
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:'k_mcc' -c 60 --csv --log-file gpurun_out/r2w/launches_l2g32.csv python bench.py --steps 6 --warmup 4 --no-cpu --no-e2e > gpurun_out/r2w/ncu.log 2>&1

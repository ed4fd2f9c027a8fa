mkdir -p gpurun_out/r3g
python bench.py --steps 40 --warmup 12 --no-cpu --no-e2e > gpurun_out/r3g/bench.json 2> gpurun_out/r3g/bench.err
python -c "
import json
d=json.load(open('gpurun_out/r3g/bench.json')); r=d['roofline']
print(' ms/step %.3f measured %.3f value %.3e kernel_frac %.3f step_frac %.3f avg_launch_ms %.3f share %.3f'%(d['ms_per_step'],r['ms_per_step_measured'],d['value'],r['frac'],r['step_frac'],r['avg_launch_ms'],r['kernel_share_of_step']))"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r3g/launches.csv python bench.py --steps 10 --warmup 5 --no-cpu --no-e2e > gpurun_out/r3g/ncu.log 2>&1

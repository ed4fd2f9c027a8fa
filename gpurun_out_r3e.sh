mkdir -p gpurun_out/r3e
python bench.py --steps 40 --warmup 12 --no-cpu --no-e2e > gpurun_out/r3e/bench.json 2> gpurun_out/r3e/bench.err
python -c "
import json
d=json.load(open('gpurun_out/r3e/bench.json')); r=d['roofline']
print(' ms/step %.3f measured %.3f value %.3e kernel_frac %.3f step_frac %.3f avg_launch_ms %.3f share %.3f'%(d['ms_per_step'],r['ms_per_step_measured'],d['value'],r['frac'],r['step_frac'],r['avg_launch_ms'],r['kernel_share_of_step']), r['reorder_in_timed_region'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_mcc_(test|collide|select_skip)' --launch-skip 30 -c 6 -o gpurun_out/r3e/ncu_mcc -f python bench.py --steps 6 --warmup 5 --no-cpu --no-e2e > gpurun_out/r3e/ncu_mcc.log 2>&1

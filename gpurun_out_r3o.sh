mkdir -p gpurun_out/r3o
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_field_scale.py tests/test_gpu_mcc.py tests/test_gpu_physics.py -m gpu -q -x > gpurun_out/r3o/pytest.log 2>&1; echo "rc=$?"
tail -4 gpurun_out/r3o/pytest.log
python bench.py --steps 48 --warmup 12 --no-cpu --no-e2e > gpurun_out/r3o/bench.json 2> gpurun_out/r3o/bench.err
python -c "
import json
d=json.load(open('gpurun_out/r3o/bench.json')); r=d['roofline']
print(' ms/step %.3f measured %.3f value %.3e kernel_frac %.3f step_frac %.3f avg_launch_ms %.3f share %.3f'%(d['ms_per_step'],r['ms_per_step_measured'],d['value'],r['frac'],r['step_frac'],r['avg_launch_ms'],r['kernel_share_of_step']), r['reorder_in_timed_region'])"
tail -2 gpurun_out/r3o/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r3o/launches.csv python bench.py --steps 10 --warmup 5 --no-cpu --no-e2e > gpurun_out/r3o/ncu.log 2>&1

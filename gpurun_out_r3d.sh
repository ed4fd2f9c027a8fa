mkdir -p gpurun_out/r3d
timeout 1500 python -m pytest tests/test_gpu_mcc.py tests/test_gpu_physics.py tests/test_gpu_c2_scripted.py -m gpu -q -x > gpurun_out/r3d/pytest.log 2>&1; echo "rc=$?"
tail -5 gpurun_out/r3d/pytest.log
python bench.py --steps 40 --warmup 12 --no-cpu --no-e2e > gpurun_out/r3d/bench.json 2> gpurun_out/r3d/bench.err
python -c "
import json
d=json.load(open('gpurun_out/r3d/bench.json')); r=d['roofline']
print(' ms/step %.3f measured %.3f value %.3e kernel_frac %.3f step_frac %.3f avg_launch_ms %.3f share %.3f'%(d['ms_per_step'],r['ms_per_step_measured'],d['value'],r['frac'],r['step_frac'],r['avg_launch_ms'],r['kernel_share_of_step']), r['reorder_in_timed_region'])"
tail -2 gpurun_out/r3d/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:'k_mcc' -c 60 --csv --log-file gpurun_out/r3d/launches_mcc.csv python bench.py --steps 6 --warmup 4 --no-cpu --no-e2e > gpurun_out/r3d/ncu.log 2>&1

mkdir -p gpurun_out/r2u
timeout 900 python -m pytest tests/test_gpu_tile.py -m gpu -q -x > gpurun_out/r2u/pytest.log 2>&1; echo "rc=$?"
tail -5 gpurun_out/r2u/pytest.log
python bench.py --steps 60 --warmup 8 --no-cpu --no-e2e > gpurun_out/r2u/bench.json 2> gpurun_out/r2u/bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2u/bench.json')); r=d['roofline']
print(' ms/step %.3f measured %.3f value %.3e kernel_frac %.3f step_frac %.3f avg_launch_ms %.3f share %.3f'%(d['ms_per_step'],r['ms_per_step_measured'],d['value'],r['frac'],r['step_frac'],r['avg_launch_ms'],r['kernel_share_of_step']))"
tail -3 gpurun_out/r2u/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/r2u/launches.csv python bench.py --steps 6 --warmup 4 --no-cpu --no-e2e > gpurun_out/r2u/ncu.log 2>&1

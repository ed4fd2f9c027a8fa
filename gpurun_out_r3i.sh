mkdir -p gpurun_out/r3i
timeout 1700 python -m pytest tests -m gpu -q -x > gpurun_out/r3i/pytest_all.log 2>&1; echo "rc=$?"
tail -6 gpurun_out/r3i/pytest_all.log
python bench.py --steps 40 --warmup 12 --no-cpu --no-e2e > gpurun_out/r3i/bench.json 2> gpurun_out/r3i/bench.err
python -c "
import json
d=json.load(open('gpurun_out/r3i/bench.json')); r=d['roofline']
print(' ms/step %.3f measured %.3f value %.3e kernel_frac %.3f step_frac %.3f avg_launch_ms %.3f share %.3f'%(d['ms_per_step'],r['ms_per_step_measured'],d['value'],r['frac'],r['step_frac'],r['avg_launch_ms'],r['kernel_share_of_step']))"
tail -2 gpurun_out/r3i/bench.err

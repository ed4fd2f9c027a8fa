mkdir -p gpurun_out/r2l
timeout 1200 python -m pytest tests/test_gpu_field_scale.py -x -q > gpurun_out/r2l/pytest_a.log 2>&1; echo "a rc=$?"
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "poisson or 100_steps" > gpurun_out/r2l/pytest_b.log 2>&1; echo "b rc=$?"
python bench.py --steps 60 --warmup 8 --no-cpu --no-e2e --sort-full 0 --sort-miss 0.0005 --sort-max 4 > gpurun_out/r2l/bench_lean_x4.json 2> gpurun_out/r2l/bench_lean_x4.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 300 --csv --log-file gpurun_out/r2l/launches.csv python bench.py --steps 10 --warmup 6 --no-cpu --no-e2e --sort-miss 0.0005 --sort-max 4 --sort-full 0 > gpurun_out/r2l/b1.log 2>&1
tail -4 gpurun_out/r2l/pytest_a.log; tail -4 gpurun_out/r2l/pytest_b.log
python -c "
import json
d=json.load(open('gpurun_out/r2l/bench_lean_x4.json')); r=d['roofline']
print(' ms/step %.3f value %.3e kernel_frac %.3f avg_launch_ms %.3f share %.3f'%(d['ms_per_step'],d['value'],r['frac'],r['avg_launch_ms'],r['kernel_share_of_step']))"
python profiles/aggregate_launches.py gpurun_out/r2l/launches.csv | head -16

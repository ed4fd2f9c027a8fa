mkdir -p gpurun_out/r2k
timeout 1200 python -m pytest tests/test_gpu_field_scale.py tests/test_gpu_tile.py tests/test_gpu_mcc.py tests/test_gpu_physics.py -x -q > gpurun_out/r2k/pytest_a.log 2>&1; echo "a rc=$?"
python bench.py --steps 60 --warmup 8 --no-cpu --no-e2e --sort-full 0 --sort-miss 0.0005 --sort-max 4 > gpurun_out/r2k/bench_lean_x4.json 2> gpurun_out/r2k/bench_lean_x4.err
tail -5 gpurun_out/r2k/pytest_a.log
python -c "
import json
d=json.load(open('gpurun_out/r2k/bench_lean_x4.json')); r=d['roofline']
print(' ms/step %.3f value %.3e kernel_frac %.3f avg_launch_ms %.3f share %.3f'%(d['ms_per_step'],d['value'],r['frac'],r['avg_launch_ms'],r['kernel_share_of_step']))"
tail -3 gpurun_out/r2k/bench_lean_x4.err

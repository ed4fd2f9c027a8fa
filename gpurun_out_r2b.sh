mkdir -p gpurun_out/r2b
timeout 600 python -m pytest tests/test_gpu_tile.py -x -q > gpurun_out/r2b/pytest_tile.log 2>&1; echo "tile rc=$?"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "fused or tiled or sort or deposit" > gpurun_out/r2b/pytest_parity.log 2>&1; echo "parity rc=$?"
python bench.py --steps 64 --warmup 8 --no-cpu --no-e2e --advance-path 1 > gpurun_out/r2b/bench_path1.json 2> gpurun_out/r2b/bench_path1.err
python bench.py --steps 64 --warmup 8 --no-cpu --no-e2e --sort-miss 0.002 --sort-max 8 --sort-full 0 > gpurun_out/r2b/bench_m002_x8.json 2> gpurun_out/r2b/bench_m002_x8.err
python bench.py --steps 64 --warmup 8 --no-cpu --no-e2e --sort-miss 0.002 --sort-max 4 --sort-full 0 > gpurun_out/r2b/bench_m002_x4.json 2> gpurun_out/r2b/bench_m002_x4.err
python bench.py --steps 64 --warmup 8 --no-cpu --no-e2e --sort-miss 0.01 --sort-max 16 --sort-full 0 > gpurun_out/r2b/bench_m01_x16.json 2> gpurun_out/r2b/bench_m01_x16.err
python bench.py --steps 64 --warmup 8 --no-cpu --no-e2e --sort-miss 0 --sort-interval 2 --sort-full 0 > gpurun_out/r2b/bench_fix2.json 2> gpurun_out/r2b/bench_fix2.err
tail -5 gpurun_out/r2b/pytest_tile.log; tail -5 gpurun_out/r2b/pytest_parity.log
for f in gpurun_out/r2b/bench_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.load(open('$f')); r=d['roofline']
    print(' ms/step %.3f value %.3e kernel_frac %.3f avg_launch_ms %.3f share %.3f'%(d['ms_per_step'],d['value'],r['frac'],r['avg_launch_ms'],r['kernel_share_of_step']), r.get('window_stats(gather_miss,deposit_miss,moves,rounds)'))
except Exception as e: print(' failed',e)
"; done
tail -3 gpurun_out/r2b/*.err

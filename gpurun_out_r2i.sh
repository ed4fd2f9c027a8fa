mkdir -p gpurun_out/r2i
timeout 900 python -m pytest tests/test_gpu_tile.py -x -q > gpurun_out/r2i/pytest_a.log 2>&1; echo "tile rc=$?"
python bench.py --steps 60 --warmup 8 --no-cpu --no-e2e --sort-full 0 --sort-miss 0.0005 --sort-max 4 > gpurun_out/r2i/bench_lean_x4.json 2> gpurun_out/r2i/bench_lean_x4.err
ncu --metrics gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum --clock-control none -k regex:"k_advance_tile" -s 10 -c 10 --csv --log-file gpurun_out/r2i/launches.csv python bench.py --steps 8 --warmup 6 --no-cpu --no-e2e --sort-miss 0.0005 --sort-max 4 --sort-full 0 > gpurun_out/r2i/b1.log 2>&1
tail -2 gpurun_out/r2i/pytest_a.log
python -c "
import json
d=json.load(open('gpurun_out/r2i/bench_lean_x4.json')); r=d['roofline']
print(' ms/step %.3f value %.3e kernel_frac %.3f avg_launch_ms %.3f share %.3f'%(d['ms_per_step'],d['value'],r['frac'],r['avg_launch_ms'],r['kernel_share_of_step']))"
grep k_advance_tile gpurun_out/r2i/launches.csv | awk -F'","' '{printf "%s ", $NF}' | tr -d '"'

mkdir -p gpurun_out/r2h
timeout 900 python -m pytest tests/test_gpu_tile.py -x -q > gpurun_out/r2h/pytest_a.log 2>&1; echo "tile rc=$?"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "fused or tiled" > gpurun_out/r2h/pytest_b.log 2>&1; echo "parity rc=$?"
for cfg in "lean_x4:--sort-miss 0.0005 --sort-max 4" "full_x4:--sort-miss 0.0005 --sort-max 4 --no-lean"; do
  name=${cfg%%:*}; args=${cfg#*:}
  python bench.py --steps 60 --warmup 8 --no-cpu --no-e2e --sort-full 0 $args > gpurun_out/r2h/bench_$name.json 2> gpurun_out/r2h/bench_$name.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_advance" -s 20 -c 20 --csv --log-file gpurun_out/r2h/launches.csv python bench.py --steps 8 --warmup 6 --no-cpu --no-e2e --sort-miss 0.0005 --sort-max 4 --sort-full 0 > gpurun_out/r2h/b1.log 2>&1
tail -3 gpurun_out/r2h/pytest_a.log; tail -3 gpurun_out/r2h/pytest_b.log
for f in gpurun_out/r2h/bench_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.load(open('$f')); r=d['roofline']
    print(' ms/step %.3f value %.3e kernel_frac %.3f avg_launch_ms %.3f share %.3f'%(d['ms_per_step'],d['value'],r['frac'],r['avg_launch_ms'],r['kernel_share_of_step']), r.get('window_stats(gather_miss,deposit_miss,moves,rounds)'))
except Exception as e: print(' failed',e)
"; done
grep k_advance_tile gpurun_out/r2h/launches.csv | awk -F'","' '{printf "%s ", $NF}' | tr -d '"'

mkdir -p gpurun_out/r3j
timeout 900 python -m pytest tests/test_gpu_tile.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r3j/pytest.log 2>&1; echo "rc=$?"
tail -4 gpurun_out/r3j/pytest.log
python bench.py --steps 40 --warmup 12 --no-cpu --no-e2e > gpurun_out/r3j/bench.json 2> gpurun_out/r3j/bench.err
python -c "
import json
d=json.load(open('gpurun_out/r3j/bench.json')); r=d['roofline']
print(' ms/step %.3f measured %.3f value %.3e kernel_frac %.3f step_frac %.3f avg_launch_ms %.3f share %.3f'%(d['ms_per_step'],r['ms_per_step_measured'],d['value'],r['frac'],r['step_frac'],r['avg_launch_ms'],r['kernel_share_of_step']))"
tail -2 gpurun_out/r3j/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_advance_tile' -c 40 --csv --log-file gpurun_out/r3j/launches_tile.csv python bench.py --steps 10 --warmup 6 --no-cpu --no-e2e > gpurun_out/r3j/ncu.log 2>&1
grep -v "^==" gpurun_out/r3j/launches_tile.csv | awk -F'","' '{printf "%s ", $NF}' | tr -d '"' | tr '\n' ' '

mkdir -p gpurun_out/r3u
timeout 900 python -m pytest tests/test_gpu_mcc.py tests/test_gpu_physics.py tests/test_gpu_c2_scripted.py -m gpu -q -x > gpurun_out/r3u/pytest.log 2>&1; echo "rc=$?"
tail -3 gpurun_out/r3u/pytest.log
python bench.py --steps 48 --warmup 8 --no-cpu --no-e2e > gpurun_out/r3u/bench.json 2> gpurun_out/r3u/bench.err
python -c "
import json
d=json.load(open('gpurun_out/r3u/bench.json')); r=d['roofline']
print(' ms/step %.3f measured %.3f value %.3e kernel_frac %.3f step_frac %.3f'%(d['ms_per_step'],r['ms_per_step_measured'],d['value'],r['frac'],r['step_frac']))"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_mcc_collide' --launch-skip 8 -c 4 --csv --log-file gpurun_out/r3u/coll.csv python bench.py --steps 4 --warmup 4 --no-cpu --no-e2e > gpurun_out/r3u/ncu.log 2>&1
grep -v "^==" gpurun_out/r3u/coll.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' '

mkdir -p gpurun_out/r3p
python bench.py > gpurun_out/r3p/bench_default.json 2> gpurun_out/r3p/bench_default.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r3p/bench_default.json')); r=d['roofline']
print(' ms/step %.3f measured %.3f value %.3e kernel_frac %.3f step_frac %.3f e2e %.3e cpu %s clocks %s launches %s'%(d['ms_per_step'],r['ms_per_step_measured'],d['value'],r['frac'],r['step_frac'],d['e2e']['value'],d['cpu_baseline'],d['clocks'],d['gpu_launches']))"
python bench.py --impl reference > gpurun_out/r3p/bench_ref.json 2> gpurun_out/r3p/bench_ref.err; echo "ref rc=$?"
tail -c 600 gpurun_out/r3p/bench_ref.json
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r3p/pytest_all.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r3p/pytest_all.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r3p/smoke.log 2>&1; tail -2 gpurun_out/r3p/smoke.log

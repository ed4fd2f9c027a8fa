mkdir -p gpurun_out/r3q
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tile.py -m gpu -q -x > gpurun_out/r3q/pytest.log 2>&1; echo "rc=$?"
tail -3 gpurun_out/r3q/pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 48 --warmup 8 --no-cpu > gpurun_out/r3q/bench_n2.json 2> gpurun_out/r3q/bench_n2.err
python -c "
import json
d=json.load(open('gpurun_out/r3q/bench_n2.json')); r=d['roofline']
print('N=2 ms/step %.3f measured %.3f value %.3e kernel_frac %.3f step_frac %.3f e2e %.3e'%(d['ms_per_step'],r['ms_per_step_measured'],d['value'],r['frac'],r['step_frac'],d['e2e']['value']))"

mkdir -p gpurun_out/r3s
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 48 --warmup 8 --no-cpu > gpurun_out/r3s/bench_n8.json 2> gpurun_out/r3s/bench_n8.err; echo "rc=$?"
wc -c gpurun_out/r3s/bench_n8.json; tail -5 gpurun_out/r3s/bench_n8.err | cut -c1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 48 --warmup 8 --no-cpu > gpurun_out/r3s/bench_n4.json 2> gpurun_out/r3s/bench_n4.err; echo "rc=$?"

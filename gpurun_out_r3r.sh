mkdir -p gpurun_out/r3r
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 48 --warmup 8 --no-cpu > gpurun_out/r3r/bench_n2.json 2> gpurun_out/r3r/bench_n2.err; echo "rc=$?"
wc -c gpurun_out/r3r/bench_n2.json; tail -30 gpurun_out/r3r/bench_n2.err | cut -c1-400

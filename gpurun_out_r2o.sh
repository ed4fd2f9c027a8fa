mkdir -p gpurun_out/r2o
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/r2o/pytest_all.log 2>&1; echo "all rc=$?"
tail -25 gpurun_out/r2o/pytest_all.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2o/bench_default.json 2> gpurun_out/r2o/bench_default.err; tail -c 2500 gpurun_out/r2o/bench_default.json; tail -5 gpurun_out/r2o/bench_default.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2o/bench_ref.json 2> gpurun_out/r2o/bench_ref.err; tail -c 600 gpurun_out/r2o/bench_ref.json

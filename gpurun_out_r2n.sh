mkdir -p gpurun_out/r2n
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2n/pytest_all.log 2>&1; echo "all rc=$?"
tail -15 gpurun_out/r2n/pytest_all.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2n/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2n/smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2n/bench_default.json 2> gpurun_out/r2n/bench_default.err; tail -c 1500 gpurun_out/r2n/bench_default.json

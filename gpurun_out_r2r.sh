mkdir -p gpurun_out/r2r
timeout 900 python -m pytest tests/test_gpu_see.py -x -q > gpurun_out/r2r/pytest_see.log 2>&1; echo "see rc=$?"
tail -30 gpurun_out/r2r/pytest_see.log
timeout 900 python -m pytest tests/test_gpu_mcc.py -x -q > gpurun_out/r2r/pytest_mcc.log 2>&1; echo "mcc rc=$?"
tail -5 gpurun_out/r2r/pytest_mcc.log

mkdir -p gpurun_out/r2p
timeout 1700 python -m pytest tests/test_gpu_c2_scripted.py tests/test_gpu_tile.py tests/test_axial.py -m gpu -q > gpurun_out/r2p/pytest.log 2>&1; echo "rc=$?"
tail -30 gpurun_out/r2p/pytest.log

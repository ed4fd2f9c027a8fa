mkdir -p gpurun_out/r3v
timeout 1700 python -m pytest tests/test_gpu_field_scale.py tests/test_axial.py tests/test_gpu_surfaces.py tests/test_problem_scripts.py -m gpu -q -x -s > gpurun_out/r3v/pytest.log 2>&1; echo "rc=$?"
grep -n "dense inverse\|passed\|failed\|Error" gpurun_out/r3v/pytest.log | tail -8

mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_composite.py tests/test_gpu_derivs.py tests/test_gpu_td.py tests/test_gpu_iterative.py -m gpu -q > gpurun_out/r2n_pytest.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/r2n_pytest.log
grep -E "^E  |FAILED|passed|failed|pytest_rc" gpurun_out/r2n_pytest.log | head -30

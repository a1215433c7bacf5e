mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_iterative.py tests/test_gpu_derivs.py tests/test_gpu_td.py -m gpu -q -k "sgpr_iterative or sgpr_cg or td_iterative" > gpurun_out/r2g_pytest.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/r2g_pytest.log
grep -E "^E  |FAILED|passed|failed|pytest_rc" gpurun_out/r2g_pytest.log | head -40

mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/r2e_pytest.log
grep -E "FAILED|passed|failed|pytest_rc" gpurun_out/r2e_pytest.log | head -30

mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2m_pytest.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/r2m_pytest.log
grep -E "^E  |FAILED|passed|failed|pytest_rc" gpurun_out/r2m_pytest.log | head -30
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2

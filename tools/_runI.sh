mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_edges.py -m gpu -q -x > gpurun_out/r2i_pytest.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/r2i_pytest.log
grep -E "^E  |FAILED|passed|failed|pytest_rc" gpurun_out/r2i_pytest.log | head -20
timeout 300 python tools/bench_small.py > gpurun_out/r2i_small_v9.json 2> gpurun_out/r2i_small.err; cat gpurun_out/r2i_small_v9.json

mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_edges.py tests/test_gpu_golden.py -m gpu -q -x > gpurun_out/r2h_pytest.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/r2h_pytest.log
grep -E "^E  |FAILED|passed|failed|pytest_rc" gpurun_out/r2h_pytest.log | head -20
timeout 300 python tools/bench_small.py > gpurun_out/r2h_small_v2.json 2> gpurun_out/r2h_small.err; cat gpurun_out/r2h_small_v2.json
GPXB_TRSV=v1 timeout 300 python tools/bench_small.py > gpurun_out/r2h_small_v1.json 2>> gpurun_out/r2h_small.err; cat gpurun_out/r2h_small_v1.json

mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2p_pytest.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/r2p_pytest.log
grep -E "^E  |FAILED|passed|failed|pytest_rc" gpurun_out/r2p_pytest.log | head -30
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2p_bench_1.json 2> gpurun_out/r2p_bench_1.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2p_bench_1.err

mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2l_bench_1.json 2> gpurun_out/r2l_bench_1.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2l_bench_1.err
GPXB_BENCH_ITER_N=1000000 GPXB_BENCH_ONLY=lml_iter_full timeout 900 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/r2l_bench_1_iter1M.json 2> gpurun_out/r2l_bench_1_iter1M.err; echo "bench rc=$?"

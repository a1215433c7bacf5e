mkdir -p gpurun_out
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu --no-extra > gpurun_out/r2t_plain.json 2>/dev/null; echo "plain rc=$?"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_bench_launches_N32768.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-extra > gpurun_out/r2t_ncu.json 2>gpurun_out/r2t_ncu.err; echo "ncu rc=$?"
wc -l gpurun_out/r2_bench_launches_N32768.csv; gzip -f gpurun_out/r2_bench_launches_N32768.csv; ls -la gpurun_out/r2_bench_launches_N32768.csv.gz

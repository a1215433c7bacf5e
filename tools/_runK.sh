mkdir -p gpurun_out
NG=${NG:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/multigpu_check.py > gpurun_out/r2k_mg_${NG}.json 2> gpurun_out/r2k_mg_${NG}.err; echo "mg rc=$?"; grep -c '"ok": true' gpurun_out/r2k_mg_${NG}.json; head -c 700 gpurun_out/r2k_mg_${NG}.json
GPXB_BENCH_ITER_N=1000000 timeout 900 $TR bench.py --gpus $NG --steps 2 --warmup 3 --no-cpu > gpurun_out/r2k_bench_${NG}.json 2> gpurun_out/r2k_bench_${NG}.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2k_bench_${NG}.err

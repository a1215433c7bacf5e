mkdir -p gpurun_out
NG=4
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $NG --steps 3 --warmup 3 > gpurun_out/r2r_bench_${NG}.json 2> gpurun_out/r2r_bench_${NG}.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r2r_bench_${NG}.err
timeout 300 $TR tools/multigpu_check.py > gpurun_out/r2r_mg_${NG}.json 2> gpurun_out/r2r_mg_${NG}.err; echo "mg rc=$?"; grep -c '"ok": true' gpurun_out/r2r_mg_${NG}.json

mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_iterative.py -m gpu -x -q -k "one_sweep or lml_iter" > gpurun_out/r2c_pytest.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/r2c_pytest.log
tail -5 gpurun_out/r2c_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/multigpu_check.py > gpurun_out/r2c_mg_peer.json 2> gpurun_out/r2c_mg_peer.err; echo "rc=$?"; tail -2 gpurun_out/r2c_mg_peer.json
GPXB_PEER=0 timeout 300 $TR tools/multigpu_check.py > gpurun_out/r2c_mg_nccl.json 2> gpurun_out/r2c_mg_nccl.err; echo "rc=$?"; tail -2 gpurun_out/r2c_mg_nccl.json
timeout 300 $TR tools/bench_sharded.py 0 0 1000000 4096 > gpurun_out/r2c_rp2_peer.json 2> gpurun_out/r2c_rp2_peer.err; tail -1 gpurun_out/r2c_rp2_peer.json
GPXB_PEER=0 timeout 300 $TR tools/bench_sharded.py 0 0 1000000 4096 > gpurun_out/r2c_rp2_nccl.json 2> gpurun_out/r2c_rp2_nccl.err; tail -1 gpurun_out/r2c_rp2_nccl.json
GPXB_RP_COLUMN=v1 GPXB_RP_MINBLOCKS=1 timeout 300 $TR tools/bench_sharded.py 0 0 1000000 4096 > gpurun_out/r2c_rp2_v1m1.json 2> gpurun_out/r2c_rp2_v1m1.err; tail -1 gpurun_out/r2c_rp2_v1m1.json
GPXB_BENCH_ITER_N=131072 timeout 600 $TR bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu > gpurun_out/r2c_bench2.json 2> gpurun_out/r2c_bench2.err; echo "bench rc=$?"

mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,GRAPH,P2P timeout 200 $TR tools/nccl_probe.py > gpurun_out/r2d_nccl_probe.log 2>&1; echo rc=$?
grep -E "^\{|via|NVLS|P2P|SHM|NET/" gpurun_out/r2d_nccl_probe.log | head -40
nvidia-smi topo -m > gpurun_out/r2d_topo.txt 2>&1; head -8 gpurun_out/r2d_topo.txt
timeout 600 python -m pytest tests/test_gpu_iterative.py -m gpu -x -q -k "1e8 or one_sweep" 2>&1 | tail -5

cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export NCCL_DEBUG=WARN
echo "== overlap, eager, launch blocking"
S2S_GRAPHS=0 CUDA_LAUNCH_BLOCKING=1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 3 --warmup 3 --no-variants > gpurun_out/job9_a.out 2> gpurun_out/job9_a.err; echo rc=$?; grep -m6 -i "S2SError\|misaligned\|illegal\|NCCL WARN" gpurun_out/job9_a.err | cut -c1-300; cut -c1-200 gpurun_out/job9_a.out
echo "== overlap, eager"
S2S_GRAPHS=0 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 3 --warmup 3 --no-variants > gpurun_out/job9_b.out 2> gpurun_out/job9_b.err; echo rc=$?; grep -m6 -i "S2SError\|misaligned\|illegal\|NCCL WARN" gpurun_out/job9_b.err | cut -c1-300; cut -c1-200 gpurun_out/job9_b.out
echo "== dp test (plain + overlap)"
timeout 600 python -m pytest tests/test_gpu_dp_nccl.py -x -q 2>&1 | tail -12 | cut -c1-300

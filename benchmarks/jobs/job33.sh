cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; export PYTHONPATH=$GRAFT_REPO_ROOT
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 20 --warmup 3 --no-variants 2>gpurun_out/job33.err | tee gpurun_out/r02_bench_n8.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=8', round(d['ms_per_step'],3), round(d['value']), d['config']['collective'][:60])"
grep -m3 -i "S2SError\|misaligned\|illegal\|NCCL error" gpurun_out/job33.err | cut -c1-300

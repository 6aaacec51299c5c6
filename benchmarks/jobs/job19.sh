cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; export PYTHONPATH=$GRAFT_REPO_ROOT
echo "== N=8 overlap inside the replayed graph"
S2S_BENCH_DP_OVERLAP=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 3 --no-variants 2>gpurun_out/job19_a.err | tee gpurun_out/r02_bench_n8_overlap.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=8 overlap', round(d['ms_per_step'],3), round(d['value']), d['config']['collective'][:60])"
grep -m3 -i "S2SError\|misaligned\|illegal\|NCCL error" gpurun_out/job19_a.err | cut -c1-300
echo "== N=8 default (plain all-reduce after the graph), with variants"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 20 --warmup 3 2>gpurun_out/job19_b.err | tee gpurun_out/r02_bench_n8.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=8', round(d['ms_per_step'],3), round(d['value']), d['config']['collective'][:50], {k:(round(v['ms_per_step'],3), round(v['value'])) for k,v in d['variants'].items()})"
grep -m3 -i "S2SError\|misaligned\|illegal\|NCCL error" gpurun_out/job19_b.err | cut -c1-300

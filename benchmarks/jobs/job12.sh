cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
echo "== dp test (plain + overlap eager)"
timeout 500 python -m pytest tests/test_gpu_dp_nccl.py -x -q 2>&1 | tail -6 | cut -c1-300
echo "== N=2 default (plain all-reduce after the graph), with variants"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/job12_a.err | tee gpurun_out/r02_bench_n2.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=2', round(d['ms_per_step'],3), round(d['value']), d['config']['collective'][:50], {k:(round(v['ms_per_step'],3), round(v['value'])) for k,v in d['variants'].items()})"
grep -m3 -i "S2SError\|misaligned\|illegal" gpurun_out/job12_a.err | cut -c1-300
echo "== N=2 overlap inside the replayed graph"
S2S_BENCH_DP_OVERLAP=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 10 --warmup 3 --no-variants 2>gpurun_out/job12_b.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=2 overlap', round(d['ms_per_step'],3), round(d['value']), d['config']['collective'][:60])"
grep -m3 -i "S2SError\|misaligned\|illegal" gpurun_out/job12_b.err | cut -c1-300

cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; export PYTHONPATH=$GRAFT_REPO_ROOT
echo "== dp test (plain, eager buckets, graph buckets)"
timeout 600 python -m pytest tests/test_gpu_dp_nccl.py -q 2>&1 | tail -4 | cut -c1-300
echo "== N=2 default (buckets inside the replayed graph), with variants"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 20 --warmup 3 2>gpurun_out/job26_a.err | tee gpurun_out/r02_bench_n2.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=2', round(d['ms_per_step'],3), round(d['value']), d['config']['collective'][:50], {k:(round(v['ms_per_step'],3), round(v['value'])) for k,v in d['variants'].items()})"
grep -m3 -i "S2SError\|misaligned\|illegal\|NCCL error" gpurun_out/job26_a.err | cut -c1-300
echo "== N=1 final"
timeout 600 python bench.py --steps 20 --warmup 3 2>gpurun_out/job26_n1.err | tee gpurun_out/r02_bench_n1.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=1', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['kernel'], {k:(round(v['ms_per_step'],3), v['launches_per_step']) for k,v in d['kernels'].items()}, {k:v.get('fp32_fma',{}).get('frac') for k,v in d['kernels'].items() if 'gru' in k}, {k:(round(v['ms_per_step'],3), round(v['value'])) for k,v in d['variants'].items()})"
tail -3 gpurun_out/job26_n1.err
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-400

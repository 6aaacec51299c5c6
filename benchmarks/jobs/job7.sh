cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_dp_nccl.py -x -q 2>&1 | tail -15
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/job7_n2.err | tee gpurun_out/r02_bench_n2.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=2', d['ms_per_step'], d['value'], d['config']['collective'][:40], {k:(round(v['ms_per_step'],3), round(v['value'])) for k,v in d['variants'].items()})"
tail -5 gpurun_out/job7_n2.err
S2S_BENCH_DP_PLAIN=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-variants 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=2 plain allreduce', d['ms_per_step'], d['value'])"
timeout 600 python bench.py --steps 10 --warmup 3 2>gpurun_out/job7_n1.err | tee gpurun_out/r02_bench_n1.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=1', d['ms_per_step'], d['value'], d['e2e']['value'], d['cpu_baseline'], {k:(round(v['ms_per_step'],3), round(v['value'])) for k,v in d['variants'].items()})"
tail -3 gpurun_out/job7_n1.err

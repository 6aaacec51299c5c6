cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
S2S_GRU_PROF=1 timeout 120 python benchmarks/gru_micro.py 32 300 256 512 2 2>&1 | tail -8
S2S_GRU_PROF=1 S2S_GRU_DBG=4 timeout 120 python benchmarks/gru_micro.py 32 300 256 512 2 2>&1 | tail -8
timeout 900 python -m pytest tests/test_gpu_timed_path.py -x -q 2>&1 | tail -5
for c in cfg2 cfg2loc; do timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config $c 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['workload'][:8], d['ms_per_step'], {k:(round(v['ms_per_step'],3), v['launches_per_step']) for k,v in d['kernels'].items()})"; done

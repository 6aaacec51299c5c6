cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; export PYTHONPATH=$GRAFT_REPO_ROOT
echo "== gpu suite"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 | cut -c1-300
echo "== bench"
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tee gpurun_out/job22_bench.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['workload'][:8], d['ms_per_step'], d['e2e']['ms_per_step'], {k:(round(v['ms_per_step'],3), v['launches_per_step']) for k,v in d['kernels'].items()}, {k:round(v['ms_per_step'],3) for k,v in d['variants'].items()})"
echo "== bench OVERLAP=0"
S2S_OVERLAP=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-variants 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['workload'][:8], d['ms_per_step'])"

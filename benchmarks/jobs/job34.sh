cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; export PYTHONPATH=$GRAFT_REPO_ROOT
echo "== timed path"; timeout 600 python -m pytest tests/test_gpu_timed_path.py -x -q 2>&1 | tail -2 | cut -c1-300
echo "== bench (loc taps unroll 5)"
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-variants 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['workload'][:8], d['ms_per_step'], {k:(round(v['ms_per_step'],3), v['launches_per_step']) for k,v in d['kernels'].items()})"

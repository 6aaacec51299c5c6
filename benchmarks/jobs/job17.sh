cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; export PYTHONPATH=$GRAFT_REPO_ROOT
echo "== full gpu suite"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 | cut -c1-300
echo "== bench N=1"
timeout 600 python bench.py --steps 20 --warmup 3 2>gpurun_out/job17_n1.err | tee gpurun_out/r02_bench_n1.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=1', d['ms_per_step'], d['value'], d['e2e']['value'], {k:(round(v['ms_per_step'],3), v['launches_per_step']) for k,v in d['kernels'].items()}, {k:(round(v['ms_per_step'],3), round(v['value'])) for k,v in d['variants'].items()})"
tail -3 gpurun_out/job17_n1.err
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2

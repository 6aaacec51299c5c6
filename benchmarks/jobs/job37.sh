cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; export PYTHONPATH=$GRAFT_REPO_ROOT
echo "== full gpu suite"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 | cut -c1-300
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
echo "== bench N=1 (defaults)"
timeout 600 python bench.py 2>gpurun_out/job37_n1.err | tee gpurun_out/r02_bench_n1_final.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=1', d['steps'], d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['kernel'], {k:(round(v['ms_per_step'],3), v['launches_per_step']) for k,v in d['kernels'].items()}, {k:(round(v['ms_per_step'],3), round(v['value'])) for k,v in d['variants'].items()})"
tail -2 gpurun_out/job37_n1.err

cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; export PYTHONPATH=$GRAFT_REPO_ROOT
echo "== full gpu suite"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 | cut -c1-300
echo "== lstm micro"; timeout 200 python benchmarks/lstm_micro.py 2>&1 | tail -8 | cut -c1-200
echo "== gru micro"; timeout 120 python benchmarks/gru_micro.py 2>&1 | tail -3
echo "== dec micro"; timeout 200 python benchmarks/dec_micro.py 2>&1 | tail -3 | cut -c1-200
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== bench N=1 (defaults)"
timeout 600 python bench.py 2>gpurun_out/job32_n1.err | tee gpurun_out/r02_bench_n1_final.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=1', d['steps'], d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['kernel'], {k:(round(v['ms_per_step'],3), v['launches_per_step']) for k,v in d['kernels'].items()}, {k:v.get('fp32_fma',{}).get('frac') for k,v in d['kernels'].items() if 'gru' in k}, {k:(round(v['ms_per_step'],3), round(v['value'])) for k,v in d['variants'].items()})"
tail -3 gpurun_out/job32_n1.err
echo "== cfg4"
timeout 600 python bench.py --config cfg4 --steps 10 --warmup 3 --no-cpu-baseline --no-variants 2>/dev/null | tee gpurun_out/r02_bench_n1_cfg4.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('cfg4', d['ms_per_step'], d['value'], d['e2e']['value'])"

cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; export PYTHONPATH=$GRAFT_REPO_ROOT
echo "== full gpu suite"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 | cut -c1-300
echo "== bench cfg4 (VGG)"
timeout 600 python bench.py --config cfg4 --steps 10 --warmup 3 --no-cpu-baseline --no-variants 2>gpurun_out/job18_c4.err | tee gpurun_out/r02_bench_n1_cfg4.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('cfg4', d['ms_per_step'], d['value'], d['e2e']['value'])"
tail -3 gpurun_out/job18_c4.err
echo "== gru micro"; timeout 200 python benchmarks/gru_micro.py 2>&1 | tail -3
echo "== dec micro"; timeout 200 python benchmarks/dec_micro.py 2>&1 | tail -6

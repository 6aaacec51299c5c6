cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; export PYTHONPATH=$GRAFT_REPO_ROOT
echo "== full gpu suite"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 | cut -c1-300
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== bench N=1 (defaults)"
timeout 600 python bench.py 2>gpurun_out/job28_n1.err | tee gpurun_out/r02_bench_n1_final.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=1', d['steps'], d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['kernel'], d['gpu_launches'], d['clocks'])"
tail -3 gpurun_out/job28_n1.err
echo "== ZP2: GRU tests"; S2S_GRU_ZP2=1 timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_nn.py tests/test_gpu_timed_path.py -x -q -k "gru or rnn or RNN or GRU or timed_configuration" 2>&1 | tail -3 | cut -c1-300
echo "== ZP2 micro"; S2S_GRU_ZP2=1 timeout 120 python benchmarks/gru_micro.py 2>&1 | tail -3
echo "== ZP2 micro B=28"; S2S_GRU_ZP2=1 timeout 120 python benchmarks/gru_micro.py 28 2>&1 | tail -3
echo "== ZP2 micro H=128"; S2S_GRU_ZP2=1 timeout 120 python benchmarks/gru_micro.py 32 300 128 256 2>&1 | tail -3
echo "== ZP2 bench"
S2S_GRU_ZP2=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['workload'][:8], d['ms_per_step'], {k:(round(v['ms_per_step'],3), v['launches_per_step']) for k,v in d['kernels'].items()}, {k:round(v['ms_per_step'],3) for k,v in d['variants'].items()})"

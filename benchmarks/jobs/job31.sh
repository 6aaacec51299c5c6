cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; export PYTHONPATH=$GRAFT_REPO_ROOT
echo "== GRU tests"; timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_nn.py -x -q -k "gru or rnn or RNN or GRU" 2>&1 | tail -2 | cut -c1-300
echo "== micro"; timeout 120 python benchmarks/gru_micro.py 2>&1 | tail -3
echo "== micro B=28"; timeout 120 python benchmarks/gru_micro.py 28 2>&1 | tail -3
echo "== micro H=128"; timeout 120 python benchmarks/gru_micro.py 32 300 128 256 2>&1 | tail -3
echo "== gru fwd trace"; S2S_GRU_TRACE=1 timeout 120 python benchmarks/gru_micro.py 2>&1 | grep "gru trace" | head -2 | cut -c1-1200
echo "== gru bwd trace"; S2S_GRU_TRACE=2 timeout 120 python benchmarks/gru_micro.py 2>&1 | grep "bwd trace" | head -2 | cut -c1-1200
echo "== timed path + model"; timeout 900 python -m pytest tests/test_gpu_timed_path.py tests/test_gpu_model.py -x -q 2>&1 | tail -2 | cut -c1-300
echo "== bench"
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['workload'][:8], d['ms_per_step'], {k:(round(v['ms_per_step'],3), v['launches_per_step']) for k,v in d['kernels'].items()}, {k:round(v['ms_per_step'],3) for k,v in d['variants'].items()})"

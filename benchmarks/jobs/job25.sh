cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; export PYTHONPATH=$GRAFT_REPO_ROOT; export S2S_GRU_GEN=5
echo "== GRU tests gen5"; timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_nn.py -x -q -k "gru or rnn or RNN or GRU" 2>&1 | tail -3 | cut -c1-300
echo "== micro gen5"; timeout 120 python benchmarks/gru_micro.py 2>&1 | tail -3
echo "== micro gen5 B=28"; timeout 120 python benchmarks/gru_micro.py 28 2>&1 | tail -3
echo "== micro gen5 B=16"; timeout 120 python benchmarks/gru_micro.py 16 2>&1 | tail -3
echo "== micro gen5 H=128"; timeout 120 python benchmarks/gru_micro.py 32 300 128 256 2>&1 | tail -3
echo "== trace gen5"; S2S_GRU_TRACE=1 timeout 120 python benchmarks/gru_micro.py 2>&1 | grep "trace" | head -3
echo "== timed path + model tests gen5"; timeout 900 python -m pytest tests/test_gpu_timed_path.py tests/test_gpu_model.py -x -q 2>&1 | tail -3 | cut -c1-300
echo "== bench gen5"
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['workload'][:8], d['ms_per_step'], {k:(round(v['ms_per_step'],3), v['launches_per_step']) for k,v in d['kernels'].items()}, {k:round(v['ms_per_step'],3) for k,v in d['variants'].items()})"

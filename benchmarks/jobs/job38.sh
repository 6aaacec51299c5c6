cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; export PYTHONPATH=$GRAFT_REPO_ROOT
echo "== GRU tests"; timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_nn.py -x -q -k "gru or rnn or RNN or GRU" 2>&1 | tail -2 | cut -c1-300
echo "== micro"; timeout 120 python benchmarks/gru_micro.py 2>&1 | tail -3
echo "== micro B=28"; timeout 120 python benchmarks/gru_micro.py 28 2>&1 | tail -3
echo "== micro H=128"; timeout 120 python benchmarks/gru_micro.py 32 300 128 256 2>&1 | tail -3

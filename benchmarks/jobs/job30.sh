cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; export PYTHONPATH=$GRAFT_REPO_ROOT
echo "== GRU tests"; timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_nn.py -x -q -k "gru or rnn or RNN or GRU" 2>&1 | tail -2 | cut -c1-300
echo "== micro"; timeout 120 python benchmarks/gru_micro.py 2>&1 | tail -3
echo "== gru bwd trace"; S2S_GRU_TRACE=2 timeout 120 python benchmarks/gru_micro.py 2>&1 | grep "bwd trace" | head -2 | cut -c1-1200

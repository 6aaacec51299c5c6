cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; export PYTHONPATH=$GRAFT_REPO_ROOT
echo "== GRU tests bwd gen5"; S2S_GRU_GEN_BWD=5 timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_nn.py -x -q -k "gru or rnn or RNN or GRU" 2>&1 | tail -2 | cut -c1-300
echo "== micro bwd gen5"; S2S_GRU_GEN_BWD=5 timeout 120 python benchmarks/gru_micro.py 2>&1 | tail -3
echo "== micro bwd gen5 B=28"; S2S_GRU_GEN_BWD=5 timeout 120 python benchmarks/gru_micro.py 28 2>&1 | tail -3
echo "== micro bwd gen5 H=128"; S2S_GRU_GEN_BWD=5 timeout 120 python benchmarks/gru_micro.py 32 300 128 256 2>&1 | tail -3

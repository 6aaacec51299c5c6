cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; export PYTHONPATH=$GRAFT_REPO_ROOT
echo "== GRU tests owners-first"; S2S_GRU_DBG=8 timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_nn.py -x -q -k "gru or rnn or RNN or GRU" 2>&1 | tail -2 | cut -c1-300
echo "== micro default"; timeout 120 python benchmarks/gru_micro.py 2>&1 | tail -3
echo "== micro owners-first"; S2S_GRU_DBG=8 timeout 120 python benchmarks/gru_micro.py 2>&1 | tail -3
echo "== micro owners-first H=128"; S2S_GRU_DBG=8 timeout 120 python benchmarks/gru_micro.py 32 300 128 256 2>&1 | tail -3

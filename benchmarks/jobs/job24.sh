cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; export PYTHONPATH=$GRAFT_REPO_ROOT
echo "== GRU tests gen3 (split accumulators)"; timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_nn.py -x -q -k "gru or rnn or RNN or GRU" 2>&1 | tail -3 | cut -c1-300
echo "== GRU tests gen3, no fence"; S2S_GRU_DBG=4 timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_nn.py -x -q -k "gru or rnn or RNN or GRU" 2>&1 | tail -3 | cut -c1-300
echo "== micro gen3"; timeout 120 python benchmarks/gru_micro.py 2>&1 | tail -3
echo "== micro gen3 no fence"; S2S_GRU_DBG=4 timeout 120 python benchmarks/gru_micro.py 2>&1 | tail -3
echo "== micro gen3 no fence B=28"; S2S_GRU_DBG=4 timeout 120 python benchmarks/gru_micro.py 28 2>&1 | tail -3
echo "== micro gen3 no fence H=128"; S2S_GRU_DBG=4 timeout 120 python benchmarks/gru_micro.py 32 300 128 256 2>&1 | tail -3
echo "== trace (no fence)"; S2S_GRU_DBG=4 S2S_GRU_TRACE=1 timeout 120 python benchmarks/gru_micro.py 2>&1 | grep "gru trace" | head -4

cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_nn.py -x -q -k "gru or rnn or RNN or GRU" 2>&1 | tail -5
for pk in 1 0; do echo "V2 PK=$pk"; S2S_GRU_PK=$pk timeout 120 python benchmarks/gru_micro.py 2>&1 | tail -3; done
echo "V2 B=28"; timeout 120 python benchmarks/gru_micro.py 28 2>&1 | tail -3
echo "V2 BG=4 (2 waves?)"; S2S_GRU_BG=4 timeout 120 python benchmarks/gru_micro.py 2>&1 | tail -3
echo "V2 BG=6"; S2S_GRU_BG=6 timeout 120 python benchmarks/gru_micro.py 2>&1 | tail -3
echo "V1"; S2S_GRU_V2=0 timeout 120 python benchmarks/gru_micro.py 2>&1 | tail -3
S2S_TEST_LOC_BWD_CLUSTER=0 timeout 900 python -m pytest tests/test_gpu_timed_path.py -x -q 2>&1 | tail -5
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-400

cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for d in 0 1 2 4 7; do echo "V2 DBG=$d"; S2S_GRU_DBG=$d timeout 120 python benchmarks/gru_micro.py 2>&1 | head -1; done
echo "H=128"; timeout 120 python benchmarks/gru_micro.py 32 300 128 256 2>&1 | head -2
echo "H=128 v1"; S2S_GRU_V2=0 timeout 120 python benchmarks/gru_micro.py 32 300 128 256 2>&1 | head -2
timeout 900 python -m pytest tests/test_gpu_timed_path.py -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests/test_gpu_model.py -x -q 2>&1 | tail -5

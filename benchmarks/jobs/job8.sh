cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_trainutils.py tests/test_gpu_nn.py tests/test_gpu_kernels.py -x -q 2>&1 | tail -15
timeout 300 python benchmarks/gru_micro.py 2>&1 | tail -3

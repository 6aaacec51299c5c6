cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; export PYTHONPATH=$GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_dp_nccl.py -q -k graph 2>&1 | grep -v "^$" | tail -40 | cut -c1-400

cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; export PYTHONPATH=$GRAFT_REPO_ROOT
echo "== gru fwd trace"; S2S_GRU_TRACE=1 timeout 120 python benchmarks/gru_micro.py 2>&1 | grep "gru trace" | head -3 | cut -c1-1600

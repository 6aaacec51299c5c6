cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; export PYTHONPATH=$GRAFT_REPO_ROOT
echo "== gru bwd trace (gen3, B=32)"; S2S_GRU_TRACE=2 timeout 120 python benchmarks/gru_micro.py 2>&1 | grep "bwd trace" | head -4

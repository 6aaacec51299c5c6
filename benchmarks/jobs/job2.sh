cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for v in "16 1 0.02" "0 1 0.0" "16 1 0.0" "0 1 0.02" "16 0 0.02"; do timeout 300 python benchmarks/debug/bisect_timed.py $v 2>&1 | tail -6; done
S2S_GRU_PIPE_BWD=0 timeout 300 python benchmarks/debug/bisect_timed.py 16 1 0.02 2>&1 | tail -6
S2S_OVERLAP=0 timeout 300 python benchmarks/debug/bisect_timed.py 16 1 0.02 2>&1 | tail -6
S2S_TC=0 timeout 300 python benchmarks/debug/bisect_timed.py 16 1 0.02 2>&1 | tail -6
S2S_DEC_CLUSTER=0 timeout 300 python benchmarks/debug/bisect_timed.py 16 1 0.02 2>&1 | tail -6

set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
nproc
S2S_TEST_LOC_BWD_CLUSTER=0 timeout 900 python -m pytest tests/test_gpu_timed_path.py -x -q 2>&1 | tail -15
for d in 0 1 2 3; do echo "GRU_DBG=$d"; S2S_GRU_DBG=$d timeout 120 python benchmarks/gru_micro.py 2>&1 | tail -3; done
echo "B=28"; timeout 120 python benchmarks/gru_micro.py 28 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 2>/dev/null | tee gpurun_out/r02_bench_base.json | cut -c1-600

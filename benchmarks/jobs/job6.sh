cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
echo "micro, warm-up added"; timeout 120 python benchmarks/gru_micro.py 2>&1 | tail -3
echo "micro, lengths given"; GRU_MICRO_LENGTHS=1 timeout 120 python benchmarks/gru_micro.py 2>&1 | tail -3
echo "micro v1"; S2S_GRU_V2=0 timeout 120 python benchmarks/gru_micro.py 2>&1 | tail -3
echo "bench PK=0"; S2S_GRU_PK=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['workload'][:8], d['ms_per_step'], {k:(round(v['ms_per_step'],3), v['launches_per_step']) for k,v in d['kernels'].items()})"
echo "bench with GRU prof"; S2S_GRU_PROF=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | grep "gru2 fwd" | tail -6

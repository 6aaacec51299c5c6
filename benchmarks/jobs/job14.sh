cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
echo "== gemm tests (MN-major operands)"; timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -q 2>&1 | grep -v "^E  \|^$" | tail -12 | cut -c1-300
if timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -x -q > /dev/null 2>&1; then export S2S_TC_MN=1; else export S2S_TC_MN=0; fi
echo "S2S_TC_MN=$S2S_TC_MN"
echo "== cluster vs per-step x3"
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_timed_path.py -q -k "per_step_chain" 2>&1 | grep "assert\|Error\|passed\|failed" | head -6 | cut -c1-300; done
echo "== same with S2S_OVERLAP=0"
S2S_OVERLAP=0 timeout 300 python -m pytest tests/test_gpu_timed_path.py -q -k "per_step_chain" 2>&1 | grep "assert\|Error\|passed\|failed" | head -6 | cut -c1-300
echo "== timed path + model tests"; timeout 900 python -m pytest tests/test_gpu_timed_path.py tests/test_gpu_model.py -q 2>&1 | tail -5 | cut -c1-300
echo "== bench OVERLAP=0"
S2S_OVERLAP=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-variants 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['workload'][:8], d['ms_per_step'], {k:(round(v['ms_per_step'],3), v['launches_per_step']) for k,v in d['kernels'].items()})"
echo "== bench MN=0"
S2S_TC_MN=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-variants 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['workload'][:8], d['ms_per_step'], {k:(round(v['ms_per_step'],3), v['launches_per_step']) for k,v in d['kernels'].items()})"
echo "== bench N=1"
timeout 600 python bench.py --steps 20 --warmup 3 2>gpurun_out/job14_n1.err | tee gpurun_out/r02_bench_n1.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=1', d['ms_per_step'], d['value'], d['e2e']['value'], {k:(round(v['ms_per_step'],3), v['launches_per_step']) for k,v in d['kernels'].items()}, {k:(round(v['ms_per_step'],3), round(v['value'])) for k,v in d['variants'].items()})"
tail -3 gpurun_out/job14_n1.err

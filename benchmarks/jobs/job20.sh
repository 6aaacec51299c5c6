cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; export PYTHONPATH=$GRAFT_REPO_ROOT
echo "== full gpu suite"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | cut -c1-300
echo "== bench N=1"
timeout 600 python bench.py --steps 20 --warmup 3 2>gpurun_out/job20_n1.err | tee gpurun_out/r02_bench_n1.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=1', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['kernel'], {k:(round(v['ms_per_step'],3), v['launches_per_step']) for k,v in d['kernels'].items()}, {k:(round(v['ms_per_step'],3), round(v['value'])) for k,v in d['variants'].items()})"
tail -3 gpurun_out/job20_n1.err
echo "== gemm micro"; timeout 300 python benchmarks/gemm_micro.py > gpurun_out/r02_gemm_micro.txt 2>&1; tail -16 gpurun_out/r02_gemm_micro.txt | cut -c1-200
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-variants --no-cpu-baseline > gpurun_out/ncu_r02_ll.log 2>&1
echo rc=$?; wc -l gpurun_out/r02_launches.csv
echo "== ncu full"
for k in gru3_fwd_kernel gru3_bwd_kernel dec_cluster_fwd_kernel dec_cluster_bwd_kernel attn_dvh_kernel attn_v1_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o /tmp/prof_r02_$k python bench.py --steps 2 --warmup 3 --no-variants --no-cpu-baseline > gpurun_out/ncu_r02_$k.log 2>&1
  echo $k rc=$?
  ncu -i /tmp/prof_r02_$k.ncu-rep --page raw --csv > gpurun_out/r02_raw_$k.csv 2>/dev/null
done
# the large forward projection and the MN-major gradient forms: the GEMM micro-benchmark under ncu (three kernels: NT projection, NN dX, TN dW)
GEMM_MICRO_ONLY="layers 2-3,dX = dA,dW_x" timeout 600 ncu --set full --clock-control none -k regex:gemm_tc_kernel -s 9 -c 12 -f -o /tmp/prof_r02_gemm python benchmarks/gemm_micro.py > gpurun_out/ncu_r02_gemm.log 2>&1
echo gemm rc=$?
ncu -i /tmp/prof_r02_gemm.ncu-rep --page raw --csv > gpurun_out/r02_raw_gemm_tc_kernel.csv 2>/dev/null
echo "== cfg5 sweep"
timeout 600 python bench.py --config cfg5 --steps 10 --warmup 3 2>gpurun_out/job20_c5.err > gpurun_out/r02_attn_sweep.json; tail -2 gpurun_out/job20_c5.err; python -c "
import json; d=json.load(open('gpurun_out/r02_attn_sweep.json')); print(d['metric'], d['value'], d['unit'])"
du -sh gpurun_out

cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
echo "== gemm tests (MN-major operands)"; timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -x -q 2>&1 | tail -8 | cut -c1-400
if timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -x -q > /dev/null 2>&1; then export S2S_TC_MN=1; else export S2S_TC_MN=0; fi
echo "S2S_TC_MN=$S2S_TC_MN"
echo "== timed path + model tests (V1 fork)"; timeout 900 python -m pytest tests/test_gpu_timed_path.py tests/test_gpu_model.py -x -q 2>&1 | tail -3
echo "== bench MN=0"
S2S_TC_MN=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-variants 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['workload'][:8], d['ms_per_step'], {k:(round(v['ms_per_step'],3), v['launches_per_step']) for k,v in d['kernels'].items()})"
echo "== bench N=1"
timeout 600 python bench.py --steps 20 --warmup 3 2>gpurun_out/job13_n1.err | tee gpurun_out/r02_bench_n1.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=1', d['ms_per_step'], d['value'], d['e2e']['value'], {k:(round(v['ms_per_step'],3), v['launches_per_step']) for k,v in d['kernels'].items()}, {k:(round(v['ms_per_step'],3), round(v['value'])) for k,v in d['variants'].items()})"
tail -3 gpurun_out/job13_n1.err
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-variants --no-cpu-baseline > gpurun_out/ncu_r02_ll.log 2>&1
echo rc=$?; wc -l gpurun_out/r02_launches.csv
echo "== ncu full"
for k in gru3_fwd_kernel gru3_bwd_kernel dec_cluster_fwd_kernel dec_cluster_bwd_kernel gemm_tc_kernel attn_dvh_kernel attn_v1_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o /tmp/prof_r02_$k python bench.py --steps 2 --warmup 3 --no-variants --no-cpu-baseline > gpurun_out/ncu_r02_$k.log 2>&1
  echo $k rc=$?
  ncu -i /tmp/prof_r02_$k.ncu-rep --page raw --csv > gpurun_out/r02_raw_$k.csv 2>/dev/null
  ncu -i /tmp/prof_r02_$k.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r02_src_$k.csv.gz
  ls -la /tmp/prof_r02_$k.ncu-rep
done
du -sh gpurun_out

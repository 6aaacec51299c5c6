cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; export PYTHONPATH=$GRAFT_REPO_ROOT
echo "== penalty gate determinism"; timeout 300 python benchmarks/debug/pen_determinism.py 2>&1 | tail -8 | cut -c1-300
echo "== timed path + model tests"; timeout 1200 python -m pytest tests/test_gpu_timed_path.py tests/test_gpu_model.py tests/test_gpu_kernels.py -q 2>&1 | tail -8 | cut -c1-300
echo "== bench N=1"
timeout 600 python bench.py --steps 20 --warmup 3 2>gpurun_out/job16_n1.err | tee gpurun_out/r02_bench_n1.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=1', d['ms_per_step'], d['value'], d['e2e']['value'], {k:(round(v['ms_per_step'],3), v['launches_per_step']) for k,v in d['kernels'].items()}, {k:(round(v['ms_per_step'],3), round(v['value'])) for k,v in d['variants'].items()})"
tail -3 gpurun_out/job16_n1.err
echo "== ncu attn_v1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_v1_kernel -s 6 -c 1 -f -o /tmp/prof_v1 python bench.py --steps 2 --warmup 3 --no-variants --no-cpu-baseline > gpurun_out/ncu_r02b_v1.log 2>&1
ncu -i /tmp/prof_v1.ncu-rep --page raw --csv > gpurun_out/r02b_raw_attn_v1_kernel.csv 2>/dev/null
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 2 --warmup 3 --no-variants --no-cpu-baseline > gpurun_out/ncu_r02b_ll.log 2>&1
echo rc=$?; wc -l gpurun_out/r02b_launches.csv

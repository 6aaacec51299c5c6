cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; export PYTHONPATH=$GRAFT_REPO_ROOT
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-variants --no-cpu-baseline > gpurun_out/ncu_r02_ll.log 2>&1
echo rc=$?; wc -l gpurun_out/r02_launches.csv
for k in gru3_fwd_kernel gru3_bwd_kernel; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o /tmp/prof_r02_$k python bench.py --steps 2 --warmup 3 --no-variants --no-cpu-baseline > gpurun_out/ncu_r02_$k.log 2>&1
  echo $k rc=$?
  ncu -i /tmp/prof_r02_$k.ncu-rep --page raw --csv > gpurun_out/r02_raw_$k.csv 2>/dev/null
done

CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/plain_c4_pair.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4_pair.csv $CMD > gpurun_out/ncu_l.log 2>&1
$CMD > gpurun_out/plain_c4_pair2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_tc_pass -s 6 -c 2 -o gpurun_out/prof_tc_c4_pair -f $CMD > gpurun_out/ncu_f.log 2>&1
tail -c 300 gpurun_out/plain_c4_pair.log; wc -l gpurun_out/launches_c4_pair.csv
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3

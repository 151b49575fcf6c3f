timeout 300 python scripts/gpu_time_models.py > gpurun_out/v6_models.log 2>&1
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/v6_pytest_all.log
python __graft_entry__.py smoke > gpurun_out/v6_smoke.log 2>&1
cat gpurun_out/v6_models.log; tail -5 gpurun_out/v6_pytest_all.log; tail -3 gpurun_out/v6_smoke.log

timeout 600 python -m pytest tests/test_tensor_path_gpu.py -x -q -m gpu 2>&1 | tail -4
timeout 200 python scripts/gpu_time_models.py 2>&1 | grep "elbo=True"

run() { echo "== $1 EW=$2"; ORI_TC_EW=$2 ORIANA_B200_LIB=$PWD/oriana_b200/lib/$1.so timeout 300 python scripts/gpu_diag_tc.py time tconly 2>&1 | tail -2; }
timeout 600 python -m pytest tests/test_tensor_path_gpu.py -x -q -m gpu 2>&1 | tail -3
( run liboriana_b200 8; run liboriana_b200 16 ) > gpurun_out/v6_variants.log 2>&1
cat gpurun_out/v6_variants.log

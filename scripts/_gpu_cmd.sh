run() { echo "== $1 EW=$2"; ORI_TC_EW=$2 ORIANA_B200_LIB=$PWD/oriana_b200/lib/$1.so timeout 300 python scripts/gpu_diag_tc.py time tconly 2>&1 | tail -2; }
( run liboriana_b200 8; run liboriana_b200 16; run variants/batch 8; run variants/late 8; run variants/late 16; run variants/batchlate 8 ) > gpurun_out/v5_variants.log 2>&1
cat gpurun_out/v5_variants.log

( time python bench.py ) > gpurun_out/bench_c4_v6.json 2> gpurun_out/bench_c4_v6.err
tail -c 2500 gpurun_out/bench_c4_v6.json; tail -5 gpurun_out/bench_c4_v6.err
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_ref_v6.json 2> gpurun_out/bench_ref_v6.err
cat gpurun_out/bench_ref_v6.json; tail -4 gpurun_out/bench_ref_v6.err

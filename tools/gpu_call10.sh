mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02n_pytest.log
python __graft_entry__.py smoke > gpurun_out/r02n_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02n_smoke.log
( time timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r02n_bench_1gpu.json 2> gpurun_out/r02n_bench_1gpu.err ) 2> gpurun_out/r02n_bench_time.txt; echo "bench rc=$?" >> gpurun_out/r02n_bench_1gpu.err
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02n_bench_reference.json 2>/dev/null ) 2>> gpurun_out/r02n_bench_time.txt
Q="--steps 4 --warmup 3 --no_sweep --no_cpu_baseline --no_e2e --no_system --no_secondary"
tail -5 gpurun_out/r02n_pytest.log; tail -3 gpurun_out/r02n_smoke.log; tail -2 gpurun_out/r02n_bench_1gpu.err; cat gpurun_out/r02n_bench_time.txt

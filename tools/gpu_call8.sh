mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_ref_golden.py tests/test_gpu_coalesce.py -q -p no:cacheprovider -k "dien or recurrence or fused_din" > gpurun_out/r02h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02h_pytest.log
timeout 600 python bench.py --config dien --steps 20 --warmup 5 --no_sweep --no_cpu_baseline --no_system --no_secondary > gpurun_out/r02h_bench_dien.json 2> gpurun_out/r02h_bench_dien.err; echo "rc=$?" >> gpurun_out/r02h_bench_dien.err
timeout 600 python bench.py --steps 20 --warmup 5 --no_system --no_sweep --no_cpu_baseline --no_e2e --secondary din,din254 > gpurun_out/r02h_bench_din.json 2> gpurun_out/r02h_bench_din.err
tail -30 gpurun_out/r02h_pytest.log; tail -3 gpurun_out/r02h_bench_dien.err

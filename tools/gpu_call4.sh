mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_ref_golden.py tests/test_gpu_packed.py tests/test_gpu_coalesce.py tests/test_zfullsize_more_gpu.py -q -p no:cacheprovider -k "din or packed or coalesce" > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no_system --no_sweep --secondary din,din254 > gpurun_out/r02d_bench_din.json 2> gpurun_out/r02d_bench_din.err; echo "bench rc=$?" >> gpurun_out/r02d_bench_din.err
tail -6 gpurun_out/r02d_pytest.log; tail -2 gpurun_out/r02d_bench_din.err

mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02k_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 --no_system --no_sweep --no_cpu_baseline --secondary wnd,ncf,din,din254,mtwnd,dien > gpurun_out/r02k_bench_short.json 2> gpurun_out/r02k_bench_short.err; echo "rc=$?" >> gpurun_out/r02k_bench_short.err
tail -8 gpurun_out/r02k_pytest.log; tail -2 gpurun_out/r02k_bench_short.err

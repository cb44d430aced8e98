mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r02a_gpu.txt 2>&1
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
python __graft_entry__.py smoke > gpurun_out/r02a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02a_smoke.log
for p in k1_probe k1_probe_u16 k1_probe_u10 k1_probe_u6 k1_probe_ef; do timeout 300 tools/bin/$p 16384 10 > gpurun_out/r02a_$p.jsonl 2>&1; done
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench_1gpu.json 2> gpurun_out/r02a_bench_1gpu.err; echo "bench rc=$?" >> gpurun_out/r02a_bench_1gpu.err
tail -40 gpurun_out/r02a_pytest.log; tail -3 gpurun_out/r02a_smoke.log; tail -3 gpurun_out/r02a_bench_1gpu.err

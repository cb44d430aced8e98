mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02j_pytest.log
python __graft_entry__.py smoke > gpurun_out/r02j_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02j_smoke.log
( time timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r02j_bench_1gpu.json 2> gpurun_out/r02j_bench_1gpu.err ) 2> gpurun_out/r02j_bench_time.txt; echo "bench rc=$?" >> gpurun_out/r02j_bench_1gpu.err
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02j_bench_reference.json 2>/dev/null ) 2>> gpurun_out/r02j_bench_time.txt
Q="--steps 4 --warmup 3 --no_sweep --no_cpu_baseline --no_e2e --no_system --no_secondary"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rnn_tc -s 8 -c 2 -o gpurun_out/r02j_rnn_tc python bench.py $Q --config dien --staged_batches 2 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_din_attention -s 4 -c 2 -o gpurun_out/r02j_din_attention python bench.py $Q --config din254 --batch 4096 --staged_batches 2 > /dev/null 2>&1
tail -5 gpurun_out/r02j_pytest.log; tail -3 gpurun_out/r02j_smoke.log; tail -2 gpurun_out/r02j_bench_1gpu.err; cat gpurun_out/r02j_bench_time.txt

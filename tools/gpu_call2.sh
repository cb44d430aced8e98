mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench_1gpu.json 2> gpurun_out/r02b_bench_1gpu.err; echo "bench rc=$?" >> gpurun_out/r02b_bench_1gpu.err
Q="--steps 4 --warmup 3 --no_sweep --no_cpu_baseline --no_e2e --no_system --no_secondary"
# launch list of the step (cold-cache, serialised: compare shares)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches_rm1.csv python bench.py $Q > /dev/null 2>&1
# full captures of the dominant kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sls_gather_persist -s 12 -c 2 -o gpurun_out/r02b_gather_rm1 python bench.py $Q > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sls_gather_persist -s 12 -c 2 -o gpurun_out/r02b_gather_rm2 python bench.py $Q --config rm2 --batch 4096 --staged_batches 2 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fc_tc_persist -s 8 -c 4 -o gpurun_out/r02b_fc_persist_wnd python bench.py $Q --config wnd --staged_batches 2 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mlp_chain -s 12 -c 4 -o gpurun_out/r02b_chain_rm1 python bench.py $Q > /dev/null 2>&1
# where the serving loop's time goes at saturation
DRS_ACCEL_PROFILE=$PWD/gpurun_out/r02b_engine_loop_cprofile.txt timeout 300 python -m deeprecsys_b200.serve --config_file deeprecsys_b200/models/configs/dlrm_rm1.json --num_batches 32 --max_mini_batch_size 1024 --batch_size_distribution normal --avg_mini_batch_size 256 --var_mini_batch_size 32 --model_accel --queue --accel_precision tf32x3 --serve_rates_ms 0.009 --serve_queries 60000 --serve_generators 3 > gpurun_out/r02b_serve_profiled.json 2> gpurun_out/r02b_serve_profiled.err
tail -5 gpurun_out/r02b_pytest.log; tail -2 gpurun_out/r02b_bench_1gpu.err; ls -la gpurun_out | grep r02b

# round 2, call 23: final-state ncu evidence — launch list of the RMC1 step, full captures of the
# gather (K1), the MLP chain and the new id-widening kernel; cProfile of the record loop
mkdir -p gpurun_out
Q="--steps 4 --warmup 3 --no_sweep --no_cpu_baseline --no_e2e --no_system --no_secondary"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02z_launches_rm1.csv python bench.py $Q > /dev/null 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sls_gather_persist -s 12 -c 2 -o gpurun_out/r02z_gather_rm1 python bench.py $Q > /dev/null 2>&1; echo "gather rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mlp_chain -s 12 -c 4 -o gpurun_out/r02z_chain_rm1 python bench.py $Q > /dev/null 2>&1; echo "chain rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_unpack_ids_blk -s 8 -c 2 -o gpurun_out/r02z_unpack_blk python bench.py --steps 4 --warmup 3 --no_sweep --no_cpu_baseline --no_system --no_secondary > /dev/null 2>&1; echo "unpack rc=$?"
DRS_ACCEL_PROFILE=$PWD/gpurun_out/r02z_record_loop_cprofile.txt timeout 300 python -m deeprecsys_b200.serve --config_file deeprecsys_b200/models/configs/dlrm_rm1.json --num_batches 32 --max_mini_batch_size 1024 --batch_size_distribution normal --avg_mini_batch_size 256 --var_mini_batch_size 32 --model_accel --queue --accel_precision tf32x3 --serve_rates_ms 0.0045 --serve_queries 60000 --serve_generators 4 > gpurun_out/r02z_serve_profiled.json 2> gpurun_out/r02z_serve_profiled.err
ls -la gpurun_out | grep r02z

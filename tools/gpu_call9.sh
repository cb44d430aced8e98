mkdir -p gpurun_out
Q="--steps 4 --warmup 3 --no_sweep --no_cpu_baseline --no_e2e --no_system --no_secondary"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02i_launches_dien.csv python bench.py $Q --config dien --staged_batches 2 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rnn_tc -s 8 -c 2 -o gpurun_out/r02i_rnn_tc python bench.py $Q --config dien --staged_batches 2 > /dev/null 2>&1
ls -la gpurun_out | grep r02i

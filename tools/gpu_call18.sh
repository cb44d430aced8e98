# round 2, call 18: record loop — group size / lanes A/B
mkdir -p gpurun_out
CFG=deeprecsys_b200/models/configs/dlrm_rm1.json
COMMON="--config_file $CFG --num_batches 32 --max_mini_batch_size 1024 --batch_size_distribution normal --avg_mini_batch_size 256 --var_mini_batch_size 32 --model_accel --queue --accel_precision tf32x3 --serve_queries 6000 --serve_generators 4"
i=0
for opt in "--accel_coalesce 32 --accel_stack_rows 8192" "--accel_coalesce 32 --accel_stack_rows 8192 --accel_lanes 3" "--accel_coalesce 16 --accel_stack_rows 4096 --accel_lanes 3" "--accel_coalesce 32 --accel_stack_rows 16384 --accel_lanes 3"; do
i=$((i+1))
timeout 300 python -m deeprecsys_b200.serve $COMMON $opt --serve_rates_ms 0.009,0.007,0.0055,0.0045,0.0038 > gpurun_out/r02s_serve_$i.json 2> gpurun_out/r02s_serve_$i.err
python - <<P
import json
d = json.load(open("gpurun_out/r02s_serve_$i.json"))
print("$opt")
for r in d["runs"]:
    print("  offered %7.0f served %7.0f qps %6.2f M inf/s p50 %.3f p99 %.3f failed %d" % (r["offered_qps"], r["qps"], r["inferences_per_s"]/1e6, r["p50_ms"], r["p99_ms"], r["failed"]))
P
done

# round 2, call 17: serving loop over packet records (C wire codec) vs the object loop
mkdir -p gpurun_out
CFG=deeprecsys_b200/models/configs/dlrm_rm1.json
COMMON="--config_file $CFG --num_batches 32 --max_mini_batch_size 1024 --batch_size_distribution normal --avg_mini_batch_size 256 --var_mini_batch_size 32 --model_accel --queue --accel_precision tf32x3 --serve_queries 6000"
for w in 0 1; do for g in 3 5; do
DRS_ACCEL_WIRE=$w timeout 300 python -m deeprecsys_b200.serve $COMMON --serve_generators $g --serve_rates_ms 0.014,0.009,0.007,0.0055,0.0045 > gpurun_out/r02r_serve_wire${w}_g${g}.json 2> gpurun_out/r02r_serve_wire${w}_g${g}.err
python - <<P
import json
d = json.load(open("gpurun_out/r02r_serve_wire${w}_g${g}.json"))
print("wire=$w gens=$g")
for r in d["runs"]:
    print("  offered %7.0f served %7.0f qps %6.2f M inf/s p50 %.3f p99 %.3f failed %d" % (r["offered_qps"], r["qps"], r["inferences_per_s"]/1e6, r["p50_ms"], r["p99_ms"], r["failed"]))
P
done; done
timeout 600 python -m pytest tests/test_input_ring.py tests/test_gpu_coalesce.py -x -q -m gpu 2>&1 | tail -3

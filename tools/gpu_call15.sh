# round 2, call 15: host packing, copies back to back; A/B streaming stores
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_hostpack.py -x -q -m gpu > gpurun_out/r02p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02p_pytest.log
tail -3 gpurun_out/r02p_pytest.log
for nt in 0 1; do
DRS_HOSTPACK_NT=$nt timeout 900 python bench.py --steps 20 --warmup 5 --no_secondary --no_system --no_sweep --no_cpu_baseline > gpurun_out/r02p_bench_e2e_nt$nt.json 2> gpurun_out/r02p_bench_e2e_nt$nt.err; echo "bench rc=$?"
python - <<P
import json
d = json.load(open("gpurun_out/r02p_bench_e2e_nt$nt.json"))
print("NT=$nt", d["e2e"]["value"], json.dumps(d["e2e"]["host_packed"], indent=1))
P
done

# round 2, call 25: host packer thread count sweep (e2e only)
mkdir -p gpurun_out
DRS_BENCH_PACK_THREADS=3,4,5,6,7,8,10,12 timeout 600 python bench.py --steps 20 --warmup 5 --no_secondary --no_system --no_sweep --no_cpu_baseline > gpurun_out/r02ad_bench_e2e_threads.json 2> gpurun_out/r02ad_bench_e2e_threads.err; echo "bench rc=$?"
python - <<P
import json
d = json.load(open("gpurun_out/r02ad_bench_e2e_threads.json"))
e = d["e2e"]
print("plain", round(e["plain"]["value"]/1e6,2), [(x["host_threads"], round(x["value"]/1e6,2), round(x["enqueue_call_ms_avg"],3)) for x in e["host_pack"]])
P

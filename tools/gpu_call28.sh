mkdir -p gpurun_out
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 5 --no_system --no_sweep --no_cpu_baseline --secondary din,ncf,wnd > gpurun_out/r02ag_bench_$i.json 2> gpurun_out/r02ag_bench_$i.err; echo "bench rc=$?"
python - <<P
import json
d = json.load(open("gpurun_out/r02ag_bench_$i.json"))
e = d["e2e"]
print("value", round(d["value"]/1e6,2), "e2e", round(e["value"]/1e6,2), e["mode"], e["host_threads"], [(x["host_threads"], round(x["value"]/1e6,2)) for x in e["host_pack"]], {k: round(v["value"]/1e6,2) for k,v in e["int64_ids"].items()})
for k, v in d["secondary"].items(): print(k, round(v["value"]/1e6,1), round(v["e2e"]["value"]/1e6,1), v["e2e"]["mode"], v["e2e"].get("steps"))
P
done

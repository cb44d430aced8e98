# round 2, call 19: what the driver runs at round end — GPU tier, smoke, reference arm, bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r02ah_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ah_pytest.log
tail -4 gpurun_out/r02ah_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02ah_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02ah_smoke.log; tail -5 gpurun_out/r02ah_smoke.log
( time python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02ah_bench_reference.json 2> gpurun_out/r02ah_bench_reference.err ) 2> gpurun_out/r02ah_bench_reference_time.txt; echo "ref rc=$?"
( time python bench.py --steps 20 --warmup 5 > gpurun_out/r02ah_bench_1gpu.json 2> gpurun_out/r02ah_bench_1gpu.err ) 2> gpurun_out/r02ah_bench_time.txt; echo "bench rc=$?"
cat gpurun_out/r02ah_bench_time.txt
python - <<'P'
import json
d = json.load(open("gpurun_out/r02ah_bench_1gpu.json"))
e = d["e2e"]
print("value", d["value"], "frac", d["roofline"]["frac"], "e2e", e["value"], e["mode"], e["host_threads"], "plain", e["plain"]["value"])
s = d["system"]
print("system", s["inferences_per_s_at_p99"], s["qps"], s["p99_ms"])
for k, v in d["secondary"].items():
    print(k, v.get("value"), v.get("e2e", {}).get("value"), v.get("e2e", {}).get("mode"), v.get("parity", {}).get("ok"), v.get("error"))
P

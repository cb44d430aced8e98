# round 2, call 21: the driver's N-GPU line (torchrun, default flags) on the final state
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02ac_bench_${N}gpu.json 2> gpurun_out/r02ac_bench_${N}gpu.err ) 2> gpurun_out/r02ac_bench_${N}gpu_time.txt; echo "bench rc=$?"
cat gpurun_out/r02ac_bench_${N}gpu_time.txt | tail -4
python - <<P
import json
d = json.load(open("gpurun_out/r02ac_bench_${N}gpu.json"))
e = d["e2e"]
print("value", d["value"], "frac", d["roofline"]["frac"], "e2e", e["value"], e["mode"])
s = d.get("system") or {}
print("system", s.get("inferences_per_s_at_p99"), s.get("qps"), s.get("p99_ms"), s.get("pairs"))
for k, v in (d.get("sharded") or {}).items():
    print(k, v.get("value"), v.get("parity"), v.get("error"))
P

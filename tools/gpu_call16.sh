# round 2, call 16: N-GPU bench (replica headline + e2e modes + one sharded leg) after host packing
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
nproc > gpurun_out/r02q_host_${N}gpu.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 20 --warmup 5 --no_system --sharded rm1:table > gpurun_out/r02q_bench_${N}gpu.json 2> gpurun_out/r02q_bench_${N}gpu.err; echo "bench rc=$?" >> gpurun_out/r02q_bench_${N}gpu.err
tail -3 gpurun_out/r02q_bench_${N}gpu.err
python - <<P
import json
d = json.load(open("gpurun_out/r02q_bench_${N}gpu.json"))
e = d["e2e"]
print("value", d["value"], "e2e", e["value"], e["mode"], e["host_threads"], "plain", e["plain"]["value"], "hp", [(x["host_threads"], x["value"]) for x in e["host_pack"]])
print("int64", e.get("int64_ids")); print("sharded", json.dumps(d.get("sharded"))[:600])
P

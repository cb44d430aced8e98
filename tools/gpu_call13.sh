mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02m_bench_${N}gpu.json 2> gpurun_out/r02m_bench_${N}gpu.err; echo "bench rc=$?" >> gpurun_out/r02m_bench_${N}gpu.err
tail -3 gpurun_out/r02m_bench_${N}gpu.err

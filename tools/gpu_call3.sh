mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r02l_gpus.txt 2>&1
nvidia-smi topo -m >> gpurun_out/r02l_gpus.txt 2>&1
python -m pytest tests/test_multigpu_gpu.py -q -p no:cacheprovider > gpurun_out/r02l_pytest_multigpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02l_pytest_multigpu.log
N=$(nvidia-smi -L | wc -l)
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02l_bench_${N}gpu.json 2> gpurun_out/r02l_bench_${N}gpu.err; echo "bench rc=$?" >> gpurun_out/r02l_bench_${N}gpu.err
tail -8 gpurun_out/r02l_pytest_multigpu.log; tail -3 gpurun_out/r02l_bench_${N}gpu.err

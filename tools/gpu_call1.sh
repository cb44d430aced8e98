mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r02a_gpu.txt 2>&1
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
python __graft_entry__.py smoke > gpurun_out/r02a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02a_smoke.log
for p in k1_probe k1_probe_u16 k1_probe_u10 k1_probe_u6 k1_probe_ef; do timeout 300 tools/bin/$p 16384 10 > gpurun_out/r02a_$p.jsonl 2>&1; done
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench_1gpu.json 2> gpurun_out/r02a_bench_1gpu.err; echo "bench rc=$?" >> gpurun_out/r02a_bench_1gpu.err
tail -40 gpurun_out/r02a_pytest.log; tail -3 gpurun_out/r02a_smoke.log; tail -3 gpurun_out/r02a_bench_1gpu.err
# A/B: MLP-chain register cap (co-residency with the gather), lanes x gather CTAs per SM
Q="--steps 20 --warmup 5 --no_secondary --no_system --no_cpu_baseline --no_e2e --no_sweep"
for v in chain80 chain112; do DRS_LIB_PATH=$PWD/tools/bin/libdrs_$v.so timeout 300 python bench.py $Q > gpurun_out/r02a_ab_$v.json 2>/dev/null; done
for v in chain80; do for c in 3 4; do DRS_GATHER_CTAS_PER_SM=$c DRS_LIB_PATH=$PWD/tools/bin/libdrs_$v.so timeout 300 python bench.py $Q --lanes 3 > gpurun_out/r02a_ab_${v}_l3_c$c.json 2>/dev/null; done; done
for l in 2 4; do for c in 3 4; do DRS_GATHER_CTAS_PER_SM=$c timeout 300 python bench.py $Q --lanes $l > gpurun_out/r02a_ab_l${l}_c$c.json 2>/dev/null; done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02a_ab_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']/1e6,2), 'M inf/s', round(d['ms_per_step'],4), 'ms', round(d['roofline']['frac'],3))
    except Exception as e: print(f, 'ERR', e)
PY

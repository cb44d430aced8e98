# round 2, call 14: engine-side host packing of the ids (drs_set_option "host_pack")
mkdir -p gpurun_out
nproc > gpurun_out/r02o_host.txt; lscpu | grep -E "Model name|^CPU\(s\)|L3" >> gpurun_out/r02o_host.txt
timeout 600 python -m pytest tests/test_gpu_hostpack.py tests/test_gpu_packed.py -x -q -m gpu > gpurun_out/r02o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02o_pytest.log
tail -5 gpurun_out/r02o_pytest.log
timeout 300 python - > gpurun_out/r02o_hostpack_probe.json 2>&1 <<'P'
import json, time, numpy as np
from deeprecsys_b200.engine import hostpack_ids
from deeprecsys_b200 import _lib
import ctypes as C
lib = _lib.load()
n = 16384 * 640
rng = np.random.RandomState(1)
ids32 = rng.randint(0, 4000000, size=n).astype(np.int32); ids64 = ids32.astype(np.int64)
out = np.empty(int(lib.drs_hostpack_words(n, 22)), np.uint32)
res = {}
for name, a in (("int32", ids32), ("int64", ids64)):
    for th in (1, 2, 4, 8, 12, 16):
        ok = C.c_int(1); best = 1e9
        for _ in range(5):
            t0 = time.perf_counter()
            lib.drs_hostpack_ids(a.ctypes.data, int(a.dtype == np.int64), n, 22, th, out.ctypes.data, C.byref(ok))
            best = min(best, time.perf_counter() - t0)
        res["%s_t%d" % (name, th)] = {"ms": best * 1e3, "Gid_per_s": n / best / 1e9}
print(json.dumps(res))
P
cat gpurun_out/r02o_hostpack_probe.json
timeout 900 python bench.py --steps 20 --warmup 5 --no_secondary --no_system --no_sweep --cpu_seconds 2 > gpurun_out/r02o_bench_e2e.json 2> gpurun_out/r02o_bench_e2e.err; echo "bench rc=$?"
python - <<'P'
import json
d = json.load(open("gpurun_out/r02o_bench_e2e.json"))
print(json.dumps(d["e2e"], indent=1))
P

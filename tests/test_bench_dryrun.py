"""bench.py's host logic end to end WITHOUT a GPU: the ctypes engine is replaced by the
oracle-backed fake (tests/fake_engine.py), tables are capped to a few thousand rows, and
bench.run_b200 runs as it does on the box — N = 1 with the `secondary` block, and world size 2
over gloo with the `sharded` legs (table-wise, row-wise, RMC3 table-wise) and their in-bench
parity checks.  What is checked is the plumbing and the JSON contract, not performance."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import sys
sys.path.insert(0, %(root)r)
import bench
import deeprecsys_b200.engine as engine_mod
from tests import fake_engine
engine_mod.AccelEngine = fake_engine.FakeAccelEngine
engine_mod.pinned_empty = fake_engine.fake_pinned_empty
engine_mod.bind_to_device_numa_node = lambda device: -1
fake_engine.TABLE_PROVIDER = lambda spec, t: bench.synth_table(spec, 123, t)
bench.DIST_BACKEND = "gloo"
bench.TABLE_ROWS_CAP = 3000
bench.SECONDARY = [(name, 48) for name, _ in bench.SECONDARY]
sys.argv = ["bench.py"] + %(argv)r
sys.exit(bench.main())
"""


def _run(argv, env=None, timeout=900):
    code = SCRIPT % {"root": ROOT, "argv": argv}
    return subprocess.Popen([sys.executable, "-c", code], cwd=ROOT,
                            env=dict(os.environ, PYTHONWARNINGS="ignore", **(env or {})),
                            stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)


def _finish(proc, timeout=900):
    out, err = proc.communicate(timeout=timeout)
    assert proc.returncode == 0, err[-4000:]
    return out


COMMON = ["--steps", "3", "--warmup", "1", "--batch", "64", "--staged_batches", "2",
          "--cpu_seconds", "0.2", "--cpu_sample", "32", "--no_system",
          "--host_pack_min_ids", "0"]          # (64-sample steps: time the packed mode anyway)


def test_single_gpu_line_carries_every_secondary_config():
    out = _finish(_run(["--gpus", "1"] + COMMON))
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1, out
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
                "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "clocks",
                "gpu_launches", "e2e", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["value"] > 0 and d["gpu_launches"] > 0
    assert d["roofline"]["bound"] == "hbm" and d["roofline"]["unit"] == "GB/s"
    assert d["roofline"]["alg_bytes_per_sample"] == 84512
    e2e = d["e2e"]
    assert e2e["plain"]["h2d_bytes_per_step"] == 64 * (8 * 80 * 4 + 8 * 4 + 128 * 4)
    assert e2e["mode"] in ("plain", "host_pack") and e2e["host_pack"][0]["host_threads"] >= 1
    # lane-interleaved 512-id blocks: 10 blocks x 16 lanes x `bits` words per table (the dry
    # run's tables are cut down, so fewer bits than the shipped 22)
    packed_id_bytes = e2e["host_pack"][0]["h2d_bytes_per_step"] - 64 * (8 * 4 + 128 * 4)
    assert packed_id_bytes % (8 * 10 * 16 * 4) == 0 and 1 <= packed_id_bytes // 5120 <= 22
    assert e2e["h2d_bytes_per_step"] == (e2e["plain"] if e2e["mode"] == "plain"
                                         else e2e["host_pack"][0])["h2d_bytes_per_step"] \
        or len(e2e["host_pack"]) > 1
    assert e2e["int64_ids"]["plain"]["value"] > 0 and e2e["int64_ids"]["host_pack"]["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
    assert d["parity"]["ok"] and d["parity"]["pooled"] == "bit-exact"
    assert d["sharded"] is None
    sec = d["secondary"]
    assert sorted(sec) == ["dien", "din", "din254", "mtwnd", "ncf", "rm2", "rm3", "wnd"]
    for name, s in sec.items():
        assert "error" not in s, (name, s)
        assert s["value"] > 0 and s["e2e"]["value"] > 0 and s["cpu_baseline"]["value"] > 0
        assert s["parity"]["ok"], (name, s["parity"])
        # (DIEN's Reshape mixes the samples of a batch: no per-sample parity there, tests/ has it)
        assert s["parity"]["pooled"] == ("skipped" if name == "dien" else "bit-exact"), name
        assert s["roofline"]["bound"] in ("hbm", "tensor") and s["roofline"]["frac"] > 0
    assert sec["rm3"]["roofline"]["unit"] == "TFLOP/s" and sec["rm2"]["roofline"]["unit"] == "GB/s"
    assert "254 tables" in sec["din254"]["config"]["workload"]
    assert "4 tables" in sec["din"]["config"]["workload"]
    # both arms describe the same workload (the driver compares the config objects)
    ref = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--steps", "1", "--warmup", "1", "--cpu_sample", "32", "--batch", "64",
                          "--staged_batches", "2"], cwd=ROOT, stdout=subprocess.PIPE,
                         stderr=subprocess.PIPE, text=True, timeout=600)
    assert ref.returncode == 0, ref.stderr[-2000:]
    want = dict(d["config"], workload=d["config"]["workload"].replace("x 3000 rows", "x 4000000 rows"))
    assert json.loads(ref.stdout.strip())["config"] == want     # (the dry run caps the tables)


def test_two_ranks_over_gloo_run_the_sharded_legs_with_parity():
    port = str(29700 + os.getpid() % 200)
    # (the unfused NCCL-exchange leg needs real device buffers: not part of the dry run)
    procs = [_run(["--gpus", "2"] + COMMON + ["--no_e2e", "--no_sweep", "--sharded",
                                             "rm1:table,rm1:row,rm3:table,rm2:table"],
                  env={"RANK": str(r), "LOCAL_RANK": str(r), "WORLD_SIZE": "2",
                       "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": port}) for r in range(2)]
    outs = [_finish(p) for p in procs]
    assert outs[1].strip() == ""                       # only rank 0 prints
    d = json.loads(outs[0].strip())
    assert d["n_gpus"] == 2 and d["config"]["parallelism"] == "replica x2"
    assert d["secondary"] is None
    legs = d["sharded"]
    assert sorted(legs) == ["rm2_table", "rm3_table", "row", "table"]
    for key, leg in legs.items():
        assert "error" not in leg, (key, leg)
        assert leg["value"] > 0 and leg["clocks"] is not None and len(leg["parity_detail"]) == 2
        assert all(p["ok"] and p["samples_checked"] > 0 for p in leg["parity_detail"]), leg
    assert legs["table"]["parity"] == "bit-exact" and legs["rm3_table"]["parity"] == "bit-exact"
    assert legs["row"]["parity"] == "bit-exact vs rank-order emulation"
    assert legs["table"]["config"]["global_batch"] == 128
    assert legs["rm3_table"]["roofline"]["bound"] == "tensor"
    assert 0 < legs["table"]["vs_replica"]

"""bench.py's reference arm runs on CPU: check the JSON contract of the line it prints
(the keys the driver reads) without a GPU.  The GPU arm's line is checked on the GPU box."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra, env=None):
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + extra, cwd=ROOT,
                          env=dict(os.environ, **(env or {})), stdout=subprocess.PIPE,
                          stderr=subprocess.PIPE, text=True, timeout=600)
    assert proc.returncode == 0, proc.stderr[-2000:]
    return proc.stdout


def test_reference_arm_prints_exactly_one_json_line_with_the_contract_keys():
    out = _run(["--impl", "reference", "--steps", "2", "--warmup", "1", "--cpu_sample", "64"])
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1, out
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "DLRM-RMC1 inferences/sec"
    assert d["unit"] == "inferences/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_reference_arm_under_torchrun_only_rank0_works_and_uses_all_cores():
    # rank != 0 exits 0 without output; rank 0 ignores torchrun's OMP_NUM_THREADS=1
    out = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1",
                "--cpu_sample", "64"], env={"RANK": "1", "WORLD_SIZE": "2", "OMP_NUM_THREADS": "1"})
    assert out.strip() == ""
    out = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1",
                "--cpu_sample", "64"], env={"RANK": "0", "WORLD_SIZE": "2", "OMP_NUM_THREADS": "1"})
    d = json.loads(out.strip())
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))

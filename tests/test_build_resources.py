"""Resource usage of the built kernels (cuobjdump -res-usage on the in-tree library): occupancy
of the HBM-bound gather is set by its register count — 64 registers x 256 threads = 4 resident
CTAs per SM.  An innocent-looking edit (a second __launch_bounds__ argument) once took it to 96
registers and the 8-GPU headline from 515 M to 360 M inferences/s; this test is the tripwire."""
import re
import shutil
import subprocess

import pytest

from deeprecsys_b200 import build as B


def _res_usage():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    try:
        out = subprocess.run([exe, "-res-usage", B.build()], stdout=subprocess.PIPE,
                             stderr=subprocess.STDOUT, text=True, timeout=300).stdout
    except (OSError, subprocess.SubprocessError):
        pytest.skip("cuobjdump not available")
    usage = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+).*?SHARED:(\d+)", out):
        usage[m.group(1)] = (int(m.group(2)), int(m.group(3)))
    if not usage:
        pytest.skip("no resource records in cuobjdump output")
    return usage


def test_gather_kernels_keep_four_ctas_per_sm():
    usage = _res_usage()
    plain = {k: v for k, v in usage.items()
             if "k_sls_gather_persist" in k and "ELb0ELb0E" in k}          # <LPR, !SHORT, !MULTI>
    assert len(plain) >= 6, sorted(usage)
    for name, (regs, _smem) in plain.items():
        assert regs <= 64, "%s uses %d registers: fewer than 4 CTAs of 256 threads per SM" % (name, regs)
    multi = {k: v for k, v in usage.items()
             if "k_sls_gather_persist" in k and "ELb0ELb1E" in k}          # coalesced serving
    for name, (regs, _smem) in multi.items():
        assert regs <= 64, (name, regs)


def test_tensor_core_kernels_fit_one_cta_per_sm_with_room_for_the_gather():
    usage = _res_usage()
    chain = [v for k, v in usage.items() if "k_mlp_chain" in k]
    assert chain
    for regs, _smem in chain:
        # 192 threads: beside two gather CTAs (2 x 16,384 registers) at most 170 registers
        assert regs <= 168, regs

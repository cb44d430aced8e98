"""Table-sharded serving across GPUs of one box (needs >= 2 GPUs): K1 stores
pooled vectors straight into the owner rank's buffer over NVLink; every rank's
output slice must match the CPU oracle and pooled slices must stay bit-exact."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world,precision", [(2, "fp32"), (2, "tf32x3"), (4, "tf32x3"),
                                             (8, "fp32"), (8, "tf32x3")])
def test_table_sharded_matches_oracle(world, precision):
    n = _ngpu()
    if n < world:
        pytest.skip("needs >= %d GPUs (have %d)" % (world, n))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           "--nproc-per-node", str(world), "--master-addr", "127.0.0.1", "--master-port", "29631",
           os.path.join(ROOT, "tests", "_shard_worker.py"), precision]
    proc = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                          text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-4000:]
    assert proc.stdout.count("DONE") == world, proc.stdout[-4000:]


@pytest.mark.parametrize("world,precision", [(2, "fp32"), (4, "tf32x3"), (8, "fp32")])
def test_row_sharded_matches_oracle(world, precision):
    """Row-wise sharding (every rank a row slice of every table): integer-exact bucketing,
    pooled vectors bit-exact vs the rank-order emulation, outputs vs the oracle."""
    n = _ngpu()
    if n < world:
        pytest.skip("needs >= %d GPUs (have %d)" % (world, n))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           "--nproc-per-node", str(world), "--master-addr", "127.0.0.1", "--master-port", "29641",
           os.path.join(ROOT, "tests", "_shard_worker.py"), precision, "row"]
    proc = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                          text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-4000:]
    assert proc.stdout.count("DONE") == world, proc.stdout[-4000:]


@pytest.mark.parametrize("world,precision", [(2, "fp32"), (4, "tf32x3"), (8, "fp32")])
def test_unfused_nccl_exchange_baseline_matches_oracle(world, precision):
    """drs_exchange_begin / all_to_all_single over NCCL / drs_exchange_end: the baseline the
    fused exchange is measured against computes the same thing."""
    n = _ngpu()
    if n < world:
        pytest.skip("needs >= %d GPUs (have %d)" % (world, n))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           "--nproc-per-node", str(world), "--master-addr", "127.0.0.1", "--master-port", "29651",
           os.path.join(ROOT, "tests", "_shard_worker.py"), precision, "nccl"]
    proc = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                          text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-4000:]
    assert proc.stdout.count("DONE") == world, proc.stdout[-4000:]

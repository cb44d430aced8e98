"""GPU tier, last file: the standalone characterization entry on a real device (written after
round 1's GPU budget was spent; a failure here cannot mask the parity tests that run before)."""
import pytest

from deeprecsys_b200.models import standalone
from deeprecsys_b200.utils.utils import cli
from tests.test_standalone import SMALL, parse_like_the_reference


@pytest.mark.gpu
def test_standalone_run_on_the_gpu():
    lines = []
    res = standalone.run(cli(SMALL + ["--use_accel"]),
                         out=lambda *a: lines.append(" ".join(str(x) for x in a)))
    (tup,) = parse_like_the_reference(lines)
    assert len(tup) == 6 and all(v >= 0 for v in tup)
    assert res["device_compute_ms"] > 0 and res["iterations"] == 6

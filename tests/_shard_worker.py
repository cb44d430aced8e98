"""torchrun worker for tests/test_multigpu_gpu.py: table-sharded serving on
WORLD_SIZE GPUs, each rank checks its slice against the CPU oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch                                    # noqa: E402
import torch.distributed as dist                # noqa: E402

from deeprecsys_b200 import _lib, shard         # noqa: E402
from deeprecsys_b200.engine import AccelEngine  # noqa: E402
from oracle import numpy_oracle as O            # noqa: E402
from tests.helpers import make_case, request    # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    precision = sys.argv[1] if len(sys.argv) > 1 else "fp32"
    for name in ("dlrm_cat", "wnd", "ncf", "din_many", "dlrm_rm3_like", "mtwnd"):
        mb = 24 if name == "dlrm_rm3_like" else 96
        c = make_case(name, num_batches=2, max_batch=mb)       # same seed on every rank
        S = shard.slice_size(mb, world)
        eng = AccelEngine(args=c.args, device=local, max_batch=S, precision=precision,
                          max_lookups_per_table=mb * max(1, c.spec.L))
        shard.connect(eng, dist)
        for t in shard.owned_tables(c.spec.T, rank, world):
            eng.load_table(t, c.tables[t])
        for (mlp_id, layer), (W, b) in c.fcs.items():
            eng.load_fc(mlp_id, layer, W, b)
        for j in range(2):
            dense, ids, lengths = request(c, j, mb)
            eng.stage_batch(j, dense, ids, lengths)
        dist.barrier()
        for rep in range(3):                                   # exercises lane reuse + consumed flags
            for batch_id, B in [(0, mb), (1, mb - 5), (0, 3)]:
                dense, ids, lengths = request(c, batch_id, B)
                st = O.forward(c.spec, c.tables, c.fcs, dense, ids, lengths)
                row0, rows = shard.sample_slice(B, rank, world)
                out, _ = eng.infer_staged(batch_id, B)
                out = out[:rows]
                np.testing.assert_allclose(out, st["out"][row0:row0 + rows], rtol=1e-4, atol=1e-5)
                if rows > 0:
                    pooled = eng.fetch_stage(_lib.STAGE_POOLED)
                    want = st["pooled"][row0:row0 + rows]
                    assert np.array_equal(pooled.view(np.uint32), want.view(np.uint32)), \
                        "%s: pooled slice not bit-exact on rank %d" % (name, rank)
                if rep == 0 and B <= S:                        # host path too (slot capacity = S)
                    out_h, _ = eng.infer(dense, ids, lengths, B)
                    np.testing.assert_allclose(out_h[:rows], st["out"][row0:row0 + rows],
                                               rtol=1e-4, atol=1e-5)
                full = shard.assemble_outputs(dist, out, B, c.spec.out_dim)
                np.testing.assert_allclose(full, st["out"], rtol=1e-4, atol=1e-5)
        dist.barrier()
        eng.close()
        if rank == 0:
            print("sharded %s ok on %d GPUs (%s)" % (name, world, precision))
    dist.destroy_process_group()
    print("RANK %d DONE" % rank)


if __name__ == "__main__":
    main()

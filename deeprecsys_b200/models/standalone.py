"""Standalone characterization run of one model on the accelerator.

Counterpart of the ``__main__`` blocks of the reference's model files
(/root/reference/models/dlrm_s_caffe2.py:571-661 and the same block in wide_and_deep.py, ncf.py,
din.py, dien.py, multi_task_wnd.py): cli -> seed -> data generator -> build the model -> run
``nepochs x nbatches`` whole batches -> six "``***``" lines (data loading / computation /
execution time, each total and per iteration).  Those lines are what the reference's scrapers
read (accelerator/predict_execution.py:10-29 ``GPU_Data.parse_gpu``,
experiments/speedup/sweep_rt.py), so a sweep script written for the reference works on this
backend by swapping the command:

    python -m deeprecsys_b200.models.standalone --config_file models/configs/dlrm_rm1.json \
        --use_accel --max_mini_batch_size 1024 --num_batches 4 --nepochs 100

One binary serves every ``model_type`` (the graphs live behind one C-ABI).  "Data loading time"
is what the reference measures around FeedBlob/EnqueueBlobs (dlrm_s_caffe2.py:553-562): here the
host->device copy of the batch (``drs_timing.h2d_ms``); computation = kernels + read-back.
``--enable_profiling`` adds a per-segment split from CUDA events on the engine's stream
(pre-gather MLP / gather(+interaction) / post-gather MLP), the place of Caffe2's per-operator
``benchmark_net`` table (:565-566).  There is no CPU mode: without ``--use_accel`` the run stops,
as does a missing GPU (no fallback).
"""
from __future__ import absolute_import, division, print_function

import sys
import time

import numpy as np


def run(args, engine_factory=None, out=print):
    """The reference's main block for ``args``; returns the six reported numbers (ms)."""
    from deeprecsys_b200.data_generator.dlrm_data import DLRMDataGenerator
    if engine_factory is None:
        from deeprecsys_b200.engine import AccelEngine as engine_factory
    np.random.seed(args.numpy_rand_seed)
    np.set_printoptions(precision=args.print_precision)
    if not args.use_accel:
        sys.exit("ERROR: this backend has no CPU mode; run the reference's models/*.py for the "
                 "CPU numbers, or pass --use_accel")
    out("Using 1 Accel(s)...")
    gen = DLRMDataGenerator(args)
    (nbatches, lX, lS_l, lS_i) = gen.generate_input_data()
    out("Generating output dataset")
    gen.generate_output_data()                      # consumes the RNG before the parameters
    out("Trying to initialize %s" % args.model_type.upper())
    eng = engine_factory(args=args, max_batch=int(args.max_mini_batch_size))
    try:
        eng.create()                                # tables -> MLPs, reference draw order
        out("Created network")
        T = eng.spec.T
        dload_ms = compute_ms = 0.0
        seg = [0.0, 0.0, 0.0]
        n_iter = int(args.nepochs) * int(nbatches)
        out("Running networks")
        t_start = time.time()
        for _ in range(int(args.nepochs)):
            for j in range(nbatches):
                n = int(np.asarray(lS_l[j][0]).shape[0])
                if args.enable_profiling:
                    eng.profile_begin(1)
                _, tm = eng.infer(lX[j] if eng.spec.m_den > 0 else None,
                                  [lS_i[j][t] for t in range(T)],
                                  [lS_l[j][t] for t in range(T)], n)
                if args.enable_profiling:
                    _, a, b, c = eng.profile_end()
                    seg = [seg[0] + a, seg[1] + b, seg[2] + c]
                dload_ms += tm.h2d_ms
                compute_ms += tm.gpu_ms + tm.d2h_ms
        total_ms = (time.time() - t_start) * 1000.
    finally:
        eng.close()
    if args.enable_profiling:
        for name, ms in zip(("pre-gather MLP", "SparseLengthsSum gather + interaction",
                             "post-gather MLP"), seg):
            out("Segment %s: %.6f ms. per iter" % (name, ms / max(1, n_iter)))
    # the reference subtracts load from wall time; keep its six lines and their order (:656-661)
    out("Total data loading time: ***", dload_ms, " ms")
    out("Total data loading time: ***", dload_ms / n_iter, " ms/iter")
    out("Total computation time: ***", (total_ms - dload_ms), " ms")
    out("Total computation time: ***", (total_ms - dload_ms) / n_iter, " ms/iter")
    out("Total execution time: ***", total_ms, " ms")
    out("Total execution time: ***", total_ms / n_iter, " ms/iter")
    return {"dload_ms": dload_ms, "compute_ms": total_ms - dload_ms, "total_ms": total_ms,
            "device_compute_ms": compute_ms, "iterations": n_iter}


def main(argv=None):
    from deeprecsys_b200.utils.utils import cli
    run(cli(argv))
    return 0


if __name__ == "__main__":
    sys.exit(main())

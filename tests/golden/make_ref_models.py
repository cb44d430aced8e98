#!/usr/bin/env python
"""Record OUTPUTS OF THE REFERENCE'S OWN MODEL CODE as golden vectors (run in the
authoring container only; /root/reference does not exist on the GPU box).

    python tests/golden/make_ref_models.py

The reference's unmodified builders (/root/reference/models/{dlrm_s_caffe2,
wide_and_deep,ncf,din,multi_task_wnd}.py) are imported and executed exactly the
way its CPU engine drives them (/root/reference/inferenceEngine.py:72-88 set-up,
:200-215 request slicing, run_queues -> run) on top of oracle/caffe2_shim — a
numpy stand-in for the Caffe2 operator runtime, which is the one piece that
cannot be installed here (torch==1.4.0+cu92, build/pip_requirements.txt:28).
Graph construction, blob names, weight draw order, tril indices, queue plumbing
and slicing are therefore the reference's code; only the operator kernels are
restated (oracle/caffe2_shim/caffe2/python/_ops.py).

For one case the script additionally runs the reference's unmodified
``inferenceEngine()`` process body (threads + Queues, ServiceRequest in,
ServiceResponse out) and asserts that it leaves the same ``prob_click`` in the
workspace as the direct flow.

Writes tests/golden/ref_models.npz: for each case the argv, the request
(batch_id, batch_size) and every activation blob of the net.
"""
import os
import sys
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "oracle", "caffe2_shim"))
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)

from caffe2.python import workspace                            # noqa: E402  (shim)
import inferenceEngine as ref_engine                           # noqa: E402  (reference, unmodified)
from data_generator.dlrm_data_caffe2 import DLRMDataGenerator  # noqa: E402  (reference)
from utils.packets import ServiceRequest                       # noqa: E402  (reference)
from utils.utils import cli as ref_cli                         # noqa: E402  (reference)

from tests.helpers import SMALL                                # noqa: E402  (argv of the small cases)

WRAPPERS = {
    "dlrm": (ref_engine.DLRM_Wrapper, "dlrm"),
    "wnd": (ref_engine.Wide_and_Deep_Wrapper, "wnd"),
    "ncf": (ref_engine.NCF_Wrapper, "ncf"),
    "din": (ref_engine.DIN_Wrapper, "din"),
    "mtwnd": (ref_engine.MT_Wide_and_Deep_Wrapper, "mtwnd"),
    "dien": (ref_engine.DIEN_Wrapper, "dien"),
}

EXTRA = {}
# the reference's MT-WnD queue mode is broken (multi_task_wnd.py:111 calls FeedBlobWrapper
# with one argument), so it is recorded through the standalone path
# (inferenceEngine.py:137-173: model.mtwnd.run(lX[j], lS_l[j], lS_i[j]) on whole batches)
STANDALONE = ("mtwnd", "dlrm_cat")

MAX_BATCH = 12
REQUESTS = [(1, 7), (0, 12), (1, 1)]          # (batch_id, batch_size)


def run_ref_cli(argv):
    old = sys.argv
    sys.argv = ["prog"] + list(argv)
    try:
        return ref_cli()
    finally:
        sys.argv = old


def case_argv(name, queue=True):
    base = SMALL.get(name) or EXTRA[name]
    return list(base) + ["--num_batches", "2", "--max_mini_batch_size", str(MAX_BATCH),
                         "--numpy_rand_seed", "123"] + (["--queue"] if queue else [])


def slice_like_engine(args, lX, lS_l, lS_i, batch_id, batch_size):
    """inferenceEngine.py:200-215, verbatim semantics."""
    lS_l_curr = np.transpose(np.array(lS_l[batch_id]))
    lS_l_curr = np.transpose(np.array(lS_l_curr[:batch_size]))
    lS_ids_curr = np.array(lS_i[batch_id])
    lS_ids_curr = np.array(lS_ids_curr[:][:, :batch_size * args.num_indices_per_lookup])
    return lS_ids_curr, lS_l_curr, lX[batch_id][:batch_size]


def activation_blobs(net):
    """The float blobs the net's ops write: pooled vectors, interaction, post-activation FC
    outputs, Sum/Concat results."""
    out = {}
    for op in net.ops:
        for name in op.output:
            v = workspace.FetchBlob(name)
            if name.endswith(("_info", "_y", "_matmul", "_flatten_all", "_cat_axis0")):
                continue                       # split infos, pre-activations, dot intermediates
            if isinstance(v, np.ndarray) and v.dtype == np.float32:
                out[name] = v.copy()
    return out


def direct_flow(name, queue=True):
    argv = case_argv(name, queue)
    args = run_ref_cli(argv)
    workspace.ResetWorkspace()
    np.random.seed(args.numpy_rand_seed)                       # inferenceEngine.py:72
    datagen = DLRMDataGenerator(args)
    nb, lX, lS_l, lS_i = datagen.generate_input_data()         # :81-83
    nb, lT = datagen.generate_output_data()                    # :84
    wrapper_cls, attr = WRAPPERS[args.model_type]
    model = wrapper_cls(args)                                  # :87
    model.create(lX[0], lS_l[0], lS_i[0], lT[0])               # :88
    net = getattr(model, attr)
    res = {}
    if queue:
        for batch_id, B in REQUESTS:
            ids, lens, dense = slice_like_engine(args, lX, lS_l, lS_i, batch_id, B)
            model.run_queues(ids, lens, dense, B)              # :211
            net.run()                                          # run_model thread, :33-43
            for k, v in activation_blobs(net.model.net).items():
                res["%s/req_%d_%d/%s" % (name, batch_id, B, k)] = v
        tag = name
    else:
        tag = name + "@standalone"
        for j in range(nb):                                    # inferenceEngine.py:142-161
            net.run(lX[j], lS_l[j], lS_i[j])
            for k, v in activation_blobs(net.model.net).items():
                res["%s/req_%d_%d/%s" % (tag, j, MAX_BATCH, k)] = v
    if args.model_type == "dien":
        # the recurrent cells' parameters come from param_init_net (XavierFill / ConstantFill,
        # dien.py:318-380 via rnn_cell.BasicRNN), not from the numpy stream: record them
        for scope in ("rnn_0", "rnn_1"):
            for key in ("i2h_w", "i2h_b", "gates_t_w", "gates_t_b"):
                res["%s/params/%s/%s" % (tag, scope, key)] = workspace.FetchBlob(
                    "%s/%s" % (scope, key)).copy()
    res[tag + "/argv"] = np.array(argv)
    res[tag + "/op_types"] = np.array([op.type for op in net.model.net.ops])
    return res, args


def engine_flow(name, batch_id, B):
    """The reference's unmodified inferenceEngine() body with real Queues."""
    from multiprocessing import Queue
    args = run_ref_cli(case_argv(name))
    workspace.ResetWorkspace()
    req_q, resp_q, ready_q = Queue(), Queue(), Queue()
    req = ServiceRequest(batch_id=batch_id, batch_size=B, epoch=0, sub_id=0, total_sub_batches=1)
    th = threading.Thread(target=ref_engine.inferenceEngine, args=(args, req_q, 0, resp_q, ready_q))
    th.daemon = True
    th.start()
    assert ready_q.get(timeout=60) is True
    req_q.put(req)
    resp = resp_q.get(timeout=60)
    out = workspace.FetchBlob("prob_click").copy()
    req_q.put(None)
    assert resp.out_batch_size == B, (resp.out_batch_size, B)
    return out


def main():
    arrays = {}
    for name in sorted(SMALL):
        if name in STANDALONE and name != "dlrm_cat":
            continue                                   # queue mode broken in the reference
        res, args = direct_flow(name)
        arrays.update(res)
        print("%-16s %-5s blobs/request=%d" % (name, args.model_type,
                                                sum(1 for k in res if "/req_1_7/" in k)))
    for name in STANDALONE:
        res, args = direct_flow(name, queue=False)
        arrays.update(res)
        print("%-16s %-5s standalone arrays=%d" % (name, args.model_type, len(res)))
    # standalone and queue mode agree where both exist (whole batch 0)
    a = arrays["dlrm_cat@standalone/req_0_%d/prob_click" % MAX_BATCH]
    b = arrays["dlrm_cat/req_0_%d/prob_click" % MAX_BATCH]
    assert np.array_equal(a, b), "standalone path != queue path"
    # the engine body itself reproduces the direct flow
    bid, B = REQUESTS[0]
    out_engine = engine_flow("dlrm_cat", bid, B)
    key = "dlrm_cat/req_%d_%d/prob_click" % (bid, B)
    assert np.array_equal(out_engine, arrays[key]), "inferenceEngine() != direct flow"
    arrays["dlrm_cat/engine_body_prob_click"] = out_engine
    print("reference inferenceEngine() body reproduces the direct flow bit-for-bit")
    path = os.path.join(HERE, "ref_models.npz")
    np.savez_compressed(path, **arrays)
    print("wrote %s (%.1f KB, %d arrays)" % (path, os.path.getsize(path) / 1024.0, len(arrays)))


if __name__ == "__main__":
    main()

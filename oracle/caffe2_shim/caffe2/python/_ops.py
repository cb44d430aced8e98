"""numpy restatement of the Caffe2 operators the reference instantiates
(operator schemas of caffe2 as bundled in torch==1.4.0; call sites cited per op).
fp32 throughout; SparseLengthsSum accumulates sequentially per column like the
CPU perfkernel (EmbeddingLookup).  TEST INFRASTRUCTURE ONLY."""
from __future__ import absolute_import, division, print_function

import numpy as np

F32 = np.float32


def _ws():
    from caffe2.python import workspace
    return workspace


def op_SparseLengthsSum(ins, outs, args):
    # models/dlrm_s_caffe2.py:319-325  SparseLengthsSum([DATA, INDICES, LENGTHS]) -> [B, D]
    W, idx, lens = ins
    W = np.asarray(W)
    idx = np.asarray(idx).astype(np.int64).ravel()
    lens = np.asarray(lens).astype(np.int64).ravel()
    if lens.sum() != idx.shape[0]:
        raise RuntimeError("SparseLengthsSum enforce: sum(LENGTHS)=%d != len(INDICES)=%d"
                           % (lens.sum(), idx.shape[0]))
    if idx.size and (idx.min() < 0 or idx.max() >= W.shape[0]):
        raise RuntimeError("SparseLengthsSum enforce: index out of range")
    out = np.zeros((lens.shape[0],) + W.shape[1:], dtype=W.dtype)
    pos = 0
    for b in range(lens.shape[0]):
        acc = np.zeros(W.shape[1:], dtype=W.dtype)
        for k in range(int(lens[b])):
            acc = acc + W[idx[pos + k]]          # one fp32 rounding per row, in index order
        pos += int(lens[b])
        out[b] = acc
    return [out]


def op_FC(ins, outs, args):
    # models/dlrm_s_caffe2.py:258-264  FC([X, W, b]): Y = X . W^T + b, W [out, in]
    X, W, b = (np.asarray(v, dtype=F32) for v in ins)
    axis = int(args.get("axis", 1))
    lead = X.shape[:axis]
    X2 = X.reshape(int(np.prod(lead)) if lead else 1, -1)
    Y = (X2 @ W.reshape(W.shape[0], -1).T).astype(F32) + b
    return [Y.astype(F32).reshape(tuple(lead) + (W.shape[0],))]


def op_Relu(ins, outs, args):
    return [np.maximum(np.asarray(ins[0]), F32(0))]


def op_Sigmoid(ins, outs, args):
    x = np.asarray(ins[0], dtype=F32)
    return [(F32(1) / (F32(1) + np.exp(-x, dtype=F32))).astype(F32)]


def op_Tanh(ins, outs, args):
    return [np.tanh(np.asarray(ins[0], dtype=F32)).astype(F32)]


def op_Concat(ins, outs, args):
    # Concat(axis=1 default for NCHW, add_axis=0) -> (concat_result, split_info)
    axis = int(args.get("axis", 1))
    xs = [np.asarray(v) for v in ins]
    if int(args.get("add_axis", 0)):
        y = np.stack(xs, axis=axis)
        info = np.ones(len(xs), dtype=np.int32)
    else:
        y = np.concatenate(xs, axis=axis)
        info = np.array([v.shape[axis] for v in xs], dtype=np.int32)
    return [y, info]


def op_BatchMatMul(ins, outs, args):
    # models/dlrm_s_caffe2.py:344
    A, B = (np.asarray(v, dtype=F32) for v in ins)
    if int(args.get("trans_a", 0)):
        A = np.swapaxes(A, -1, -2)
    if int(args.get("trans_b", 0)):
        B = np.swapaxes(B, -1, -2)
    return [np.matmul(A, B).astype(F32)]


def op_Flatten(ins, outs, args):
    x = np.asarray(ins[0])
    axis = int(args.get("axis", 1))
    return [x.reshape(int(np.prod(x.shape[:axis])) if axis > 0 else 1, -1)]


def op_BatchGather(ins, outs, args):
    # models/dlrm_s_caffe2.py:350  out[b, j] = DATA[b, INDICES[j]]
    data, idx = np.asarray(ins[0]), np.asarray(ins[1]).astype(np.int64)
    return [np.take(data, idx, axis=1)]


def op_Sum(ins, outs, args):
    # ncf.py:303, din.py:263,283: n-ary elementwise sum, left to right
    acc = np.asarray(ins[0]).copy()
    for v in ins[1:]:
        acc = acc + np.asarray(v)
    return [acc]


def op_Cast(ins, outs, args):
    to = args.get("to")
    from caffe2.proto.caffe2_pb2 import TensorProto as TP
    dt = {TP.FLOAT: np.float32, TP.INT32: np.int32, TP.INT64: np.int64}[to]
    return [np.asarray(ins[0]).astype(dt)]


def op_Reshape(ins, outs, args):
    x = np.asarray(ins[0])
    shape = list(args["shape"])
    for i, s in enumerate(shape):
        if s == 0:
            shape[i] = x.shape[i]
    return [x.reshape(shape), np.array(x.shape, dtype=np.int64)]


def op_Softmax(ins, outs, args):
    x = np.asarray(ins[0], dtype=F32)
    axis = int(args.get("axis", 1))
    x2 = x.reshape(int(np.prod(x.shape[:axis])), -1)
    e = np.exp(x2 - x2.max(axis=1, keepdims=True), dtype=F32)
    return [(e / e.sum(axis=1, keepdims=True, dtype=F32)).astype(F32).reshape(x.shape)]


def op_Add(ins, outs, args):
    return [np.asarray(ins[0]) + np.asarray(ins[1])]


# ---- parameter initialisers (brew / rnn_cell put these into model.param_init_net) -------------
def op_XavierFill(ins, outs, args):
    # Caffe2 XavierFill: U(-sqrt(3 / fan_in), +sqrt(3 / fan_in)), fan_in = size / dim0.
    # Drawn from numpy's global RNG here (Caffe2 uses its own generator).
    shape = tuple(int(v) for v in args["shape"])
    fan_in = int(np.prod(shape)) // shape[0]
    scale = np.sqrt(3.0 / fan_in)
    return [np.random.uniform(-scale, scale, size=shape).astype(F32)]


def op_ConstantFill(ins, outs, args):
    shape = tuple(int(v) for v in args["shape"])
    return [np.full(shape, args.get("value", 0.0), dtype=F32)]


def op_BasicRNNShim(ins, outs, args):
    # rnn_cell.BasicRNN(activation="tanh"), see caffe2/python/rnn_cell.py in this shim
    x, seq_len, h0, wi, bi, wg, bg = (np.asarray(v) for v in ins)
    x = x.astype(F32)
    N, B, _ = x.shape
    H = wg.shape[0]
    h = h0.astype(F32).reshape(-1, H)[-B:] if h0.ndim == 3 else h0.astype(F32).reshape(B, H)
    i2h = (x.reshape(N * B, -1) @ wi.T.astype(F32) + bi).astype(F32).reshape(N, B, H)
    seq_len = np.asarray(seq_len).astype(np.int64).ravel()
    all_h = np.zeros((N, B, H), dtype=F32)
    for t in range(N):
        gates = ((h @ wg.T.astype(F32)).astype(F32) + bg).astype(F32)     # FC(h_{t-1})
        gates = (gates + i2h[t]).astype(F32)                               # Sum(gates_t, input_t)
        h_new = np.tanh(gates).astype(F32)
        valid = (seq_len > t)[:, None]
        if int(args.get("drop_states", 0)):
            h = np.where(valid, h_new, F32(0)).astype(F32)
        else:
            h = np.where(valid, h_new, h).astype(F32)
        all_h[t] = h
    return [all_h, h.reshape(1, B, H).copy()]


# ---- blob queues (dlrm_s_caffe2.py:180-211, 257, 307-314) ----------------------------------
def op_CreateBlobsQueue(ins, outs, args):
    import queue
    return [queue.Queue(maxsize=int(args.get("capacity", 0)))]


def op_EnqueueBlobs(ins, outs, args):
    q = ins[0]
    for v in ins[1:]:
        q.put(np.array(v, copy=True))
    return [ins[1 + i] for i in range(len(outs))]


def op_DequeueBlobs(ins, outs, args):
    q = ins[0]
    return [q.get() for _ in outs]            # blocks like the operator does


def run_op(op, blobs):
    fn = globals().get("op_" + op.type)
    if fn is None:
        raise NotImplementedError("caffe2 shim: operator %s not restated" % op.type)
    try:
        ins = [blobs[name] for name in op.input]
    except KeyError as exc:
        raise RuntimeError("caffe2 shim: blob %s not found (op %s)" % (exc, op.type))
    res = fn(ins, op.output, op.args)
    for name, val in zip(op.output, res):
        blobs[name] = val

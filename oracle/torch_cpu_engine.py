"""CPU ORACLE / CPU BASELINE — TEST INFRASTRUCTURE ONLY (see numpy_oracle.py).

torch-CPU restatement of the net the reference's CPU engine runs
(/root/reference/inferenceEngine.py + models/*.py on Caffe2): the same operator
graph with ``embedding_bag(mode="sum")`` standing in for SparseLengthsSum (same
perfkernel lineage, bit-identical to sequential fp32 accumulation — checked in
tests/test_oracle.py) and ``addmm`` (MKL sgemm) for FC.  It exists because
``caffe2.python`` (torch==1.4.0+cu92, build/pip_requirements.txt:28) cannot be
imported in this image, so this is the *port* that is timed as the CPU baseline
(bench.py ``cpu_baseline`` and ``--impl reference``) and that stands in for
``inferenceEngine`` in the protocol tests.  Parity status: unpinned (no
reference outputs exist), cross-checked against numpy_oracle.py.
"""
from __future__ import absolute_import, division, print_function

import time

import numpy as np
import torch
import torch.nn.functional as F

MLP_BOT, MLP_TOP, MLP_FINAL, MLP_ATTN_BASE, MLP_TASK_BASE, MLP_RNN_BASE = 0, 1, 2, 16, 4096, 8192


class TorchCpuModel(object):
    def __init__(self, spec, tables, fcs):
        self.spec = spec
        self.tables = [torch.from_numpy(np.ascontiguousarray(W)) for W in tables]
        self.fcs = {k: (torch.from_numpy(np.ascontiguousarray(W)),
                        torch.from_numpy(np.ascontiguousarray(b))) for k, (W, b) in fcs.items()}

    def _layers(self, mlp_id):
        out, k = [], 0
        while (mlp_id, k) in self.fcs:
            out.append(self.fcs[(mlp_id, k)])
            k += 1
        return out

    @staticmethod
    def _mlp(x, layers, sigmoid_last=False, sigmoid_layer=0):
        n = len(layers)
        for i, (W, b) in enumerate(layers):
            x = torch.addmm(b, x, W.t())
            sig = (sigmoid_last and i == n - 1) or (sigmoid_layer and i == sigmoid_layer - 1)
            x = torch.sigmoid(x) if sig else torch.relu(x)
        return x

    def pool(self, ids, lengths):
        out = []
        for W, i, l in zip(self.tables, ids, lengths):
            i = torch.as_tensor(np.asarray(i), dtype=torch.int64)
            l = torch.as_tensor(np.asarray(l), dtype=torch.int64)
            off = torch.zeros(l.shape[0], dtype=torch.int64)
            if l.shape[0] > 1:
                off[1:] = torch.cumsum(l, 0)[:-1]
            out.append(F.embedding_bag(i, W, off, mode="sum"))
        return out

    @torch.no_grad()
    def forward(self, dense, ids, lengths):
        s = self.spec
        pooled = self.pool(ids, lengths)
        if s.model_type == "dlrm":
            x = self._mlp(torch.as_tensor(dense), self._layers(MLP_BOT))
            if s.interaction_op == "dot":
                T3 = torch.stack([x] + pooled, dim=1)
                Z = torch.bmm(T3, T3.transpose(1, 2))
                Fn = T3.shape[1]
                off = 1 if s.interaction_itself else 0
                li = torch.tensor([i for i in range(Fn) for j in range(i + off)])
                lj = torch.tensor([j for i in range(Fn) for j in range(i + off)])
                R = torch.cat([x, Z[:, li, lj]], dim=1)
            else:
                R = torch.cat([x] + pooled, dim=1)
            return self._mlp(R, self._layers(MLP_TOP), sigmoid_last=True)
        if s.model_type == "wnd":
            R = torch.cat([torch.as_tensor(dense)] + pooled, dim=1)
            return self._mlp(R, self._layers(MLP_TOP), sigmoid_last=True)
        if s.model_type == "mtwnd":
            R = torch.cat([torch.as_tensor(dense)] + pooled, dim=1)
            shared = self._mlp(R, self._layers(MLP_TOP))
            heads = [self._mlp(shared, self._layers(MLP_TASK_BASE + k),
                               sigmoid_layer=s.task_sigmoid_layer) for k in range(s.n_tasks)]
            return torch.cat(heads, dim=1)
        if s.model_type == "dien":
            T, N = s.T, s.T - 3
            x = torch.cat(pooled[1:T - 2], dim=1).reshape(N, -1, s.D)
            for cell in range(2):
                (wi, bi), (wg, bg) = self._layers(MLP_RNN_BASE + cell)
                i2h = torch.addmm(bi, x.reshape(-1, x.shape[2]), wi.t()).reshape(N, -1, wg.shape[0])
                h = torch.zeros(i2h.shape[1], wg.shape[0])
                outs = []
                for t in range(N):
                    h = torch.tanh(torch.addmm(bg, h, wg.t()) + i2h[t])
                    outs.append(h)
                x = torch.stack(outs, dim=0)
            R = torch.cat([x[-1], pooled[0], pooled[T - 2], pooled[T - 1]], dim=1)
            return self._mlp(R, self._layers(MLP_TOP))
        if s.model_type == "ncf":
            zmf = pooled[0] + pooled[1]
            tower = self._mlp(torch.cat([pooled[2], pooled[3]], dim=1), self._layers(MLP_TOP))
            return self._mlp(torch.cat([zmf, tower], dim=1), self._layers(MLP_FINAL))
        if s.model_type == "din":
            T = s.T
            ad, atten = pooled[T - 2], None
            for i in range(T - 3):
                eu = pooled[1 + i]
                o = self._mlp(torch.cat([eu, ad, eu + ad], dim=1), self._layers(MLP_ATTN_BASE + i))
                atten = o if atten is None else atten + o
            R = torch.cat([pooled[0], atten, ad, pooled[T - 1]], dim=1)
            return self._mlp(R, self._layers(MLP_TOP))
        raise ValueError(s.model_type)


def time_forward(model, dense, ids, lengths, min_seconds=10.0, max_seconds=30.0, threads=None):
    """Run forward on the given sample repeatedly for ~min_seconds of CPU work.
    Returns (samples_per_second, reps, seconds, threads)."""
    if threads:
        torch.set_num_threads(int(threads))
    n = int(np.asarray(lengths[0]).shape[0])
    ids_t = [torch.as_tensor(np.asarray(i), dtype=torch.int64) for i in ids]
    model.forward(dense, ids_t, lengths)                  # warm-up
    reps, t0 = 0, time.perf_counter()
    while True:
        model.forward(dense, ids_t, lengths)
        reps += 1
        dt = time.perf_counter() - t0
        if dt >= min_seconds or dt >= max_seconds:
            break
    return n * reps / dt, reps, dt, torch.get_num_threads()

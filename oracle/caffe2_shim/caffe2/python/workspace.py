"""caffe2.python.workspace stand-in: one global blob dict, eager numpy execution.
TEST INFRASTRUCTURE ONLY."""
from __future__ import absolute_import, division, print_function

import numpy as np

from caffe2.python import _import_c_extension as C  # noqa: F401  (workspace.C.benchmark_net)
from caffe2.python import _ops, core

_blobs = {}
_nets = {}


def GlobalInit(_args):
    return True


def ResetWorkspace():
    _blobs.clear()
    _nets.clear()
    return True


def FeedBlob(name, arr, device_option=None):
    _blobs[str(name)] = np.array(arr)             # FeedBlob copies into the workspace
    return True


def FetchBlob(name):
    return _blobs[str(name)]


def HasBlob(name):
    return str(name) in _blobs


def Blobs():
    return sorted(_blobs)


def _as_net(net):
    if isinstance(net, core.Net):
        return net
    if isinstance(net, core._Proto):
        return net._net
    if hasattr(net, "net") and isinstance(net.net, core.Net):   # ModelHelper
        return net.net
    return _nets[str(net)]


def RunNetOnce(net):
    for op in _as_net(net).ops:
        _ops.run_op(op, _blobs)
    return True


def CreateNet(net, overwrite=False):
    n = _as_net(net)
    _nets[n.Name()] = n
    return True


def RunNet(net, num_iter=1):
    n = _as_net(net)
    for _ in range(num_iter):
        for op in n.ops:
            _ops.run_op(op, _blobs)
    return True

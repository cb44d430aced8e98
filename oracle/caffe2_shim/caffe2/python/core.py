"""caffe2.python.core stand-in: nets are op lists built by attribute access,
executed by workspace.RunNet*.  TEST INFRASTRUCTURE ONLY."""
from __future__ import absolute_import, division, print_function

import contextlib

from caffe2.proto import caffe2_pb2

DataType = caffe2_pb2.TensorProto


class BlobReference(str):
    """A blob name (the reference concatenates these with strings)."""
    def Net(self):
        return getattr(self, "_net", None)


class _Device(object):
    def __init__(self, device_type, device_id=0):
        self.device_type, self.device_id = device_type, device_id


def DeviceOption(device_type, device_id=0, **_):
    return _Device(device_type, device_id)


@contextlib.contextmanager
def DeviceScope(_option):
    yield


@contextlib.contextmanager
def NameScope(_prefix):
    yield


class Op(object):
    def __init__(self, type_, inputs, outputs, args):
        self.type, self.input, self.output, self.args = type_, inputs, outputs, args


class _Proto(object):
    def __init__(self, net):
        self._net = net
        self.type = "simple"
        self.num_workers = 1
        self.name = net._name

    @property
    def op(self):
        return self._net.ops


def _names(x):
    if x is None:
        return []
    if isinstance(x, (str, bytes)):
        return [str(x)]
    return [str(v) for v in x]


class Net(object):
    _counter = 0

    def __init__(self, name):
        self._name = str(name)
        self.ops = []
        self._proto = _Proto(self)

    def Name(self):
        return self._name

    def Proto(self):
        return self._proto

    def NextName(self, prefix="blob"):
        Net._counter += 1
        return "%s/%s_%d" % (self._name, prefix, Net._counter)

    def _add(self, type_, inputs, outputs, kwargs):
        ins = _names(inputs)
        if outputs is None:
            outs = [self.NextName(type_)]
        elif isinstance(outputs, int):
            outs = [self.NextName(type_) for _ in range(outputs)]
        else:
            outs = _names(outputs)
        self.ops.append(Op(type_, ins, outs, dict(kwargs)))
        refs = [BlobReference(o) for o in outs]
        if len(refs) == 1:
            return refs[0]
        return tuple(refs)

    def __getattr__(self, type_):
        if type_.startswith("_") or not type_[0].isupper():
            raise AttributeError(type_)

        def make(inputs=None, outputs=None, **kwargs):
            return self._add(type_, inputs, outputs, kwargs)
        return make

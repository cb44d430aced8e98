"""caffe2.python.brew stand-in — the three helpers DIEN uses (models/dien.py:336-355).
Parameters are created the way brew does it: fill ops appended to model.param_init_net
(XavierFill for weights, ConstantFill 0 for biases), executed when the reference runs
``workspace.RunNetOnce(model.param_init_net)``.  TEST INFRASTRUCTURE ONLY."""
from __future__ import absolute_import, division, print_function


def fc(model, blob_in, blob_out, dim_in, dim_out, axis=1, **kwargs):
    if blob_out is None:
        blob_out = model.net.NextName("fc")
    w, b = str(blob_out) + "_w", str(blob_out) + "_b"
    model.param_init_net.XavierFill([], w, shape=[dim_out, dim_in])
    model.param_init_net.ConstantFill([], b, shape=[dim_out], value=0.0)
    model.params.extend([w, b])
    return model.net.FC([blob_in, w, b], blob_out, axis=axis)


def softmax(model, blob_in, blob_out=None, **kwargs):
    return model.net.Softmax(blob_in, blob_out, **kwargs)


def sum(model, blob_in, blob_out, **kwargs):          # noqa: A001  (brew's own name)
    return model.net.Sum(blob_in, blob_out)

"""caffe2.python.brew — only DIEN uses it (models/dien.py:336-355); not restated."""


def _unsupported(*_a, **_k):
    raise NotImplementedError("caffe2 shim: brew helpers (DIEN) are not restated")


fc = softmax = sum = _unsupported

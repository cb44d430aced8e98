"""caffe2.python.rnn_cell — only DIEN uses it (models/dien.py:318,370); not restated."""


def BasicRNN(*_a, **_k):
    raise NotImplementedError("caffe2 shim: rnn_cell.BasicRNN (DIEN) is not restated")

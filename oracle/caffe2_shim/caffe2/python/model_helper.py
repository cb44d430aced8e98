"""caffe2.python.model_helper stand-in.  TEST INFRASTRUCTURE ONLY."""
from caffe2.python import core


class ModelHelper(object):
    def __init__(self, name=None, init_params=True, **_):
        self.name = name or "model"
        self.net = core.Net(self.name)
        self.param_init_net = core.Net(self.name + "_init")
        self.init_params = init_params
        self.params = []

    def Proto(self):
        return self.net.Proto()

"""Device-type constants the reference passes to core.DeviceOption."""
CPU = 0
CUDA = 1


class TensorProto(object):
    FLOAT = 1
    INT32 = 2
    INT64 = 10

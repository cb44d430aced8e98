"""caffe2 stand-in (oracle/caffe2_shim/README.md) — test infrastructure only."""
__shim__ = True

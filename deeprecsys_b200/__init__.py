"""deeprecsys_b200 — B200-native accelerator backend for DeepRecSys' per-query
inference hot path (SparseLengthsSum gathers -> interaction -> MLPs).

Layout:
  csrc/        CUDA kernels (sm_100a) and the C-ABI implementation
  lib/         built libdrs_b200.so (git-ignored)
  _lib.py      ctypes binding of include/drs_b200.h (no fallback)
  engine.py    AccelEngine: spec -> C-ABI; mirrors *_Wrapper.run_queues / *_Net.run
  spec.py      model dims + parameter draw order of the reference builders
  accelInferenceEngine.py   drop-in process body for DeepRecSys.py
  utils/, data_generator/, models/configs/   the reference's spec surface
"""
__version__ = "0.1.0"

"""CPU ORACLE — TEST INFRASTRUCTURE ONLY (see numpy_oracle.py).

Two stand-ins used by the protocol tests (tests/test_protocol.py), which run the
*unmodified* reference DeepRecSys.py / loadGenerator.py / scheduler.py:

* ``inferenceEngine`` — caffe2-free restatement of the reference CPU engine
  process (/root/reference/inferenceEngine.py:62-232): same 5-argument
  signature, same seed -> datagen -> parameter order, same request slicing
  (:200-215), same ServiceResponse fields (:216-230), torch-CPU arithmetic.
* ``OracleBackend`` — replaces deeprecsys_b200.accelInferenceEngine.AccelBackend
  in the CPU test tier (there is no GPU in the authoring container), so the
  accel engine's Queue protocol can be exercised end to end.  It is never
  reachable from the product path: only tests install it.
"""
from __future__ import absolute_import, division, print_function

import os
import time

import numpy as np

from deeprecsys_b200.data_generator.dlrm_data import DLRMDataGenerator, slice_request
from deeprecsys_b200.spec import ModelSpec, draw_all_parameters
from oracle.torch_cpu_engine import TorchCpuModel

try:
    from utils.packets import ServiceResponse
except ImportError:
    from deeprecsys_b200.utils.packets import ServiceResponse


class OracleBackend(object):
    def __init__(self, args):
        self.args = args
        gen = DLRMDataGenerator(args)
        (self.nbatches, self.lX, self.lS_l, self.lS_i) = gen.generate_input_data()
        (_, self.lT) = gen.generate_output_data()
        self.spec = ModelSpec(args)
        tables, fcs = draw_all_parameters(self.spec)
        self.model = TorchCpuModel(self.spec, tables, fcs)
        self.dump_dir = os.environ.get("DRS_ACCEL_DUMP_DIR")

    def serve(self, batch_id, batch_size):
        d, l, i = slice_request(self.lX, self.lS_l, self.lS_i, batch_id, batch_size, self.spec.L)
        dense = d if self.spec.m_den > 0 else None
        return self.model.forward(dense, list(i), list(l)).numpy()

    def serve_inputs(self, request):
        """A TensorServiceRequest: the inputs come with the request (inline arrays or a slot of
        a shared-memory ring) instead of from the pre-generated dataset."""
        import torch
        B = int(request.batch_size)
        if request.ring is not None:
            from deeprecsys_b200.utils.input_ring import SharedInputRing
            if not hasattr(self, "_rings"):
                self._rings = {}
            ring = self._rings.get(request.ring["path"])
            if ring is None:
                ring = self._rings[request.ring["path"]] = SharedInputRing.attach(request.ring)
            dense, ids, lengths = ring.read(request.slot, B, request.counts)
        else:
            dense, ids, lengths = request.dense, request.ids, request.lengths
        dense = np.array(dense[:B]) if self.spec.m_den > 0 else None
        ids = [torch.as_tensor(np.array(i), dtype=torch.int64) for i in ids]
        lengths = [np.array(l[:B], dtype=np.int32) for l in lengths]
        if request.ring is not None:
            ring.release(request.slot)          # copied out: the producer may reuse the slot
        return self.model.forward(dense, ids, lengths).numpy()

    def close(self):
        pass


def inferenceEngine(args, requestQueue=None, engine_id=None, responseQueue=None,
                    inferenceEngineReadyQueue=None):
    np.random.seed(args.numpy_rand_seed)
    backend = OracleBackend(args)
    if requestQueue is None:
        t0 = time.time()
        for _ in range(args.nepochs):
            for j in range(backend.nbatches):
                backend.serve(j, args.max_mini_batch_size)
        print("Total execution time: ***", (time.time() - t0) * 1000., " ms")
        return
    inferenceEngineReadyQueue.put(True)
    while True:
        request = requestQueue.get()
        if request is None:
            responseQueue.put(None)
            return
        start = time.time()
        out = backend.serve(request.batch_id, request.batch_size)
        end = time.time()
        responseQueue.put(ServiceResponse(
            consumer_id=engine_id, epoch=request.epoch, batch_id=request.batch_id,
            batch_size=request.batch_size, arrival_time=request.arrival_time,
            process_start_time=start, queue_end_time=end, inference_end_time=end,
            out_batch_size=out.shape[0], total_sub_batches=request.total_sub_batches,
            exp_packet=request.exp_packet, sub_id=request.sub_id))

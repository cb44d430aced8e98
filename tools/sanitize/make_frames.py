"""Frames for wire_asan.cpp: 200 pickled ServiceRequests with mixed field types -> frames.bin"""
import os
import pickle
import random
import struct
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from deeprecsys_b200.utils.packets import ServiceRequest  # noqa: E402

random.seed(1)
blob = b""
for i in range(200):
    r = ServiceRequest(batch_id=random.choice([i, 2 ** 40 + i, None, True]),
                       epoch=random.randint(-70000, 70000),
                       arrival_time=random.choice([None, 1.5 * i]),
                       batch_size=random.randint(0, 70000), sub_id=random.choice([0, -2 ** 33]),
                       total_sub_batches=1, exp_packet=random.choice([True, False, None]))
    d = pickle.dumps(r)
    blob += struct.pack("!i", len(d)) + d
open(sys.argv[1], "wb").write(blob)

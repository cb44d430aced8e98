"""Wire packets exchanged over the multiprocessing Queues.

Mirrors the attribute surface of the reference packets
(/root/reference/utils/packets.py:6-22 ``ServiceRequest`` and :32-59
``ServiceResponse``).  The orchestrator logs ``response.__dict__``
(DeepRecSys.py:138) and joins sub-batches by (epoch, batch_id, exp_packet)
(DeepRecSys.py:101-118), so the attribute *names* are the contract: they are
kept exactly, including the reference's quirk that the ``process_start_time``
keyword is stored as ``queue_start_time``.

Packets carry no tensors: a request names a pre-generated
batch by ``batch_id`` and asks for its first ``batch_size`` rows.
``TensorServiceRequest`` is this backend's extension for queries that bring
their own inputs; the reference classes stay as they are.
"""

_REQUEST_FIELDS = ("batch_id", "batch_size", "epoch", "arrival_time",
                   "total_sub_batches", "sub_id", "exp_packet")


class ServiceRequest(object):
    """One (sub-)query: 'run rows [0, batch_size) of batch ``batch_id``'."""

    def __init__(self, batch_id=None, epoch=None, arrival_time=None,
                 batch_size=None, sub_id=None, total_sub_batches=None,
                 exp_packet=None):
        given = locals()
        for name in _REQUEST_FIELDS:
            setattr(self, name, given[name])

    def __str__(self):
        # The reference's __str__ raises (str + tuple); this one works.
        return "Request[%s] -> arrival_time %s" % (
            (self.epoch, self.batch_id, self.batch_size), self.arrival_time)


class TensorServiceRequest(ServiceRequest):
    """A request that carries its own inputs instead of naming a pre-generated batch.

    Not in the reference (its packets only name ``batch_id``/``batch_size``,
    utils/packets.py:6-22): an engine that really computes can serve data the client sends.
    Every reference attribute is kept — the orchestrator joins and logs by them — and one of two
    payloads is added:
      * ``ring`` + ``slot`` + ``counts``: the tensors sit in slot ``slot`` of the shared-memory
        ring described by ``ring`` (utils/input_ring.py); ``counts[t]`` = ids of table t;
      * ``dense`` / ``ids`` / ``lengths``: the arrays themselves, pickled with the request
        (simple, slow for large queries).
    The CPU engines of the reference ignore the extra attributes and serve the request from
    ``batch_id`` as before."""

    def __init__(self, dense=None, ids=None, lengths=None, ring=None, slot=None, counts=None,
                 **reference_fields):
        super(TensorServiceRequest, self).__init__(**reference_fields)
        self.dense, self.ids, self.lengths = dense, ids, lengths
        self.ring, self.slot, self.counts = ring, slot, counts

    def has_inputs(self):
        return self.ring is not None or self.ids is not None


class ServiceResponse(object):
    """Timing record for one served request (no tensors)."""

    def __init__(self, consumer_id=None, epoch=None, batch_id=None,
                 batch_size=None, arrival_time=None, process_start_time=None,
                 queue_end_time=None, inference_end_time=None,
                 out_batch_size=None, sub_id=None, total_sub_batches=None,
                 exp_packet=None):
        self.consumer_id = consumer_id
        self.epoch, self.batch_id, self.batch_size = epoch, batch_id, batch_size
        self.arrival_time = arrival_time
        # reference stores the process_start_time kwarg under this name
        self.queue_start_time = process_start_time
        self.queue_end_time = queue_end_time
        self.inference_end_time = inference_end_time
        self.out_batch_size = out_batch_size
        self.total_sub_batches = total_sub_batches
        self.exp_packet = exp_packet
        self.sub_id = sub_id

    def __str__(self):
        key = (self.epoch, self.batch_id, self.batch_size, self.consumer_id)
        return ("Response[%s] -> arrival %s -> queue_start %s -> queue_end %s "
                "-> inference_end %s" % (key, self.arrival_time,
                                         self.queue_start_time,
                                         self.queue_end_time,
                                         self.inference_end_time))

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

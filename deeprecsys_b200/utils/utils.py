"""Flag / JSON spec surface of the accelerator path.

Same flag names, types, defaults and override rules as the reference's
``cli()`` (/root/reference/utils/utils.py:15-165) so that the reference's
run scripts and ``models/configs/*.json`` drive this backend unchanged:

* every flag of utils.py:22-128 exists with the same type and default
  (the flag table below is data, not a transcription of the argparse calls);
* ``--model_type din`` expands ``arch_embedding_size`` with
  ``user_behavior_tables`` copies of the behaviour table *before* the JSON is
  applied (utils.py:132-149), so a ``--config_file`` overrides the expansion
  (SURVEY fact 7);
* the JSON file wins over the command line and each value is coerced with the
  type of the existing attribute; unknown keys raise AttributeError
  (utils.py:154-160);
* numpy's legacy global RNG is seeded last (utils.py:163).

Flags added for the B200 backend (defaults keep old scripts unchanged) are
listed in ``ACCEL_FLAGS``.
"""
from __future__ import absolute_import, division, print_function

import argparse
import json
import sys

import numpy as np

_T = "store_true"

# (name, type-or-"store_true", default)
REFERENCE_FLAGS = (
    # model
    ("arch_sparse_feature_size", int, 2),
    ("arch_embedding_size", str, "4-3-2"),
    ("arch_mlp_bot", str, "4-3-2"),
    ("arch_mlp_top", str, "4-2-1"),
    ("arch_mlp_tasks", str, "4-2-1"),
    ("num_multi_tasks", int, 1),
    ("hidden_size", int, 64),
    ("arch_interaction_op", str, "dot"),
    ("arch_interaction_itself", _T, False),
    ("inter_op_workers", int, 1),
    ("sls_workers", int, 1),
    ("fc_workers", int, 1),
    ("model_type", str, "dlrm"),
    ("user_behavior_tables", int, 1000),
    # inference
    ("inference_only", _T, True),
    ("save_proto_types_shapes", _T, False),
    ("output_log_file", str, None),
    # dataset
    ("num_batches", int, 0),
    ("mini_batch_size", int, 1),
    ("max_mini_batch_size", int, 1),
    ("avg_mini_batch_size", float, 1),
    ("var_mini_batch_size", float, 1),
    ("batch_size_distribution", str, "fixed"),
    ("batch_dist_file", str, "config/batch_distribution.txt"),
    ("sub_task_batch_size", int, 16),
    ("data_generation", str, "random"),
    ("data_trace_file", str, "./input/dist_emb_j.log"),
    ("data_set", str, "kaggle"),
    ("raw_data_file", str, ""),
    ("processed_data_file", str, ""),
    ("data_randomize", str, "total"),
    ("data_trace_enable_padding", bool, False),
    ("num_indices_per_lookup", int, 10),
    ("num_indices_per_lookup_fixed", bool, False),
    # serving system
    ("queue", _T, False),
    ("inference_engines", int, 1),
    ("avg_arrival_rate", float, 10),
    ("target_latency", float, 10),
    ("req_granularity", int, 64),
    ("batch_configs", str, "32-64-128-256-512-1024"),
    ("tune_batch_qps", _T, False),
    ("tune_accel_qps", _T, False),
    ("accel_configs", str, "128-256-512"),
    ("stable_region", float, 0.10),
    ("max_arr_range", float, 100),
    ("min_arr_range", float, 1),
    ("arr_steps", int, 20),
    ("sched_timeout", int, 100),
    # hardware
    ("use_accel", _T, False),
    ("model_accel", _T, False),
    ("accel_request_size_thres", int, 1024),
    ("model_name", str, ""),
    ("accel_root_dir", str, "accelerator/"),
    # activations / loss (unused by inference, kept for script compatibility)
    ("activation_function", str, "relu"),
    ("loss_function", str, "mse"),
    ("loss_threshold", float, 0.0),
    ("round_targets", bool, False),
    # run control
    ("nepochs", int, 1),
    ("learning_rate", float, 0.01),
    ("print_precision", int, 5),
    ("numpy_rand_seed", int, 123),
    ("sync_dense_params", bool, True),
    ("caffe2_net_type", str, "simple"),
    ("engine", str, "TBB"),
    # debugging
    ("print_freq", int, 1),
    ("print_time", _T, False),
    ("debug_mode", _T, False),
    ("enable_profiling", _T, False),
    ("plot_compute_graph", _T, False),
    ("log_file", str, "log/output.log"),
    # experiment
    ("config_file", str, None),
)

# B200 backend additions. None of them changes behaviour at its default.
ACCEL_FLAGS = (
    ("accel_device", int, 0),            # CUDA ordinal used by the accel engine
    ("accel_precision", str, "fp32"),    # fp32 (parity) | tf32x3 | tf32 (tensor-core MLP)
    ("accel_shard", str, "replica"),     # replica: --accel_gpus replicas behind the accel slot.
                                         # table/row sharding needs lock-step ranks and is driven
                                         # by bench.py / torchrun, not by the Queue-fed engine
    ("accel_upload_inputs", _T, False),  # H2D per query instead of pre-staged batches
    ("accel_gpus", int, 1),              # GPUs behind the single accel slot (replicas; sibling
                                         # processes share accelRequestQueue / responseQueue)
    ("accel_coalesce", int, 32),         # requests already waiting on the accel queue that one
                                         # launch may serve together (1 = one launch per request;
                                         # drs_submit_many takes at most 32)
    ("accel_stack_rows", int, 16384),    # most samples one coalesced launch stacks
    ("accel_lanes", int, 3),             # launches in flight on the GPU (0 = library default, 2).
                                         # B200, RMC1, queries ~ normal(256, 32): 16 / 4096 / 2
                                         # saturate at 134 k queries/s, 32 / 16384 / 3 at 219 k
                                         # (profiles/r02r_*, r02s_*)
    ("accel_responder", int, 1),         # 1: ServiceResponses are built and sent by a helper
                                         # process fed through a shared-memory ring
    ("accel_host_pack", int, 0),         # host threads that bit-pack the ids of queries carrying
                                         # their own inputs before the copy (0 = ids travel as
                                         # they are; used for queries of >= 1,048,576 lookups)
    ("data_trace_sort_lookups", _T, False),  # --data_generation synthetic: sort the ids inside
                                         # every lookup (the random generator's np.unique does)
)


def debugPrint(args, system_tag, message):
    """Reference utils.py:8-12: print only under --debug_mode, flushed."""
    if args.debug_mode:
        print("[%s] %s" % (system_tag, message))
        sys.stdout.flush()


def build_parser(with_accel_flags=True):
    parser = argparse.ArgumentParser(description="DeepRecBench (B200 accel backend)")
    flags = REFERENCE_FLAGS + (ACCEL_FLAGS if with_accel_flags else ())
    for name, kind, default in flags:
        if kind is _T:
            parser.add_argument("--" + name, action="store_true", default=default)
        else:
            parser.add_argument("--" + name, type=kind, default=default)
    return parser


def expand_din_tables(arch_embedding_size, user_behavior_tables):
    """DIN: '<profile>-<behaviour>-<ad>-<ctx>' -> profile, N+1 behaviour tables,
    ad, ctx (reference utils.py:132-149 prepends N copies to the tail list)."""
    sizes = [int(s) for s in arch_embedding_size.split("-")]
    profile, behaviour, tail = sizes[0], sizes[1], sizes[1:]
    expanded = [profile] + [behaviour] * user_behavior_tables + tail
    return "-".join(str(s) for s in expanded)


def apply_config_file(args, path):
    """JSON is master: overrides argv, coerced by the current attribute's type."""
    with open(path, "r") as f:
        config = json.load(f)
    for key in config:
        caster = type(getattr(args, key))      # AttributeError on unknown keys
        setattr(args, key, caster(config[key]))
    return args


def cli(argv=None, with_accel_flags=True):
    args = build_parser(with_accel_flags).parse_args(argv)
    if args.model_type == "din":
        args.arch_embedding_size = expand_din_tables(args.arch_embedding_size,
                                                     args.user_behavior_tables)
    if args.config_file:
        apply_config_file(args, args.config_file)
    np.random.seed(args.numpy_rand_seed)
    return args

#!/usr/bin/env python
"""Generate the golden fixtures from the REFERENCE's own python (run in the
authoring container only; /root/reference does not exist on the GPU box).

    python tests/golden/make_golden.py

Writes (small, committed):
  cli_namespaces.json   vars(cli()) of the reference's utils/utils.py for every
                        shipped config and a few CLI-only variants
  datagen_cases.npz     outputs of the reference's DLRMDataGenerator
                        (generate_input_data + generate_output_data) for small
                        seeded cases, plus the next 4 doubles of the RNG stream
  rm1_probe.json        first values of the full-size dlrm_rm1.json dataset, seed 123
  packets.json          attribute names of the reference's ServiceRequest/Response
  trace_cases.npz       outputs of the reference's data_generator/trace_generator.py
                        (trace_generate_lru, seeded python + numpy streams) and
                        trace_profile.py (trace_profile) for small cases
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REF)

from data_generator.dlrm_data_caffe2 import DLRMDataGenerator  # noqa: E402  (reference)
from data_generator import trace_generator as ref_tracegen      # noqa: E402  (reference)
from data_generator import trace_profile as ref_traceprof       # noqa: E402  (reference)
from utils import packets as ref_packets                       # noqa: E402  (reference)
from utils.utils import cli as ref_cli                         # noqa: E402  (reference)

CFG = os.path.join(REF, "models", "configs")

CLI_CASES = {
    "defaults": [],
    "rm1": ["--config_file", CFG + "/dlrm_rm1.json"],
    "rm2": ["--config_file", CFG + "/dlrm_rm2.json"],
    "rm3": ["--config_file", CFG + "/dlrm_rm3.json"],
    "wnd": ["--config_file", CFG + "/wide_and_deep.json"],
    "ncf": ["--config_file", CFG + "/ncf.json"],
    "din_json": ["--config_file", CFG + "/din.json"],
    "dien": ["--config_file", CFG + "/dien.json"],
    "mtwnd": ["--config_file", CFG + "/mtwnd.json"],
    "din_cli_254": ["--model_type", "din", "--arch_embedding_size", "1000-100-2000-3000",
                    "--user_behavior_tables", "250"],
    "din_cli_then_json": ["--model_type", "din", "--arch_embedding_size", "7-8-9-10",
                          "--user_behavior_tables", "3", "--config_file", CFG + "/din.json"],
    "system_run": ["--queue", "--model_accel", "--inference_engines", "2",
                   "--config_file", CFG + "/dlrm_rm1.json", "--num_batches", "32",
                   "--max_mini_batch_size", "1024", "--avg_mini_batch_size", "165",
                   "--var_mini_batch_size", "16", "--batch_size_distribution", "normal",
                   "--accel_request_size_thres", "165", "--avg_arrival_rate", "5",
                   "--tune_accel_qps", "--debug_mode"],
}

DATAGEN_CASES = {
    # name -> argv (kept small: these arrays are committed)
    "dlrm_small": ["--arch_embedding_size", "50-60-70", "--arch_sparse_feature_size", "8",
                   "--arch_mlp_bot", "6-8", "--arch_mlp_top", "16-1", "--num_indices_per_lookup", "4",
                   "--num_batches", "2", "--max_mini_batch_size", "8"],
    "dup_heavy": ["--arch_embedding_size", "40-30", "--arch_sparse_feature_size", "4",
                  "--arch_mlp_bot", "3-4", "--arch_mlp_top", "8-1", "--num_indices_per_lookup", "12",
                  "--num_batches", "2", "--max_mini_batch_size", "16", "--numpy_rand_seed", "7"],
    "wnd_l1": ["--model_type", "wnd", "--arch_embedding_size", "100-100-100",
               "--arch_sparse_feature_size", "4", "--arch_mlp_bot", "5", "--arch_mlp_top", "8-1",
               "--arch_interaction_op", "cat", "--num_indices_per_lookup", "1",
               "--num_batches", "3", "--max_mini_batch_size", "4"],
    "round_targets": ["--arch_embedding_size", "20-20", "--arch_sparse_feature_size", "4",
                      "--arch_mlp_bot", "3-4", "--arch_mlp_top", "8-1", "--num_indices_per_lookup", "2",
                      "--num_batches", "1", "--max_mini_batch_size", "6", "--round_targets", "True"],
}


# name -> (table_size, list_sd, probabilities, trace length, enable_padding, seed)
TRACE_CASES = {
    "mostly_new": (400, [0, 3, 7, 20], [0.9, 0.05, 0.03, 0.02], 300, False, 11),
    "high_reuse": (500, [0, 1, 2, 3, 5, 8, 13, 40, 100],
                   [0.3, 0.2, 0.1, 0.1, 0.1, 0.05, 0.05, 0.05, 0.05], 4000, False, 7),
    "high_reuse_padded": (500, [0, 1, 2, 3, 5, 8, 13, 40, 100],
                          [0.3, 0.2, 0.1, 0.1, 0.1, 0.05, 0.05, 0.05, 0.05], 4000, True, 7),
    "table_exhausted": (150, [0, 1, 2, 5, 30], [0.5, 0.2, 0.1, 0.1, 0.1], 2000, False, 3),
}
# name -> (ids in [0, hi), trace length, max_stack_distance, seed)
PROFILE_CASES = {
    "narrow_window": (50, 1500, 20, 5),
    "wide_window": (30, 1000, 5000, 6),
    "window_one": (5, 300, 1, 8),
    "sparse_reuse": (400, 1500, 100, 9),
}


def cumulative(probs):
    c, acc = [], 0.0
    for p in probs:
        acc += p
        c.append(acc)
    c[-1] = 1.0
    return c


def trace_goldens():
    import random
    arrays = {}
    for name, (rows, list_sd, probs, k, pad, seed) in TRACE_CASES.items():
        cumm = cumulative(probs)
        random.seed(seed)
        np.random.seed(seed)
        tr = ref_tracegen.trace_generate_lru(rows, list_sd, cumm, k, pad)
        arrays["gen/%s/params" % name] = np.array([rows, k, int(pad), seed], dtype=np.int64)
        arrays["gen/%s/list_sd" % name] = np.array(list_sd, dtype=np.int64)
        arrays["gen/%s/cumm_sd" % name] = np.array(cumm, dtype=np.float64)
        arrays["gen/%s/trace" % name] = np.array(tr, dtype=np.uint64)
        arrays["gen/%s/rng_tail" % name] = np.random.random_sample(2)
    for name, (hi, n, window, seed) in PROFILE_CASES.items():
        tr = np.random.RandomState(seed).randint(0, hi, n).astype(np.int32)
        sd, la = ref_traceprof.trace_profile(tr, window)
        arrays["prof/%s/params" % name] = np.array([hi, n, window, seed], dtype=np.int64)
        arrays["prof/%s/stack_distances" % name] = np.array(sd, dtype=np.int64)
        arrays["prof/%s/line_accesses" % name] = np.array(la, dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "trace_cases.npz"), **arrays)


def run_ref_cli(argv):
    old = sys.argv
    sys.argv = ["prog"] + list(argv)
    try:
        return ref_cli()
    finally:
        sys.argv = old


def main():
    out = {}
    for name, argv in CLI_CASES.items():
        ns = vars(run_ref_cli(argv))
        rel = [a.replace(CFG, "<CONFIGS>") for a in argv]
        ns = {k: (v.replace(CFG, "<CONFIGS>") if isinstance(v, str) else v) for k, v in ns.items()}
        out[name] = {"argv": rel, "namespace": ns}
    with open(os.path.join(HERE, "cli_namespaces.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)

    arrays = {}
    for name, argv in DATAGEN_CASES.items():
        args = run_ref_cli(argv)                    # seeds numpy
        g = DLRMDataGenerator(args)
        nb, lX, lS_l, lS_i = g.generate_input_data()
        _, lT = g.generate_output_data()
        arrays[name + "/argv"] = np.array(argv)
        for j in range(nb):
            arrays["%s/X%d" % (name, j)] = lX[j]
            arrays["%s/len%d" % (name, j)] = np.array(lS_l[j])
            arrays["%s/idx%d" % (name, j)] = np.array(lS_i[j])
            arrays["%s/T%d" % (name, j)] = lT[j]
        arrays[name + "/rng_tail"] = np.random.random_sample(4)
    np.savez_compressed(os.path.join(HERE, "datagen_cases.npz"), **arrays)

    args = run_ref_cli(["--config_file", CFG + "/dlrm_rm1.json", "--num_batches", "1",
                        "--max_mini_batch_size", "4"])
    g = DLRMDataGenerator(args)
    nb, lX, lS_l, lS_i = g.generate_input_data()
    probe = {"argv": ["--config_file", "<CONFIGS>/dlrm_rm1.json", "--num_batches", "1",
                      "--max_mini_batch_size", "4"],
             "lX0_row0_first4": [float(v) for v in lX[0][0, :4]],
             "idx_table0_first8": [int(v) for v in lS_i[0][0][:8]],
             "idx_table7_last4": [int(v) for v in lS_i[0][7][-4:]],
             "idx_dtype_after_np_array": str(np.array(lS_i[0]).dtype),
             "len_dtype_after_np_array": str(np.array(lS_l[0]).dtype)}
    with open(os.path.join(HERE, "rm1_probe.json"), "w") as f:
        json.dump(probe, f, indent=1)

    req = ref_packets.ServiceRequest()
    resp = ref_packets.ServiceResponse()
    with open(os.path.join(HERE, "packets.json"), "w") as f:
        json.dump({"ServiceRequest": sorted(vars(req)), "ServiceResponse": sorted(vars(resp))},
                  f, indent=1)
    trace_goldens()
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()

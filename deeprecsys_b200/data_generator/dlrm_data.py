"""Synthetic query data for the accelerator path.

Produces *bit-identical* data, and leaves numpy's legacy global RNG in the
*same state*, as the reference generator
(/root/reference/data_generator/dlrm_data_caffe2.py:69-124 for inputs,
:128-148 for targets) — every engine process of DeepRecSys regenerates the
dataset locally from the seed and packets only name ``(batch_id, batch_size)``
(SURVEY fact 3), so the accel engine must consume the RNG exactly like the CPU
engines do, because the model weights are drawn from the same stream afterwards.

What the reference does per batch j (all draws from the global MT19937):
  1. dense ``rand(n, m_den)`` cast to f32                                (:89)
  2. for each table, for each of the n samples: draw ``random(L)``, map to
     ``unique(round(r * (rows-1)).astype(int32))`` and *redraw the whole
     group until it has exactly L distinct ids*                       (:105-110)
and, after all input batches, one ``rand(n, 1)`` target per batch       (:145).

The per-sample python loop is replaced by an equivalent block formulation:
the stream is a sequence of "rows" of L doubles; sample k of a table receives
the k-th row whose L rounded ids are all distinct.  Drawing exactly as many
rows as are still missing, filtering, and repeating consumes precisely the
same doubles as the reference's redraw loop (the last block drawn contains only
accepted rows, so nothing is over-drawn).  tests/test_spec_surface.py checks this
against the reference generator itself and against tests/golden fixtures.

Layout returned (numpy instead of python lists, same values):
  lX[j]    float32 [n, m_den]
  lS_l[j]  int32   [T, n]          lengths (== L)
  lS_i[j]  int32   [T, n*L]        ids, sorted ascending inside each sample
"""
from __future__ import absolute_import, division, print_function

import numpy as np


def _parse_dash_list(text):
    return np.array([int(tok) for tok in str(text).split("-")], dtype=np.int64)


def _draw_distinct_groups(rows, n, L):
    """n groups of L distinct sorted int32 ids in [0, rows-1] for one table."""
    out = np.empty((n, L), dtype=np.int32)
    filled = 0
    while filled < n:
        need = n - filled
        r = np.random.random_sample((need, L))
        ids = np.round(r * (rows - 1)).astype(np.int32)
        ids.sort(axis=1)
        if L > 1:
            ok = (ids[:, 1:] != ids[:, :-1]).all(axis=1)
        else:
            ok = np.ones(need, dtype=bool)
        good = ids[ok]
        out[filled:filled + len(good)] = good
        filled += len(good)
    return out


class DLRMDataGenerator(object):
    """Same constructor / method names as the reference class."""

    def __init__(self, args):
        self.args = args

    # reference :33-60 (only the "random" branch is live code there; its "synthetic" branch calls
    # functions that are not in scope and cannot run — see generate_synthetic_input_data below)
    def generate_input_data(self):
        a = self.args
        ln_bot = _parse_dash_list(a.arch_mlp_bot)
        ln_emb = _parse_dash_list(a.arch_embedding_size)
        if a.data_generation == "synthetic":
            return self.generate_synthetic_input_data(
                a.num_batches, a.max_mini_batch_size, a.round_targets,
                a.num_indices_per_lookup, a.num_indices_per_lookup_fixed,
                int(ln_bot[0]), ln_emb, a.data_trace_file, a.data_trace_enable_padding)
        if a.data_generation != "random":
            raise SystemExit("ERROR: --data-generation=%s is not supported"
                             % a.data_generation)
        return self.generate_random_input_data(
            a.num_batches, a.max_mini_batch_size, a.round_targets,
            a.num_indices_per_lookup, a.num_indices_per_lookup_fixed,
            int(ln_bot[0]), ln_emb)

    def generate_output_data(self):
        a = self.args
        return self.generate_random_output_data(
            a.num_batches, a.max_mini_batch_size, round_targets=a.round_targets)

    def generate_random_input_data(self, num_batches, mini_batch_size,
                                   round_targets, num_indices_per_lookup,
                                   num_indices_per_lookup_fixed, m_den, ln_emb):
        n, L, T = int(mini_batch_size), int(num_indices_per_lookup), len(ln_emb)
        lX, lS_l, lS_i = [], [], []
        for _ in range(num_batches):
            lX.append(np.random.rand(n, m_den).astype(np.float32))
            ids = np.empty((T, n * L), dtype=np.int32)
            for t, rows in enumerate(ln_emb):
                ids[t] = _draw_distinct_groups(int(rows), n, L).reshape(-1)
            lS_i.append(ids)
            lS_l.append(np.full((T, n), L, dtype=np.int32))
        return num_batches, lX, lS_l, lS_i

    def generate_synthetic_input_data(self, num_batches, mini_batch_size, round_targets,
                                      num_indices_per_lookup, num_indices_per_lookup_fixed,
                                      m_den, ln_emb, trace_file, enable_padding=False):
        """``--data_generation synthetic``: ids with temporal locality.

        The reference's branch (dlrm_data_caffe2.py:48-56,152-227) is dead code — it calls
        module-level names that only exist as class attributes — so there is no behaviour to
        reproduce; what is built here is the flow its tools describe (data_generator/README.md:
        trace_profile.py -> profile/sd_cumm -> trace_generator.py -> one long trace per table of
        num_batches * mini_batch_size * pooling_factor ids), wired into the generator:

          * table t reads its stack-distance profile from ``trace_file`` with every "j" replaced
            by t (the reference's convention, :196), or from ``trace_file`` itself when that
            per-table file does not exist (one profile shared by all tables, e.g. the shipped
            profile/sd_cumm);
          * a two-line profile (trace_profile.py's output) takes the order of first references
            from one ``np.random.permutation(rows)`` per table, drawn in table order before the
            first batch — the seeded numpy stream, where the reference's script uses python's
            unseeded ``random.sample`` — a three-line profile (dlrm_data_caffe2.py:380-393)
            carries that order itself;
          * per batch: dense ``rand(n, m_den)`` as in the random mode, then for each table the
            next n*L ids of that table's LRU-stack trace (one uniform draw per id); reuse
            carries across samples and batches;
          * lookups keep the fixed length L the engines' request slicing needs
            (inferenceEngine.py:204-206), in trace order; an id may repeat inside a lookup
            (SparseLengthsSum then adds that row twice).  ``--data_trace_sort_lookups`` sorts
            the ids inside every lookup, as the random generator's ``np.unique`` does
            (same multiset of rows per sample, ascending walk through the table).
        """
        import os
        from deeprecsys_b200.data_generator.trace_generator import (LruTraceGenerator,
                                                                    read_dist_from_file)
        if not num_indices_per_lookup_fixed:
            raise SystemExit("ERROR: --data_generation=synthetic needs "
                             "--num_indices_per_lookup_fixed (requests are sliced as B*L ids)")
        n, L, T = int(mini_batch_size), int(num_indices_per_lookup), len(ln_emb)
        traces = []
        for t, rows in enumerate(ln_emb):
            path = str(trace_file).replace("j", str(t))
            if not os.path.exists(path):
                path = str(trace_file)
            dist = read_dist_from_file(path)
            if len(dist) == 3:
                order = np.asarray(dist[0], dtype=np.int64)
                if order.size and (order.min() < 0 or order.max() >= int(rows)):
                    print("WARNING: distribution is inconsistent with embedding "
                          "table size (using mod to recover and continue)")      # ref :207-215
                    order = np.mod(order, int(rows))
            else:
                order = np.random.permutation(int(rows))
            traces.append(LruTraceGenerator(order, dist[-2], dist[-1], enable_padding))
        lX, lS_l, lS_i = [], [], []
        for _ in range(num_batches):
            lX.append(np.random.rand(n, m_den).astype(np.float32))
            ids = np.empty((T, n * L), dtype=np.int32)
            for t in range(T):
                ids[t] = traces[t].generate(n * L)
            if getattr(self.args, "data_trace_sort_lookups", False):
                ids.reshape(T, n, L).sort(axis=2)
            lS_i.append(ids)
            lS_l.append(np.full((T, n), L, dtype=np.int32))
        return num_batches, lX, lS_l, lS_i

    def generate_random_output_data(self, num_batches, mini_batch_size,
                                    num_targets=1, round_targets=False):
        lT = []
        for _ in range(num_batches):
            P = np.random.rand(mini_batch_size, num_targets).astype(np.float32)
            if round_targets:
                P = np.round(P).astype(np.int32)
            lT.append(P)
        return num_batches, lT


def slice_request(lX, lS_l, lS_i, batch_id, batch_size, L):
    """Request -> tensors, as the CPU engine does (inferenceEngine.py:200-215):
    the first ``batch_size`` samples of batch ``batch_id``; ``sub_id`` is not
    an offset.  Returns (dense [B, m_den], lengths [T, B], ids [T, B*L])."""
    B = int(batch_size)
    return (lX[batch_id][:B], lS_l[batch_id][:, :B], lS_i[batch_id][:, :B * L])

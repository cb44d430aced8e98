"""Model spec of the accel path: the integers the reference's builders derive
from the CLI/JSON flags, and the parameter draw order they imply.

Dimension rules per model (reference file:line):
  dlrm  models/dlrm_s_caffe2.py:404-440  num_fea = T+1; num_int = cat: F*D,
        dot: F(F-1)/2 + D (or F(F+1)/2 + D with --arch_interaction_itself);
        ln_top = [num_int] + arch_mlp_top; requires ln_bot[-1] == D.
  wnd   models/wide_and_deep.py:332-350  no bottom layers; num_int = T*D + m_den.
  ncf   models/ncf.py:348-388            4 tables, L = 1; tower = [2D] + top[:-1];
        final FC = [D + top[-2], top[-1]].
  din   models/din.py:246-258,346-359    attention MLP = [3D] + arch_mlp_bot + [D]
        per behaviour table (tables 1..T-3); top input = 4D.
  dien  models/dien.py:296-433           table roles as DIN; the N = T-3 behaviour embeddings
        feed two stacked tanh BasicRNN cells (D -> H, H -> H, H = --hidden_size); the last hidden
        state is concatenated with profile / candidate-ad / context embeddings -> top MLP
        [H + 3D] + arch_mlp_top, all ReLU.  The cells' weights are re-initialised by Caffe2's own
        XavierFill when the reference runs model.param_init_net (dien.py:502), so they do NOT
        come from the numpy stream: the reference's numpy randn draws for them are consumed (to
        keep the stream aligned for the top MLP) and XavierFill-distributed weights are drawn
        from a private RandomState(numpy_rand_seed).
  mtwnd models/multi_task_wnd.py:304-316,340-362  WnD trunk with an all-ReLU shared top MLP,
        then num_multi_tasks heads of widths arch_mlp_tasks (own weights each); every head is
        built with sigmoid_layer = len(ln_top) - 1, i.e. the Sigmoid lands on task layer
        len(arch_mlp_top) whatever the head's depth.  Output = heads side by side.

Parameter draw order (global legacy numpy RNG, after datagen inputs+targets):
tables in index order (uniform(-sqrt(1/n), sqrt(1/n)), drawn float64 then cast,
dlrm_s_caffe2.py:297-299), then MLP layers in graph-construction order, each
W ~ normal(0, sqrt(2/(m+n)), (m, n)) then b ~ normal(0, sqrt(1/m), m)
(dlrm_s_caffe2.py:245-249).
"""
from __future__ import absolute_import, division, print_function

import numpy as np

from . import _lib


def _ints(text):
    return [int(tok) for tok in str(text).split("-")]


class ModelSpec(object):
    def __init__(self, args):
        self.model_type = args.model_type
        if self.model_type not in _lib.MODEL_TYPES:
            raise ValueError("model_type %r is not on the accel path (supported: %s)"
                             % (self.model_type, sorted(_lib.MODEL_TYPES)))
        self.table_rows = _ints(args.arch_embedding_size)
        self.T = len(self.table_rows)
        self.D = int(args.arch_sparse_feature_size)
        self.L = int(args.num_indices_per_lookup)
        self.interaction_op = args.arch_interaction_op
        self.interaction_itself = bool(args.arch_interaction_itself)
        arch_bot, arch_top = _ints(args.arch_mlp_bot), _ints(args.arch_mlp_top)
        self.datagen_m_den = arch_bot[0]          # what the generator draws (always)
        T, D = self.T, self.D

        if self.model_type == "dlrm":
            F = T + 1
            if self.interaction_op == "dot":
                pairs = F * (F + 1) // 2 if self.interaction_itself else F * (F - 1) // 2
                num_int = pairs + arch_bot[-1]
            elif self.interaction_op == "cat":
                num_int = F * arch_bot[-1]
            else:
                raise ValueError("--arch-interaction-op=%s is not supported"
                                 % self.interaction_op)
            if D != arch_bot[-1]:
                raise ValueError("arch-sparse-feature-size %d does not match last dim of "
                                 "bottom mlp %d" % (D, arch_bot[-1]))
            self.m_den, self.ln_bot, self.ln_top = arch_bot[0], arch_bot, [num_int] + arch_top
            self.mlps = [(_lib.MLP_BOT, self.ln_bot), (_lib.MLP_TOP, self.ln_top)]
            self.sigmoid_last = True
        elif self.model_type == "wnd":
            if self.interaction_op != "cat" or len(arch_bot) != 1:
                raise ValueError("wide&deep needs cat interaction and a single arch_mlp_bot width")
            self.m_den, self.ln_bot = arch_bot[0], [arch_bot[0]]
            self.ln_top = [T * D + arch_bot[0]] + arch_top
            self.mlps = [(_lib.MLP_TOP, self.ln_top)]
            self.sigmoid_last = True
        elif self.model_type == "ncf":
            if T != 4 or self.L != 1 or self.interaction_op != "cat":
                raise ValueError("NCF needs 4 tables, 1 lookup per table, cat interaction")
            self.m_den, self.ln_bot = 0, []
            self.ln_top = [2 * D] + arch_top
            tower = self.ln_top[:-1]
            self.mlps = [(_lib.MLP_TOP, tower), (_lib.MLP_FINAL, [D + tower[-1], self.ln_top[-1]])]
            self.sigmoid_last = False
        elif self.model_type == "dien":
            if T < 4:
                raise ValueError("DIEN needs at least 4 embedding tables")        # dien.py:457
            self.m_den, self.ln_bot = 0, []
            self.hidden_size = int(args.hidden_size)
            self.n_steps = T - 3
            self.ln_top = [self.hidden_size + 3 * D] + arch_top                  # dien.py:426-429
            self.rnn = [(D, self.hidden_size), (self.hidden_size, self.hidden_size)]
            self.mlps = [(_lib.MLP_TOP, self.ln_top)]
            self.sigmoid_last = False
            self.seed = int(getattr(args, "numpy_rand_seed", 0))
        elif self.model_type == "mtwnd":
            if self.interaction_op != "cat" or len(arch_bot) != 1:
                raise ValueError("multi-task wide&deep needs cat interaction and a single "
                                 "arch_mlp_bot width")
            self.m_den, self.ln_bot = arch_bot[0], [arch_bot[0]]
            self.ln_top = [T * D + arch_bot[0]] + arch_top
            self.ln_task = _ints(args.arch_mlp_tasks)
            self.n_tasks = int(args.num_multi_tasks)
            if self.ln_top[-1] != self.ln_task[0]:
                raise ValueError("Shared top layer and task MLP layers must have same "
                                 "input/output dimension")          # multi_task_wnd.py:362
            if self.n_tasks < 1:
                raise ValueError("num_multi_tasks must be >= 1")
            self.task_sigmoid_layer = len(self.ln_top) - 1         # sigmoid_top, :396
            if self.task_sigmoid_layer > len(self.ln_task) - 1:
                self.task_sigmoid_layer = 0                        # never reached: all ReLU
            self.mlps = [(_lib.MLP_TOP, self.ln_top)]
            self.mlps += [(_lib.MLP_TASK_BASE + i, self.ln_task) for i in range(self.n_tasks)]
            self.sigmoid_last = False
        else:  # din
            if T < 4:
                raise ValueError("DIN needs at least 4 embedding tables")
            self.m_den = 0
            self.ln_bot = [3 * D] + arch_bot + [D]
            self.ln_top = [4 * D] + arch_top
            self.n_units = T - 3
            self.mlps = [(_lib.MLP_ATTN_BASE + i, self.ln_bot) for i in range(self.n_units)]
            self.mlps.append((_lib.MLP_TOP, self.ln_top))
            self.sigmoid_last = False
        self.out_dim = self.mlps[-1][1][-1]
        if self.model_type == "mtwnd":
            self.out_dim = self.n_tasks * self.ln_task[-1]

    # ---- algorithmic work per sample (DESIGN.md / SURVEY 8d)
    def gather_bytes_per_sample(self, fused=True):
        rows = self.T * self.L * self.D * 4
        idx = self.T * self.L * 4
        lens = self.T * 4
        pooled = 0 if fused else self.T * self.D * 4
        return rows + idx + lens + pooled

    def mlp_flops_per_sample(self):
        total = 0
        for mlp_id, ln in self.mlps:
            total += sum(2 * ln[i] * ln[i + 1] for i in range(len(ln) - 1))
        if self.model_type == "dlrm" and self.interaction_op == "dot":
            total += 2 * (self.T + 1) ** 2 * self.D
        return total

    def describe(self):
        return {"model_type": self.model_type, "tables": self.T, "rows": self.table_rows,
                "D": self.D, "L": self.L, "ln_bot": self.ln_bot, "ln_top": self.ln_top,
                "interaction": self.interaction_op, "out_dim": self.out_dim}


def draw_table(rows, dim):
    """One embedding table from the global legacy RNG (dlrm_s_caffe2.py:297-299)."""
    bound = np.sqrt(1 / rows)
    return np.random.uniform(low=-bound, high=bound, size=(rows, dim)).astype(np.float32)


def draw_fc(n_in, n_out):
    """One FC layer (W [out,in], b [out]) from the global RNG (dlrm_s_caffe2.py:245-249)."""
    W = np.random.normal(0.0, np.sqrt(2 / (n_out + n_in)), size=(n_out, n_in)).astype(np.float32)
    b = np.random.normal(0.0, np.sqrt(1 / n_out), size=n_out).astype(np.float32)
    return W, b


def iter_parameters(spec):
    """Yield ('table', t, W) then ('fc', mlp_id, layer, W, b) in the reference's
    draw order.  A generator so that multi-GB tables can be uploaded and freed
    one at a time."""
    for t, rows in enumerate(spec.table_rows):
        yield ("table", t, draw_table(rows, spec.D))
    if spec.model_type == "dien":
        # dien.py:307-310, 359-362: four randn draws per cell, fed and then overwritten by the
        # framework's initialiser -> consume them, then draw XavierFill-distributed weights
        # (U(+-sqrt(3 / fan_in)), zero biases) from a private stream.  Cell l is addressed as
        # mlp id MLP_RNN_BASE + l with "layer" 0 = i2h, "layer" 1 = gates_t.
        private = np.random.RandomState(spec.seed)
        for cell, (n_in, H) in enumerate(spec.rnn):
            np.random.randn(H, H); np.random.randn(H); np.random.randn(H, n_in); np.random.randn(H)
            for layer, fan_in in ((0, n_in), (1, H)):
                scale = np.sqrt(3.0 / fan_in)
                W = private.uniform(-scale, scale, size=(H, fan_in)).astype(np.float32)
                yield ("fc", _lib.MLP_RNN_BASE + cell, layer, W, np.zeros(H, np.float32))
    for mlp_id, ln in spec.mlps:
        for layer in range(len(ln) - 1):
            W, b = draw_fc(ln[layer], ln[layer + 1])
            yield ("fc", mlp_id, layer, W, b)


def draw_all_parameters(spec):
    """Materialise everything (small configs / tests)."""
    tables, fcs = [], {}
    for item in iter_parameters(spec):
        if item[0] == "table":
            tables.append(item[2])
        else:
            fcs[(item[1], item[2])] = (item[3], item[4])
    return tables, fcs

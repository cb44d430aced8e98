"""caffe2.python.rnn_cell stand-in: BasicRNN as DIEN calls it (models/dien.py:318-328,370-378).

Restated from the published cell (caffe2/python/rnn_cell.py, BasicRNNCell):
    prepare_input : i2h = FC(input [N, B, dim_in], <scope>/i2h_w, <scope>/i2h_b, axis=2)
    step t        : gates_t = FC(h_{t-1}, <scope>/gates_t_w, <scope>/gates_t_b, axis=2)
                    gates_t = Sum(gates_t, i2h[t]);  h_t = Tanh(gates_t)
                    rows with t >= seq_lengths[b] keep h_{t-1}  (drop_states=False)
    returns (all hidden states [N, B, H], last hidden state [1, B, H])
The initial state may be [B, H] or [1, B, H] (RecurrentNetwork accepts 2-D initial states).
Parameters are registered with XavierFill / ConstantFill in model.param_init_net exactly like
brew.fc does, so blobs the caller fed under the same names beforehand are re-initialised when
param_init_net runs — which is what Caffe2 does to the reference's numpy-fed rnn_*/ weights.
TEST INFRASTRUCTURE ONLY."""
from __future__ import absolute_import, division, print_function


def BasicRNN(model, input_blob, seq_lengths, initial_states, dim_in, dim_out, scope,
             activation=None, forward_only=False, drop_states=False, **kwargs):
    assert activation == "tanh", "only the tanh cell DIEN uses is restated"
    names = {}
    for key, shape in (("i2h_w", [dim_out, dim_in]), ("gates_t_w", [dim_out, dim_out])):
        names[key] = "%s/%s" % (scope, key)
        model.param_init_net.XavierFill([], names[key], shape=shape)
    for key in ("i2h_b", "gates_t_b"):
        names[key] = "%s/%s" % (scope, key)
        model.param_init_net.ConstantFill([], names[key], shape=[dim_out], value=0.0)
    model.params.extend(names.values())
    all_h, last_h = model.net.BasicRNNShim(
        [input_blob, seq_lengths, initial_states[0], names["i2h_w"], names["i2h_b"],
         names["gates_t_w"], names["gates_t_b"]],
        ["%s/hidden_t_all" % scope, "%s/hidden_t_last" % scope], drop_states=int(drop_states))
    return all_h, last_h

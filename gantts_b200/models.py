"""Drop-in for reference ``gantts/models.py``: same class names, constructor signatures, forward
signatures and ``state_dict`` keys, with the arithmetic in hand-written sm_100a CUDA."""
import torch
from torch import nn

from . import _lib
from . import config
from . import ops
from . import rnn


class AbstractModel(object):
    """Interface for VC and TTS models (reference gantts/models.py:11-18)."""

    def include_parameter_generation(self):
        """Whether model includes parameter generation or not."""
        return False


def _mlp(x, hidden, last, p, training, last_act, engine=None):
    """Hidden stack + last_linear.  Tensor-core engine: ONE fused stack op (planes-resident
    activations); SIMT engine: per-layer exact-fp32 ops (validation path)."""
    if (engine or config.engine) == "tc":
        layers = list(hidden) + [last]
        return ops.mlp_stack(x, [l.weight for l in layers], [l.bias for l in layers], p=p, training=training,
                             last_act=last_act)
    for layer in hidden:
        x = ops.linear_act(x, layer.weight, layer.bias, _lib.ACT_LEAKY_DROPOUT, p=p, training=training,
                           engine=engine)
    return ops.linear_act(x, last.weight, last.bias, last_act, engine=engine)


class MLP(AbstractModel, nn.Module):
    """Generator or discriminator MLP (reference gantts/models.py:121-141):
    ``x = Dropout(LeakyReLU_0.01(Linear(x)))`` per hidden layer, ``last_linear``, optional sigmoid.
    state_dict keys: ``layers.{i}.weight/bias``, ``last_linear.weight/bias``."""

    def __init__(self, in_dim=118, out_dim=1, num_hidden=2, hidden_dim=256,
                 dropout=0.5, last_sigmoid=True, bidirectional=None):
        # bidirectional is dummy
        super(MLP, self).__init__()
        in_sizes = [in_dim] + [hidden_dim] * (num_hidden - 1)
        out_sizes = [hidden_dim] * num_hidden
        self.layers = nn.ModuleList(
            [nn.Linear(in_size, out_size) for (in_size, out_size) in zip(in_sizes, out_sizes)])
        self.last_linear = nn.Linear(hidden_dim, out_dim)
        self.dropout_p = float(dropout)
        self.last_sigmoid = last_sigmoid
        self.engine = None          # None -> gantts_b200.config.engine

    def forward(self, x, lengths=None):
        act = _lib.ACT_SIGMOID if self.last_sigmoid else _lib.ACT_NONE
        return _mlp(x, self.layers, self.last_linear, self.dropout_p, self.training, act, self.engine)


class In2OutHighwayNet(AbstractModel, nn.Module):
    """Input-to-output highway network for VC (reference gantts/models.py:21-69): returns
    ``(y_hat, x_static + sigmoid(T(x_static)) * MLPG(R, y_hat))``.
    state_dict keys: ``T.*``, ``H.{i}.*``, ``last_linear.*``."""

    def __init__(self, in_dim=118, out_dim=118, static_dim=118 // 2,
                 num_hidden=3, hidden_dim=512, dropout=0.5):
        super(In2OutHighwayNet, self).__init__()
        self.static_dim = static_dim
        self.T = nn.Linear(static_dim, static_dim)
        in_sizes = [in_dim] + [hidden_dim] * (num_hidden - 1)
        out_sizes = [hidden_dim] * num_hidden
        self.H = nn.ModuleList(
            [nn.Linear(in_size, out_size) for (in_size, out_size) in zip(in_sizes, out_sizes)])
        self.last_linear = nn.Linear(hidden_dim, out_dim)
        self.dropout_p = float(dropout)
        self.engine = None

    def include_parameter_generation(self):
        return True

    def forward(self, x, R, lengths=None):
        x = x.unsqueeze(0) if x.dim() == 2 else x
        x_static = x[:, :, :self.static_dim]
        Tx = ops.linear_act(x_static, self.T.weight, self.T.bias, _lib.ACT_SIGMOID, engine=self.engine)
        h = _mlp(x, self.H, self.last_linear, self.dropout_p, self.training, _lib.ACT_NONE, self.engine)
        Gx = ops.unit_variance_mlpg(R, h)
        return h, ops.highway_combine(x_static, Tx, Gx)


class In2OutRNNHighwayNet(AbstractModel, nn.Module):
    """LSTM variant of the highway network (reference gantts/models.py:72-118).  NOTE: like the
    reference it returns its INPUT ``x`` as the first output (``:118``).
    state_dict keys: ``T.*``, ``lstm.weight_ih_l{k}[_reverse]`` ..., ``hidden2out.*``."""

    def __init__(self, in_dim=118, out_dim=118, static_dim=118 // 2,
                 num_hidden=3, hidden_dim=512, bidirectional=False, dropout=0.5):
        super(In2OutRNNHighwayNet, self).__init__()
        self.static_dim = static_dim
        self.num_direction = 2 if bidirectional else 1
        self.T = nn.Linear(static_dim, static_dim)
        self.lstm = nn.LSTM(in_dim, hidden_dim, num_hidden, batch_first=True,
                            bidirectional=bidirectional, dropout=dropout)
        self.hidden2out = nn.Linear(hidden_dim * self.num_direction, out_dim)
        self.engine = None

    def include_parameter_generation(self):
        return True

    def forward(self, x, R, lengths=None):
        x = x.unsqueeze(0) if x.dim() == 2 else x
        x_static = x[:, :, :self.static_dim]
        Tx = ops.linear_act(x_static, self.T.weight, self.T.bias, _lib.ACT_SIGMOID, engine=self.engine)
        output = rnn.lstm_forward(self.lstm, x, lengths, self.training, self.engine)
        output = ops.linear_act(output, self.hidden2out.weight, self.hidden2out.bias, _lib.ACT_NONE,
                                engine=self.engine)
        Gx = ops.unit_variance_mlpg(R, output)
        return x, ops.highway_combine(x_static, Tx, Gx)


class _LSTMNet(AbstractModel, nn.Module):
    """pack -> nn.LSTM -> pad -> Linear (-> sigmoid), reference gantts/models.py:170-213."""
    _rnn_attr = "lstm"

    def __init__(self, in_dim=118, out_dim=118, num_hidden=2, hidden_dim=256,
                 bidirectional=False, dropout=0, last_sigmoid=False):
        super(_LSTMNet, self).__init__()
        self.num_direction = 2 if bidirectional else 1
        setattr(self, self._rnn_attr, nn.LSTM(in_dim, hidden_dim, num_hidden, batch_first=True,
                                              bidirectional=bidirectional, dropout=dropout))
        self.hidden2out = nn.Linear(hidden_dim * self.num_direction, out_dim)
        self.last_sigmoid = last_sigmoid
        self.engine = None

    def forward(self, sequence, lengths):
        output = rnn.lstm_forward(getattr(self, self._rnn_attr), sequence, lengths, self.training, self.engine)
        act = _lib.ACT_SIGMOID if self.last_sigmoid else _lib.ACT_NONE
        return ops.linear_act(output, self.hidden2out.weight, self.hidden2out.bias, act, engine=self.engine)


class LSTMRNN(_LSTMNet):
    """reference gantts/models.py:193-213; state_dict prefix ``lstm.``."""
    _rnn_attr = "lstm"


class GRURNN(_LSTMNet):
    """reference gantts/models.py:170-190 -- despite the name an nn.LSTM stored as ``.gru``."""
    _rnn_attr = "gru"


class SRURNN(AbstractModel, nn.Module):
    """SRU generator (reference gantts/models.py:144-167; the hparams default for TTS).  The upstream
    ``cuda_functional.SRU`` is replaced by ``gantts_b200.rnn.SRU`` (same recurrence, parameters under
    ``gru.rnn_lst.{i}.weight/bias``); like the reference it IGNORES ``lengths``."""

    def __init__(self, in_dim=118, out_dim=118, num_hidden=2, hidden_dim=256,
                 bidirectional=False, dropout=0, last_sigmoid=False,
                 use_relu=0, rnn_dropout=0.0):
        super(SRURNN, self).__init__()
        self.num_direction = 2 if bidirectional else 1
        self.gru = rnn.SRU(in_dim, hidden_dim, num_hidden, bidirectional=bidirectional, dropout=dropout,
                           use_relu=use_relu, rnn_dropout=rnn_dropout)
        self.hidden2out = nn.Linear(hidden_dim * self.num_direction, out_dim)
        self.last_sigmoid = last_sigmoid
        self.engine = None

    def forward(self, sequence, lengths):
        output = self.gru(sequence, engine=self.engine)
        act = _lib.ACT_SIGMOID if self.last_sigmoid else _lib.ACT_NONE
        return ops.linear_act(output, self.hidden2out.weight, self.hidden2out.bias, act, engine=self.engine)

"""CPU restatement (torch CPU fp32) of the reference GAN-step hot path.  TEST INFRASTRUCTURE.

Each function cites the reference file:line it follows (reference = r9y9/gantts @ fb1e75f).  The
reference itself is Python and cannot travel to the GPU box, so this port is what the ``-m gpu``
parity tests, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
leg run there.  It is PINNED: ``tests/golden/make_golden.py`` runs the unmodified reference
(``oracle.reference_loader``) and this port on the same seeded inputs in the build container and
commits the reference outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks the
port against those vectors (bit-exact for masks/indexing, <=1e-6 relative for float outputs: the
port issues the same torch CPU ops in the same order as the reference).

Written as plain functions over explicit weight lists (no nn.Module copies of the reference
classes): parameters are passed as ``[(W0, b0), (W1, b1), ...]`` with ``W`` laid out ``[out, in]``
exactly like the reference's ``nn.Linear`` state_dict entries.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import nnmnkwii_port as nn_port

LEAKY_SLOPE = 0.01          # nn.LeakyReLU() default, reference gantts/models.py:37,132
BCE_EPS = 1e-20             # reference train.py:246,285


# ----------------------------------------------------------------------------- seqloss.py
def sequence_mask(lengths, max_len=None):
    """reference gantts/seqloss.py:9-20 -- ``arange(max_len)[None] < len[:, None]`` as float."""
    lengths = torch.as_tensor(lengths).long().view(-1)
    if max_len is None:
        max_len = int(lengths.max())
    rng = torch.arange(0, int(max_len)).long().unsqueeze(0).expand(lengths.numel(), int(max_len))
    return (rng < lengths.unsqueeze(1)).float()


def masked_mse(inp, target, lengths=None, mask=None, max_len=None):
    """reference gantts/seqloss.py:27-43 -- sum(((in - tgt) * m)^2) / sum(m), m is (B,T,1)."""
    if lengths is None and mask is None:
        raise RuntimeError("Should provide either lengths or mask")
    if mask is None:
        mask = sequence_mask(lengths, max_len).unsqueeze(-1)
    m = mask.expand_as(inp)
    loss = F.mse_loss(inp * m, target * m, reduction="sum")
    return loss / mask.sum()


# ------------------------------------------------------------------------- multistream.py
def get_static_stream_sizes(stream_sizes, has_dynamic_features, num_windows):
    """reference gantts/multistream.py:46-53."""
    out = np.array(stream_sizes)
    sel = np.asarray(has_dynamic_features, dtype=bool)
    out[sel] = out[sel] / num_windows
    return out


def select_streams(inputs, stream_sizes=(60, 1, 1, 1), streams=(True, True, True, True)):
    """reference gantts/multistream.py:33-43 -- column gather of enabled streams."""
    starts = np.hstack(([0], np.cumsum(stream_sizes)[:-1]))
    parts = [inputs[:, :, int(s):int(s) + int(n)]
             for s, n, on in zip(starts, stream_sizes, streams) if on]
    return torch.cat(parts, dim=-1)


def get_static_features(inputs, num_windows, stream_sizes=(180, 3, 1, 3),
                        has_dynamic_features=(True, True, False, True),
                        streams=(True, True, True, True)):
    """reference gantts/multistream.py:56-79 -- static columns of a static+delta tensor."""
    D = inputs.size(-1)
    if stream_sizes is None or (len(stream_sizes) == 1 and has_dynamic_features[0]):
        return inputs[:, :, :D // num_windows]
    if len(stream_sizes) == 1 and not has_dynamic_features[0]:
        return inputs
    starts = np.hstack(([0], np.cumsum(stream_sizes)[:-1]))
    parts = []
    for s, n, dyn, on in zip(starts, stream_sizes, has_dynamic_features, streams):
        if not on:
            continue
        width = int(n) // num_windows if dyn else int(n)
        parts.append(inputs[:, :, int(s):int(s) + width])
    return torch.cat(parts, dim=-1)


def multi_stream_mlpg(inputs, R, stream_sizes=(180, 3, 1, 3),
                      has_dynamic_features=(True, True, False, True),
                      streams=(True, True, True, True)):
    """reference gantts/multistream.py:82-123 -- per-stream dense-R MLPG, static streams copied."""
    if inputs.size(-1) != sum(stream_sizes):
        raise RuntimeError("You probably have specified wrong dimention params.")
    starts = np.hstack(([0], np.cumsum(stream_sizes)[:-1]))
    ends = np.cumsum(stream_sizes)
    parts = []
    for s, e, dyn, on in zip(starts, ends, has_dynamic_features, streams):
        if not on:
            continue
        x = inputs[:, :, int(s):int(e)]
        parts.append(nn_port.unit_variance_mlpg(R, x) if dyn else x)
    return torch.cat(parts, dim=-1)


# ------------------------------------------------------------------------------ models.py
def mlp_forward(x, layers, dropout_p=0.0, training=False, last_sigmoid=False, masks=None):
    """reference gantts/models.py:137-141 -- x = Dropout(LeakyReLU(Linear(x))) per hidden
    layer, then last_linear (+ sigmoid).  ``layers`` = [(W, b), ...]; the last pair is
    ``last_linear``.

    ``masks`` (test hook, SURVEY.md 7 hard part 4): one tensor per hidden layer holding the dropout
    multiplier {0, 1/(1-p)} of every element; when given it REPLACES ``F.dropout`` (same place in
    the chain: Linear -> LeakyReLU -> Dropout), so a train-mode run of the product can be compared
    against this port with the product's own keep decisions injected (torch's Philox stream cannot
    be reproduced on the device)."""
    for l, (W, b) in enumerate(layers[:-1]):
        h = F.leaky_relu(F.linear(x, W, b), LEAKY_SLOPE)
        x = h * masks[l].view_as(h) if masks is not None else F.dropout(h, dropout_p, training)
    W, b = layers[-1]
    x = F.linear(x, W, b)
    return torch.sigmoid(x) if last_sigmoid else x


def in2out_highway_forward(x, R, gate, layers, static_dim, dropout_p=0.0, training=False, masks=None):
    """reference gantts/models.py:54-69 -- returns (y_hat, x_static + sigmoid(T x_static) * MLPG(y_hat))."""
    x = x.unsqueeze(0) if x.dim() == 2 else x
    x_static = x[:, :, :static_dim]
    Tx = torch.sigmoid(F.linear(x_static, gate[0], gate[1]))
    h = mlp_forward(x, layers, dropout_p, training, last_sigmoid=False, masks=masks)
    Gx = nn_port.unit_variance_mlpg(R, h)
    return h, x_static + Tx * Gx


def in2out_rnn_highway_forward(x, R, lengths, gate, lstm, hidden2out, static_dim):
    """reference gantts/models.py:92-118 -- pack -> nn.LSTM -> pad -> hidden2out -> MLPG; returns
    ``(x, x_static + sigmoid(T x_static) * Gx)``: the FIRST output is the input itself (``:118``)."""
    x = x.unsqueeze(0) if x.dim() == 2 else x
    x_static = x[:, :, :static_dim]
    Tx = torch.sigmoid(F.linear(x_static, gate[0], gate[1]))
    if lengths is not None:
        packed = torch.nn.utils.rnn.pack_padded_sequence(x, [int(l) for l in lengths], batch_first=True)
        out, _ = lstm(packed)
        out, _ = torch.nn.utils.rnn.pad_packed_sequence(out, batch_first=True)
    else:
        out, _ = lstm(x)
    out = F.linear(out, hidden2out[0], hidden2out[1])
    Gx = nn_port.unit_variance_mlpg(R, out)
    return x, x_static + Tx * Gx


def lstm_forward(x, lengths, lstm, hidden2out, last_sigmoid=False):
    """reference gantts/models.py:204-213 (LSTMRNN) / :181-190 (GRURNN, also an nn.LSTM):
    pack -> nn.LSTM -> pad -> Linear (-> sigmoid).  ``lstm`` is a torch ``nn.LSTM`` (the oracle
    for the recurrent kernel is torch's own CPU LSTM, as in the reference)."""
    lengths = [int(l) for l in lengths]
    packed = torch.nn.utils.rnn.pack_padded_sequence(x, lengths, batch_first=True)
    out, _ = lstm(packed)
    out, _ = torch.nn.utils.rnn.pad_packed_sequence(out, batch_first=True)
    out = F.linear(out, hidden2out[0], hidden2out[1])
    return torch.sigmoid(out) if last_sigmoid else out


# ------------------------------------------------------------------- train.py step functions
def get_selected_static_stream(y_static, hp):
    """reference train.py:232-242 -- adversarial streams, first ``mask_nth`` mgc columns dropped."""
    sizes = get_static_stream_sizes(hp["stream_sizes"], hp["has_dynamic_features"],
                                    hp["num_windows"])
    sel = select_streams(y_static, sizes, streams=hp["adversarial_streams"])
    if hp.get("mask_nth_mgc_for_adv_loss", 0) > 0:
        sel = sel[:, :, hp["mask_nth_mgc_for_adv_loss"]:]
    return sel


def bce_real(D, mask, T, eps=BCE_EPS):
    """reference train.py:269 -- -(log(D + eps) * m).sum() / T."""
    return -(torch.log(D + eps) * mask).sum() / T


def bce_fake(D, mask, T, eps=BCE_EPS):
    """reference train.py:270 -- -(log(1 - D + eps) * m).sum() / T."""
    return -(torch.log(1 - D + eps) * mask).sum() / T


def clip_grad_norm(grads, max_norm=1.0):
    """torch.nn.utils.clip_grad_norm_ semantics (reference train.py:275,317)."""
    total = torch.sqrt(sum((g.detach() ** 2).sum() for g in grads))
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for g in grads:
        g.mul_(coef)
    return total


def adagrad_step(params, grads, state_sums, lr=0.01, weight_decay=1e-7, eps=1e-10):
    """torch.optim.Adagrad (lr_decay=0, initial_accumulator_value=0) as configured in reference
    hparams.py:223-227,240-244 and stepped at train.py:276,318."""
    with torch.no_grad():
        for p, g, s in zip(params, grads, state_sums):
            g = g + weight_decay * p
            s.addcmul_(g, g, value=1.0)
            p.addcdiv_(g, s.sqrt() + eps, value=-lr)


class AdamStepper(object):
    """torch.optim.Adam (amsgrad off) as configured by reference hparams.py:125-130 for the duration model
    (lr 1e-3, betas (0.5, 0.9), weight_decay 0, eps 1e-8) and stepped at train.py:276,318:
    g' = g + wd p; m = b1 m + (1 - b1) g'; v = b2 v + (1 - b2) g'^2;
    p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps).  Callable as ``stepper(params, grads)``;
    pass it to ``gan_step(..., d_opt=..., g_opt=...)`` in place of the default Adagrad."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.lr, self.b1, self.b2, self.eps, self.wd = float(lr), float(betas[0]), float(betas[1]), float(eps), float(weight_decay)
        self.m = [torch.zeros_like(p) for p in params]
        self.v = [torch.zeros_like(p) for p in params]
        self.t = 0

    def __call__(self, params, grads):
        self.t += 1
        c1, c2 = 1.0 - self.b1 ** self.t, 1.0 - self.b2 ** self.t
        with torch.no_grad():
            for p, g, m, v in zip(params, grads, self.m, self.v):
                g = g + self.wd * p
                m.mul_(self.b1).add_(g, alpha=1.0 - self.b1)
                v.mul_(self.b2).addcmul_(g, g, value=1.0 - self.b2)
                p.addcdiv_(m, v.sqrt() / math.sqrt(c2) + self.eps, value=-self.lr / c1)


class GanStepState(object):
    """Weights + Adagrad accumulators for one MLP-G / MLP-D pair (plain tensors)."""

    def __init__(self, g_layers, d_layers):
        self.g = [(W.clone().requires_grad_(True), b.clone().requires_grad_(True)) for W, b in g_layers]
        self.d = [(W.clone().requires_grad_(True), b.clone().requires_grad_(True)) for W, b in d_layers]
        self.g_sum = [torch.zeros_like(t) for pair in self.g for t in pair]
        self.d_sum = [torch.zeros_like(t) for pair in self.d for t in pair]

    def g_params(self):
        return [t for pair in self.g for t in pair]

    def d_params(self):
        return [t for pair in self.d for t in pair]


def apply_generator(model_out, x, R, hp, include_parameter_generation=False):
    """reference train.py:336-355 given the generator's raw output: models that include parameter
    generation return ``(y_hat, y_hat_static)`` themselves; generic models return ``y_hat`` which is
    (front-)padded to the input length if pad_packed_sequence shortened it (``:347-349``, a no-op
    whenever the longest utterance spans the padded length, as in train.py's own batches) and goes
    through ``multi_stream_mlpg``."""
    if include_parameter_generation:
        return model_out
    y_hat = model_out
    if y_hat.size(1) != x.size(1):
        y_hat = F.pad(y_hat.unsqueeze(0), (0, 0, x.size(1) - y_hat.size(-2), 0)).squeeze(0)
    return y_hat, multi_stream_mlpg(y_hat, R, hp["stream_sizes"], hp["has_dynamic_features"])


def gan_step(g_forward, g_params, g_sum, d_layers, d_sum, x, y, lengths, R, hp, w_d=1.0, mse_w=0.0,
             mge_w=1.0, adv_w=1.0, dropout_d=0.0, training=True, lr=0.01, weight_decay=1e-7, update=True,
             d_masks=None, d_opt=None, g_opt=None):
    """One mini-batch of the reference train_loop body (train.py:528-580) for ANY generator:
    ``g_forward()`` -> ``(y_hat, y_hat_static)`` is the result of ``apply_generator`` (train.py:336-355)
    with autograd history on ``g_params``; ``d_layers`` is the MLP discriminator ``[(W, b), ...]`` or
    None.  ``d_masks`` = {"real": [...], "fake": [...], "adv": [...]} injects dropout multipliers into
    the three discriminator forwards (see ``mlp_forward``).

    Order of operations and quirks preserved: single zero_grad at the
    top; y_hat_static is NOT detached in the discriminator update, so ``loss_d.backward`` also
    deposits the fake-term gradient on the generator; the discriminator steps before the third
    D forward used by the adversarial loss; gradients of both backwards accumulate on G before
    its clip + Adagrad step.  ``training=False`` with ``update=False`` is the "test" phase of
    train.py:481-486 (forwards and losses only).  Returns a dict of python floats and the
    generator outputs."""
    nw = hp["num_windows"]
    y_static = get_static_features(y, nw, hp["stream_sizes"], hp["has_dynamic_features"])   # :528-529
    mask = sequence_mask(lengths, x.size(1)).unsqueeze(-1)                                   # :535
    d_params = [t for pair in d_layers for t in pair] if d_layers is not None else []
    for p in list(g_params) + d_params:                                                      # :538-539
        p.grad = None
    y_hat, y_hat_static = g_forward()                                                        # :542
    out = {}
    T = mask.sum().item()
    cond = hp.get("discriminator_linguistic_condition", False)
    dm = d_masks or {}
    if w_d > 0 and d_layers is not None:
        # update_discriminator, train.py:245-279
        real_in = get_selected_static_stream(y_static, hp)
        fake_in = get_selected_static_stream(y_hat_static, hp)
        if cond:
            real_in = torch.cat((x, real_in), -1)
            fake_in = torch.cat((x, fake_in), -1)
        D_real = mlp_forward(real_in, d_layers, dropout_d, training, last_sigmoid=True, masks=dm.get("real"))
        out["real_correct"] = ((D_real > 0.5).float() * mask).sum().item()
        D_fake = mlp_forward(fake_in, d_layers, dropout_d, training, last_sigmoid=True, masks=dm.get("fake"))
        out["fake_correct"] = ((D_fake < 0.5).float() * mask).sum().item()
        loss_real = bce_real(D_real, mask, T)
        loss_fake = bce_fake(D_fake, mask, T)
        loss_d = loss_real + loss_fake
        if update:
            loss_d.backward(retain_graph=True)
            dg = [p.grad for p in d_params]
            out["d_grad_norm"] = float(clip_grad_norm(dg, 1.0))
            if d_opt is not None:
                d_opt(d_params, dg)                      # e.g. AdamStepper (hparams.py:125-130)
            else:
                adagrad_step(d_params, dg, d_sum, lr, weight_decay)
        out.update(loss_d=loss_d.item(), loss_fake_d=loss_fake.item(), loss_real_d=loss_real.item())
    # update_generator, train.py:282-320
    loss_mge = masked_mse(y_hat_static, y_static, mask=mask)
    loss_mse = masked_mse(y_hat, y, mask=mask)
    if adv_w > 0 and w_d > 0 and d_layers is not None:
        fake_in = get_selected_static_stream(y_hat_static, hp)
        if cond:
            fake_in = torch.cat((x, fake_in), -1)
        D_adv = mlp_forward(fake_in, d_layers, dropout_d, training, last_sigmoid=True, masks=dm.get("adv"))
        loss_adv = bce_real(D_adv, mask, T)
    else:
        loss_adv = y.new_zeros(1)
        adv_w = 0.0
    loss_g = (mse_w * loss_mse + mge_w * loss_mge) + adv_w * loss_adv
    if update:
        loss_g.backward()
        g_params = list(g_params)
        gg = [p.grad if p.grad is not None else torch.zeros_like(p) for p in g_params]
        out["g_grad_norm"] = float(clip_grad_norm(gg, 1.0))
        if g_opt is not None:
            g_opt(g_params, gg)
        else:
            adagrad_step(g_params, gg, g_sum, lr, weight_decay)
    out.update(loss_mse=loss_mse.item(), loss_mge=loss_mge.item(), loss_adv=float(loss_adv.detach()),
               loss_g=float(loss_g.detach()))
    return out, y_hat.detach(), y_hat_static.detach()


def gan_step_mlp(state, x, y, lengths, R, hp, w_d=1.0, mse_w=0.0, mge_w=1.0, adv_w=1.0,
                 dropout_g=0.0, dropout_d=0.0, training=True, lr=0.01, weight_decay=1e-7,
                 update=True, masks=None, d_opt=None, g_opt=None):
    """``gan_step`` for an MLP generator (reference ``MLP`` class, models.py:121-141) held in a
    ``GanStepState``.  ``masks`` = {"g": [...], "real": [...], "fake": [...], "adv": [...]} injects the
    dropout multipliers of the generator forward and of the three discriminator forwards."""
    masks = masks or {}

    def g_forward():
        y_hat = mlp_forward(x, state.g, dropout_g, training, last_sigmoid=False, masks=masks.get("g"))
        return apply_generator(y_hat, x, R, hp)

    return gan_step(g_forward, state.g_params(), state.g_sum, state.d if w_d > 0 else None, state.d_sum,
                    x, y, lengths, R, hp, w_d=w_d, mse_w=mse_w, mge_w=mge_w, adv_w=adv_w, dropout_d=dropout_d,
                    training=training, lr=lr, weight_decay=weight_decay, update=update, d_masks=masks,
                    d_opt=d_opt, g_opt=g_opt)


class GeneratorOracle(object):
    """CPU generator + Adagrad state for ``gan_step`` built from a reference ``state_dict`` (same key
    names as the reference classes, models.py:40-48,84-86,128-131,198-200).  ``kind``:
    "mlp" (MLP), "highway" (In2OutHighwayNet), "rnn_highway" (In2OutRNNHighwayNet), "lstm" (LSTMRNN;
    GRURNN with ``rnn_attr="gru"``).  Recurrent kinds run torch's own CPU ``nn.LSTM`` on packed
    sequences exactly like the reference."""

    def __init__(self, kind, sd, static_dim=None, num_hidden=None, hidden_dim=None, bidirectional=False,
                 rnn_attr="lstm"):
        self.kind, self.static_dim = kind, static_dim
        t = lambda k: torch.as_tensor(np.asarray(sd[k])).clone().float().requires_grad_(True)
        self.named = {}
        if kind in ("mlp", "highway"):
            pre = "layers" if kind == "mlp" else "H"
            n = len([k for k in sd if k.startswith(pre + ".") and k.endswith(".weight")])
            self.layers = [(t("%s.%d.weight" % (pre, i)), t("%s.%d.bias" % (pre, i))) for i in range(n)]
            self.layers.append((t("last_linear.weight"), t("last_linear.bias")))
            for i in range(n):
                self.named["%s.%d.weight" % (pre, i)], self.named["%s.%d.bias" % (pre, i)] = self.layers[i]
            self.named["last_linear.weight"], self.named["last_linear.bias"] = self.layers[-1]
        else:
            in_dim = np.asarray(sd[rnn_attr + ".weight_ih_l0"]).shape[1]
            self.lstm = torch.nn.LSTM(in_dim, hidden_dim, num_hidden, batch_first=True, bidirectional=bidirectional)
            self.lstm.load_state_dict({k[len(rnn_attr) + 1:]: torch.as_tensor(np.asarray(v)).float()
                                       for k, v in sd.items() if k.startswith(rnn_attr + ".")})
            self.lstm.train()
            for k, p in self.lstm.named_parameters():
                self.named[rnn_attr + "." + k] = p
            self.h2o = (t("hidden2out.weight"), t("hidden2out.bias"))
            self.named["hidden2out.weight"], self.named["hidden2out.bias"] = self.h2o
        if kind in ("highway", "rnn_highway"):
            self.gate = (t("T.weight"), t("T.bias"))
            self.named["T.weight"], self.named["T.bias"] = self.gate
        self.sums = [torch.zeros_like(p) for p in self.params()]

    def params(self):
        return list(self.named.values())

    def include_parameter_generation(self):
        return self.kind in ("highway", "rnn_highway")

    def forward(self, x, R, lengths, hp, dropout_p=0.0, training=False, masks=None):
        """(y_hat, y_hat_static) as reference apply_generator (train.py:336-355) returns them."""
        if self.kind == "mlp":
            out = mlp_forward(x, self.layers, dropout_p, training, last_sigmoid=False, masks=masks)
        elif self.kind == "highway":
            out = in2out_highway_forward(x, R, self.gate, self.layers, self.static_dim, dropout_p, training, masks)
        elif self.kind == "rnn_highway":
            out = in2out_rnn_highway_forward(x, R, lengths, self.gate, self.lstm, self.h2o, self.static_dim)
        else:
            out = lstm_forward(x, lengths, self.lstm, self.h2o)
        return apply_generator(out, x, R, hp, self.include_parameter_generation())


def discriminator_layers(sd):
    """[(W, b), ...] (requires_grad) of a reference ``MLP`` state_dict (``layers.i``, ``last_linear``)."""
    t = lambda k: torch.as_tensor(np.asarray(sd[k])).clone().float().requires_grad_(True)
    n = len([k for k in sd if k.startswith("layers.") and k.endswith(".weight")])
    return [(t("layers.%d.weight" % i), t("layers.%d.bias" % i)) for i in range(n)] + \
           [(t("last_linear.weight"), t("last_linear.bias"))]


# ------------------------------------------------------------------------------ SRU (unpinned)
def sru_layer_forward(x, W, b, bidirectional=False, use_tanh=False, use_relu=True, mask_x=None, mask_h=None):
    """SRU v1 layer (Lei et al. 2017, github.com/taolei87/sru ``cuda_functional.py``; NOT vendored
    in the reference tree -- restated from the published recurrence, **parity unpinned**):

        U = x W ; per direction and hidden unit j, with k = 3 (n_in == n_out) or 4 gates
        f_t = sigmoid(U_f + b_f) ; r_t = sigmoid(U_r + b_r)
        c_t = f_t * c_{t-1} + (1 - f_t) * U_x
        h_t = r_t * g(c_t) + (1 - r_t) * x'_t     (x' = x if k == 3 else U_x')

    ``x``: (T, B, n_in); ``W``: (n_in, dirs*k*d) laid out ``[..., dir, d, k]`` (k fastest) as in
    the upstream kernel; ``b``: (dirs*2*d,) = [f-bias | r-bias] per direction.

    Train mode of upstream ``SRUCell.forward``: ``mask_x`` (B, n_in) is the variational ``rnn_dropout``
    multiplier applied to the GEMM input ONLY (``u = (x * mask_x) @ W``; the highway term keeps the unmasked
    ``x``); ``mask_h`` (B, dirs*d) is the ``dropout`` multiplier on ``g(c_t)`` inside the recurrence
    (``h_t = r_t * g(c_t) * mask_h + (1 - r_t) * x'_t``)."""
    T, B, n_in = x.shape
    dirs = 2 if bidirectional else 1
    d = b.numel() // (2 * dirs)
    k = W.shape[1] // (d * dirs)
    xin = x * mask_x.unsqueeze(0) if mask_x is not None else x
    U = (xin.reshape(-1, n_in) @ W).view(T, B, dirs, d, k)
    bias = b.view(dirs, 2, d)
    act = torch.tanh if use_tanh else (torch.relu if use_relu else (lambda v: v))
    outs = []
    for di in range(dirs):
        c = x.new_zeros(B, d)
        hs = [None] * T
        order = range(T) if di == 0 else range(T - 1, -1, -1)
        for t in order:
            u = U[t, :, di]
            f = torch.sigmoid(u[..., 1] + bias[di, 0])
            r = torch.sigmoid(u[..., 2] + bias[di, 1])
            c = f * c + (1 - f) * u[..., 0]
            xp = x[t][:, di * d:(di + 1) * d] if k == 3 else u[..., 3]
            gc = act(c) * mask_h[:, di * d:(di + 1) * d] if mask_h is not None else act(c)
            hs[t] = r * gc + (1 - r) * xp
        outs.append(torch.stack(hs, 0))
    return torch.cat(outs, -1)


# ------------------------------------------------------------------------------------------------------------------
# Logging metrics of the step (reference train.py:358-432): inv_scale :358-381, split_streams :384-397,
# compute_distortions :399-432.  numpy float64; `hp` needs .name and, for "acoustic", .windows / .stream_sizes /
# .has_dynamic_features, for "vc" .order.  nnmnkwii.preprocessing.inv_scale(x, m, s) = x * s + m.
def inv_scale_streams(mgc, lf0, vuv, bap, Y_mean, Y_std, hp):
    nw = len(hp.windows)
    mgc_dim, lf0_dim, vuv_dim, bap_dim = hp.stream_sizes                     # static + dynamic domain (:360)
    lf0_0 = mgc_dim
    vuv_0 = lf0_0 + lf0_dim
    bap_0 = vuv_0 + vuv_dim
    mgc = mgc * Y_std[:mgc_dim // nw] + Y_mean[:mgc_dim // nw]               # :368
    lf0 = lf0 * Y_std[lf0_0:lf0_0 + lf0_dim // nw] + Y_mean[lf0_0:lf0_0 + lf0_dim // nw]
    bap = bap * Y_std[bap_0:bap_0 + bap_dim // nw] + Y_mean[bap_0:bap_0 + bap_dim // nw]
    vuv = vuv * Y_std[vuv_0] + Y_mean[vuv_0]
    return mgc, lf0, (vuv > 0.5).astype(np.int64), bap                       # :376-379


def split_streams_np(y_static, Y_mean, Y_std, hp):
    from . import nnmnkwii_port  # noqa: F401  (same package: the metrics live next to the MLPG port)
    sizes = get_static_stream_sizes(hp.stream_sizes, hp.has_dynamic_features, len(hp.windows))   # :386
    a = np.cumsum([0] + list(sizes))
    y = np.asarray(y_static, dtype=np.float64)
    return inv_scale_streams(y[:, :, a[0]:a[1]], y[:, :, a[1]:a[2]], y[:, :, a[2]], y[:, :, a[3]:], Y_mean, Y_std, hp)


def compute_distortions(y_static, y_hat_static, Y_mean, Y_std, lengths, hp):
    from . import nnmnkwii_port as M
    Y_mean, Y_std = np.asarray(Y_mean, dtype=np.float64), np.asarray(Y_std, dtype=np.float64)
    if hp.name == "acoustic":                                                 # :400-417
        mgc, lf0, vuv, bap = split_streams_np(y_static, Y_mean, Y_std, hp)
        mgc_h, lf0_h, vuv_h, bap_h = split_streams_np(y_hat_static, Y_mean, Y_std, hp)
        try:
            f0_mse = M.lf0_mean_squared_error(lf0, vuv, lf0_h, vuv_h, lengths=lengths, linear_domain=True)
        except ZeroDivisionError:
            f0_mse = float("nan")
        return {"mcd": M.melcd(mgc[:, :, 1:], mgc_h[:, :, 1:], lengths=lengths),
                "bap_mcd": M.melcd(bap, bap_h, lengths=lengths) / 10.0,
                "f0_rmse": float(np.sqrt(f0_mse)),
                "vuv_err": M.vuv_error(vuv, vuv_h, lengths=lengths)}
    if hp.name == "duration":                                                 # :418-422
        a = np.asarray(y_static, dtype=np.float64) * Y_std + Y_mean
        b = np.asarray(y_hat_static, dtype=np.float64) * Y_std + Y_mean
        return {"dur_rmse": float(np.sqrt(M.mean_squared_error(a, b, lengths=lengths)))}
    if hp.name == "vc":                                                       # :423-428
        d = hp.order
        a = np.asarray(y_static, dtype=np.float64) * Y_std[:d] + Y_mean[:d]
        b = np.asarray(y_hat_static, dtype=np.float64) * Y_std[:d] + Y_mean[:d]
        return {"mcd": M.melcd(a, b, lengths=lengths)}
    raise AssertionError("unknown hp.name %r" % (hp.name,))

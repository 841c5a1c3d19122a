"""Fused GAN step: ONE C call (gantts_gan_step) per mini-batch for an MLP generator + MLP
discriminator -- the whole of reference train.py:528-580 enqueued on the current stream without a
single host synchronisation.  Not drop-in for train.py (which owns its step
functions); offered next to the compatible modular path (gantts_b200.step.GanTrainer).

Data parallel: utterance shards, the two flat gradient buffers are SUM all-reduced (NCCL via
torch.distributed on the same stream) between the phases of the step; losses are normalised by the
GLOBAL number of valid frames.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from . import multistream
from . import ops
from . import parallel

LOSS_NAMES = ("loss_d", "loss_fake_d", "loss_real_d", "loss_mse", "loss_mge", "loss_adv", "loss_g",
              "real_correct", "fake_correct", "frames", "d_grad_norm", "g_grad_norm")


def _fill_mlp(desc, model, p, last_act):
    layers = list(model.layers) + [model.last_linear]
    if len(layers) > _lib.MAX_LAYERS:
        raise RuntimeError("gantts_b200: at most %d layers" % _lib.MAX_LAYERS)
    desc.num_layers = len(layers)
    desc.dims[0] = layers[0].weight.shape[1]
    for i, l in enumerate(layers):
        ops.require_cuda(l.weight, l.bias)
        if not (l.weight.is_contiguous() and l.bias.is_contiguous()):
            raise RuntimeError("gantts_b200: parameters must be contiguous")
        desc.dims[i + 1] = l.weight.shape[0]
        desc.W[i] = l.weight.data_ptr()
        desc.b[i] = l.bias.data_ptr()
    desc.slope, desc.dropout_p, desc.last_act, desc.seed = ops.LEAKY_SLOPE, float(p), int(last_act), 0
    return layers


class FusedGanStep(object):
    def __init__(self, model_g, model_d, hp, B, T, w_d=1.0, mse_w=0.0, mge_w=1.0, lr=0.01, weight_decay=1e-7,
                 max_norm=1.0, process_group=None, seed=None, optimizer="Adagrad", optimizer_params=None):
        """``optimizer`` / ``optimizer_params`` mirror ``getattr(optim, hp.optimizer_g)(params, **hp.optimizer_g_params)``
        of reference train.py:784-789 (one setting for both models): "Adagrad" (lr, weight_decay, eps; the defaults
        are hparams.py:201-206) or "Adam" (lr, betas, eps, weight_decay; hparams.py:125-130)."""
        lib = _lib.load()
        if optimizer not in ("Adagrad", "Adam"):
            raise RuntimeError("FusedGanStep: no native optimiser %r (Adagrad and Adam are the ones hparams.py uses)" % optimizer)
        self.optimizer = optimizer
        okw = dict(optimizer_params or {})
        if optimizer == "Adam":
            lr, weight_decay = okw.get("lr", 1e-3), okw.get("weight_decay", 0.0)
        else:
            lr, weight_decay = okw.get("lr", lr), okw.get("weight_decay", weight_decay)
        self.g, self.d, self.hp, self.pg = model_g, model_d, hp, process_group
        self.B, self.T = int(B), int(T)
        dev = next(model_g.parameters()).device
        self.device = dev
        parallel.broadcast_parameters(model_g, group=process_group)
        parallel.broadcast_parameters(model_d, group=process_group)
        c = _lib.GanStepT()
        c.B, c.T = self.B, self.T
        self._g_layers = _fill_mlp(c.g, model_g, model_g.dropout_p, _lib.ACT_NONE)
        self._d_layers = _fill_mlp(c.d, model_d, model_d.dropout_p, _lib.ACT_SIGMOID)
        if model_g.last_sigmoid or not model_d.last_sigmoid:
            raise RuntimeError("FusedGanStep: generator must be linear-output, discriminator sigmoid-output")
        self._sums, self._sqs = [], []      # Adagrad: state_sum | Adam: exp_avg, exp_avg_sq (model.parameters() order)
        for layers, sw, sb, qw, qb in ((self._g_layers, c.g_sumW, c.g_sumb, c.g_sqW, c.g_sqb),
                                       (self._d_layers, c.d_sumW, c.d_sumb, c.d_sqW, c.d_sqb)):
            for i, l in enumerate(layers):
                a, b = torch.zeros_like(l.weight), torch.zeros_like(l.bias)
                self._sums += [a, b]
                sw[i], sb[i] = a.data_ptr(), b.data_ptr()
                if optimizer == "Adam":
                    a2, b2 = torch.zeros_like(l.weight), torch.zeros_like(l.bias)
                    self._sqs += [a2, b2]
                    qw[i], qb[i] = a2.data_ptr(), b2.data_ptr()
        nw = len(hp.windows)
        entries, n_static = multistream.mlpg_stream_entries(hp.stream_sizes, hp.has_dynamic_features,
                                                            [True] * len(hp.stream_sizes), nw)
        c.streams = _lib.make_streams(entries)
        c.windows = _lib.make_windows(hp.windows)
        self._table = ops.mlpg_table(hp.windows, self.T, dev)
        c.mlpg_table = self._table.data_ptr()
        scols = multistream.static_feature_columns(nw, hp.stream_sizes, hp.has_dynamic_features,
                                                   [True] * len(hp.stream_sizes))
        c.n_static, c.n_static_cols = n_static, len(scols)
        for i, v in enumerate(scols):
            c.static_cols[i] = v
        sizes = multistream.get_static_stream_sizes(hp.stream_sizes, hp.has_dynamic_features, nw)
        acols = multistream.select_stream_columns(sizes, hp.adversarial_streams)
        if hp.mask_nth_mgc_for_adv_loss > 0:
            acols = acols[hp.mask_nth_mgc_for_adv_loss:]
        c.n_adv = len(acols)
        for i, v in enumerate(acols):
            c.adv_cols[i] = v
        c.d_conditioned = 1 if hp.discriminator_linguistic_condition else 0
        c.lr_g = c.lr_d = float(lr)
        c.wd_g = c.wd_d = float(weight_decay)
        c.max_norm = float(max_norm)
        if optimizer == "Adam":
            betas = okw.get("betas", (0.9, 0.999))
            self._betas = (float(betas[0]), float(betas[1]))          # as given (the C struct holds them as float32)
            c.optimizer, c.beta1, c.beta2, c.eps = _lib.OPT_ADAM, float(betas[0]), float(betas[1]), float(okw.get("eps", 1e-8))
        else:
            c.optimizer, c.eps = _lib.OPT_ADAGRAD, float(okw.get("eps", 1e-10))
        c.opt_step = 1
        c.w_d, c.mse_w, c.mge_w, c.adv_w = float(w_d), float(mse_w), float(mge_w), 1.0
        self.cfg = c
        nbytes = lib.gantts_gan_step_workspace_bytes(ctypes.byref(c))
        if nbytes == 0:
            raise RuntimeError("gantts_b200 gan_step config rejected: %s" % lib.gantts_last_error_string().decode())
        self._ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        self.losses = torch.zeros(len(LOSS_NAMES), dtype=torch.float32, device=dev)
        self.y_hat = torch.empty(self.B, self.T, c.g.dims[c.g.num_layers], dtype=torch.float32, device=dev)
        self.y_hat_static = torch.empty(self.B, self.T, n_static, dtype=torch.float32, device=dev)
        self._seed = int(seed) if seed is not None else ops.draw_seed() & ((1 << 60) - 1)
        self._step = 0
        self._grad_views = {}

    def grad_buffer(self, which):
        """Flat fp32 gradient buffer (0 = generator, 1 = discriminator) as a tensor view."""
        if which not in self._grad_views:
            lib = _lib.load()
            ptr, cnt = ctypes.c_void_p(), ctypes.c_int64()
            _lib.check(lib.gantts_gan_step_grad_buffer(ctypes.byref(self.cfg), self._ws.data_ptr(), which,
                                                       ctypes.byref(ptr), ctypes.byref(cnt)))
            off = ptr.value - self._ws.data_ptr()
            self._grad_views[which] = self._ws[off:off + 4 * cnt.value].view(torch.float32)
        return self._grad_views[which]

    def _call(self, phases, x, y, lengths, inv_frames, seed):
        lib = _lib.load()
        _lib.check(lib.gantts_gan_step(ctypes.byref(self.cfg), phases, x.data_ptr(), y.data_ptr(),
                                       lengths.data_ptr(), inv_frames, seed, self.y_hat.data_ptr(),
                                       self.y_hat_static.data_ptr(), self.losses.data_ptr(), self._ws.data_ptr(),
                                       self._ws.numel(), ops._stream()))

    def step(self, x, y, lengths, frames=None, adv_w=1.0, train=None):
        """x (B,T,d_in), y (B,T,d_out) contiguous CUDA float32; lengths CUDA int64 (B,); frames = GLOBAL
        number of valid frames (host number; checked against the device-side count when the losses are read,
        see loss_dict).  Returns the device tensor of 12 loss scalars.

        ``train=None`` follows the models like the reference's train_loop does (train.py:481-486): both models
        in ``.train()`` -> training step; both in ``.eval()`` -> the "test" phase (forwards and losses only,
        dropout off, parameters and Adagrad state untouched)."""
        ops.require_cuda(x, y)
        if not (x.is_contiguous() and y.is_contiguous()):
            raise RuntimeError("FusedGanStep: x and y must be contiguous")
        if tuple(x.shape[:2]) != (self.B, self.T) or tuple(y.shape[:2]) != (self.B, self.T):
            raise RuntimeError("FusedGanStep: batch shape differs from the configured (B, T)")
        if not lengths.is_cuda or lengths.dtype != torch.int64:
            raise RuntimeError("FusedGanStep: lengths must be a CUDA int64 tensor")
        self.cfg.adv_w = float(adv_w)
        for desc, layers in ((self.cfg.g, self._g_layers), (self.cfg.d, self._d_layers)):
            for i, l in enumerate(layers):      # parameters may have been re-allocated (load_state_dict keeps them)
                desc.W[i], desc.b[i] = l.weight.data_ptr(), l.bias.data_ptr()
        if train is None:
            if self.g.training != self.d.training:
                raise RuntimeError("FusedGanStep: generator and discriminator disagree on train()/eval()")
            train = self.g.training
        world = torch.distributed.get_world_size(self.pg) if (torch.distributed.is_available()
                                                              and torch.distributed.is_initialized()) else 1
        if frames is None:
            # single process: the step derives 1 / mask.sum() on the device from `lengths` (no host number to trust)
            if world > 1:
                raise RuntimeError("FusedGanStep: data-parallel steps need frames = the GLOBAL number of valid frames")
            self._frames_claim, inv = None, 0.0
        else:
            self._frames_claim = float(frames)
            inv = 1.0 / float(frames)
        if not train:
            self._call(_lib.STEP_EVAL, x, y, lengths, inv, 0)
            return self.losses
        seed = (self._seed + self._step) & ((1 << 61) - 1)
        self.last_seed = seed
        self._step += 1
        self.cfg.opt_step = self._step          # Adam's bias corrections: the number of the step being taken
        if world == 1:
            self._call(7, x, y, lengths, inv, seed)
        else:
            self._call(1, x, y, lengths, inv, seed)
            parallel.allreduce_sum_(self.grad_buffer(1), self.pg)
            self._call(2, x, y, lengths, inv, seed)
            parallel.allreduce_sum_(self.grad_buffer(0), self.pg)
            self._call(4, x, y, lengths, inv, seed)
        return self.losses

    def loss_dict(self):
        """Host copy of the 12 loss scalars (one synchronising read).  Single process: also verifies the `frames`
        the caller passed to step() against the device-side sum of the mask -- a wrong value would silently
        rescale every loss and gradient."""
        v = dict(zip(LOSS_NAMES, self.losses.tolist()))
        world = torch.distributed.get_world_size(self.pg) if (torch.distributed.is_available()
                                                              and torch.distributed.is_initialized()) else 1
        claim = getattr(self, "_frames_claim", None)
        if world == 1 and claim is not None and v["frames"] != claim:
            raise RuntimeError("FusedGanStep: step() was told frames=%g but the lengths sum to %g valid frames"
                               % (claim, v["frames"]))
        return v

    # ---- checkpoint / resume (reference train.py:162-171 save_checkpoint, :174-199 load_checkpoint round-trip
    # optimizer.state_dict(); the layout below is torch.optim.Adagrad's, one entry per parameter in
    # model.parameters() order, so the files are interchangeable with the reference's)
    def _opt_state(self, lo, hi, lr, wd):
        n = hi - lo
        step = torch.tensor(float(self._step))
        if self.optimizer == "Adam":
            return {"state": {i: {"step": step.clone(), "exp_avg": self._sums[lo + i].detach().clone(),
                                  "exp_avg_sq": self._sqs[lo + i].detach().clone()} for i in range(n)},
                    "param_groups": [{"lr": lr, "betas": self._betas,
                                      "eps": float(self.cfg.eps), "weight_decay": wd, "amsgrad": False, "maximize": False,
                                      "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                                      "params": list(range(n))}]}
        return {"state": {i: {"step": step.clone(), "sum": self._sums[lo + i].detach().clone()} for i in range(n)},
                "param_groups": [{"lr": lr, "lr_decay": 0, "eps": float(self.cfg.eps), "weight_decay": wd,
                                  "initial_accumulator_value": 0, "foreach": None, "maximize": False,
                                  "differentiable": False, "fused": None, "params": list(range(n))}]}

    def state_dict(self):
        ng, n = 2 * len(self._g_layers), len(self._sums)
        return {"optimizer_g": self._opt_state(0, ng, float(self.cfg.lr_g), float(self.cfg.wd_g)),
                "optimizer_d": self._opt_state(ng, n, float(self.cfg.lr_d), float(self.cfg.wd_d)),
                "step": self._step, "seed": self._seed}

    def load_state_dict(self, sd):
        ng, n = 2 * len(self._g_layers), len(self._sums)
        for key, lo, hi in (("optimizer_g", 0, ng), ("optimizer_d", ng, n)):
            st = sd[key]["state"]
            for i in range(hi - lo):
                e = st.get(i, st.get(str(i)))
                if e is None:
                    raise RuntimeError("FusedGanStep.load_state_dict: %s has no state for parameter %d" % (key, i))
                if self.optimizer == "Adam":
                    self._sums[lo + i].copy_(e["exp_avg"])
                    self._sqs[lo + i].copy_(e["exp_avg_sq"])
                else:
                    self._sums[lo + i].copy_(e["sum"])
            grp = (sd[key].get("param_groups") or [{}])[0]
            if self.optimizer == "Adam" and "betas" in grp:
                self._betas = (float(grp["betas"][0]), float(grp["betas"][1]))
                self.cfg.beta1, self.cfg.beta2 = self._betas
            if key == "optimizer_g":
                self.cfg.lr_g = float(grp.get("lr", self.cfg.lr_g))
                self.cfg.wd_g = float(grp.get("weight_decay", self.cfg.wd_g))
            else:
                self.cfg.lr_d = float(grp.get("lr", self.cfg.lr_d))
                self.cfg.wd_d = float(grp.get("weight_decay", self.cfg.wd_d))
        self._step = int(sd.get("step", self._step))
        self._seed = int(sd.get("seed", self._seed))

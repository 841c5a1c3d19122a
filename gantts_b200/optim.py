"""clip_grad_norm_(params, max_norm) + torch.optim.Adagrad.step() / torch.optim.Adam.step() as ONE native pass
(gantts_grad_sumsq + gantts_clip_adagrad_step / gantts_clip_adam_step), replacing reference
train.py:275-276,317-318 (optimisers of hparams.py:223-227,240-244,125-130)."""
import ctypes

import torch

from . import _lib
from . import ops


class _ClipOptimizer(object):
    """Flat-buffer optimiser base: ``.grad`` of every parameter is a view into ``flat_grad`` (one buffer => one
    NCCL all-reduce per model under data parallelism); global-norm clipping precedes the update."""

    def __init__(self, params, max_norm=1.0):
        self.params = [p for p in params]
        if not self.params:
            raise RuntimeError("%s: empty parameter list" % type(self).__name__)
        for p in self.params:
            ops.require_cuda(p)
            if not p.is_contiguous():
                raise RuntimeError("%s: parameters must be contiguous" % type(self).__name__)
        self.max_norm = float(max_norm)
        dev = self.params[0].device
        self.total = sum(p.numel() for p in self.params)
        self.flat_grad = torch.zeros(self.total, dtype=torch.float32, device=dev)
        self.sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
        self._grads = self._views(self.flat_grad)
        for p, g in zip(self.params, self._grads):
            p.grad = g
        n = len(self.params)
        self._n = n
        self._sizes = (ctypes.c_int64 * n)(*[p.numel() for p in self.params])
        self._ws = torch.empty(_lib.load().gantts_optim_workspace_bytes(), dtype=torch.uint8, device=dev)
        self.steps = 0

    def _views(self, flat):
        out, off = [], 0
        for p in self.params:
            n = p.numel()
            out.append(flat[off:off + n].view_as(p))
            off += n
        return out

    def _ptrs(self, tensors):
        return (ctypes.c_void_p * self._n)(*[t.data_ptr() for t in tensors])

    def zero_grad(self):
        self.flat_grad.zero_()
        for p, g in zip(self.params, self._grads):
            if p.grad is not g:
                p.grad = g

    def _sumsq(self):
        lib = _lib.load()
        for p, g in zip(self.params, self._grads):
            if p.grad is not g:
                raise RuntimeError("%s: .grad was re-bound; use zero_grad() of this optimizer" % type(self).__name__)
        _lib.check(lib.gantts_grad_sumsq(self._ptrs(self._grads), self._sizes, self._n, self.sumsq.data_ptr(),
                                         self._ws.data_ptr(), self._ws.numel(), ops._stream()))

    def grad_norm(self):
        """Device tensor: total gradient norm seen by the last step (before clipping)."""
        return self.sumsq.sqrt()


class ClipAdagrad(_ClipOptimizer):
    """Adagrad (lr_decay=0, initial_accumulator_value=0, eps=1e-10) preceded by global-norm clipping."""

    def __init__(self, params, lr=0.01, weight_decay=0.0, max_norm=1.0, eps=1e-10):
        super(ClipAdagrad, self).__init__(params, max_norm)
        self.lr, self.weight_decay, self.eps = float(lr), float(weight_decay), float(eps)
        self.flat_sum = torch.zeros_like(self.flat_grad)
        self._sums = self._views(self.flat_sum)

    def step(self):
        lib = _lib.load()
        self._sumsq()
        _lib.check(lib.gantts_clip_adagrad_step(self._ptrs(self.params), self._ptrs(self._grads),
                                                self._ptrs(self._sums), self._sizes, self._n,
                                                self.sumsq.data_ptr(), self.max_norm, self.lr,
                                                self.weight_decay, self.eps, ops._stream()))
        self.steps += 1

    def state_dict(self):
        """torch.optim.Adagrad layout, so reference train.py:162-171 save_checkpoint / load_checkpoint and a
        torch.optim.Adagrad over the same parameters can exchange optimiser state with this class."""
        return {"state": {i: {"step": torch.tensor(float(self.steps)), "sum": s.detach().clone()}
                          for i, s in enumerate(self._sums)},
                "param_groups": [{"lr": self.lr, "lr_decay": 0, "eps": self.eps, "weight_decay": self.weight_decay,
                                  "initial_accumulator_value": 0, "foreach": None, "maximize": False,
                                  "differentiable": False, "fused": None, "params": list(range(self._n))}]}

    def load_state_dict(self, sd):
        if "state" in sd:
            st = sd["state"]
            for i, s in enumerate(self._sums):
                e = st.get(i, st.get(str(i)))
                if e is None:
                    raise RuntimeError("ClipAdagrad.load_state_dict: no state for parameter %d" % i)
                s.copy_(e["sum"])
                self.steps = int(e.get("step", self.steps))
            groups = sd.get("param_groups") or [{}]
            self.lr = float(groups[0].get("lr", self.lr))
            self.weight_decay = float(groups[0].get("weight_decay", self.weight_decay))
            self.eps = float(groups[0].get("eps", self.eps))
        else:                                   # round-1 layout
            for s, v in zip(self._sums, sd["sum"]):
                s.copy_(v)
            self.steps = int(sd.get("steps", 0))


class ClipAdam(_ClipOptimizer):
    """torch.optim.Adam (amsgrad off) preceded by global-norm clipping: the duration model's optimiser
    (reference hparams.py:125-130: lr 1e-3, betas (0.5, 0.9), weight_decay 0)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), weight_decay=0.0, max_norm=1.0, eps=1e-8):
        super(ClipAdam, self).__init__(params, max_norm)
        self.lr, self.betas, self.weight_decay, self.eps = float(lr), (float(betas[0]), float(betas[1])), \
            float(weight_decay), float(eps)
        self.flat_m, self.flat_v = torch.zeros_like(self.flat_grad), torch.zeros_like(self.flat_grad)
        self._m, self._v = self._views(self.flat_m), self._views(self.flat_v)

    def step(self):
        lib = _lib.load()
        self._sumsq()
        self.steps += 1
        _lib.check(lib.gantts_clip_adam_step(self._ptrs(self.params), self._ptrs(self._grads), self._ptrs(self._m),
                                             self._ptrs(self._v), self._sizes, self._n, self.sumsq.data_ptr(),
                                             self.max_norm, self.lr, self.betas[0], self.betas[1], self.weight_decay,
                                             self.eps, self.steps, ops._stream()))

    def state_dict(self):
        return {"state": {i: {"step": torch.tensor(float(self.steps)), "exp_avg": m.detach().clone(),
                              "exp_avg_sq": v.detach().clone()} for i, (m, v) in enumerate(zip(self._m, self._v))},
                "param_groups": [{"lr": self.lr, "betas": self.betas, "eps": self.eps,
                                  "weight_decay": self.weight_decay, "amsgrad": False, "maximize": False,
                                  "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                                  "params": list(range(self._n))}]}

    def load_state_dict(self, sd):
        st = sd["state"]
        for i, (m, v) in enumerate(zip(self._m, self._v)):
            e = st.get(i, st.get(str(i)))
            m.copy_(e["exp_avg"])
            v.copy_(e["exp_avg_sq"])
            self.steps = int(e.get("step", self.steps))
        g = (sd.get("param_groups") or [{}])[0]
        self.lr = float(g.get("lr", self.lr))
        self.betas = tuple(float(b) for b in g.get("betas", self.betas))


def make_optimizer(name, params, **kw):
    """``getattr(optim, hp.optimizer_g)(params, **hp.optimizer_g_params)`` of reference train.py:784-789 for the
    native classes."""
    if name == "Adagrad":
        return ClipAdagrad(params, **kw)
    if name == "Adam":
        return ClipAdam(params, **kw)
    raise RuntimeError("gantts_b200: no native optimiser %r (Adagrad and Adam are the ones hparams.py uses)" % name)

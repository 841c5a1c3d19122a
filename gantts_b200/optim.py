"""clip_grad_norm_(params, max_norm) + torch.optim.Adagrad.step() as ONE native pass
(gantts_grad_sumsq + gantts_clip_adagrad_step), replacing reference train.py:275-276,317-318."""
import ctypes

import torch

from . import _lib
from . import ops


class ClipAdagrad(object):
    """Adagrad (lr_decay=0, initial_accumulator_value=0, eps=1e-10) preceded by global-norm
    clipping, over flat views of the parameters.  ``.grad`` of every parameter is a view into
    ``flat_grad`` (one buffer => one NCCL all-reduce per model under data parallelism)."""

    def __init__(self, params, lr=0.01, weight_decay=0.0, max_norm=1.0, eps=1e-10):
        self.params = [p for p in params]
        if not self.params:
            raise RuntimeError("ClipAdagrad: empty parameter list")
        for p in self.params:
            ops.require_cuda(p)
            if not p.is_contiguous():
                raise RuntimeError("ClipAdagrad: parameters must be contiguous")
        self.lr, self.weight_decay, self.max_norm, self.eps = float(lr), float(weight_decay), float(max_norm), float(eps)
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_sum = torch.zeros(total, dtype=torch.float32, device=dev)
        self.sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
        self._grads, self._sums = [], []
        off = 0
        for p in self.params:
            n = p.numel()
            g = self.flat_grad[off:off + n].view_as(p)
            p.grad = g
            self._grads.append(g)
            self._sums.append(self.flat_sum[off:off + n].view_as(p))
            off += n
        n = len(self.params)
        self._n = n
        self._sizes = (ctypes.c_int64 * n)(*[p.numel() for p in self.params])
        self._ws = torch.empty(_lib.load().gantts_optim_workspace_bytes(), dtype=torch.uint8, device=dev)
        self.steps = 0

    def _ptrs(self, tensors):
        return (ctypes.c_void_p * self._n)(*[t.data_ptr() for t in tensors])

    def zero_grad(self):
        self.flat_grad.zero_()
        for p, g in zip(self.params, self._grads):
            if p.grad is not g:
                p.grad = g

    def step(self):
        lib = _lib.load()
        for p, g in zip(self.params, self._grads):
            if p.grad is not g:
                raise RuntimeError("ClipAdagrad: .grad was re-bound; use zero_grad() of this optimizer")
        st = ops._stream()
        _lib.check(lib.gantts_grad_sumsq(self._ptrs(self._grads), self._sizes, self._n, self.sumsq.data_ptr(),
                                         self._ws.data_ptr(), self._ws.numel(), st))
        _lib.check(lib.gantts_clip_adagrad_step(self._ptrs(self.params), self._ptrs(self._grads),
                                                self._ptrs(self._sums), self._sizes, self._n,
                                                self.sumsq.data_ptr(), self.max_norm, self.lr,
                                                self.weight_decay, self.eps, st))
        self.steps += 1

    def grad_norm(self):
        """Device tensor: total gradient norm seen by the last step (before clipping)."""
        return self.sumsq.sqrt()

    def state_dict(self):
        return {"sum": [s.clone() for s in self._sums], "steps": self.steps,
                "lr": self.lr, "weight_decay": self.weight_decay}

    def load_state_dict(self, sd):
        for s, v in zip(self._sums, sd["sum"]):
            s.copy_(v)
        self.steps = int(sd.get("steps", 0))

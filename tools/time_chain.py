"""Micro-timing of the discriminator stack (gantts_mlp_fwd / gantts_mlp_bwd through ops.mlp_stack) at the cfg2 row
counts: CUDA events around forward and backward, 30 iterations after 20 warm-ups.  Run once per environment setting
(GANTTS_B200_CHAIN=0|1, GANTTS_B200_CHAIN_DBG=...): the switches are read once per process."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__
__graft_entry__.build()
from gantts_b200 import ops, _lib

dev = torch.device("cuda:0")
torch.manual_seed(0)
tag = "CHAIN=%s DBG=%s" % (os.environ.get("GANTTS_B200_CHAIN", "default"), os.environ.get("GANTTS_B200_CHAIN_DBG", "0"))
iters = int(os.environ.get("ITERS", "30"))


def run(M, dims, p, need_gx, need_gw):
    Ws = [(torch.randn(o, i) / i ** 0.5).to(dev).requires_grad_(need_gw) for i, o in zip(dims[:-1], dims[1:])]
    bs = [torch.zeros(o, device=dev, requires_grad=need_gw) for o in dims[1:]]
    x = torch.randn(M, dims[0], device=dev, requires_grad=need_gx)
    g = torch.randn(M, 1, device=dev)
    f = lambda: ops.mlp_stack(x, Ws, bs, p=p, training=p > 0, last_act=_lib.ACT_SIGMOID, seed=5)
    for _ in range(20):
        f().backward(g)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = tb = 0.0
    for _ in range(iters):
        e[0].record()
        y = f()
        e[1].record()
        y.backward(g)
        e[2].record()
        torch.cuda.synchronize()
        tf += e[0].elapsed_time(e[1])
        tb += e[1].elapsed_time(e[2])
    fl = 2.0 * M * sum(a * b for a, b in zip(dims[:-1], dims[1:]))
    print("%-22s M=%6d gx=%d gw=%d  fwd %7.1f us (%4.0f TF exec)  bwd %7.1f us" % (
        tag, M, need_gx, need_gw, tf / iters * 1e3, 3 * fl / (tf / iters * 1e-3) / 1e12, tb / iters * 1e3), flush=True)


D = [58, 256, 256, 256, 1]
run(64000, D, 0.5, True, True)      # stacked real | fake pass (weight gradients + input gradient)
run(32000, D, 0.5, True, False)     # adversarial pass (input gradient only)

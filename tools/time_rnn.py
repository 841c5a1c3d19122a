"""Timing of the recurrent generators (cfg3-like VC BiLSTM, B=16 x T=2000) fwd+bwd on the GPU."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__
__graft_entry__.build()
import gantts_b200
dev = torch.device("cuda:0")
torch.manual_seed(0)


def run(name, model, B, T, I, n=3):
    model.to(dev).train()
    x = torch.randn(B, T, I, device=dev, requires_grad=True)
    lens = [T] * B
    g = None
    for it in range(n + 1):
        if it == 1:
            torch.cuda.synchronize(); t0 = time.perf_counter()
        y = model(x, lens)
        if g is None:
            g = torch.randn_like(y)
        y.backward(g)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record(); y = model(x, lens); e[1].record(); y.backward(g); e[2].record()
    torch.cuda.synchronize()
    print("%-46s B=%d T=%d: fwd+bwd %.1f ms (fwd %.1f, bwd %.1f) -> %.3g frames/s" % (
        name, B, T, dt * 1e3, e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), B * T / dt), flush=True)


run("LSTMRNN 177-(3x512 bi)-177 [cfg3 width]", gantts_b200.models.LSTMRNN(177, 177, 3, 512, bidirectional=True), 16, 2000, 177)
run("LSTMRNN 425-(3x512 bi)-187 [cfg5 width]", gantts_b200.models.LSTMRNN(425, 187, 3, 512, bidirectional=True), 64, 1500, 425, n=2)
run("SRURNN 425-(6x512 bi)-187 [hparams default]", gantts_b200.models.SRURNN(425, 187, 6, 512, bidirectional=True, use_relu=1), 32, 1000, 425)

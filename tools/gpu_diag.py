"""One-shot GPU diagnostics: error metrics of every CUDA op against the CPU oracle.  Prints a
table and never raises, so that a single gpurun call gives the full picture."""
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import WINDOWS, rel_err  # noqa: E402
from oracle import gantts_port as gp  # noqa: E402
from oracle import nnmnkwii_port as nnp  # noqa: E402
import gantts_b200  # noqa: E402
from gantts_b200 import ops, _lib  # noqa: E402

dev = torch.device("cuda:0")
results = []


def report(name, fn):
    t0 = time.time()
    try:
        out = fn()
        torch.cuda.synchronize()
        results.append((name, "ok", out))
        print("[diag] %-46s %s  (%.2fs)" % (name, out, time.time() - t0), flush=True)
    except Exception as e:  # noqa
        results.append((name, "EXC", repr(e)))
        print("[diag] %-46s EXC %r" % (name, e), flush=True)
        traceback.print_exc()
        try:
            torch.cuda.synchronize()
        except Exception as e2:  # noqa
            print("[diag] device unusable after failure: %r" % (e2,), flush=True)
            raise SystemExit(3)


def fmt(**kw):
    return " ".join("%s=%.2e" % (k, v) for k, v in kw.items())


def t_mlpg(B, T, nw=3):
    def f():
        torch.manual_seed(0)
        wins = WINDOWS[:nw]
        x = torch.randn(B, T, 187 if nw == 3 else 2 * 20)
        if nw == 3:
            ss, dyn = [180, 3, 1, 3], [True, True, False, True]
        else:
            ss, dyn = [40], [True]
        R = torch.from_numpy(nnp.unit_variance_mlpg_matrix(wins, T))
        xr = x.clone().requires_grad_(True)
        yr = gp.multi_stream_mlpg(xr, R, ss, dyn, [True] * len(ss))
        g = torch.randn_like(yr)
        yr.backward(g)
        xg = x.to(dev).requires_grad_(True)
        yg = gantts_b200.multistream.multi_stream_mlpg(xg, R.to(dev), ss, dyn, [True] * len(ss))
        yg.backward(g.to(dev))
        return fmt(fwd=rel_err(yg.detach().cpu().numpy(), yr.detach().numpy()),
                   bwd=rel_err(xg.grad.cpu().numpy(), xr.grad.numpy()))
    return f


def t_linear(M, K, N, act, engine, p=0.0):
    def f():
        torch.manual_seed(1)
        x = torch.randn(M, K)
        W = torch.randn(N, K) / np.sqrt(K)
        b = torch.randn(N) * 0.1
        xr, Wr, br = [t.clone().double().requires_grad_(True) for t in (x, W, b)]
        z = torch.nn.functional.linear(xr, Wr, br)
        if act == 1:
            yr = torch.nn.functional.leaky_relu(z, 0.01)
        elif act == 2:
            yr = torch.sigmoid(z)
        else:
            yr = z
        g = torch.randn(M, N)
        yr.backward(g.double())
        xg, Wg, bg = [t.to(dev).requires_grad_(True) for t in (x, W, b)]
        yg = ops.linear_act(xg, Wg, bg, act, p=p, training=p > 0, engine=engine)
        yg.backward(g.to(dev))
        return fmt(y=rel_err(yg.detach().cpu().numpy(), yr.detach().numpy()),
                   gx=rel_err(xg.grad.cpu().numpy(), xr.grad.numpy()),
                   gW=rel_err(Wg.grad.cpu().numpy(), Wr.grad.numpy()),
                   gb=rel_err(bg.grad.cpu().numpy(), br.grad.numpy()))
    return f


def t_losses():
    torch.manual_seed(2)
    B, T, D = 7, 53, 63
    a, b = torch.randn(B, T, D), torch.randn(B, T, D)
    lens = torch.LongTensor(sorted(np.random.RandomState(0).randint(20, T + 1, B), reverse=True))
    ar = a.clone().requires_grad_(True)
    lr_ = gp.masked_mse(ar, b, lengths=lens, max_len=T)
    lr_.backward()
    ag = a.to(dev).requires_grad_(True)
    lg = gantts_b200.seqloss.MaskedMSELoss()(ag, b.to(dev), lengths=lens.to(dev), max_len=T)
    lg.backward()
    m_ok = bool((gantts_b200.seqloss.sequence_mask(lens.to(dev), T).cpu() == gp.sequence_mask(lens, T)).all())
    # BCE
    Dv = torch.rand(B, T, 1)
    Dv[0, 0, 0] = 1.0
    Dv[0, 1, 0] = 0.0
    mask = gp.sequence_mask(lens, T).unsqueeze(-1)
    Tn = mask.sum().item()
    out = {}
    for kind, fn in ((0, gp.bce_real), (1, gp.bce_fake)):
        dr = Dv.clone().requires_grad_(True)
        l = fn(dr, mask, Tn)
        l.backward()
        dg = Dv.to(dev).requires_grad_(True)
        o = ops.masked_bce(dg, mask.to(dev), kind)
        (o[0] / Tn).backward()
        out["bce%d" % kind] = abs(o[0].item() / Tn - l.item()) / abs(l.item())
        gr, gg = dr.grad.numpy(), dg.grad.cpu().numpy()
        fin = np.isfinite(gr) & (np.abs(gr) < 1e10)
        out["bce%dg" % kind] = rel_err(gg[fin], gr[fin])
    return "mask_exact=%s " % m_ok + fmt(mse=abs(lg.item() - lr_.item()) / abs(lr_.item()),
                                         mse_g=rel_err(ag.grad.cpu().numpy(), ar.grad.numpy()), **out)


def t_gather():
    x = torch.randn(4, 9, 187)
    xg = x.to(dev).requires_grad_(True)
    a = gantts_b200.multistream.get_static_features(xg, 3)
    r = gp.get_static_features(x, 3)
    a.sum().backward()
    s = gantts_b200.multistream.select_streams(a.detach(), [60, 1, 1, 1], [True, False, False, True])
    sr = gp.select_streams(r, [60, 1, 1, 1], [True, False, False, True])
    return "static_exact=%s select_exact=%s grad_sum=%.1f" % (
        bool((a.detach().cpu() == r).all()), bool((s.cpu() == sr).all()), xg.grad.sum().item())


def t_optim():
    torch.manual_seed(3)
    import ctypes
    lib = _lib.load()
    ps = [torch.randn(40, 30), torch.randn(40), torch.randn(7, 40)]
    gs = [torch.randn_like(p) for p in ps]
    ss = [torch.rand_like(p) for p in ps]
    pr = [p.clone() for p in ps]
    gr = [g.clone() for g in gs]
    sr = [s.clone() for s in ss]
    gp.clip_grad_norm(gr, 1.0)
    gp.adagrad_step(pr, gr, sr)
    pd, gd, sd = [[t.to(dev) for t in L] for L in (ps, gs, ss)]
    n = len(ps)
    arr = lambda L: (ctypes.c_void_p * n)(*[t.data_ptr() for t in L])
    sizes = (ctypes.c_int64 * n)(*[t.numel() for t in ps])
    sumsq = torch.zeros(1, device=dev)
    ws = torch.empty(lib.gantts_optim_workspace_bytes(), dtype=torch.uint8, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.gantts_grad_sumsq(arr(gd), sizes, n, sumsq.data_ptr(), ws.data_ptr(), ws.numel(), st))
    _lib.check(lib.gantts_clip_adagrad_step(arr(pd), arr(gd), arr(sd), sizes, n, sumsq.data_ptr(),
                                            1.0, 0.01, 1e-7, 1e-10, st))
    return fmt(p=max(rel_err(a.cpu().numpy(), b.numpy()) for a, b in zip(pd, pr)),
               s=max(rel_err(a.cpu().numpy(), b.numpy()) for a, b in zip(sd, sr)))


def main():
    print("device:", torch.cuda.get_device_name(0), "supported:", _lib.load().gantts_device_supported())
    report("gather/static/select", t_gather)
    report("losses (mask, mse, bce)", t_losses)
    report("optim clip+adagrad", t_optim)
    report("mlpg B=3 T=37", t_mlpg(3, 37))
    report("mlpg B=2 T=100 nw=2", t_mlpg(2, 100, 2))
    report("mlpg B=4 T=1000", t_mlpg(4, 1000))
    for eng in ("simt", "tc"):
        for (M, K, N, act) in [(300, 20, 32, 1), (1000, 425, 512, 1), (1111, 512, 187, 0), (999, 58, 256, 1),
                               (640, 256, 1, 2), (4096, 512, 512, 1)]:
            report("linear %s M=%d K=%d N=%d act=%d" % (eng, M, K, N, act), t_linear(M, K, N, act, eng))
    bad = [r for r in results if r[1] != "ok"]
    print("[diag] done: %d ok, %d failed" % (len(results) - len(bad), len(bad)))


if __name__ == "__main__":
    main()

"""Same-box library bar for the MLP GAN step (VERDICT r01 missing #3): the reference ALGORITHM on torch-CUDA
-- nn.Linear-style matmuls on cuBLAS, eager element-wise ops, dense-R MLPG matmul, F.dropout, the reference's .item() syncs --
i.e. what the reference's own GPU path runs, next to gantts_gan_step on the same box at cfg2 (B=32, T=1000, 425->187,
D 58-256-256-256-1, dropout 0.5).  Not a test (not collected); lives under tests/ because it drives the checker's step
function (oracle.gantts_port.gan_step_mlp) with CUDA tensors.  Prints a markdown table (kept as
profiles/r02_step_vs_torch_eager.md).  python tests/time_torch_eager_step.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__  # noqa: E402
__graft_entry__.build()
import bench  # noqa: E402
from oracle import gantts_port as gp, nnmnkwii_port as nnp  # noqa: E402

dev = torch.device("cuda:0")
w = bench.WORKLOADS["cfg2"]
B, T = w["B"], w["T"]
hp = dict(w["hp"])

_mask = gp.sequence_mask
gp.sequence_mask = lambda lengths, max_len=None: _mask(lengths, max_len).to(dev)      # the checker builds its mask on the CPU


def timeit(fn, warm=3, iters=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


torch.manual_seed(1234)
mg, md = bench.build_models(w, dev)
names = ["layers.0", "layers.1", "layers.2", "last_linear"]


def layers_of(m):
    sd = m.state_dict()
    return [(sd[n + ".weight"].detach().clone(), sd[n + ".bias"].detach().clone()) for n in names]


state = gp.GanStepState(layers_of(mg), layers_of(md))
host = bench.make_batches(w, 1234, 2, pinned=False)
batches = [(x.to(dev), y.to(dev)) for x, y in host]
lens = [T] * B
R = torch.from_numpy(nnp.unit_variance_mlpg_matrix(bench.WINDOWS, T)).to(dev)
k = [0]


def eager_step():
    x, y = batches[k[0] % 2]
    k[0] += 1
    gp.gan_step_mlp(state, x, y, lens, R, hp, w_d=1.0, mse_w=0.0, mge_w=1.0, dropout_g=0.5, dropout_d=0.5, training=True)


rows = []
for tf32 in (False, True):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    rows.append(("torch-CUDA eager, cuBLAS %s" % ("TF32 (outside the 1e-4 bar)" if tf32 else "fp32"), timeit(eager_step)))
torch.backends.cuda.matmul.allow_tf32 = False

from gantts_b200 import fused, step as gstep  # noqa: E402
hpo = gstep.HParams(windows=bench.WINDOWS, stream_sizes=hp["stream_sizes"], has_dynamic_features=hp["has_dynamic_features"],
                    adversarial_streams=hp["adversarial_streams"], mask_nth_mgc_for_adv_loss=hp["mask_nth_mgc_for_adv_loss"],
                    discriminator_linguistic_condition=False)
fs = fused.FusedGanStep(mg, md, hpo, B, T, w_d=1.0, mse_w=0.0, mge_w=1.0)
lengths = torch.full((B,), T, dtype=torch.int64, device=dev)


def ours():
    x, y = batches[k[0] % 2]
    k[0] += 1
    fs.step(x, y, lengths, frames=B * T)


t_ours = timeit(ours, warm=5, iters=50)
print("| GAN step at cfg2 (B=32 x T=1000, 425->187, D 58-256-256-256-1, dropout 0.5, both optimiser steps) | ms/step | frames/s | vs ours |")
print("|---|---|---|---|")
for name, ms in rows:
    print("| %s: the reference algorithm (dense-R MLPG matmul with R resident, F.dropout, its %d host syncs) | %.2f | %.2f M | %.1fx slower |"
          % (name, 10, ms, B * T / ms / 1e3, ms / t_ours))
print("| gantts_gan_step (bf16x3 tcgen05 GEMMs, fused epilogues, MLPG substitution, no host sync) | %.3f | %.2f M | 1.0 |"
      % (t_ours, B * T / t_ours / 1e3))

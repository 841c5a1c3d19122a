"""torchrun --nproc-per-node N tools/ddp_check.py : utterance-sharded FusedGanStep over NCCL gives the same
losses / gradient norms / updated weights as the single-process global batch (dropout 0)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__  # noqa: E402
from gantts_b200 import parallel  # noqa: E402

rank, world, local = parallel.init_from_env()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if rank == 0:
    __graft_entry__.build()
if world > 1:
    torch.distributed.barrier()
import gantts_b200  # noqa: E402
from gantts_b200 import fused, step as gstep  # noqa: E402


def models():
    torch.manual_seed(7)
    g = gantts_b200.models.MLP(64, 187, 3, 128, dropout=0.0, last_sigmoid=False).to(dev)
    d = gantts_b200.models.MLP(58, 1, 3, 64, dropout=0.0, last_sigmoid=True).to(dev)
    return g, d


torch.manual_seed(3)
B, T = 4 * world, 96
lens = sorted([T] + [int(v) for v in torch.randint(T // 2, T, (B - 1,))], reverse=True)
x = torch.rand(B, T, 64)
y = torch.randn(B, T, 187)
for b, n in enumerate(lens):
    x[b, n:] = 0
    y[b, n:] = 0
frames = float(sum(lens))
# sharded run
idx = parallel.shard_indices(B, rank, world)
g, d = models()
fs = fused.FusedGanStep(g, d, gstep.TTS_ACOUSTIC, len(idx), T)
for it in range(2):
    fs.step(x[idx].to(dev), y[idx].to(dev), torch.tensor([lens[i] for i in idx], device=dev), frames=frames)
torch.cuda.synchronize()
shard_losses = fs.losses.clone()
# local loss sums are per shard; reduce the additive ones for comparison
tot = shard_losses.clone()
if world > 1:
    torch.distributed.all_reduce(tot)
if rank == 0:
    torch.distributed.destroy_process_group() if False else None
    g2, d2 = models()
    # single process, global batch (no process group involvement: world-size-1 path)
    import gantts_b200.parallel as par
    saved = par.allreduce_sum_
    par.allreduce_sum_ = lambda t, group=None: t
    fs2 = fused.FusedGanStep.__new__(fused.FusedGanStep)
    os.environ["GANTTS_DDP_CHECK"] = "1"
print("rank", rank, "sharded losses", [round(v, 6) for v in shard_losses.tolist()[:7]], flush=True)
# every rank also computes the global batch locally with all-reduce disabled
import gantts_b200.fused as F  # noqa: E402
F.parallel.allreduce_sum_ = lambda t, group=None: t
_ws = torch.distributed.get_world_size
torch.distributed.get_world_size = lambda group=None: 1
g2, d2 = models()
fs2 = fused.FusedGanStep(g2, d2, gstep.TTS_ACOUSTIC, B, T)
for it in range(2):
    fs2.step(x.to(dev), y.to(dev), torch.tensor(lens, device=dev), frames=frames)
torch.cuda.synchronize()
torch.distributed.get_world_size = _ws
ref = fs2.losses
# additive quantities: loss sums (already normalised by the global frame count) add up across shards
names = fused.LOSS_NAMES
ok = True
for k in (0, 1, 2, 4, 5, 6, 7, 8, 9):
    a, b = float(tot[k]), float(ref[k])
    rel = abs(a - b) / max(abs(b), 1e-12)
    if rel > 2e-4:
        ok = False
    if rank == 0:
        print("%-12s sharded-sum %.6f  global %.6f  rel %.2e" % (names[k], a, b, rel))
# grad norms and weights are identical on every rank after the all-reduce
for k in (10, 11):
    a, b = float(shard_losses[k]), float(ref[k])
    rel = abs(a - b) / abs(b)
    ok = ok and rel < 2e-4
    if rank == 0:
        print("%-12s sharded %.6f  global %.6f  rel %.2e" % (names[k], a, b, rel))
wd = max(float((p - q).abs().max()) for p, q in zip(g.parameters(), g2.parameters()))
if rank == 0:
    print("max |W_sharded - W_global| over generator params: %.3e (lr = 1e-2)" % wd)
    print("DDP CHECK", "PASS" if ok else "FAIL")
if world > 1:
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()

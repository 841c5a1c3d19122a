"""world_size-2 gloo test (CPU) of the data-parallel host logic: utterance sharding, GLOBAL
valid-frame normalisation and SUM all-reduce give the single-process global-batch gradient."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import WINDOWS, TTS_HP


def _grads(state, x, y, lens, R, T_global):
    from oracle import gantts_port as gp
    mask = gp.sequence_mask(lens, x.size(1)).unsqueeze(-1)
    y_static = gp.get_static_features(y, 3)
    y_hat = gp.mlp_forward(x, state.g)
    y_hat_static = gp.multi_stream_mlpg(y_hat, R)
    sse = ((y_hat_static * mask - y_static * mask) ** 2).sum()
    loss = sse / T_global
    grads = torch.autograd.grad(loss, state.g_params())
    return torch.cat([g.reshape(-1) for g in grads]), mask.sum()


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    from gantts_b200 import parallel
    from oracle import gantts_port as gp
    from oracle import nnmnkwii_port as nnp
    r, w, _ = parallel.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    B, Tn = 6, 24
    g_layers = [(torch.randn(16, 10) * 0.3, torch.zeros(16)), (torch.randn(187, 16) * 0.3, torch.zeros(187))]
    state = gp.GanStepState(g_layers, [(torch.randn(1, 58), torch.zeros(1))])
    x, y = torch.rand(B, Tn, 10), torch.randn(B, Tn, 187)
    lens = [24, 22, 20, 17, 13, 12]
    R = torch.from_numpy(nnp.unit_variance_mlpg_matrix(WINDOWS, Tn))
    idx = parallel.shard_indices(B, rank, world)
    # rule 2: shards keep the GLOBAL padded length; rule 1: normalise by the GLOBAL frame count
    t_local = torch.tensor([float(sum(lens[i] for i in idx))])
    t_global = parallel.allreduce_sum_(t_local.clone())
    g_local, _ = _grads(state, x[idx], y[idx], [lens[i] for i in idx], R, float(t_global))
    g_sum = parallel.allreduce_sum_(g_local.clone())
    g_ref, t_ref = _grads(state, x, y, lens, R, float(sum(lens)))
    q.put((rank, float(t_global), float(t_ref), float((g_sum - g_ref).abs().max() / g_ref.abs().max())))
    dist.destroy_process_group()


def test_sharded_gradients_equal_global_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in procs]
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    for rank, tg, tr, err in res:
        assert tg == tr == 108.0
        assert err < 1e-5, err


def test_shard_indices_partition():
    from gantts_b200 import parallel
    for B, w in ((32, 8), (7, 2), (5, 4)):
        all_idx = sorted(i for r in range(w) for i in parallel.shard_indices(B, r, w))
        assert all_idx == list(range(B))

"""Utterance-sharded data parallelism (new functionality: the reference is single-process,
SURVEY.md 2.1 / 8e).  One process per GPU; every op on the path is independent across utterances,
so the only exchange is a SUM all-reduce of the flat gradient buffers (NCCL over NVLink/NVSwitch)
plus one scalar all-reduce of the valid-frame count.

Parity rules for "same result as the single-process global batch":
  1. local losses are normalised by the GLOBAL valid-frame count and gradients are SUMMED;
  2. every shard is padded to the GLOBAL max_len (MLPG solves over the padded length);
  3. clipping happens after the reduction.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's RANK/WORLD_SIZE/MASTER_* (no-op for 1 process).
    Returns (rank, world_size, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_indices(batch_size, rank, world):
    """Utterances b == rank (mod world) of the length-sorted batch: balances lengths across ranks."""
    return list(range(rank, batch_size, world))


def allreduce_sum_(tensor, group=None):
    """In-place SUM all-reduce when a multi-rank group is active; identity otherwise."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
    return tensor


def broadcast_parameters(module, src=0, group=None):
    """Identical initial weights on every rank."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        for p in module.parameters():
            dist.broadcast(p.data, src=src, group=group)

"""torchrun --nproc-per-node N tools/diag_multi.py : per-repeat timing of the data-parallel fused step on every rank --
device time (CUDA events) and host enqueue time of each 50-step loop, first without and then with bench.py's NVML clock
sampler thread, then with the two all-reduces replaced by no-ops (same three C calls).  Diagnoses where a slow repeat
comes from (host starvation, the sampler, NCCL)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__  # noqa: E402
import bench  # noqa: E402
from gantts_b200 import parallel  # noqa: E402

rank, world, local = parallel.init_from_env()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if rank == 0:
    __graft_entry__.build()
if world > 1:
    torch.distributed.barrier()
from gantts_b200 import _lib, fused, step as gstep  # noqa: E402

w = bench.WORKLOADS["cfg2"]
torch.manual_seed(1234)
mg, md = bench.build_models(w, dev)
hpd = w["hp"]
hp = gstep.HParams(windows=bench.WINDOWS, stream_sizes=hpd["stream_sizes"], has_dynamic_features=hpd["has_dynamic_features"],
                   adversarial_streams=hpd["adversarial_streams"], mask_nth_mgc_for_adv_loss=hpd["mask_nth_mgc_for_adv_loss"],
                   discriminator_linguistic_condition=False)
lengths = torch.full((w["B"],), w["T"], dtype=torch.int64, device=dev)
frames = w["B"] * w["T"] * world
fs = fused.FusedGanStep(mg, md, hp, w["B"], w["T"], w_d=1.0, mse_w=0.0, mge_w=1.0)
host = bench.make_batches(w, 1234 + rank, 4, pinned=False)
res = [(x.to(dev), y.to(dev)) for x, y in host]
print("rank %d: cpu_count %s affinity %d OMP_NUM_THREADS=%s" % (rank, os.cpu_count(), len(os.sched_getaffinity(0)),
                                                              os.environ.get("OMP_NUM_THREADS")), flush=True)
for _ in range(12):
    parallel.allreduce_sum_(fs.grad_buffer(0))
    parallel.allreduce_sum_(fs.grad_buffer(1))
fs.grad_buffer(0).zero_()
fs.grad_buffer(1).zero_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def repeat(tag, n=50):
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    for i in range(n):
        fs.step(res[i % 4][0], res[i % 4][1], lengths, frames=frames)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("rank %d %-22s device %.3f ms/step  host enqueue %.3f ms/step  wall %.3f ms/step" % (
        rank, tag, e0.elapsed_time(e1) / n, (t1 - t0) / n * 1e3, (t2 - t0) / n * 1e3), flush=True)


for i in range(10):
    fs.step(res[i % 4][0], res[i % 4][1], lengths, frames=frames)
for k in range(5):
    repeat("plain #%d" % k)
sampler = bench.ClockSampler(local)
if rank == 0:
    sampler.start()
for k in range(5):
    repeat("rank0 sampler on #%d" % k)
if rank == 0:
    print("sampler:", sampler.stop(), flush=True)
for k in range(3):
    repeat("sampler off #%d" % k)
# same three C calls per step, no all-reduce in between
saved = parallel.allreduce_sum_
parallel.allreduce_sum_ = lambda t, group=None: t
for k in range(3):
    repeat("no all-reduce #%d" % k)
parallel.allreduce_sum_ = saved
if world > 1:
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()

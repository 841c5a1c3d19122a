"""bench.py -- GAN-step frames/sec of the B200 hot path (BASELINE.json metric) + roofline + CPU baseline.

    python bench.py --gpus N --steps K --warmup W            # our arm (N>1 under torchrun)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port)

A "step" is one full GAN training step (reference train.py:538-580: zero_grad, generator forward,
MLPG, discriminator update, generator update, both clip+Adagrad steps) over one synthetic batch of
the workload BASELINE.json quotes the metric on: cfg2 = TTS acoustic MLP generator
(425-512-512-512-187) + MLP discriminator on the 58 masked mgc dims (58-256-256-256-1), B=32 x
T=1000 padded frames PER GPU (weak scaling), dropout 0.5 in train mode, full-length utterances.

One JSON line on stdout (rank 0); see the task contract for the keys.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(1, os.path.join(ROOT, "compat"))

WORKLOAD = dict(B=32, T=1000, d_in=425, d_out=187, g_hidden=512, g_layers=3, d_in_adv=58, d_hidden=256,
                d_layers=3, dropout_g=0.5, dropout_d=0.5)
NUM_BATCHES = 4          # ring of distinct synthetic batches (4 x 78 MB > 126 MB L2)


def algorithmic_flops_per_frame(w):
    """MAC/frame = 3 F_G - k0 + 8 F_D - 256 d_in  (de-duplicated GAN step)."""
    g_dims = [w["d_in"]] + [w["g_hidden"]] * w["g_layers"] + [w["d_out"]]
    d_dims = [w["d_in_adv"]] + [w["d_hidden"]] * w["d_layers"] + [1]
    F_G = sum(a * b for a, b in zip(g_dims[:-1], g_dims[1:]))
    F_D = sum(a * b for a, b in zip(d_dims[:-1], d_dims[1:]))
    k0 = g_dims[0] * g_dims[1]
    return 2.0 * (3 * F_G - k0 + 8 * F_D - w["d_hidden"] * w["d_in_adv"])


def make_batches(w, seed, n, pinned):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        x = torch.rand(w["B"], w["T"], w["d_in"], generator=g) * 0.98 + 0.01
        y = torch.randn(w["B"], w["T"], w["d_out"], generator=g)
        if pinned:
            x, y = x.pin_memory(), y.pin_memory()
        out.append((x, y))
    return out


class ClockSampler(object):
    """SM clock / throttle reasons sampled DURING the timed region (NVML polling thread; falls back to an
    `nvidia-smi -lms 100` subprocess when the NVML bindings are unavailable)."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nvml = index, [], None, None
        self.sm, self.mask, self.max_mhz, self._stop = [], 0, None, False

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            # CUDA_VISIBLE_DEVICES-relative index -> NVML handle through the PCI bus id
            bus = torch.cuda.get_device_properties(self.index).pci_bus_id if hasattr(
                torch.cuda.get_device_properties(self.index), "pci_bus_id") else None
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index) if bus is None else None
            if self.handle is None:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    h = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if pynvml.nvmlDeviceGetPciInfo(h).bus == bus:
                        self.handle = h
                        break
                if self.handle is None:
                    self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self._stop:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                self.mask |= int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
            except Exception:
                pass
            time.sleep(0.005)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self._stop = True
            self.thread.join(timeout=1.0)
            reasons = sorted(v for k, v in self.REASONS.items() if self.mask & k)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                    "reasons": reasons, "samples": len(self.sm), "how": "NVML poll every 5 ms"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "how": "nvidia-smi -lms 100"}


# ------------------------------------------------------------------------------ CPU baseline
def cpu_reference_step_runner(w, sample_B, threads):
    """The reference's CPU path (oracle port of gantts/*.py + train.py step functions, torch CPU
    fp32, dense-R MLPG exactly like the reference) on a bounded sample: `sample_B` utterances of the
    workload's T and widths.  Returns a closure running one step and the frames it processes."""
    from oracle import gantts_port as gp
    from oracle import nnmnkwii_port as nnp
    torch.set_num_threads(threads)
    torch.manual_seed(1234)
    mk = lambda i, o: (torch.nn.Linear(i, o).weight.detach().clone(), torch.zeros(o))
    g_dims = [w["d_in"]] + [w["g_hidden"]] * w["g_layers"] + [w["d_out"]]
    d_dims = [w["d_in_adv"]] + [w["d_hidden"]] * w["d_layers"] + [1]
    state = gp.GanStepState([mk(a, b) for a, b in zip(g_dims[:-1], g_dims[1:])],
                            [mk(a, b) for a, b in zip(d_dims[:-1], d_dims[1:])])
    hp = dict(stream_sizes=[180, 3, 1, 3], has_dynamic_features=[True, True, False, True],
              adversarial_streams=[True, False, False, False], mask_nth_mgc_for_adv_loss=2, num_windows=3,
              discriminator_linguistic_condition=False)
    windows = [(0, 0, np.array([1.0])), (1, 1, np.array([-0.5, 0.0, 0.5])), (1, 1, np.array([1.0, -2.0, 1.0]))]
    (x, y), = make_batches(dict(w, B=sample_B), 1234, 1, False)
    lens = [w["T"]] * sample_B
    R = torch.from_numpy(nnp.unit_variance_mlpg_matrix(windows, w["T"]))   # memoised: not re-timed per step

    def run():
        gp.gan_step_mlp(state, x, y, lens, R, hp, w_d=1.0, mse_w=0.0, mge_w=1.0, adv_w=1.0,
                        dropout_g=w["dropout_g"], dropout_d=w["dropout_d"], training=True)
    return run, sample_B * w["T"]


def time_cpu_baseline(w, steps, warmup, sample_B=4):
    """Reference CPU arm.  torch's CPU kernels do not scale to every core of a large host on these
    shapes (128 threads were 3x slower than 32 on the B200 box), so a one-step probe picks the fastest
    thread count among {16, 32, 64, all} and the timed run uses it; `cores` reports that choice."""
    ncpu = os.cpu_count() or 1
    cands = sorted(set(min(c, ncpu) for c in (16, 32, 64, ncpu)))
    best, best_dt = cands[0], None
    run, frames = cpu_reference_step_runner(w, sample_B, cands[0])
    run()                                   # first-touch / allocator warm-up
    for c in cands:
        torch.set_num_threads(c)
        t0 = time.perf_counter()
        run()
        dt = time.perf_counter() - t0
        if best_dt is None or dt < best_dt:
            best, best_dt = c, dt
    torch.set_num_threads(best)
    for _ in range(max(0, warmup - 1)):
        run()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    dt = (time.perf_counter() - t0) / steps
    return {"value": frames / dt, "unit": "frames/s", "cores": best, "kind": "port",
            "sample": "%d of %d utterances x T=%d per step, %d timed steps, dense-R MLPG with R prebuilt, "
                      "torch %s CPU fp32, thread count picked from %s of %d host cores by a one-step probe"
                      % (sample_B, w["B"], w["T"], steps, torch.__version__, cands, ncpu),
            "ms_per_step": dt * 1e3}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOAD
    steps, warmup = max(1, min(args.steps, 8)), max(1, min(args.warmup, 2))
    cb = time_cpu_baseline(w, steps, warmup)
    line = {"impl": "reference", "metric": "gan_step_frames_per_sec", "value": cb["value"], "unit": "frames/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": cb["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(w, "cpu"), "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit_json_line(line)


def workload_config(w, engine):
    return {"workload": "cfg2: TTS acoustic MLP G 425-512-512-512-187 + MLP D 58-256-256-256-1 (mgc adv, "
                        "mask_nth=2), B=32 T=1000 per GPU, MGE(MLPG)+ADV, dropout 0.5, Adagrad",
            "global_batch_per_gpu": w["B"], "seq_len": w["T"], "engine": engine,
            "l2": "per-step working set ~1.3 GB and a ring of %d distinct input batches (%d MB) exceed the "
                  "126 MB L2; no explicit flush" % (NUM_BATCHES, NUM_BATCHES * 78),
            "parallelism": "utterance-sharded data parallel, one NCCL SUM all-reduce per model per step"}


_REAL_STDOUT_FD = None


def redirect_stdout_to_stderr():
    global _REAL_STDOUT_FD
    if _REAL_STDOUT_FD is None:
        sys.stdout.flush()
        _REAL_STDOUT_FD = os.dup(1)
        os.dup2(2, 1)


def emit_json_line(line):
    text = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT_FD is None:
        os.write(1, text)
    else:
        os.write(_REAL_STDOUT_FD, text)


# ------------------------------------------------------------------------------------ B200 arm
def run_b200_arm(args):
    if os.environ.get("GANTTS_B200_DBG", "0") not in ("", "0"):
        raise SystemExit("bench.py: GANTTS_B200_DBG is a phase-timing switch that skips work inside the kernels; "
                         "refusing to produce a benchmark line with it set")
    import __graft_entry__
    from gantts_b200 import parallel
    # stdout carries exactly ONE JSON line.  NCCL prints its version banner with a bare printf on fd 1 during
    # the first communicator init (NCCL_DEBUG_FILE does not catch it), so fd 1 points at stderr for the whole
    # run and the JSON line is written to the saved descriptor (emit_json_line).
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    redirect_stdout_to_stderr()
    rank, world, local = parallel.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if rank == 0:
        __graft_entry__.build()
    if world > 1:
        torch.distributed.barrier()
    import gantts_b200
    from gantts_b200 import _lib, step as gstep, config
    from nnmnkwii.paramgen import unit_variance_mlpg_matrix   # compat shim (product side, memoised)
    lib = _lib.load()
    config.engine = args.engine
    w = WORKLOAD
    torch.manual_seed(1234)
    mg = gantts_b200.models.MLP(w["d_in"], w["d_out"], w["g_layers"], w["g_hidden"], dropout=w["dropout_g"],
                                last_sigmoid=False).to(dev).train()
    md = gantts_b200.models.MLP(w["d_in_adv"], 1, w["d_layers"], w["d_hidden"], dropout=w["dropout_d"],
                                last_sigmoid=True).to(dev).train()
    hp = gstep.TTS_ACOUSTIC
    lengths = torch.full((w["B"],), w["T"], dtype=torch.int64, device=dev)
    frames_global = w["B"] * w["T"] * world

    class Runner(object):
        """path 'fused': gantts_gan_step (one C call per mini-batch); 'modular': GanTrainer."""

        def __init__(self):
            if args.path == "fused":
                from gantts_b200 import fused
                self.fs = fused.FusedGanStep(mg, md, hp, w["B"], w["T"], w_d=1.0, mse_w=0.0, mge_w=1.0)
            else:
                self.tr = gstep.GanTrainer(mg, md, hp, w_d=1.0, mse_w=0.0, mge_w=1.0)
                self.R = torch.from_numpy(unit_variance_mlpg_matrix(hp.windows, w["T"])).to(dev)

        def step(self, x, y, lengths, R=None):
            if args.path == "fused":
                losses = self.fs.step(x, y, lengths, frames=frames_global)
                return {"loss_d": losses[0], "loss_mge": losses[4], "loss_adv": losses[5], "loss_g": losses[6]}, None, None
            return self.tr.step(x, y, lengths, self.R)

    trainer = Runner()
    R = None
    host = make_batches(w, 1234 + rank, NUM_BATCHES, pinned=True)
    resident = [(x.to(dev), y.to(dev)) for x, y in host]
    frames_per_step = w["B"] * w["T"] * world

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    # ---------------- value: inputs resident in HBM
    for i in range(args.warmup):
        trainer.step(*resident[i % NUM_BATCHES], lengths, R)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = lib.gantts_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        out, _, _ = trainer.step(*resident[i % NUM_BATCHES], lengths, R)
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = lib.gantts_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms_total / args.steps
    value = frames_per_step / (ms_per_step * 1e-3)
    loss_g = float(out["loss_g"])
    # ---------------- roofline pass: the same K steps again with CUDA events around every GEMM / MLPG launch
    # (kept out of the region `value` is timed on: the event records sit between consecutive kernels)
    import ctypes
    lib.gantts_profile_enable(1)
    for i in range(args.steps):
        trainer.step(*resident[i % NUM_BATCHES], lengths, R)
    torch.cuda.synchronize()
    lib.gantts_profile_enable(0)
    pms, pwork, pn = (ctypes.c_double * 8)(), (ctypes.c_double * 8)(), (ctypes.c_longlong * 8)()
    _lib.check(lib.gantts_profile_collect(pms, pwork, pn))

    # ---------------- e2e: host buffers in, losses out, copies inside the timed region
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [(torch.empty_like(resident[0][0]), torch.empty_like(resident[0][1])) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    loss_host = torch.empty(4, dtype=torch.float32).pin_memory()

    def prefetch(i):
        slot = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])
            bufs[slot][0].copy_(host[i % NUM_BATCHES][0], non_blocking=True)
            bufs[slot][1].copy_(host[i % NUM_BATCHES][1], non_blocking=True)
            ready[slot].record(copy_stream)

    def e2e_loop(n):
        prefetch(0)
        for i in range(n):
            if i + 1 < n:
                prefetch(i + 1)
            slot = i % 2
            torch.cuda.current_stream().wait_event(ready[slot])
            o, _, _ = trainer.step(bufs[slot][0], bufs[slot][1], lengths, R)
            consumed[slot].record()
            loss_host.copy_(torch.stack([o["loss_g"], o["loss_d"], o["loss_mge"], o["loss_adv"]]), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for ev in consumed:
        ev.record()
    e2e_loop(max(2, min(args.warmup, 4)))
    barrier()
    t0 = time.perf_counter()
    e0.record()
    e2e_loop(args.steps)
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    h2d = sum(t.numel() * 4 for t in host[0])

    if rank != 0:
        return
    # ---------------- roofline of the dominant kernel (tcgen05 GEMM, K-major instance)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained") or 1400.0
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PFLOP/s sustained (of fallback)"
    kinds = {0: "gemm_bf16x3_kernel<K-major>", 1: "gemm_bf16x3_kernel<MN-major>", 2: "mlpg_fwd_kernel", 3: "mlpg_bwd_kernel"}
    per_kind = {}
    for k, name in kinds.items():
        if pn[k]:
            per_kind[name] = {"ms_per_step": pms[k] / args.steps, "launches_per_step": pn[k] / args.steps,
                              "work_per_step": pwork[k] / args.steps}
    dom = max((0, 1), key=lambda k: pms[k])
    ach = (pwork[dom] / (pms[dom] * 1e-3)) / 1e12 if pms[dom] else 0.0
    traffic = None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "r01_top_kernel.json")))
        traffic = prof.get("dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {"kernel": kinds[dom], "bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": ach / peak_tf, "traffic": traffic, "peak_source": peak_src,
                "note": "achieved = algorithmic fp32-equivalent FLOPs (2MNK per GEMM) / CUDA-event time of the "
                        "launches (second pass of the same K steps with an event pair around every launch); the "
                        "bf16x3 split executes 3 tensor-core MMAs per algorithmic product, so executed bf16 pipe "
                        "rate = 3 x achieved",
                "tensor_pipe_frac_executed": 3.0 * ach / peak_tf,
                "share_of_step": (pms[dom] / args.steps) / ms_per_step,
                "kernels": per_kind}
    cb = time_cpu_baseline(w, steps=3, warmup=1) if world == 1 and not args.no_cpu_baseline else None
    line = {"metric": "gan_step_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (GEMMs: bf16x3 split on tcgen05, fp32 accumulate)" if args.engine == "tc" else "f32",
            "data": "synthetic", "config": workload_config(w, args.engine), "clocks": clocks,
            "e2e": {"value": frames_per_step / (e2e_ms * 1e-3), "unit": "frames/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 16,
                    "how": "pinned host x,y -> double-buffered cudaMemcpyAsync on a copy stream -> one step through the "
                           "public API (%s) -> 4 loss scalars D2H; copies inside the timed region" % args.path},
            "gpu_launches": int(launches), "gpu_launches_per_step": launches / args.steps,
            "algorithmic_gflop_per_step": algorithmic_flops_per_frame(w) * w["B"] * w["T"] / 1e9,
            "step_tflops_algorithmic": algorithmic_flops_per_frame(w) * frames_per_step / (ms_per_step * 1e-3) / 1e12 / world,
            "roofline": roofline, "loss_g_last": loss_g,
            "path": "gantts_gan_step (one C call per mini-batch)" if args.path == "fused" else "GanTrainer (python-orchestrated native ops)"}
    if cb is not None:
        line["cpu_baseline"] = cb
    emit_json_line(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--engine", default=os.environ.get("GANTTS_B200_ENGINE", "tc"), choices=["tc", "simt"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--path", default="fused", choices=["fused", "modular"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py — candidates/s of the DAN forward (BASELINE.json metric) on N B200s, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--precision bf16|fp32] [--batch B]

A "step" is one pass of the hot path (encoder -> conv stack -> pooling -> FC -> heads, dl4vc/model.py:434-961) over one
batch of B synthetic PROD-shaped pileup candidates (100 reads x 201 positions, all read rows populated = the dense
worst case). Candidates are independent, so N GPUs shard by candidate with no collective ("scaling": "weak").

  value     whole-job candidates/s, inputs already resident in HBM (device pointers -> dan_forward)
  e2e       the same metric through the host-buffer call: pinned uint8 host tensors -> H2D -> dan_forward -> D2H of the
            (B,27) head matrix, every step, all inside the timed region
  roofline  dominant kernel class (conv stack) timed with CUDA events on the launching stream by the library's
            profile hook; tensor-pipe bound (2.65e5 FLOP/B >> ridge), peak = MEASURED_PEAKS.json
  cpu_baseline / --impl reference
            the reference's own torch-CPU op sequence (oracle/dan_torch_cpu.py — the reference is pure PyTorch and
            /root/reference does not exist on the GPU box) on all host cores, bounded sample of the same workload

The per-step batch (default 4144 candidates = 252 MB of uint8 input) is larger than the 126 MB L2, so no L2 flush is
needed between timed iterations.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "candidate variants/sec (DAN forward)"
UNIT = "candidates/s"
WORKLOAD = "PROD DAN forward (call_variants.sh:101-147 flag set: 7 conv layers x128ch, highway 32, FC 73856-1024-256), "\
           "100 reads x 201 bp, dense pileups, 1M-candidate job processed in per-step batches"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"bf16_burst": float(d["bf16_tflops"]), "bf16_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                "hbm_gbs": float(d["hbm_gbs"]), "source": "measured"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.1):
        super().__init__(daemon=True)
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.power = [], set(), []
        self.max_mhz = None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:  # NVML unavailable: record why
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                except Exception:
                    pass
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(self.period)

    def stop(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=2)
        s = sorted(self.samples)
        med = s[len(s) // 2] if s else None
        pw = sorted(self.power)
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s),
                "power_w": round(pw[len(pw) // 2], 1) if pw else None}


# ---------------------------------------------------------------------------------------------------------
def cpu_reference_rate(cfg, sd, sample, batch, repeats, warmup=1):
    """candidates/s of the reference's torch-CPU op sequence on this host (all cores)."""
    import torch

    from oracle import dan_torch_cpu

    torch.set_num_threads(os.cpu_count() or 1)
    arrs = sample.slice(0, batch).arrays()
    r, q, s, ref, rm, vm = arrs
    for _ in range(warmup):
        dan_torch_cpu.forward(cfg, sd, r, ref, q, s, rm, vm)
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        dan_torch_cpu.forward(cfg, sd, r, ref, q, s, rm, vm)
        times.append(time.perf_counter() - t0)
    return batch * len(times) / sum(times), times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch

    from dl4vc_b200.config import prod_config
    from dl4vc_b200.synth import make_pileups
    from dl4vc_b200.weights import synth_state_dict

    cfg = prod_config()
    sd = synth_state_dict(cfg, seed=1)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample = make_pileups(256, seed=20261018, coverage="full")
    # size a step so that the whole run stays within a few minutes: probe 16 candidates, aim at ~2 s per step
    rate_probe, _ = cpu_reference_rate(cfg, sd, sample, 16, 1, warmup=1)
    per_step = int(min(256, max(16, (rate_probe * 2.0) // 16 * 16)))
    total_steps = args.steps + args.warmup
    while per_step > 16 and per_step * total_steps / rate_probe > 240:
        per_step -= 16
    from oracle import dan_torch_cpu

    r, q, s, ref, rm, vm = sample.slice(0, per_step).arrays()
    for _ in range(args.warmup):
        dan_torch_cpu.forward(cfg, sd, r, ref, q, s, rm, vm)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        dan_torch_cpu.forward(cfg, sd, r, ref, q, s, rm, vm)
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "candidates_per_step": per_step, "model": "PROD", "reads": cfg.num_reads, "window": cfg.read_len},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} timed steps x {per_step} dense PROD candidates, torch {torch.__version__} CPU ops "
                                   "(oracle/dan_torch_cpu.py = the reference's own op sequence, fp32)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch

    from dl4vc_b200 import _lib
    from dl4vc_b200.config import prod_config
    from dl4vc_b200.factory import build_model
    from dl4vc_b200.synth import make_pileups
    from dl4vc_b200.weights import synth_state_dict

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (the DAN forward has no CPU path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries the ONE JSON line only: NCCL prints its version banner (and, at NCCL_DEBUG=INFO, its log) to fd 1 when the
        # communicator is created, so fd 1 points at stderr until the first collective has run
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    lib = _lib.load_library()     # fails loudly if the CUDA library is missing
    cfg = prod_config()
    sd = synth_state_dict(cfg, seed=1)
    model = build_model(cfg, sd, device=dev, precision=args.precision)
    if args.pass_candidates:
        model.set_pass_candidates(args.pass_candidates)

    # ---- synthetic inputs: `unique` distinct candidates per rank, tiled to the step batch ---------------
    B = args.batch
    unique = min(B, 512)
    sample = make_pileups(unique, seed=20261018 + rank, coverage="full")
    reps = (B + unique - 1) // unique
    host = [torch.from_numpy(np.ascontiguousarray(np.concatenate([a] * reps, axis=0)[:B])).pin_memory() for a in sample.arrays()]
    h_r, h_q, h_s, h_ref, h_rm, h_vm = host
    d_r, d_q, d_s, d_ref, d_rm, d_vm = [t.to(dev) for t in host]
    in_bytes = sum(t.numel() for t in host)
    out_host = torch.empty((B, _lib.NUM_HEAD_OUTPUTS), dtype=torch.float32).pin_memory()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms

    def step_resident():
        return model.forward_heads(d_r, d_ref, d_q, d_s, d_rm, d_vm)

    def step_e2e():
        # pinned host uint8 -> chunked H2D on the library's side stream, overlapped with the kernels -> D2H of the (B,27) result
        return model.forward_heads_host(h_r, h_ref, h_q, h_s, h_rm, h_vm, out=out_host)

    # ---- warm-up, then the timed regions ------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local)
    sampler.start()
    ms = timed(step_resident, args.steps)
    launches = model.last_launch_count * args.steps
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    # dominant-kernel timing: the same steps with the library's event hook on
    lib.dan_profile_enable(1)
    prof_steps = min(args.steps, 3)
    timed(step_resident, prof_steps)
    import ctypes as C
    ms_cls = (C.c_double * _lib.PROF_NUM_CLASSES)()
    n_cls = (C.c_int * _lib.PROF_NUM_CLASSES)()
    _lib.check(lib.dan_profile_read(ms_cls, n_cls, _lib.PROF_NUM_CLASSES), "dan_profile_read")
    lib.dan_profile_enable(0)
    clocks = sampler.stop()

    total_cands = B * world * args.steps
    value = total_cands / (ms * 1e-3)
    e2e_value = total_cands / (ms_e2e * 1e-3)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peaks = load_peaks()
    R, P, Cc = cfg.num_reads, cfg.read_len, cfg.channels
    # conv-stack class = conv(1x3) + residual 1x1 + bottleneck 1x1 MACs; compression/FC/heads run in the GEMM class
    macs_conv_stack = 0
    cin = cfg.in_channels
    for l in range(1, cfg.total_conv_layers + 1):
        macs_conv_stack += R * P * Cc * cin * 3
        cin = Cc
        if cfg.is_residual(l):
            macs_conv_stack += R * P * Cc * Cc
        if cfg.highway:
            macs_conv_stack += R * P * Cc * cfg.bottleneck
    flops_total = 2 * cfg.macs_per_candidate()
    cls_names = [lib.dan_profile_class_name(i).decode() for i in range(_lib.PROF_NUM_CLASSES)]
    cls_ms = {cls_names[i]: ms_cls[i] / prof_steps for i in range(_lib.PROF_NUM_CLASSES)}
    cls_n = {cls_names[i]: n_cls[i] // prof_steps for i in range(_lib.PROF_NUM_CLASSES)}
    dom = "conv_stack"
    dom_ms_step = cls_ms[dom]
    dom_launches = max(cls_n[dom], 1)
    flops_per_launch = 2.0 * macs_conv_stack * B / dom_launches
    achieved = flops_per_launch / (dom_ms_step / dom_launches * 1e-3) / 1e12 if dom_ms_step > 0 else 0.0
    peak = peaks["bf16_sustained"] if args.precision == "bf16" else None
    roofline = {
        "bound": "tensor", "kernel": "dan_stack_kernel (conv stack class)", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
        "frac": (achieved / peak) if peak else None,
        # dram__bytes_read.sum + dram__bytes_write.sum of dan_stack_kernel from the committed `ncu --set full` capture
        # (profiles/r01e_stack_kernel_ncu_full_summary.csv: 1.384 GB + 2.500 GB for the two segment launches of a 148-candidate pass
        # = 26.24 MB per candidate), averaged per launch like `achieved`
        "traffic": 26.24e6 * B / dom_launches, "traffic_source": "ncu --set full, profiles/r01e (26.2 MB per candidate over both segment launches)",
        "peak_source": f"{peaks['source']} bf16_tflops_sustained (kernel timed inside a long step); burst {peaks['bf16_burst']}",
        "launches_per_step": dom_launches, "avg_launch_ms": dom_ms_step / dom_launches,
        "share_of_step": dom_ms_step / sum(cls_ms.values()) if sum(cls_ms.values()) > 0 else None,
        "class_ms_per_step": cls_ms, "class_launches_per_step": cls_n,
        "whole_forward_tflops": value / world * flops_total / 1e12,
        "whole_forward_frac": (value / world * flops_total / 1e12 / peak) if peak else None,
        "hbm_frac_algorithmic": value / world * 60923 / (peaks["hbm_gbs"] * 1e9),
    }
    if args.precision == "fp32":
        # the 1e-4-parity path runs on the CUDA-core FMA pipe and has no per-class event hooks: whole-forward FLOP
        # rate against the fp32 FMA peak of the part (148 SMs x 128 lanes x 2 FLOP x max SM clock)
        fma_peak = 148 * 128 * 2 * (clocks.get("sm_max_mhz") or 1965) * 1e6 / 1e12
        ach = value / world * flops_total / 1e12
        roofline = {"bound": "fp32_fma (CUDA cores)", "kernel": "whole forward (sgemm_taps_kernel dominates)", "achieved": ach, "peak": fma_peak,
                    "unit": "TFLOP/s", "frac": ach / fma_peak, "traffic": None, "peak_source": "nominal: 148 SMs x 128 FMA lanes x 2 x sm_max_mhz"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "candidates_per_step_per_gpu": B, "model": "PROD", "reads": R, "window": P,
                   "precision": args.precision,
                   "l2_policy": (f"inputs per step {in_bytes/1e6:.0f} MB > 126 MB L2, no flush" if in_bytes > 126e6 else
                                 f"inputs per step {in_bytes/1e6:.0f} MB; intermediates of a step ({B * 31.3:.0f} MB written and re-read) exceed the 126 MB L2, no flush"),
                   "parallelism": f"candidate-sharded x{world}, no collective"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": B * _lib.NUM_HEAD_OUTPUTS * 4,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
    }
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        nb = 64
        rate, times = cpu_reference_rate(cfg, sd, sample, nb, repeats=2, warmup=1)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"2 timed batches x {nb} of the same dense PROD candidates (1 warm-up), torch {torch.__version__} "
                                          f"CPU fp32, oracle/dan_torch_cpu.py = the reference's op sequence; batch times {['%.2f' % t for t in times]} s"}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=4144, help="candidates per step per GPU (28 passes of 148 candidates)")
    ap.add_argument("--pass-candidates", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())

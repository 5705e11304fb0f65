#!/usr/bin/env python
"""Headline benchmark: stage-4 256x256 train images/sec (BASELINE.json `metric`, configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch-per-gpu B] [--impl ours|reference]

One "step" = one full trainer step of `UncertaintyGuidedProgressiveTrainer` at stage 4 on one batch
of synthetic MoNuSeg-shaped input: PGUNet4 forward (train-mode BN), PGUNet3 eval forward for the
uncertainty map, fused uncertainty-weighted BCE loss + metrics, backward (dgrad + wgrad + BN/ReLU/
pool/upsample adjoints), gradient all-reduce when N > 1, RMSprop step.  Nothing is skipped or cached.

Prints ONE JSON line (rank 0).  `value` = images/s with inputs resident in HBM; `e2e` = the same
through the public trainer API from pinned HOST buffers (H2D of the batch and D2H of the step
metrics inside the timed region).  `roofline` is for the dominant kernel class (the tcgen05
implicit-GEMM convolution, forward + dgrad launches) timed per launch with CUDA events in an
instrumented step.  `cpu_baseline` / `--impl reference` time the CPU oracle port of the reference
step (oracle/ug_oracle.py; the Python reference itself cannot travel to the GPU box) on a bounded
sample.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GF_PER_IMG_MODEL = 185.3070        # conv GEMM GFLOP / image, stage-4 train step
GF_PER_IMG_UNC = 39.9796           # stage-3 uncertainty forward
STAGE, RES = 4, 256


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"bf16_tflops": d.get("bf16_tflops", 1590.0), "bf16_tflops_sustained": d.get("bf16_tflops_sustained", 1400.0),
                "hbm_gbs": d.get("hbm_gbs", 6650.0), "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._stop_evt = index, [], set(), None, threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def synth_batch(B, seed, device=None, pinned=False):
    """The reference's own DummyDataset distribution (demo_uncertainty_guided.py:29-36)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 3, RES, RES, generator=g)
    y = torch.randint(0, 2, (B, 1, RES, RES), generator=g).float()
    if pinned:
        return x.pin_memory(), y.pin_memory()
    return (x.to(device), y.to(device)) if device else (x, y)


def cpu_reference_step_rate(batch: int, steps: int, warmup: int):
    """Times the CPU oracle port of the reference trainer step (fwd + unc + loss + bwd + RMSprop)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ug_oracle as O
    sd = O.seeded_state_dict(4, 3, 1, seed=4)
    sd_prev = O.seeded_state_dict(3, 3, 1, seed=3)
    state = {k: torch.zeros_like(v) for k, v in sd.items() if v.is_floating_point()}
    x, y = synth_batch(batch, 1234)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        out = O.train_step(sd, x, y, 4, sd_prev)
        for k, g in out["grads"].items():
            O.rmsprop_step(sd[k], g, state[k], 1e-4)
        for k, v in out["bn"].items():
            sd[k] = v
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    return batch / dt, dt, torch.get_num_threads()


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = args.cpu_batch
    steps = max(1, min(args.steps, 3))
    ips, dt, cores = cpu_reference_step_rate(batch, steps, min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": "stage-4 256x256 train images/sec", "value": ips, "unit": "images/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: PGUNet4(3,1) stage-4 256x256 MoNuSeg-shaped trainer step "
                               "(train fwd + PGUNet3 uncertainty fwd + fused loss + bwd + RMSprop)",
                   "stage": 4, "resolution": 256, "batch_per_gpu": args.batch_per_gpu,
                   "global_batch": args.batch_per_gpu * args.gpus, "parallelism": "cpu",
                   "cpu_sample_batch": batch, "gflop_per_image": GF_PER_IMG_MODEL + GF_PER_IMG_UNC},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} step(s) of the stage-4 trainer step at batch {batch} "
                                   "(oracle/ug_oracle.py, fp32, torch CPU)"},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the sm_100a path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import ug_pg_unet_b200 as U
    from ug_pg_unet_b200 import ops

    B = args.batch_per_gpu
    torch.manual_seed(0)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        trainer = U.UncertaintyGuidedProgressiveTrainer(3, 1, device=dev, uncertainty_alpha=1.0, log_interval=0)
        for s in (2, 3, 4):
            trainer.transfer_weights(s - 1, s)
    trainer.current_stage, trainer.current_model = STAGE, trainer.models[STAGE]
    trainer.setup_optimizer(STAGE)
    trainer.current_model.train()
    trainer.models[STAGE - 1].eval()

    x, y = synth_batch(B, 1234 + rank, device=dev)
    hx, hy = synth_batch(B, 1234 + rank, pinned=True)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---------------- device-resident timing ----------------
    for _ in range(args.warmup):
        trainer.train_step(x, y, STAGE)
    sync_all()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        trainer.train_step(x, y, STAGE)
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()

    # ---------------- end-to-end through the public API, host buffers ----------------
    class OneBatchLoader:
        def __init__(self, n): self.n = n
        def __len__(self): return self.n
        def __iter__(self):
            for _ in range(self.n):
                yield hx, hy
    with contextlib.redirect_stdout(io.StringIO()):
        trainer.train_epoch(OneBatchLoader(max(1, min(3, args.warmup))), STAGE)
        sync_all()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        res = trainer.train_epoch(OneBatchLoader(args.steps), STAGE)     # reads metrics back (D2H) per epoch
        e3.record()
        sync_all()
    ms_e2e = e2.elapsed_time(e3)

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()
    ips = world * B * args.steps / (ms / 1e3)
    ips_e2e = world * B * args.steps / (ms_e2e / 1e3)

    # ---------------- per-kernel roofline (instrumented step, rank 0) ----------------
    peaks = load_peaks()
    roof = None
    # kernels per step: counted on one eagerly launched step (the CUDA graph replays exactly these)
    ops.reset_launch_count()
    if rank != 0:
        trainer._eager_step(x, y, STAGE)
    if rank == 0:
        # same kernels as the captured graph, launched one by one on ONE stream (the step's two side
        # streams are switched off here so that each launch's CUDA events time that kernel alone)
        trainer.overlap_streams = False
        ops.start_profile()
        trainer._eager_step(x, y, STAGE)
        torch.cuda.synchronize()
        prof = ops.stop_profile()
        trainer.overlap_streams = True
    launches = ops.launch_count() * args.steps
    if rank == 0:
        conv = [(f, t_ms) for (name, f, t_ms) in prof if name == "conv_gemm"]
        wg = [(f, t_ms) for (name, f, t_ms) in prof if name == "conv_wgrad"]
        peak = peaks["bf16_tflops_sustained"]

        def tfs(pairs):
            t = sum(t_ms for _, t_ms in pairs) / 1e3
            return (sum(f for f, _ in pairs) / t / 1e12) if t > 0 else 0.0

        # dominant kernel = the heaviest single launch of the step: up4.conv.conv_op.0 forward
        # (128 -> 64 channels at 256x256: 2*B*256*256*64*9*128 FLOP, conv_halo_pair_kernel<64>)
        dom_f = 2.0 * B * RES * RES * 64 * 9 * 128
        dom = [(f, t_ms) for (f, t_ms) in conv if abs(f - dom_f) < 1.0][:1] or conv[:1]
        ach = tfs(dom)
        all_fd, all_wg = tfs(conv), tfs(wg)
        t_conv_ms = sum(t for _, t in conv) + sum(t for _, t in wg)
        roof = {"bound": "tensor",
                "kernel": "conv_halo_pair_kernel<64>: up4.conv.conv_op.0 forward, 128->64 ch @256x256 "
                          "(tcgen05 cta_group::2 implicit GEMM, halo-staged A)",
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                "peak_source": peaks["source"] + ", sustained bf16 (kernel timed inside a long step); burst peak "
                               f"{peaks['bf16_tflops']:.1f}",
                "flop_per_launch": dom[0][0] if dom else None, "ms_per_launch": dom[0][1] if dom else None,
                # dram__bytes_read.sum + dram__bytes_write.sum of this launch at batch 16, one
                # `ncu --set full` capture (profiles/r01_halo_pair64_up4_full.txt); algorithmic bytes 403 MB
                "traffic": 376.5e6 if B == 16 else None, "traffic_unit": "bytes/launch",
                "all_conv_fwd_dgrad": {"achieved": all_fd, "frac": all_fd / peak, "launches_per_step": len(conv),
                                       "ms_per_step_in_kernel": sum(t for _, t in conv)},
                "all_conv_wgrad": {"achieved": all_wg, "frac": all_wg / peak, "launches_per_step": len(wg),
                                   "ms_per_step_in_kernel": sum(t for _, t in wg)},
                "conv_share_of_step": t_conv_ms / (ms / args.steps) if ms > 0 else None,
                "step_conv_tflops": ips / world * (GF_PER_IMG_MODEL + GF_PER_IMG_UNC) / 1e3,
                "step_frac_of_peak": ips / world * (GF_PER_IMG_MODEL + GF_PER_IMG_UNC) / 1e3 / peak}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        c_ips, c_dt, cores = cpu_reference_step_rate(args.cpu_batch, 1, 1)
        cpu = {"value": c_ips, "unit": "images/s", "cores": cores, "kind": "port",
               "sample": f"1 timed + 1 warm-up stage-4 trainer step at batch {args.cpu_batch} "
                         "(oracle/ug_oracle.py, fp32, torch CPU)"}

    if rank == 0:
        line = {
            "metric": "stage-4 256x256 train images/sec", "value": ips, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "configs[1]: PGUNet4(3,1) stage-4 256x256 MoNuSeg-shaped trainer step "
                                   "(train fwd + PGUNet3 uncertainty fwd + fused loss + bwd + RMSprop)",
                       "stage": 4, "resolution": 256, "batch_per_gpu": B, "global_batch": B * world,
                       "parallelism": f"dp{world}", "l2": "per-step working set (~2 GB of activations at batch 16) "
                                                          "is far larger than the 126 MB L2; no explicit flush",
                       "gflop_per_image": GF_PER_IMG_MODEL + GF_PER_IMG_UNC},
            "e2e": {"value": ips_e2e, "unit": "images/s",
                    "h2d_bytes_per_step": int(hx.numel() * 4 + hy.numel() * 4),
                    "d2h_bytes_per_step": int((5 + 3 * B) * 8), "ms_per_step": ms_e2e / args.steps,
                    "final_loss": res[0]},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # Leave without tearing NCCL down: communicators referenced by captured CUDA graphs made
        # destroy_process_group() hang at exit on this stack.  Everything is flushed and all ranks
        # have passed the final barrier, so a hard exit is safe.
        sync_all()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch-per-gpu", type=int, default=16)
    ap.add_argument("--cpu-batch", type=int, default=2)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

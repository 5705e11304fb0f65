#!/usr/bin/env python
"""Headline benchmark: stage-4 256x256 train images/sec (BASELINE.json `metric`, configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch-per-gpu B] [--impl ours|reference]

One "step" = one full trainer step of `UncertaintyGuidedProgressiveTrainer` at stage 4 on one batch
of synthetic MoNuSeg-shaped input: PGUNet4 forward (train-mode BN), PGUNet3 eval forward for the
uncertainty map, fused uncertainty-weighted BCE loss + metrics, backward (dgrad + wgrad + BN/ReLU/
pool/upsample adjoints), gradient all-reduce when N > 1, RMSprop step.  Nothing is skipped or cached.

Prints ONE JSON line (rank 0):
  value      images/s with inputs resident in HBM (CUDA events around K graph replays, max over ranks)
  e2e        the same through the public trainer API (`train_epoch`) from pinned HOST fp32 buffers - H2D of
             every batch and D2H of the step metrics inside the timed region; `e2e_u8` = the loader-byte
             path (uint8 HWC images + uint8 masks, 4x fewer bytes).  Both run three epochs of K steps and
             report the median epoch (`epochs_ms` keeps all three)
  roofline   the dominant kernel CLASS (tcgen05 implicit-GEMM convolution, forward + dgrad launches):
             algorithmic FLOP / device time.  Device time per launch comes from a CUDA-graph replay of the
             step on ONE stream with event-record nodes between the kernel nodes (no host code between the
             two timestamps of a launch); sub-keys give wgrad, the best launch, the whole step and, under
             `hbm`, every HBM-bound kernel class with its algorithmic bytes and GB/s against the measured
             copy bandwidth.
  stock_gpu  the reference's step in STOCK PyTorch on the same GPU (cuDNN TF32 and bf16 autocast
             channels_last; tools/stock_torch_step.py) - the kernel to beat
  also       BASELINE.json configs[3] (Herlev 7-class stage-4 step, 256^2, batch 16/GPU) and configs[4]
             (segmentation at batch 64/GPU), a few replays each
  cpu_baseline / `--impl reference`: the CPU oracle port of the reference step (oracle/ug_oracle.py; the
             Python reference itself cannot travel to the GPU box) on a bounded sample, all host threads.
"""
from __future__ import annotations

import argparse
import contextlib
import io
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
WORKLOAD = ("configs[1]: PGUNet4(3,1) stage-4 256x256 MoNuSeg-shaped trainer step "
            "(train fwd + PGUNet3 uncertainty fwd + fused loss + bwd + RMSprop)")


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
            time.sleep(0.02)

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


def synth_batch_u8(B, seed):
    """Loader bytes: uint8 HWC images and uint8 0/1 masks in pinned host memory."""
    import torch
    g = torch.Generator().manual_seed(seed)
    x = torch.randint(0, 256, (B, RES, RES, 3), generator=g, dtype=torch.uint8)
    y = torch.randint(0, 2, (B, 1, RES, RES), generator=g, dtype=torch.uint8)
    return x.pin_memory(), y.pin_memory()


def host_threads() -> int:
    """All the host threads the CPU arm may use: torchrun exports OMP_NUM_THREADS=1 to its workers, which
    would handicap the CPU reference 10-30x - the arm sets its thread count itself."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_reference_step_rate(batch: int, steps: int, warmup: int):
    """Times the CPU oracle port of the reference trainer step (fwd + unc + loss + bwd + RMSprop)."""
    import torch
    torch.set_num_threads(host_threads())
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
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = args.cpu_batch
    steps = max(1, min(args.steps, 3))
    warm = min(args.warmup, 1)
    ips, dt, cores = cpu_reference_step_rate(batch, steps, warm)
    line = {
        "impl": "reference", "metric": "stage-4 256x256 train images/sec", "value": ips, "unit": "images/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "stage": 4, "resolution": 256, "batch_per_gpu": args.batch_per_gpu,
                   "global_batch": args.batch_per_gpu * args.gpus, "parallelism": "cpu",
                   "cpu_sample_batch": batch, "gflop_per_image": GF_PER_IMG_MODEL + GF_PER_IMG_UNC,
                   "bounded_sample": f"CPU arm: --steps/--warmup clamped to {steps}/{warm} steps of batch {batch} "
                                     "(a batch-16 step is ~30 s on 8 cores); images/s is per-image work, "
                                     "independent of the sample's batch size"},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} timed + {warm} warm-up stage-4 trainer step(s) at batch {batch} "
                                   "(oracle/ug_oracle.py, fp32, torch CPU, all host threads)"},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# kernel classes of the instrumented step (ops._timed names) -> what bounds them
TENSOR_CLASSES = {"conv_gemm": "conv forward + dgrad (tcgen05 implicit GEMM)", "conv_wgrad": "conv wgrad (tcgen05)"}


def instrumented_step(trainer, ops, x, y, replays: int = 5):
    """Device time of every instrumented launch: the step is captured on ONE stream into a CUDA graph with
    external event-record nodes around each launch and replayed `replays` times; per launch the median."""
    import torch
    trainer.overlap_streams = False
    try:
        trainer.optimizer.zero_grad(set_to_none=True)
        torch.cuda.synchronize()
        ops.start_profile(external=True)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            trainer._eager_step(x, y, STAGE)
        evs = list(ops.profile_events())
        samples = [[] for _ in evs]
        for _ in range(replays):
            g.replay()
            torch.cuda.synchronize()
            for i, (_, _, _, a, b) in enumerate(evs):
                samples[i].append(a.elapsed_time(b))
        ops.stop_profile()
        del g
        from ug_pg_unet_b200.optim import FusedRMSprop
        if isinstance(trainer.optimizer, FusedRMSprop):
            trainer.optimizer.note_replayed_steps(replays - 1)
        out = []
        for (name, fl, by, _, _), s in zip(evs, samples):
            s = sorted(s)
            out.append((name, fl, by, s[len(s) // 2]))
        return out, "cuda-graph replay, external event-record nodes, one stream, median of %d" % replays
    except Exception as e:      # fall back to events around eager launches (host-inflated, flagged)
        try:
            ops.stop_profile()
        except Exception:
            pass
        torch.cuda.synchronize()
        ops.start_profile(external=False)
        trainer._eager_step(x, y, STAGE)
        return ops.stop_profile(), f"eager launches (graph instrumentation failed: {type(e).__name__}: {e})"
    finally:
        trainer.overlap_streams = True


def timed_replays(step_fn, steps, warmup, sync_all, dev):
    import torch
    for _ in range(warmup):
        step_fn()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step_fn()
    e1.record()
    sync_all()
    return e0.elapsed_time(e1)


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

    def allmax(*vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    # ---------------- device-resident timing ----------------
    for _ in range(max(args.warmup, 4)):          # 3 eager warm-ups + the capture pass come first
        trainer.train_step(x, y, STAGE)
    sync_all()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms = timed_replays(lambda: trainer.train_step(x, y, STAGE), args.steps, 0, sync_all, dev)
    clocks = sampler.stop()

    # ---------------- end-to-end through the public API, host buffers ----------------
    class OneBatchLoader:
        def __init__(self, n, batch):
            self.n, self.batch = n, batch

        def __len__(self):
            return self.n

        def __iter__(self):
            for _ in range(self.n):
                yield self.batch

    def e2e_epochs(batch):
        with contextlib.redirect_stdout(io.StringIO()):
            trainer.train_epoch(OneBatchLoader(max(4, min(6, args.warmup + 1)), batch), STAGE)   # warm-up / capture
            sync_all()
            # K steps, three times: the host side of this region (pinned copies, allocator, the box's other
            # tenants) shows a stall of tens of ms in about one epoch out of ten, which is not the path's
            # throughput - the MEDIAN epoch is reported and all three are kept in the line (`epochs_ms`)
            runs = []
            for _ in range(3):
                e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e2.record()
                res = trainer.train_epoch(OneBatchLoader(args.steps, batch), STAGE)     # reads metrics back (D2H)
                e3.record()
                sync_all()
                runs.append((e2.elapsed_time(e3), res))
        runs.sort(key=lambda r: r[0])
        return runs[1][0], runs[1][1], [r[0] for r in runs]

    ms_e2e, res, all_e2e = e2e_epochs((hx, hy))
    ux, uy = synth_batch_u8(B, 4321 + rank)
    ms_u8, res_u8, all_u8 = e2e_epochs((ux, uy))

    ms, ms_e2e, ms_u8 = allmax(ms, ms_e2e, ms_u8)
    ips = world * B * args.steps / (ms / 1e3)
    ips_e2e = world * B * args.steps / (ms_e2e / 1e3)
    ips_u8 = world * B * args.steps / (ms_u8 / 1e3)

    # ---------------- per-kernel roofline (instrumented graph; every rank replays it, rank 0 reports) -------
    peaks = load_peaks()
    ops.reset_launch_count()
    prof, how = instrumented_step(trainer, ops, x, y)
    # launches per step: counted while the instrumented step was captured (the headline graph replays the same)
    launches = ops.launch_count() * args.steps
    roof = None
    if rank == 0 and args.dump_launches:
        with open(args.dump_launches, "w") as f:
            f.write(f"# per-launch device times of one batch-{B} step ({how})\n"
                    "# idx name us GFLOP MB TFLOP/s GB/s\n")
            for i, (n, fl, by, t) in enumerate(prof):
                f.write(f"{i:4d} {n:22s} {t * 1e3:9.1f} {fl / 1e9:9.2f} {by / 1e6:9.1f} "
                        f"{(fl / (t / 1e3) / 1e12 if t > 0 else 0):8.1f} {(by / (t / 1e3) / 1e9 if t > 0 else 0):8.0f}\n")
    if rank == 0:
        peak = peaks["bf16_tflops_sustained"]
        step_ms = ms / args.steps

        def agg(name):
            rows = [(f, by, t) for (n, f, by, t) in prof if n == name]
            t = sum(r[2] for r in rows)
            return {"launches_per_step": len(rows), "ms_per_step_in_kernel": t,
                    "flop": sum(r[0] for r in rows), "bytes": sum(r[1] for r in rows)}

        fd, wg = agg("conv_gemm"), agg("conv_wgrad")
        tf = lambda a: a["flop"] / (a["ms_per_step_in_kernel"] / 1e3) / 1e12 if a["ms_per_step_in_kernel"] > 0 else 0.0
        best = max(((f / (t / 1e3) / 1e12, f, t) for (n, f, by, t) in prof if n == "conv_gemm" and f > 5e10),
                   default=(0.0, 0.0, 0.0))
        hbm = {}
        for name in sorted({n for (n, f, by, t) in prof if n not in TENSOR_CLASSES}):
            a = agg(name)
            if a["ms_per_step_in_kernel"] <= 0:
                continue
            gbs = a["bytes"] / (a["ms_per_step_in_kernel"] / 1e3) / 1e9
            hbm[name] = {"launches_per_step": a["launches_per_step"], "us_per_step": a["ms_per_step_in_kernel"] * 1e3,
                         "algorithmic_mb": a["bytes"] / 1e6, "achieved_gbs": gbs, "frac": gbs / peaks["hbm_gbs"]}
        t_all = sum(t for (_, _, _, t) in prof)
        all_tf = (fd["flop"] + wg["flop"]) / ((fd["ms_per_step_in_kernel"] + wg["ms_per_step_in_kernel"]) / 1e3) / 1e12
        roof = {"bound": "tensor",
                "kernel": "conv forward + dgrad class: conv_halo(_pair)_kernel / conv_pair_kernel / conv_gemm_kernel "
                          "(tcgen05 implicit GEMM, TMA-staged operands, TMEM accumulators), all launches of one step",
                "achieved": tf(fd), "peak": peak, "unit": "TFLOP/s", "frac": tf(fd) / peak,
                "peak_source": peaks["source"] + ", sustained bf16 (kernels timed inside a long step); burst peak "
                               f"{peaks['bf16_tflops']:.1f}",
                "flop_per_step": fd["flop"], "ms_per_step_in_kernel": fd["ms_per_step_in_kernel"],
                "launches_per_step": fd["launches_per_step"],
                "timing": how,
                # dram__bytes_read.sum + dram__bytes_write.sum summed over the class's launches at batch 16, from the
                # ncu metrics pass profiles/r02_ncu_launches_by_kernel.txt (same bench command, 49 launches of one step)
                "traffic": CONV_CLASS_DRAM_BYTES if B == 16 else None, "traffic_unit": "bytes/step (class)",
                "algorithmic_bytes": fd["bytes"],
                "wgrad": {"achieved": tf(wg), "frac": tf(wg) / peak, **{k: wg[k] for k in ("launches_per_step", "ms_per_step_in_kernel")}},
                "all_conv": {"achieved": all_tf, "frac": all_tf / peak},
                "best_launch": {"achieved": best[0], "frac": best[0] / peak, "flop": best[1], "ms": best[2]},
                "kernel_time_sum_ms": t_all,
                "conv_share_of_kernel_time": (fd["ms_per_step_in_kernel"] + wg["ms_per_step_in_kernel"]) / t_all if t_all else None,
                "step_conv_tflops": ips / world * (GF_PER_IMG_MODEL + GF_PER_IMG_UNC) / 1e3,
                "step_frac_of_peak": ips / world * (GF_PER_IMG_MODEL + GF_PER_IMG_UNC) / 1e3 / peak,
                "hbm": {"peak_gbs": peaks["hbm_gbs"], "kernels": hbm}}

    # ---------------- other BASELINE.json configs: a few replays each ----------------
    also = {}
    if not args.no_also:
        ksteps = max(3, min(args.steps, 8))
        # configs[4]: segmentation at batch 64 per GPU
        x64, y64 = synth_batch(64, 99 + rank, device=dev)
        for _ in range(4):
            trainer.train_step(x64, y64, STAGE)
        t64 = timed_replays(lambda: trainer.train_step(x64, y64, STAGE), ksteps, 0, sync_all, dev)
        (t64,) = allmax(t64)
        also["seg_stage4_b64"] = {"workload": "configs[4]: stage-4 256x256 segmentation step, batch 64/GPU",
                                  "ms_per_step": t64 / ksteps, "images_per_s": world * 64 * ksteps / (t64 / 1e3),
                                  "steps": ksteps, "step_frac_of_peak":
                                      64 * ksteps / (t64 / 1e3) * (GF_PER_IMG_MODEL + GF_PER_IMG_UNC) / 1e3 /
                                      peaks["bf16_tflops_sustained"]}
        del x64, y64
        # configs[3]: Herlev-shaped 7-class classification, stage 4 at 256x256, batch 16 per GPU
        import ug_pg_unet_b200.herlev as H
        with contextlib.redirect_stdout(io.StringIO()):
            ht = H.HerlevTrainer({"device": dev, "epochs_per_stage": 1, "num_classes": 7, "uncertainty_alpha": 1.0,
                                  "log_interval": 0, "stage4_resolution": 256})
            ht.setup_optimizer_scheduler(4)
            ht.transfer_weights(3, 4)
        ht.models[4].train()
        ht.models[3].eval()
        gh = torch.Generator().manual_seed(77 + rank)
        hxd = torch.randn(B, 3, RES, RES, generator=gh).to(dev)
        hyd = torch.randint(0, 7, (B,), generator=gh).to(dev)
        for _ in range(4):
            ht.train_step(hxd, hyd, 4)
        th = timed_replays(lambda: ht.train_step(hxd, hyd, 4), ksteps, 0, sync_all, dev)
        (th,) = allmax(th)
        # encoder-only train step: 3 x 26.80 GF forward-equivalents - first-conv dgrad + stage-3 encoder forward
        also["herlev_stage4_b16"] = {"workload": "configs[3]: Herlev-shaped 7-class step, stage 4 at 256x256 "
                                                 "(encoder + GAP + MLP head, stage-3 entropy weighting, Adam), "
                                                 f"batch {B}/GPU", "ms_per_step": th / ksteps,
                                     "images_per_s": world * B * ksteps / (th / 1e3), "steps": ksteps}
        ht.close()
        del ht, hxd, hyd

    # ---------------- stock PyTorch on the same GPU: the kernel to beat ----------------
    stock = None
    if rank == 0 and world == 1 and not args.no_stock:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import stock_torch_step as S
        stock = {}
        for mode in ("tf32", "bf16"):
            try:
                s_ips, s_ms, _ = S.time_stock_step(mode, B, max(3, min(args.steps, 10)), 3, dev)
                stock[mode] = {"images_per_s": s_ips, "ms_per_step": s_ms, "ours_over_stock": ips / s_ips}
            except Exception as e:
                stock[mode] = {"error": f"{type(e).__name__}: {e}"}
            torch.cuda.empty_cache()
        stock["what"] = ("reference step in stock PyTorch (cuDNN/ATen) on this GPU, batch %d: tf32 = fp32 tensors with "
                         "TF32 convolutions (PyTorch default), bf16 = autocast + channels_last; cudnn.benchmark on" % B)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        c_ips, c_dt, cores = cpu_reference_step_rate(args.cpu_batch, 1, 1)
        cpu = {"value": c_ips, "unit": "images/s", "cores": cores, "kind": "port",
               "sample": f"1 timed + 1 warm-up stage-4 trainer step at batch {args.cpu_batch} "
                         "(oracle/ug_oracle.py, fp32, torch CPU, all host threads)"}

    if rank == 0:
        line = {
            "metric": "stage-4 256x256 train images/sec", "value": ips, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "stage": 4, "resolution": 256, "batch_per_gpu": B, "global_batch": B * world,
                       "parallelism": f"dp{world}", "l2": "per-step working set (~2 GB of activations at batch 16) "
                                                          "is far larger than the 126 MB L2; no explicit flush",
                       "gflop_per_image": GF_PER_IMG_MODEL + GF_PER_IMG_UNC},
            "e2e": {"value": ips_e2e, "unit": "images/s",
                    "h2d_bytes_per_step": int(hx.numel() * 4 + hy.numel() * 4),
                    "d2h_bytes_per_step": int((5 + 3 * B) * 8), "ms_per_step": ms_e2e / args.steps,
                    "final_loss": res[0], "input": "pinned fp32 NCHW images + fp32 masks (what the reference's loaders yield)",
                    "stat": f"median of 3 epochs of {args.steps} steps", "epochs_ms": all_e2e},
            "e2e_u8": {"value": ips_u8, "unit": "images/s", "h2d_bytes_per_step": int(ux.numel() + uy.numel()),
                       "d2h_bytes_per_step": int((5 + 3 * B) * 8), "ms_per_step": ms_u8 / args.steps,
                       "final_loss": res_u8[0], "input": "pinned uint8 HWC image bytes + uint8 masks (loader-byte path)",
                       "stat": f"median of 3 epochs of {args.steps} steps", "epochs_ms": all_u8},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roof,
            "stock_gpu": stock,
            "also": also,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    # ---------------- teardown: graphs first, then the communicator ----------------
    trainer.close()
    sync_all()
    if world > 1:
        sys.stdout.flush()
        sys.stderr.flush()
        done = threading.Event()

        def _destroy():
            try:
                dist.destroy_process_group()
            finally:
                done.set()

        th_ = threading.Thread(target=_destroy, daemon=True)
        th_.start()
        if not done.wait(timeout=30.0):
            # should not happen once the graphs are gone; never let a bench run hang the box
            sys.stderr.write("bench.py: destroy_process_group() did not return within 30 s; hard exit\n")
            sys.stderr.flush()
            os._exit(0)


# Sum of dram__bytes_read.sum + dram__bytes_write.sum over the forward + dgrad convolution launches of one
# batch-16 step (ncu metrics pass, profiles/r02_ncu_launches_by_kernel.txt: 3 372.8 MB against 4 744 MB algorithmic - consecutive
# kernels meet in the 126 MB L2).
CONV_CLASS_DRAM_BYTES = 3117.7e6


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch-per-gpu", type=int, default=16)
    ap.add_argument("--cpu-batch", type=int, default=2)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stock", action="store_true")
    ap.add_argument("--no-also", action="store_true")
    ap.add_argument("--dump-launches", default=None, help="write the per-launch table of the instrumented step here")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

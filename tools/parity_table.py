"""Measured parity of the sm_100a path against the reference-generated fixtures (tests/golden/*.pt),
per stage and batch size, beside north_star's tolerances (logits / gradients rel 2e-2, loss 1e-3,
masks bit-exact) and beside the bf16-STORAGE floor of the oracle (fp32 arithmetic, tensors rounded where
the CUDA path stores bf16).  Used by the `-m gpu` tests (the assertions quote these numbers) and as a
script on the GPU box:

    python tools/parity_table.py [--out profiles/r02_parity_table.md]

Nothing here reads /root/reference: inputs of the batch-16 fixtures are regenerated from their seeds and
checked against the checksums stored in the fixture.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")
TRAINER_FIXTURES = ["trainer_stage1_nc1", "trainer_stage1_nc2_config1", "trainer_stage2_nc1", "trainer_stage3_nc1",
                    "trainer_stage4_nc1", "trainer_stage3_nc1_b16", "trainer_stage4_nc1_b16"]
HERLEV_FIXTURES = ["herlev_stage1_k7", "herlev_stage2_k7", "herlev_stage3_k7_b16", "herlev_stage4_k7_b16_256",
                   "herlev_stage4_k7_b16_224"]
ZERO_GRAD_SUFFIX = ("conv_op.0.bias", "conv_op.3.bias")   # conv bias in front of train-mode BN: analytically 0


def rel(a, b):
    a, b = a.detach().double().flatten().cpu(), b.detach().double().flatten().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def cos(a, b):
    a, b = a.detach().double().flatten().cpu(), b.detach().double().flatten().cpu()
    return (a @ b / (a.norm() * b.norm() + 1e-300)).item()


def against_sample(t: torch.Tensor, s: dict):
    """(rel error, cosine, norm ratio) of the full tensor `t` against a fixture entry written by
    tools/make_goldens.py::sample (whole tensor for small ones, strided 4096-sample + full norm otherwise)."""
    f = t.detach().double().flatten().cpu()
    if "full" in s:
        ref, v, nr = s["full"].double(), f, None
    else:
        ref, v = s["vals"].double(), f[s["idx"]]
        nr = (f.norm() / (s["norm"] + 1e-300)).item()
    r = ((v - ref).norm() / (ref.norm() + 1e-30)).item()
    c = (v @ ref / (v.norm() * ref.norm() + 1e-300)).item()
    return r, c, (nr if nr is not None else (v.norm() / (ref.norm() + 1e-300)).item())


def regenerate_inputs(fx):
    """x, y of a fixture: stored (small fixtures) or regenerated from the seed (batch-16 fixtures)."""
    import ug_oracle as O
    if "x" in fx:
        return fx["x"], (fx["y"].float() if fx["y"].dtype == torch.uint8 else fx["y"])
    if "K" in fx:      # Herlev
        g = torch.Generator().manual_seed(2000 + fx["seed"])
        x = torch.randn(fx["B"], 3, fx["R"], fx["R"], generator=g)
        y = torch.randint(0, fx["K"], (fx["B"],), generator=g)
        assert torch.equal(y, fx["y"])
    else:
        R = O.STAGE_RESOLUTION[fx["stage"]]
        g = torch.Generator().manual_seed(1000 + fx["seed"])
        x = torch.randn(fx["B"], 3, R, R, generator=g)
        y = torch.randint(0, 2, (fx["B"], fx["nc"], R, R), generator=g).float()
        assert int(y.sum().item()) == fx["y_checksum"]
    assert abs(x.double().sum().item() - fx["x_checksum"]) < 1e-6, "seeded inputs differ from the fixture's"
    return x, y


def _grad_rows(named_grads, fx, floors_rel, floors_cos):
    rows = {}
    for k, g in named_grads:
        if k.endswith(ZERO_GRAD_SUFFIX) or k not in fx["grads"] or g is None:
            continue
        r, c, nr = against_sample(g, fx["grads"][k])
        rows[k] = {"rel": r, "cos": c, "norm_ratio": nr, "floor_rel": floors_rel[k], "floor_cos": floors_cos[k]}
    return rows


def _summary(rows):
    rels = [v["rel"] for v in rows.values()]
    worst = max(rows, key=lambda k: rows[k]["rel"])
    return {"n": len(rows), "median_rel": statistics.median(rels), "worst_rel": rows[worst]["rel"], "worst_key": worst,
            "worst_floor_rel": rows[worst]["floor_rel"], "min_cos": min(v["cos"] for v in rows.values()),
            "median_floor_rel": statistics.median(v["floor_rel"] for v in rows.values()),
            "n_le_2e-2": sum(r <= 2e-2 for r in rels),
            "max_ratio_to_floor": max(v["rel"] / max(v["floor_rel"], 1e-4) for v in rows.values())}


def measure_trainer(name: str, dev) -> dict:
    """One `uncertainty_guided_forward_pass` + backward of the CUDA path on the fixture's weights / inputs."""
    import ug_oracle as O
    import ug_pg_unet_b200 as U

    fx = torch.load(os.path.join(GOLDEN, name + ".pt"))
    stage, nc = fx["stage"], fx["nc"]
    x, y = regenerate_inputs(fx)
    sd = O.seeded_state_dict(stage, 3, nc, seed=stage)
    sd_prev = O.seeded_state_dict(stage - 1, 3, nc, seed=stage - 1) if stage > 1 else None
    if "emu" in fx:      # batch-16 fixtures carry the floors (the oracle needs ~1 min per pass there)
        emu = fx["emu"]
        fl_logits, fl_final, fl_base = emu["logits_rel"], emu["final_loss"], emu["base_loss"]
        fl_rel, fl_cos, fl_mask = emu["grads_rel"], emu["grads_cos"], emu["mask_mismatch"]
    else:
        ref = O.train_step(sd, x, y, stage, sd_prev)
        with O.bf16_storage(True):
            e = O.train_step(sd, x, y, stage, sd_prev)
        fl_logits, fl_final, fl_base = rel(e["logits"], fx["logits"]), e["final_loss"], e["base_loss"]
        fl_rel = {k: rel(e["grads"][k], ref["grads"][k]) for k in ref["grads"]}
        fl_cos = {k: cos(e["grads"][k], ref["grads"][k]) for k in ref["grads"]}
        fl_mask = ((e["logits"] > 0) != (fx["logits"] > 0)).float().mean().item()

    model = U.STAGE_MODELS[stage](3, nc).to(dev)
    model.load_state_dict(sd)
    model.train()
    prev = None
    if stage > 1:
        prev = U.STAGE_MODELS[stage - 1](3, nc).to(dev)
        prev.load_state_dict(sd_prev)
    ugl = U.UncertaintyGuidedLoss(dev)
    crit = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor([5.0], device=dev), reduction="none")
    xd, yd = x.to(dev), y.to(dev)
    logits = model(xd)
    A = ugl.generate_uncertainty_map(xd, prev, 32 << (stage - 2), 32 << (stage - 1)) if stage > 1 else None
    final, base = ugl.apply_uncertainty_weighted_loss(crit, logits, yd, A, 1.0)
    final.backward()
    torch.cuda.synchronize()

    zr = fx["logits"]
    zc = logits.detach().cpu()
    # A predicted-mask pixel can only be required to match where the reference logit is not within the
    # logit noise of the threshold.  The excluded band is |z_ref| <= max(5 % of max|z_ref|, 6 sigma) with
    # sigma = the rms logit error of the oracle's bf16-storage emulation (NOT of the kernels under test).
    rms_floor = fl_logits * zr.double().pow(2).mean().sqrt().item()
    band = max(0.05 * zr.abs().max().item(), 6.0 * rms_floor)
    safe = zr.abs() > band
    mism = (zc > 0) != (zr > 0)
    out = {
        "name": name, "stage": stage, "B": fx["B"], "nc": nc,
        "logits_rel": rel(zc, zr), "logits_floor": fl_logits,
        "final_loss": final.item(), "final_loss_ref": fx["final_loss"],
        "final_loss_rel": abs(final.item() - fx["final_loss"]) / abs(fx["final_loss"]),
        "final_loss_floor_rel": abs(fl_final - fx["final_loss"]) / abs(fx["final_loss"]),
        "base_loss": base, "base_loss_ref": fx["base_loss"],
        "base_loss_rel": abs(base - fx["base_loss"]) / abs(fx["base_loss"]),
        "base_loss_floor_rel": abs(fl_base - fx["base_loss"]) / abs(fx["base_loss"]),
        "mask_mismatch": mism.float().mean().item(), "mask_mismatch_floor": fl_mask,
        "mask_excluded_frac": 1.0 - safe.float().mean().item(), "mask_band_over_max": band / zr.abs().max().item(),
        "mask_mismatch_safe": int(mism[safe].sum().item()),
        # a flipped pixel is always one whose reference logit is small: the largest |z_ref| among them
        "mask_flip_max_abs_zref_over_max": (zr.abs()[mism].max() / zr.abs().max()).item() if mism.any() else 0.0,
        "grads": _grad_rows(((k, p.grad) for k, p in model.named_parameters()), fx, fl_rel, fl_cos),
    }
    if A is not None and "uncertainty" in fx:
        out["uncertainty_max_abs"] = (A.cpu() - fx["uncertainty"].float()).abs().max().item()
    out["bn_worst_rel"] = max(rel(model.state_dict()[k], v) for k, v in fx["bn"].items()
                              if not k.endswith("num_batches_tracked"))
    out["grad_summary"] = _summary(out["grads"])
    out["_objects"] = (model, logits, A, final, base, ugl, fx)
    return out


def measure_herlev(name: str, dev) -> dict:
    import ug_oracle as O
    import ug_pg_unet_b200.herlev as H

    fx = torch.load(os.path.join(GOLDEN, name + ".pt"))
    stage, K = fx["stage"], fx["K"]
    x, y = regenerate_inputs(fx)
    sd = O.seeded_herlev_state_dict(stage, K, seed=stage)
    sd_prev = O.seeded_herlev_state_dict(stage - 1, K, seed=stage - 1) if stage > 1 else None
    if "emu" in fx:
        emu = fx["emu"]
        fl_logits, fl_final, fl_base = emu["logits_rel"], emu["final_loss"], emu["base_loss"]
        fl_rel, fl_cos = emu["grads_rel"], emu["grads_cos"]
    else:
        ref = O.herlev_train_step(sd, x, y, stage, sd_prev)
        with O.bf16_storage(True):
            e = O.herlev_train_step(sd, x, y, stage, sd_prev)
        fl_logits, fl_final, fl_base = rel(e["logits"], fx["logits"]), e["final_loss"], e["base_loss"]
        fl_rel = {k: rel(e["grads"][k], ref["grads"][k]) for k in ref["grads"] if ref["grads"][k] is not None}
        fl_cos = {k: cos(e["grads"][k], ref["grads"][k]) for k in ref["grads"] if ref["grads"][k] is not None}
    cfg = {"device": dev, "epochs_per_stage": 1, "num_classes": K, "uncertainty_alpha": 1.0, "log_interval": 0}
    if stage == 4:
        cfg["stage4_resolution"] = fx.get("R", 256)
    with contextlib.redirect_stdout(io.StringIO()):
        tr = H.HerlevTrainer(cfg)
    tr.models[stage].load_state_dict(sd)
    if stage > 1:
        tr.models[stage - 1].load_state_dict(sd_prev)
        tr.models[stage - 1].eval()
    model = tr.models[stage]
    model.train()
    for m in model.classifier:
        if isinstance(m, torch.nn.Dropout):
            m.eval()
    xd, yd = x.to(dev), y.to(dev)
    final, metrics = tr.uncertainty_guided_forward_pass(xd, yd, stage)
    final.backward()
    torch.cuda.synchronize()
    zc = metrics["output"].detach().float().cpu()
    out = {
        "name": name, "stage": stage, "B": fx["B"], "K": K, "R": x.shape[-1],
        "logits_rel": rel(zc, fx["logits"]), "logits_floor": fl_logits,
        "final_loss": metrics["final_loss"], "final_loss_ref": fx["final_loss"],
        "final_loss_rel": abs(metrics["final_loss"] - fx["final_loss"]) / abs(fx["final_loss"]),
        "final_loss_floor_rel": abs(fl_final - fx["final_loss"]) / abs(fx["final_loss"]),
        "base_loss_rel": abs(metrics["base_loss"] - fx["base_loss"]) / abs(fx["base_loss"]),
        "base_loss_floor_rel": abs(fl_base - fx["base_loss"]) / abs(fx["base_loss"]),
        "argmax_equal": bool(torch.equal(zc.argmax(1), fx["logits"].argmax(1))),
        # a sample's class may only differ where the reference's top-2 margin is inside the logit noise:
        # 5 sigma of the bf16-storage emulation's rms logit error (of a difference of two logits)
        "argmax_flips": int((zc.argmax(1) != fx["logits"].argmax(1)).sum()),
        "argmax_margin_band": 5.0 * 2 ** 0.5 * fl_logits * fx["logits"].double().pow(2).mean().sqrt().item(),
        "argmax_flip_max_margin": max([0.0] + [
            (t[0] - t[1]).item() for t, a, b in zip(fx["logits"].topk(2, 1).values, zc.argmax(1), fx["logits"].argmax(1))
            if a != b]),
        "correct": int((zc.argmax(1) == y).sum()), "correct_ref": fx["correct"],
        "unc_mean_abs": abs(metrics["uncertainty_weight_mean"] - fx["uncertainty_weight_mean"]),
        "unc_std_abs": abs(metrics["uncertainty_weight_std"] - fx["uncertainty_weight_std"]),
        "no_grad_ok": all(p.grad is None for k, p in model.named_parameters() if k in fx["no_grad_keys"]),
        "grads": _grad_rows(((k, p.grad) for k, p in model.named_parameters() if k not in fx["no_grad_keys"]),
                            fx, fl_rel, fl_cos),
    }
    out["bn_worst_rel"] = max(rel(model.state_dict()[k], v) for k, v in fx["bn"].items()
                              if not k.endswith("num_batches_tracked"))
    out["grad_summary"] = _summary(out["grads"])
    out["_objects"] = (tr, model, metrics, fx, xd, yd)
    return out


def measure_stock_gpu(name: str, dev, mode: str) -> dict:
    """The reference's own step in STOCK PyTorch on the GPU (tools/stock_torch_step.StockPGUNet: nn.Conv2d /
    nn.BatchNorm2d / F.max_pool2d / F.interpolate, cuDNN kernels) on the fixture's weights and inputs, in
    `mode` = "fp32" (TF32 off), "tf32" (PyTorch's default on this GPU) or "bf16" (autocast + channels_last):
    how far the reference's OWN GPU paths land from the fp32 CPU fixture."""
    import torch.nn.functional as F
    import ug_oracle as O
    import stock_torch_step as S

    fx = torch.load(os.path.join(GOLDEN, name + ".pt"))
    stage, nc = fx["stage"], fx["nc"]
    x, y = regenerate_inputs(fx)
    sd = O.seeded_state_dict(stage, 3, nc, seed=stage)
    sd_prev = O.seeded_state_dict(stage - 1, 3, nc, seed=stage - 1) if stage > 1 else None

    def build(st, state):
        m = S.StockPGUNet(st, nc)
        keymap = {"inc.conv.": "inc."}
        for i, dn in enumerate(O.down_names(st)):
            keymap[f"{dn}.mpconv.1."] = f"downs.{i}."
        for j in range(st):
            keymap[f"up{j + 1}.conv."] = f"ups.{j}."
        heads = O.head_names(st)
        for j, hn in enumerate(heads):
            keymap[f"{hn}.conv."] = f"heads.{j}."
        new = {}
        for k, v in state.items():
            for a, b in keymap.items():
                if k.startswith(a):
                    new[b + k[len(a):]] = v
                    break
            else:
                raise KeyError(k)
        m.load_state_dict(new, strict=True)
        inv = {b: a for a, b in keymap.items()}
        return m.to(dev), inv

    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = (mode == "tf32")
    try:
        fmt = torch.channels_last if mode == "bf16" else torch.contiguous_format
        model, inv = build(stage, sd)
        model = model.to(memory_format=fmt).train()
        xd = x.to(dev).contiguous(memory_format=fmt)
        yd = y.to(dev)
        cast = torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16"))
        with cast:
            out = model(xd)
            A = None
            if stage > 1:
                prev, _ = build(stage - 1, sd_prev)
                prev = prev.to(memory_format=fmt).eval()
                with torch.no_grad():
                    R = x.shape[-1]
                    xp = F.interpolate(xd, size=(R // 2, R // 2), mode="bilinear", align_corners=True)
                    pr = torch.sigmoid(prev(xp).float())
                    pr = F.interpolate(pr, size=(R, R), mode="bilinear", align_corners=True)
                    A = 1.0 - 2.0 * torch.abs(pr - 0.5)
        crit = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor([5.0], device=dev), reduction="none")
        pl = crit(out.float(), yd)
        loss = torch.mean(pl if A is None else pl * (1.0 + A).detach())
        loss.backward()
        torch.cuda.synchronize()
        rows = {}
        for k, p in model.named_parameters():
            pref = max((b for b in inv if k.startswith(b)), key=len)
            rk = inv[pref] + k[len(pref):]
            if rk.endswith(ZERO_GRAD_SUFFIX) or rk not in fx["grads"]:
                continue
            r, c, nr = against_sample(p.grad.float(), fx["grads"][rk])
            rows[rk] = {"rel": r, "cos": c, "norm_ratio": nr, "floor_rel": 0.0, "floor_cos": 1.0}
        zc = out.detach().float().cpu()
        return {"name": name, "mode": mode, "stage": stage, "B": fx["B"], "logits_rel": rel(zc, fx["logits"]),
                "final_loss_rel": abs(loss.item() - fx["final_loss"]) / abs(fx["final_loss"]),
                "mask_mismatch": ((zc > 0) != (fx["logits"] > 0)).float().mean().item(),
                "median_grad_rel": statistics.median(v["rel"] for v in rows.values()),
                "worst_grad_rel": max(v["rel"] for v in rows.values()),
                "n_le_2e-2": sum(v["rel"] <= 2e-2 for v in rows.values()), "n": len(rows),
                "n_le_1e-3": sum(v["rel"] <= 1e-3 for v in rows.values())}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def stock_table(rows) -> str:
    lines = ["| fixture | stock PyTorch mode | logits rel | final loss rel | mask mismatch | grad tensors <= 2e-2 | "
             "grad tensors <= 1e-3 | median grad rel | worst grad rel |", "|---|---|---|---|---|---|---|---|---|"]
    for r in rows:
        lines.append(f"| {r['name']} | {r['mode']} | {_fmt(r['logits_rel'])} | {_fmt(r['final_loss_rel'])} | "
                     f"{_fmt(r['mask_mismatch'])} | {r['n_le_2e-2']}/{r['n']} | {r['n_le_1e-3']}/{r['n']} | "
                     f"{_fmt(r['median_grad_rel'])} | {_fmt(r['worst_grad_rel'])} |")
    return "\n".join(lines)


def _fmt(v):
    return f"{v:.2e}"


def table(results) -> str:
    lines = ["| fixture | stage | B | logits rel (floor) | final loss rel (floor) | base loss rel (floor) | "
             "mask / argmax mismatch (floor; fraction of pixels in the excluded near-threshold band; flips outside it) | "
             "grad tensors <= 2e-2 | median grad rel (floor) | worst grad rel (floor) [tensor] | min cos | "
             "max rel / floor |",
             "|---|---|---|---|---|---|---|---|---|---|---|---|"]
    for r in results:
        gs = r["grad_summary"]
        if "mask_mismatch" in r:
            mask = (f"{r['mask_mismatch']:.2e} ({r['mask_mismatch_floor']:.2e}; {r['mask_excluded_frac']:.3f}; "
                    f"{r['mask_mismatch_safe']})")
        else:
            mask = "argmax equal" if r["argmax_equal"] else (
                f"{r['argmax_flips']} argmax flip(s), ref top-2 margin {r['argmax_flip_max_margin']:.1e} "
                f"(noise band {r['argmax_margin_band']:.1e})")
        lines.append(
            f"| {r['name']} | {r['stage']} | {r['B']} | {_fmt(r['logits_rel'])} ({_fmt(r['logits_floor'])}) | "
            f"{_fmt(r['final_loss_rel'])} ({_fmt(r['final_loss_floor_rel'])}) | "
            f"{_fmt(r['base_loss_rel'])} ({_fmt(r['base_loss_floor_rel'])}) | {mask} | "
            f"{gs['n_le_2e-2']}/{gs['n']} | {_fmt(gs['median_rel'])} ({_fmt(gs['median_floor_rel'])}) | "
            f"{_fmt(gs['worst_rel'])} ({_fmt(gs['worst_floor_rel'])}) [{gs['worst_key']}] | {gs['min_cos']:.4f} | "
            f"{gs['max_ratio_to_floor']:.2f} |")
    return "\n".join(lines)


def main():
    out_path = None
    if "--out" in sys.argv:
        out_path = sys.argv[sys.argv.index("--out") + 1]
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    results = []
    for n in TRAINER_FIXTURES:
        r = measure_trainer(n, dev)
        r.pop("_objects")
        results.append(r)
        print(n, "logits", _fmt(r["logits_rel"]), "floor", _fmt(r["logits_floor"]), r["grad_summary"], flush=True)
    for n in HERLEV_FIXTURES:
        r = measure_herlev(n, dev)
        r.pop("_objects")
        results.append(r)
        print(n, "logits", _fmt(r["logits_rel"]), "floor", _fmt(r["logits_floor"]), r["grad_summary"], flush=True)
    txt = table(results)
    print(txt)
    stock = []
    if "--no-stock" not in sys.argv:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        for n in ("trainer_stage1_nc1", "trainer_stage2_nc1", "trainer_stage3_nc1_b16", "trainer_stage4_nc1_b16"):
            for mode in ("fp32", "tf32", "bf16"):
                stock.append(measure_stock_gpu(n, dev, mode))
                print(stock[-1], flush=True)
        txt += ("\n\nThe reference's step in STOCK PyTorch on the same GPU (cuDNN; the reference's own GPU paths) "
                "against the same fp32 CPU fixtures:\n\n" + stock_table(stock))
    opts = os.environ.get("UGP_OPTIONS", "")
    if out_path:
        with open(out_path, "w") as f:
            f.write(f"UGP_OPTIONS='{opts}'\n\n" + txt + "\n\nper-tensor gradients (rel, floor rel, cos):\n")
            for r in results:
                f.write(f"\n{r['name']}\n")
                for k, v in r["grads"].items():
                    f.write(f"  {k:44s} rel {v['rel']:.3e}  floor {v['floor_rel']:.3e}  cos {v['cos']:.5f}  "
                            f"floor cos {v['floor_cos']:.5f}  |g|/|g_ref| {v['norm_ratio']:.4f}\n")
        with open(os.path.splitext(out_path)[0] + ".json", "w") as f:
            json.dump({"cuda_path": results, "stock_pytorch_gpu": stock}, f, indent=1)


if __name__ == "__main__":
    main()

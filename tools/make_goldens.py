"""Generates tests/golden/*.pt by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only (the reference does not travel to the GPU box):

    python tools/make_goldens.py

Weights come from oracle.ug_oracle.seeded_state_dict (a pure function of key/shape/seed), loaded
into the reference's own nn.Modules with load_state_dict(strict=True) — which also proves the
key/shape table of the oracle matches the reference.  Each fixture stores the inputs and the
reference's outputs (logits, losses, metrics, uncertainty map, BN buffers after the step, and
gradients: full tensors for small parameters, a strided sample + norm for the big conv weights) so
the files stay small.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, "/root/reference")

import torch
import torch.nn.functional as F

import ug_oracle as O

with contextlib.redirect_stdout(io.StringIO()):
    import UG_unet as REF
    import uncertainty_guided_trainer as REFT

OUT = os.path.join(ROOT, "tests", "golden")
SAMPLE = 4096


def sample(t: torch.Tensor):
    f = t.flatten()
    if f.numel() <= SAMPLE:
        return {"full": f.clone()}
    idx = torch.linspace(0, f.numel() - 1, SAMPLE).long()
    return {"idx": idx, "vals": f[idx].clone(), "norm": f.double().norm().item(), "numel": f.numel()}


def trainer_case(stage: int, B: int, nc: int, seed: int):
    torch.manual_seed(seed)
    R = O.STAGE_RESOLUTION[stage]
    g = torch.Generator().manual_seed(1000 + seed)
    x = torch.randn(B, 3, R, R, generator=g)
    y = torch.randint(0, 2, (B, nc, R, R), generator=g).float()
    with contextlib.redirect_stdout(io.StringIO()):
        tr = REFT.UncertaintyGuidedProgressiveTrainer(3, nc, device="cpu", uncertainty_alpha=1.0)
    tr.models[stage].load_state_dict(O.seeded_state_dict(stage, 3, nc, seed=stage), strict=True)
    if stage > 1:
        tr.models[stage - 1].load_state_dict(O.seeded_state_dict(stage - 1, 3, nc, seed=stage - 1), strict=True)
    tr.current_stage = stage
    tr.current_model = tr.models[stage]
    tr.setup_optimizer(stage)
    tr.current_model.train()
    if stage > 1:
        tr.models[stage - 1].eval()
    tr.optimizer.zero_grad()
    final_loss, metrics = tr.uncertainty_guided_forward_pass(x, y, stage)
    final_loss.backward()
    fx = {
        "stage": stage, "B": B, "nc": nc, "x": x, "y": y.to(torch.uint8),
        "logits": metrics["output"].detach().clone(),
        "final_loss": metrics["final_loss"], "base_loss": metrics["base_loss"],
        "uncertainty_weight_mean": metrics["uncertainty_weight_mean"],
        "uncertainty_weight_std": metrics["uncertainty_weight_std"],
        "grads": {k: sample(p.grad) for k, p in tr.current_model.named_parameters()},
        "bn": {k: v.clone() for k, v in tr.current_model.state_dict().items()
               if k.endswith(("running_mean", "running_var", "num_batches_tracked"))},
    }
    if stage > 1:
        A = tr.uncertainty_loss.generate_uncertainty_map(x, tr.models[stage - 1], O.STAGE_RESOLUTION[stage - 1], R)
        fx["uncertainty"] = A.to(torch.float16)
    if nc == 1:
        pred = tr.get_predictions(metrics["output"].detach())
        fx["dice"] = tr.dice_coefficient(pred.float(), y.squeeze(1)).item()
        fx["accuracy"] = float(tr.calculate_accuracy(pred, y.squeeze(1).long()))
    # one optimiser step, then a sample of the updated weights
    tr.optimizer.step()
    fx["after_step"] = {k: sample(p.detach()) for k, p in tr.current_model.named_parameters()}
    # eval-mode logits with the post-step weights / buffers
    tr.current_model.eval()
    with torch.no_grad():
        fx["eval_logits"] = tr.current_model(x).clone()
    return fx


def seeded_batch(stage: int, B: int, nc: int, seed: int):
    """Inputs of the batch-16 fixtures: regenerated from the seed wherever they are needed (CPU generator,
    same torch build in the build container and on the GPU box) instead of being stored; the fixture
    keeps a checksum."""
    R = O.STAGE_RESOLUTION[stage]
    g = torch.Generator().manual_seed(1000 + seed)
    x = torch.randn(B, 3, R, R, generator=g)
    y = torch.randint(0, 2, (B, nc, R, R), generator=g).float()
    return x, y


def _rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return (a @ b / (a.norm() * b.norm() + 1e-300)).item()


def trainer_case_b16(stage: int, B: int, nc: int, seed: int):
    """BASELINE.json configs[1] at its real batch size: one trainer step of the UNMODIFIED reference
    (uncertainty_guided_trainer.py:143-184, 203-218) at batch 16.  Inputs are not stored (seeded_batch);
    gradients are stored as a 4096-element strided sample + the full norm per tensor.  Beside the
    reference's outputs the fixture records, per tensor, how far the oracle's bf16-STORAGE emulation
    (fp32 arithmetic, tensors rounded where the sm_100a path stores bf16) lands from the reference -
    the measured floor the GPU tests quote where north_star's 2e-2 cannot be met by a bf16 path."""
    torch.manual_seed(seed)
    x, y = seeded_batch(stage, B, nc, seed)
    with contextlib.redirect_stdout(io.StringIO()):
        tr = REFT.UncertaintyGuidedProgressiveTrainer(3, nc, device="cpu", uncertainty_alpha=1.0)
    sd = O.seeded_state_dict(stage, 3, nc, seed=stage)
    sd_prev = O.seeded_state_dict(stage - 1, 3, nc, seed=stage - 1) if stage > 1 else None
    tr.models[stage].load_state_dict(sd, strict=True)
    if stage > 1:
        tr.models[stage - 1].load_state_dict(sd_prev, strict=True)
    tr.current_stage = stage
    tr.current_model = tr.models[stage]
    tr.setup_optimizer(stage)
    tr.current_model.train()
    if stage > 1:
        tr.models[stage - 1].eval()
    tr.optimizer.zero_grad()
    final_loss, metrics = tr.uncertainty_guided_forward_pass(x, y, stage)
    final_loss.backward()
    logits = metrics["output"].detach().clone()
    grads = {k: p.grad.detach().clone() for k, p in tr.current_model.named_parameters()}
    fx = {
        "stage": stage, "B": B, "nc": nc, "seed": seed,
        "x_checksum": x.double().sum().item(), "y_checksum": int(y.sum().item()),
        "logits": logits,
        "final_loss": metrics["final_loss"], "base_loss": metrics["base_loss"],
        "uncertainty_weight_mean": metrics["uncertainty_weight_mean"],
        "uncertainty_weight_std": metrics["uncertainty_weight_std"],
        "grads": {k: sample(g) for k, g in grads.items()},
        "bn": {k: v.clone() for k, v in tr.current_model.state_dict().items()
               if k.endswith(("running_mean", "running_var", "num_batches_tracked"))},
    }
    if stage > 1:
        A = tr.uncertainty_loss.generate_uncertainty_map(x, tr.models[stage - 1], O.STAGE_RESOLUTION[stage - 1], O.STAGE_RESOLUTION[stage])
        fx["uncertainty"] = A.to(torch.float16)
    pred = tr.get_predictions(logits)
    fx["dice"] = tr.dice_coefficient(pred.float(), y.squeeze(1)).item()
    fx["accuracy"] = float(tr.calculate_accuracy(pred, y.squeeze(1).long()))
    # the oracle restatement must reproduce the reference here too (checked again in tests/test_oracle_golden.py
    # at the small sizes); and its bf16-storage emulation gives the per-tensor floor
    ref = O.train_step(sd, x, y, stage, sd_prev)
    fx["oracle_vs_reference"] = {"logits": _rel(ref["logits"], logits),
                                 "worst_grad": max(_rel(ref["grads"][k], g) for k, g in grads.items()
                                                   if not k.endswith(("conv_op.0.bias", "conv_op.3.bias")))}
    with O.bf16_storage(True):
        emu = O.train_step(sd, x, y, stage, sd_prev)
    fx["emu"] = {
        "logits_rel": _rel(emu["logits"], logits),
        "final_loss": emu["final_loss"], "base_loss": emu["base_loss"],
        "mask_mismatch": ((emu["logits"] > 0) != (logits > 0)).float().mean().item(),
        "grads_rel": {k: _rel(emu["grads"][k], g) for k, g in grads.items()},
        "grads_cos": {k: _cos(emu["grads"][k], g) for k, g in grads.items()},
    }
    return fx


def herlev_case_b16(stage: int, B: int, K: int, seed: int, R: int):
    """Herlev stage 3 / 4 steps of the UNMODIFIED reference at batch 16 (Herlev/train_herlev.py:216-296),
    at R x R inputs (256 = configs[3]; 224 = the reference's own Herlev stage-4 size); same recipe as
    herlev_case, inputs regenerated from the seed, bf16-storage floors recorded like trainer_case_b16."""
    sys.path.insert(0, "/root/reference/Herlev")
    with contextlib.redirect_stdout(io.StringIO()):
        import train_herlev as REFH
    g = torch.Generator().manual_seed(2000 + seed)
    x = torch.randn(B, 3, R, R, generator=g)
    y = torch.randint(0, K, (B,), generator=g)
    cfg = {"device": "cpu", "epochs_per_stage": 1, "num_classes": K, "uncertainty_alpha": 1.0}
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        tr = REFH.HerlevTrainer(cfg)
    if R != O.STAGE_RESOLUTION[stage]:
        # the reference trainer reads its resolutions from stage_configs (train_herlev.py:137-143)
        tr.stage_configs[stage]["resolution"] = R
    sd = O.seeded_herlev_state_dict(stage, K, seed=stage)
    sd_prev = O.seeded_herlev_state_dict(stage - 1, K, seed=stage - 1)
    tr.models[stage].load_state_dict(sd, strict=True)
    tr.models[stage - 1].load_state_dict(sd_prev, strict=True)
    tr.models[stage - 1].eval()
    model = tr.models[stage]
    model.train()
    for m in model.classifier:
        if isinstance(m, torch.nn.Dropout):
            m.eval()
    final_loss, metrics = tr.uncertainty_guided_forward_pass(x, y, stage)
    final_loss.backward()
    logits = metrics["output"].detach().clone()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    fx = {
        "stage": stage, "B": B, "K": K, "R": R, "seed": seed, "x_checksum": x.double().sum().item(), "y": y,
        "prev_resolution": tr.stage_configs[stage - 1]["resolution"],
        "logits": logits, "final_loss": metrics["final_loss"], "base_loss": metrics["base_loss"],
        "uncertainty_weight_mean": metrics["uncertainty_weight_mean"],
        "uncertainty_weight_std": metrics["uncertainty_weight_std"],
        "correct": int((logits.argmax(1) == y).sum()),
        "grads": {k: sample(g) for k, g in grads.items()},
        "no_grad_keys": [k for k, p in model.named_parameters() if p.grad is None],
        "bn": {k: v.clone() for k, v in model.state_dict().items()
               if k.endswith(("running_mean", "running_var", "num_batches_tracked"))},
    }
    ref = O.herlev_train_step(sd, x, y, stage, sd_prev, prev_resolution=fx["prev_resolution"])
    fx["oracle_vs_reference"] = {"logits": _rel(ref["logits"], logits),
                                 "worst_grad": max(_rel(ref["grads"][k], g) for k, g in grads.items()
                                                   if not k.endswith(("conv_op.0.bias", "conv_op.3.bias")))}
    with O.bf16_storage(True):
        emu = O.herlev_train_step(sd, x, y, stage, sd_prev, prev_resolution=fx["prev_resolution"])
    fx["emu"] = {
        "logits_rel": _rel(emu["logits"], logits), "final_loss": emu["final_loss"], "base_loss": emu["base_loss"],
        "grads_rel": {k: _rel(emu["grads"][k], g) for k, g in grads.items()},
        "grads_cos": {k: _cos(emu["grads"][k], g) for k, g in grads.items()},
    }
    return fx


def primitive_case():
    """Primitive semantics the kernels must reproduce."""
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, 8, 8, 8, generator=g)
    x[0, 0, :2, :2] = 0.0                        # all-equal window -> index of the first element
    x[0, 1, :2, :2] = torch.tensor([[1.0, 3.0], [3.0, 2.0]])
    pooled, idx = F.max_pool2d(x, 2, return_indices=True)
    up = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
    z = torch.randn(2, 1, 8, 8, generator=g) * 3
    y = torch.randint(0, 2, (2, 1, 8, 8), generator=g).float()
    crit = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor([5.0]), reduction="none")
    prev1 = REF.PGUNet1(3, 1)
    prev1.load_state_dict(O.seeded_state_dict(1, 3, 1, seed=1))
    cur = torch.randn(2, 3, 64, 64, generator=g)
    A = REF.UncertaintyGuidedLoss("cpu").generate_uncertainty_map(cur, prev1, 32, 64)
    with contextlib.redirect_stdout(io.StringIO()):
        pu = REF.ProgressiveUNet(3, 1)
        sd1 = O.seeded_state_dict(1, 3, 1, seed=1)
        sd2 = O.seeded_state_dict(2, 3, 1, seed=2)
        moved = pu.transfer_weights(sd1, sd2, 2)
    return {
        "pool_x": x, "pool_out": pooled, "pool_idx": idx, "up_out": up,
        "bce_z": z, "bce_y": y, "bce_out": crit(z, y),
        "unc_x": cur, "unc_A": A,
        "transfer_1_to_2": {k: sample(v) for k, v in moved.items() if v.is_floating_point()},
    }


def herlev_case(stage: int, B: int, K: int, seed: int):
    """One `HerlevTrainer.uncertainty_guided_forward_pass` + backward of the UNMODIFIED reference
    (Herlev/train_herlev.py) with the classifier's Dropout layers switched to eval (they draw from the
    global RNG; everything else is as in train_epoch)."""
    sys.path.insert(0, "/root/reference/Herlev")
    with contextlib.redirect_stdout(io.StringIO()):
        import train_herlev as REFH
    R = O.STAGE_RESOLUTION[stage]
    g = torch.Generator().manual_seed(2000 + seed)
    x = torch.randn(B, 3, R, R, generator=g)
    y = torch.randint(0, K, (B,), generator=g)
    cfg = {"device": "cpu", "epochs_per_stage": 1, "num_classes": K, "uncertainty_alpha": 1.0}
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        tr = REFH.HerlevTrainer(cfg)
    tr.models[stage].load_state_dict(O.seeded_herlev_state_dict(stage, K, seed=stage), strict=True)
    if stage > 1:
        tr.models[stage - 1].load_state_dict(O.seeded_herlev_state_dict(stage - 1, K, seed=stage - 1), strict=True)
        tr.models[stage - 1].eval()
    model = tr.models[stage]
    model.train()
    for m in model.classifier:
        if isinstance(m, torch.nn.Dropout):
            m.eval()
    final_loss, metrics = tr.uncertainty_guided_forward_pass(x, y, stage)
    final_loss.backward()
    return {
        "stage": stage, "B": B, "K": K, "x": x, "y": y,
        "logits": metrics["output"].detach().clone(),
        "final_loss": metrics["final_loss"], "base_loss": metrics["base_loss"],
        "uncertainty_weight_mean": metrics["uncertainty_weight_mean"],
        "uncertainty_weight_std": metrics["uncertainty_weight_std"],
        "correct": int((metrics["output"].argmax(1) == y).sum()),
        "grads": {k: sample(p.grad) for k, p in model.named_parameters() if p.grad is not None},
        "no_grad_keys": [k for k, p in model.named_parameters() if p.grad is None],
        "bn": {k: v.clone() for k, v in model.state_dict().items()
               if k.endswith(("running_mean", "running_var", "num_batches_tracked"))},
    }


def legacy_case():
    """ConvTranspose2d(k=2, s=2) primitive + one forward/backward of the reference's legacy `UNet`
    (UG_unet.py:430-462) at 1 x 3 x 256 x 256 (the smallest size whose 16x16 bottleneck the GEMM tiles
    cover), mean-BCE loss.  Weights: oracle.seeded_legacy_state_dict loaded with strict=True."""
    with contextlib.redirect_stdout(io.StringIO()):
        import UG_unet_parts as REFP
    g = torch.Generator().manual_seed(31)
    # primitive: the UpSample block alone
    up = REFP.UpSample(128, 64)
    sd_up = {k[len("up_convolution_4."):]: v for k, v in O.seeded_legacy_state_dict(3, 1, seed=5).items()
             if k.startswith("up_convolution_4.")}
    up.load_state_dict(sd_up, strict=True)
    up.train()
    x1 = torch.randn(1, 128, 16, 16, generator=g, requires_grad=True)
    x2 = torch.randn(1, 64, 32, 32, generator=g, requires_grad=True)
    t_up = up.up(x1)
    y_up = up(x1, x2)
    gy = torch.randn(y_up.shape, generator=g)
    (y_up * gy).sum().backward()
    prim = {"x1": x1.detach().clone(), "x2": x2.detach().clone(), "gy": gy, "tconv_out": t_up.detach().clone(),
            "out": y_up.detach().clone(), "dx1": x1.grad.clone(), "dx2": x2.grad.clone(),
            "grads": {k: p.grad.clone() for k, p in up.named_parameters()}}
    # full legacy UNet
    net = REF.UNet(3, 1)
    net.load_state_dict(O.seeded_legacy_state_dict(3, 1, seed=6), strict=True)
    net.train()
    x = torch.randn(1, 3, 256, 256, generator=g)
    y = torch.randint(0, 2, (1, 1, 256, 256), generator=g).float()
    logits = net(x)
    loss = F.binary_cross_entropy_with_logits(logits, y)
    loss.backward()
    fx = {"prim": prim, "x": x, "y": y.to(torch.uint8), "logits": logits.detach().clone(), "loss": loss.item(),
          "grads": {k: sample(p.grad) for k, p in net.named_parameters()},
          "bn": {k: sample(v.float()) for k, v in net.state_dict().items()
                 if k.endswith(("running_mean", "running_var"))}}
    net.eval()
    with torch.no_grad():
        fx["eval_logits"] = net(x).clone()
    return fx


def evaluator_case():
    """The reference's MoNuSegEvaluator (MoNuSegImprove/test_monuseg.py) driven on tensors: the
    predict_image / evaluate_dataset arithmetic (:176-201, :236-258) and calculate_metrics (:264-297)
    with a seeded stage-2 model on four 64 x 64 uint8 images."""
    import numpy as np
    sys.path.insert(0, "/root/reference/MoNuSegImprove")
    with contextlib.redirect_stdout(io.StringIO()):
        import test_monuseg as REFE
    g = torch.Generator().manual_seed(41)
    imgs = torch.randint(0, 256, (4, 64, 64, 3), generator=g, dtype=torch.uint8)
    gts = (torch.rand(4, 64, 64, generator=g) > 0.6).float()
    ev = REFE.MoNuSegEvaluator.__new__(REFE.MoNuSegEvaluator)
    ev.device = "cpu"
    ev.model = REF.PGUNet2(3, 1)
    ev.model.load_state_dict(O.seeded_state_dict(2, 3, 1, seed=8), strict=True)
    ev.model.eval()
    masks, confs, per, logits = [], [], [], []
    for i in range(imgs.shape[0]):
        image_tensor = torch.from_numpy(np.array(imgs[i].numpy())).permute(2, 0, 1).float() / 255.0   # :180
        image_tensor = image_tensor.unsqueeze(0)
        with torch.no_grad():
            output = ev.model(image_tensor)
            probs = torch.sigmoid(output)
            pred_mask = (probs > 0.5).float()
        masks.append(pred_mask[0].to(torch.uint8))
        confs.append(probs.mean().item())
        logits.append(output[0])
        m = ev.calculate_metrics(pred_mask.squeeze().numpy(), gts[i].numpy())
        per.append({k: float(v) for k, v in m.items()})
    avg = {k: float(np.mean([m[k] for m in per])) for k in per[0]}
    std = {k: float(np.std([m[k] for m in per])) for k in per[0]}
    return {"images": imgs, "gts": gts.to(torch.uint8), "masks": torch.stack(masks), "confidence": confs,
            "per_image": per, "avg": avg, "std": std, "logits": torch.stack(logits)}


def herlev_evaluator_case():
    """The reference's HerlevEvaluator.calculate_metrics (Herlev/test_herlev.py:217-260; sklearn
    classification_report / confusion_matrix underneath) on seeded labels / predictions with one class
    absent from the labels and one never predicted.  test_herlev.py imports matplotlib / seaborn for its
    plots at module level; they are not installed here and are stubbed out - no plotting code runs."""
    import types
    import numpy as np
    for m in ("matplotlib", "matplotlib.pyplot", "seaborn"):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.path.insert(0, "/root/reference/Herlev")
    with contextlib.redirect_stdout(io.StringIO()):
        import test_herlev as REFHE
    K = 7
    ev = REFHE.HerlevEvaluator.__new__(REFHE.HerlevEvaluator)
    ev.num_classes = K
    ev.class_names = [f"c{i}" for i in range(K)]
    rng = np.random.RandomState(3)
    y = rng.randint(0, K - 1, 200)              # class 6 never occurs
    p = rng.randint(1, K, 200)                  # class 0 never predicted
    p[:60] = y[:60]
    p = np.where(p == 0, 1, p)
    probs = rng.rand(200, K)
    probs /= probs.sum(1, keepdims=True)
    m = ev.calculate_metrics(y, p, probs)
    rep = m["classification_report"]
    return {"K": K, "y": torch.from_numpy(y), "p": torch.from_numpy(p), "accuracy": float(m["accuracy"]),
            "confusion_matrix": torch.tensor(m["confusion_matrix"]),
            "per_class": {i: {k: float(v) for k, v in rep[f"c{i}"].items()} for i in range(K) if f"c{i}" in rep},
            "macro_avg": {k: float(v) for k, v in rep["macro avg"].items()},
            "weighted_avg": {k: float(v) for k, v in rep["weighted avg"].items()},
            "per_class_accuracy": {int(k[1:]): float(v) for k, v in m["per_class_accuracy"].items()}}


B16_TRAINER = {"trainer_stage3_nc1_b16": (3, 16, 1, 33), "trainer_stage4_nc1_b16": (4, 16, 1, 34)}
B16_HERLEV = {"herlev_stage3_k7_b16": (3, 16, 7, 43, 128), "herlev_stage4_k7_b16_256": (4, 16, 7, 44, 256),
              "herlev_stage4_k7_b16_224": (4, 16, 7, 45, 224)}


def main_b16():
    """`--b16`: the batch-16 fixtures (BASELINE.json configs[1] and configs[3] at their real batch size)."""
    import time
    for name, (stage, B, K, seed, R) in B16_HERLEV.items():
        t0 = time.time()
        fx = herlev_case_b16(stage, B, K, seed, R)
        path = os.path.join(OUT, name + ".pt")
        torch.save(fx, path)
        print(f"{name}: final_loss={fx['final_loss']:.6f} emu logits rel {fx['emu']['logits_rel']:.3e} "
              f"oracle-vs-ref {fx['oracle_vs_reference']} -> {os.path.getsize(path) / 1024:.0f} KiB "
              f"({time.time() - t0:.0f} s)", flush=True)
    for name, (stage, B, nc, seed) in B16_TRAINER.items():
        t0 = time.time()
        fx = trainer_case_b16(stage, B, nc, seed)
        path = os.path.join(OUT, name + ".pt")
        torch.save(fx, path)
        print(f"{name}: final_loss={fx['final_loss']:.6f} emu logits rel {fx['emu']['logits_rel']:.3e} "
              f"oracle-vs-ref {fx['oracle_vs_reference']} -> {os.path.getsize(path) / 1024:.0f} KiB "
              f"({time.time() - t0:.0f} s)", flush=True)


def main():
    os.makedirs(OUT, exist_ok=True)
    if "--b16" in sys.argv:
        return main_b16()
    if "--herlev-evaluator-only" in sys.argv or len(sys.argv) == 1:
        torch.save(herlev_evaluator_case(), os.path.join(OUT, "herlev_evaluator.pt"))
        print("herlev_evaluator.pt written")
        if "--herlev-evaluator-only" in sys.argv:
            return
    if "--evaluator-only" in sys.argv or len(sys.argv) == 1:
        torch.save(evaluator_case(), os.path.join(OUT, "evaluator.pt"))
        print("evaluator.pt written")
        if "--evaluator-only" in sys.argv:
            return
    if "--legacy-only" in sys.argv or "--all" in sys.argv or len(sys.argv) == 1:
        fx = legacy_case()
        path = os.path.join(OUT, "legacy_unet.pt")
        torch.save(fx, path)
        print(f"legacy_unet: loss={fx['loss']:.6f} -> {os.path.getsize(path) / 1024:.0f} KiB")
        if "--legacy-only" in sys.argv:
            return
    for name, (stage, B, K, seed) in {"herlev_stage1_k7": (1, 4, 7, 21), "herlev_stage2_k7": (2, 4, 7, 22)}.items():
        fx = herlev_case(stage, B, K, seed)
        path = os.path.join(OUT, name + ".pt")
        torch.save(fx, path)
        print(f"{name}: final_loss={fx['final_loss']:.6f} -> {os.path.getsize(path) / 1024:.0f} KiB")
    if "--herlev-only" in sys.argv:
        return
    cases = {
        "trainer_stage1_nc1": (1, 2, 1, 11),
        "trainer_stage1_nc2_config1": (1, 4, 2, 12),     # BASELINE.json configs[0] shape, smaller batch
        "trainer_stage2_nc1": (2, 2, 1, 13),
        "trainer_stage3_nc1": (3, 1, 1, 14),
        "trainer_stage4_nc1": (4, 1, 1, 15),
    }
    for name, (stage, B, nc, seed) in cases.items():
        fx = trainer_case(stage, B, nc, seed)
        path = os.path.join(OUT, name + ".pt")
        torch.save(fx, path)
        print(f"{name}: final_loss={fx['final_loss']:.6f} -> {os.path.getsize(path) / 1024:.0f} KiB")
    torch.save(primitive_case(), os.path.join(OUT, "primitives.pt"))
    print("primitives.pt written")


if __name__ == "__main__":
    main()

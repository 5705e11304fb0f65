"""Pins the CPU oracle (oracle/ug_oracle.py) against vectors produced by the unmodified reference
(tools/make_goldens.py).  CPU only."""
import os

import pytest
import torch

import ug_oracle as O

CASES = ["trainer_stage1_nc1", "trainer_stage1_nc2_config1", "trainer_stage2_nc1", "trainer_stage3_nc1",
         "trainer_stage4_nc1"]


def _check_sample(t, s, rtol, atol=0.0):
    f = t.flatten()
    if "full" in s:
        torch.testing.assert_close(f, s["full"], rtol=rtol, atol=atol)
    else:
        assert f.numel() == s["numel"]
        torch.testing.assert_close(f[s["idx"]], s["vals"], rtol=rtol, atol=atol)
        assert abs(f.double().norm().item() - s["norm"]) <= rtol * s["norm"] + atol


@pytest.mark.parametrize("name", CASES)
def test_train_step_matches_reference(golden_dir, name):
    fx = torch.load(os.path.join(golden_dir, name + ".pt"))
    stage, nc = fx["stage"], fx["nc"]
    if stage == 4 and os.environ.get("UGP_FAST_TESTS"):
        pytest.skip("fast mode")
    sd = O.seeded_state_dict(stage, 3, nc, seed=stage)
    sd_prev = O.seeded_state_dict(stage - 1, 3, nc, seed=stage - 1) if stage > 1 else None
    out = O.train_step(sd, fx["x"], fx["y"].float(), stage, sd_prev)
    torch.testing.assert_close(out["logits"], fx["logits"], rtol=1e-4, atol=1e-4)
    assert abs(out["final_loss"] - fx["final_loss"]) < 1e-5
    assert abs(out["base_loss"] - fx["base_loss"]) < 1e-5
    assert abs(out["uncertainty_weight_mean"] - fx["uncertainty_weight_mean"]) < 1e-6
    assert abs(out["uncertainty_weight_std"] - fx["uncertainty_weight_std"]) < 1e-6
    if "uncertainty" in fx:
        torch.testing.assert_close(out["uncertainty"], fx["uncertainty"].float(), rtol=0, atol=1e-3)
    if nc == 1:
        assert abs(out["dice"] - fx["dice"]) < 1e-6
        assert abs(out["accuracy"] - fx["accuracy"]) < 1e-7
    gmax = max(g.abs().max().item() for g in out["grads"].values())
    for k, s in fx["grads"].items():
        _check_sample(out["grads"][k], s, rtol=2e-3, atol=2e-5 * gmax)
    for k, v in fx["bn"].items():
        torch.testing.assert_close(out["bn"][k], v, rtol=1e-5, atol=1e-6)
    # optimiser step (RMSprop lr from the stage config, wd 1e-4)
    lr = {1: 3e-4}.get(stage, 1e-4)
    for k, s in fx["after_step"].items():
        p = sd[k].clone()
        O.rmsprop_step(p, out["grads"][k], torch.zeros_like(p), lr)
        # a conv bias in front of train-mode BN has a pure rounding-noise gradient, and the first
        # RMSprop step is lr * 10 * sign(g): its sign is arbitrary there
        noise = k.endswith(("conv_op.0.bias", "conv_op.3.bias"))
        _check_sample(p, s, rtol=1e-4, atol=21 * lr if noise else 2e-5)


def test_primitives(golden_dir):
    import torch.nn.functional as F

    fx = torch.load(os.path.join(golden_dir, "primitives.pt"))
    pooled, idx = F.max_pool2d(fx["pool_x"], 2, return_indices=True)
    assert torch.equal(pooled, fx["pool_out"]) and torch.equal(idx, fx["pool_idx"])
    assert idx[0, 0, 0, 0].item() == 0 and idx[0, 1, 0, 0].item() == 1      # first max wins
    torch.testing.assert_close(O.bce_with_logits(fx["bce_z"], fx["bce_y"], 5.0), fx["bce_out"], rtol=1e-6, atol=1e-6)
    A = O.uncertainty_map(fx["unc_x"], O.seeded_state_dict(1, 3, 1, seed=1), 1, 64)
    torch.testing.assert_close(A, fx["unc_A"], rtol=1e-5, atol=1e-6)
    assert float(A.min()) >= 0.0 and float(A.max()) <= 1.0
    moved, copied = O.transfer_weights(O.seeded_state_dict(1, 3, 1, seed=1), O.seeded_state_dict(2, 3, 1, seed=2))
    assert len(copied) == 42
    for k, s in fx["transfer_1_to_2"].items():
        _check_sample(moved[k], s, rtol=0, atol=0)


def test_loss_affine_in_alpha():
    """final(alpha) = base + alpha * mean(l*A): demo_uncertainty_guided.py:155-201."""
    g = torch.Generator().manual_seed(0)
    z = torch.randn(2, 1, 16, 16, generator=g)
    y = torch.randint(0, 2, (2, 1, 16, 16), generator=g).float()
    A = torch.rand(2, 1, 16, 16, generator=g)
    f0, b0 = O.weighted_loss(z, y, A, 0.0)
    f1, _ = O.weighted_loss(z, y, A, 1.0)
    f2, _ = O.weighted_loss(z, y, A, 2.0)
    assert abs(f0.item() - b0.item()) < 1e-7
    assert abs((f2 - f1).item() - (f1 - f0).item()) < 1e-6
    fs1, bs1 = O.weighted_loss(z, y, None, 1.0)
    assert abs(fs1.item() - bs1.item()) < 1e-7


def test_state_dict_tables():
    counts = {1: 10048769, 2: 12551042, 3: 13175491, 4: 13395780}
    for stage, n in counts.items():
        shapes = O.state_dict_shapes(stage, 3, 1)
        params = sum(int(torch.tensor(s).prod()) if len(s) else 1 for k, s in shapes.items()
                     if not k.endswith(("running_mean", "running_var", "num_batches_tracked")))
        assert params == n
    assert len(O.state_dict_shapes(4, 3, 1)) == 134


@pytest.mark.parametrize("name", ["herlev_stage1_k7", "herlev_stage2_k7"])
def test_herlev_step_matches_reference(golden_dir, name):
    """oracle.herlev_train_step vs the unmodified Herlev/train_herlev.py step (classifier Dropout inactive)."""
    fx = torch.load(os.path.join(golden_dir, name + ".pt"))
    stage, K = fx["stage"], fx["K"]
    sd = O.seeded_herlev_state_dict(stage, K, seed=stage)
    sd_prev = O.seeded_herlev_state_dict(stage - 1, K, seed=stage - 1) if stage > 1 else None
    out = O.herlev_train_step(sd, fx["x"], fx["y"], stage, sd_prev)
    torch.testing.assert_close(out["logits"], fx["logits"], rtol=1e-4, atol=1e-5)
    assert abs(out["final_loss"] - fx["final_loss"]) < 1e-5
    assert abs(out["base_loss"] - fx["base_loss"]) < 1e-5
    assert abs(out["uncertainty_weight_mean"] - fx["uncertainty_weight_mean"]) < 1e-6
    assert abs(out["uncertainty_weight_std"] - fx["uncertainty_weight_std"]) < 1e-6
    assert out["correct"] == fx["correct"]
    gmax = max(g.abs().max().item() for g in out["grads"].values() if g is not None)
    for k, s in fx["grads"].items():
        _check_sample(out["grads"][k], s, rtol=2e-3, atol=2e-5 * gmax)
    # decoder / heads / down4 never run in the classification model: no gradient
    for k in fx["no_grad_keys"]:
        assert out["grads"][k] is None, k
    for k, v in fx["bn"].items():
        if k in out["bn"]:
            torch.testing.assert_close(out["bn"][k], v, rtol=1e-5, atol=1e-6)
        else:                                      # untouched blocks keep their buffers
            torch.testing.assert_close(sd[k], v, rtol=0, atol=0)


def test_legacy_unet_matches_reference(golden_dir):
    """oracle.conv_transpose2x2 / legacy_unet_forward vs the unmodified UpSample and UNet
    (UG_unet_parts.py:57-67, UG_unet.py:430-462)."""
    import torch.nn.functional as F

    fx = torch.load(os.path.join(golden_dir, "legacy_unet.pt"))
    shapes = O.legacy_unet_shapes(3, 1)
    assert len(shapes) == 136
    params = sum(int(torch.tensor(s).prod()) if len(s) else 1 for k, s in shapes.items()
                 if not k.endswith(("running_mean", "running_var", "num_batches_tracked")))
    assert params == 31043521
    # the UpSample block alone: ConvTranspose2d(k2,s2) -> cat([up, skip]) -> DoubleConv
    p = fx["prim"]
    sd_up = {k[len("up_convolution_4."):]: v for k, v in O.seeded_legacy_state_dict(3, 1, seed=5).items()
             if k.startswith("up_convolution_4.")}
    x1 = p["x1"].clone().requires_grad_(True)
    x2 = p["x2"].clone().requires_grad_(True)
    w = sd_up["up.weight"].clone().requires_grad_(True)
    b = sd_up["up.bias"].clone().requires_grad_(True)
    t = O.conv_transpose2x2(x1, w, b)
    torch.testing.assert_close(t, p["tconv_out"], rtol=1e-5, atol=1e-6)
    # closed form of the stride == kernel case: each output pixel has one contributing input pixel
    manual = torch.einsum("nchw,cdij->ndhiwj", p["x1"], sd_up["up.weight"]).reshape(1, 64, 32, 32) \
        + sd_up["up.bias"].view(1, -1, 1, 1)
    torch.testing.assert_close(manual, p["tconv_out"], rtol=1e-4, atol=1e-5)
    y = O.double_conv(torch.cat([t, x2], 1), sd_up, "conv", True)
    torch.testing.assert_close(y, p["out"], rtol=1e-4, atol=1e-4)
    (y * p["gy"]).sum().backward()
    torch.testing.assert_close(x1.grad, p["dx1"], rtol=2e-3, atol=2e-4)
    torch.testing.assert_close(x2.grad, p["dx2"], rtol=2e-3, atol=2e-4)
    torch.testing.assert_close(w.grad, p["grads"]["up.weight"], rtol=2e-3, atol=2e-4)
    torch.testing.assert_close(b.grad, p["grads"]["up.bias"], rtol=2e-3, atol=2e-3)
    if os.environ.get("UGP_FAST_TESTS"):
        return
    # the whole legacy UNet, train-mode forward + backward of mean BCE
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v)
          for k, v in O.seeded_legacy_state_dict(3, 1, seed=6).items()}
    bn = {}
    logits = O.legacy_unet_forward(sd, fx["x"], True, bn)
    torch.testing.assert_close(logits, fx["logits"], rtol=1e-4, atol=1e-4)
    loss = F.binary_cross_entropy_with_logits(logits, fx["y"].float())
    assert abs(loss.item() - fx["loss"]) < 1e-5
    loss.backward()
    gmax = max(v.grad.abs().max().item() for v in sd.values() if v.requires_grad)
    for k, s in fx["grads"].items():
        _check_sample(sd[k].grad, s, rtol=2e-3, atol=2e-5 * gmax)
    for k, s in fx["bn"].items():
        _check_sample(bn[k], s, rtol=1e-5, atol=1e-6)
    with torch.no_grad():
        sd_eval = {k: v.detach() for k, v in sd.items()}
        sd_eval.update(bn)
        torch.testing.assert_close(O.legacy_unet_forward(sd_eval, fx["x"], False), fx["eval_logits"], rtol=1e-4,
                                   atol=1e-4)


def test_evaluator_matches_reference(golden_dir):
    """oracle.seg_metrics / evaluate_segmentation vs the unmodified MoNuSegEvaluator arithmetic
    (MoNuSegImprove/test_monuseg.py:176-201, 236-258, 264-297)."""
    fx = torch.load(os.path.join(golden_dir, "evaluator.pt"))
    sd = O.seeded_state_dict(2, 3, 1, seed=8)
    masks, conf, per, avg, std, logits = O.evaluate_segmentation(sd, 2, fx["images"], fx["gts"].float())
    torch.testing.assert_close(logits, fx["logits"], rtol=1e-4, atol=1e-4)
    assert torch.equal(masks.to(torch.uint8), fx["masks"])
    for i, m in enumerate(fx["per_image"]):
        for k, v in m.items():
            assert abs(per[i][k] - v) < 1e-6, (i, k)        # the reference accumulates in float32
        assert abs(float(conf[i]) - fx["confidence"][i]) < 1e-6
    for k in fx["avg"]:
        assert abs(avg[k] - fx["avg"][k]) < 1e-6 and abs(std[k] - fx["std"][k]) < 1e-6


def test_classification_evaluator_matches_reference(golden_dir):
    """oracle.classification_metrics vs the unmodified HerlevEvaluator.calculate_metrics (sklearn report and
    confusion matrix underneath; Herlev/test_herlev.py:217-260), incl. a class absent from the labels and a
    class that is never predicted (zero_division=0)."""
    fx = torch.load(os.path.join(golden_dir, "herlev_evaluator.pt"))
    out = O.classification_metrics(fx["y"], fx["p"], fx["K"])
    assert abs(out["accuracy"] - fx["accuracy"]) < 1e-12
    assert torch.equal(out["confusion_matrix"], fx["confusion_matrix"])
    for k, rep in fx["per_class"].items():
        assert abs(out["per_class"][k]["precision"] - rep["precision"]) < 1e-12
        assert abs(out["per_class"][k]["recall"] - rep["recall"]) < 1e-12
        assert abs(out["per_class"][k]["f1"] - rep["f1-score"]) < 1e-12
        assert out["per_class"][k]["support"] == int(rep["support"])
    for name in ("macro_avg", "weighted_avg"):
        for m, key in (("precision", "precision"), ("recall", "recall"), ("f1", "f1-score")):
            assert abs(out[name][m] - fx[name][key]) < 1e-12
    assert out["per_class_accuracy"].keys() == fx["per_class_accuracy"].keys()
    for k, v in fx["per_class_accuracy"].items():
        assert abs(out["per_class_accuracy"][k] - v) < 1e-12


# --------------------------------------------------------------------------- batch-16 fixtures (tools/make_goldens.py --b16)
def _regen(fx):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import parity_table as P
    return P.regenerate_inputs(fx)


def test_batch16_fixtures_record_the_oracle_pin(golden_dir):
    """Every batch-16 fixture stores how far the oracle restatement was from the unmodified reference when
    the fixture was generated (same process): logits identical, gradients at fp32 summation-order noise."""
    for name in ("trainer_stage3_nc1_b16", "trainer_stage4_nc1_b16", "herlev_stage3_k7_b16",
                 "herlev_stage4_k7_b16_256", "herlev_stage4_k7_b16_224"):
        fx = torch.load(os.path.join(golden_dir, name + ".pt"))
        assert fx["B"] == 16
        assert fx["oracle_vs_reference"]["logits"] <= 1e-6 and fx["oracle_vs_reference"]["worst_grad"] <= 1e-4, name
        assert set(fx["emu"]["grads_rel"]) == set(fx["grads"])


def test_oracle_reproduces_the_stage3_batch16_fixture(golden_dir):
    """Re-runs the oracle at batch 16 (stage 3: ~10 s; the stage-4 twin is covered by the recorded pin above
    and by the batch-1 stage-4 case) on inputs regenerated from the fixture's seed."""
    fx = torch.load(os.path.join(golden_dir, "trainer_stage3_nc1_b16.pt"))
    x, y = _regen(fx)
    sd = O.seeded_state_dict(3, 3, 1, seed=3)
    out = O.train_step(sd, x, y, 3, O.seeded_state_dict(2, 3, 1, seed=2))
    torch.testing.assert_close(out["logits"], fx["logits"], rtol=1e-4, atol=1e-4)
    assert abs(out["final_loss"] - fx["final_loss"]) < 1e-5 and abs(out["base_loss"] - fx["base_loss"]) < 1e-5
    assert abs(out["dice"] - fx["dice"]) < 1e-6 and abs(out["accuracy"] - fx["accuracy"]) < 1e-9
    for k, s in fx["grads"].items():
        if k.endswith(("conv_op.0.bias", "conv_op.3.bias")):
            continue
        _check_sample(out["grads"][k], s, rtol=2e-3, atol=1e-7)


@pytest.mark.parametrize("name", ["herlev_stage3_k7_b16", "herlev_stage4_k7_b16_224"])
def test_oracle_reproduces_the_herlev_batch16_fixtures(golden_dir, name):
    fx = torch.load(os.path.join(golden_dir, name + ".pt"))
    x, y = _regen(fx)
    stage, K = fx["stage"], fx["K"]
    out = O.herlev_train_step(O.seeded_herlev_state_dict(stage, K, seed=stage), x, y, stage,
                              O.seeded_herlev_state_dict(stage - 1, K, seed=stage - 1),
                              prev_resolution=fx["prev_resolution"])
    torch.testing.assert_close(out["logits"], fx["logits"], rtol=1e-4, atol=1e-5)
    assert abs(out["final_loss"] - fx["final_loss"]) < 1e-5 and out["correct"] == fx["correct"]
    for k, s in fx["grads"].items():
        if k.endswith(("conv_op.0.bias", "conv_op.3.bias")):
            continue
        _check_sample(out["grads"][k], s, rtol=2e-3, atol=1e-7)


def test_precision_emulations_bracket_the_reference():
    """bf16-storage and TF32-operand emulations of the oracle: both perturb the fp32 result, TF32 less."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 3, 32, 32, generator=g)
    y = torch.randint(0, 2, (2, 1, 32, 32), generator=g).float()
    sd = O.seeded_state_dict(1, 3, 1, seed=1)
    ref = O.train_step(sd, x, y, 1)
    with O.bf16_storage(True):
        b = O.train_step(sd, x, y, 1)
    with O.tf32_operands(True):
        t = O.train_step(sd, x, y, 1)
    rel = lambda a, r: ((a.double() - r.double()).norm() / r.double().norm()).item()
    eb, et = rel(b["logits"], ref["logits"]), rel(t["logits"], ref["logits"])
    assert 0 < et < eb < 5e-2 and et < 5e-3
    v = O._tf32(torch.tensor([1.0 + 2 ** -11, 1.0 + 2 ** -12 + 2 ** -20]))
    assert v.tolist() == [1.0 + 2 ** -10, 1.0 + 2 ** -10] or v.tolist() == [1.0 + 2 ** -10, 1.0]

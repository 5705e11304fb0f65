"""CPU-only checks of the host side: module surface / state_dict parity with the reference tables,
weight transfer, trainer metric arithmetic, the no-CPU-fallback contract and the C-ABI exports."""
import os
import re

import pytest
import torch

import ug_oracle as O
import ug_pg_unet_b200 as U

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("stage", [1, 2, 3, 4])
@pytest.mark.parametrize("nc", [1, 2, 7])
def test_state_dict_keys_and_shapes_match_reference_tables(stage, nc):
    model = U.STAGE_MODELS[stage](3, nc)
    sd = model.state_dict()
    shapes = O.state_dict_shapes(stage, 3, nc)          # pinned against the reference in test_oracle_golden
    assert list(sd.keys()) == list(shapes.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == shapes[k], k
    assert sd["inc.conv.conv_op.1.num_batches_tracked"].dtype == torch.int64
    # reference-format state dicts load strictly, and round-trip through torch.save
    model.load_state_dict(O.seeded_state_dict(stage, 3, nc, seed=3), strict=True)


def test_module_attribute_names():
    m = U.PGUNet4(3, 1)
    for name in ("inc", "down1", "down2", "down3", "down4", "up1", "up2", "up3", "up4", "outc1", "outc2", "outc3", "outc4"):
        assert hasattr(m, name)
    assert hasattr(U.PGUNet1(3, 1), "outc")
    assert isinstance(m.down1.mpconv[0], torch.nn.MaxPool2d)
    assert isinstance(m.inc.conv.conv_op[0], torch.nn.Conv2d) and isinstance(m.inc.conv.conv_op[1], torch.nn.BatchNorm2d)
    pu = U.ProgressiveUNet(3, 1)
    assert pu.stage_resolutions == {1: 32, 2: 64, 3: 128, 4: 256} and pu.current_stage == 1
    assert set(pu.stages) == {1, 2, 3, 4} and pu.stage4 is pu.stages[4]


def test_progressive_unet_signatures_and_errors():
    pu = U.ProgressiveUNet(in_channels=3, out_channels=1, stage=2, task_type="segmentation")   # README form
    assert pu.current_stage == 2 and pu.get_current_resolution() == 64 and pu.num_classes == 1
    with pytest.raises(ValueError, match="Stage must be 1, 2, 3, or 4"):
        pu.set_stage(5)
    with pytest.raises(TypeError):
        U.ProgressiveUNet(3)
    ul = U.UncertaintyGuidedLoss(task_type="segmentation")
    assert isinstance(ul, torch.nn.Module) and ul.device == "cuda"


def test_transfer_weights_matches_reference_semantics(golden_dir):
    sd1 = O.seeded_state_dict(1, 3, 1, seed=1)
    sd2 = O.seeded_state_dict(2, 3, 1, seed=2)
    moved = U.transfer_state(sd1, sd2)
    ref, copied = O.transfer_weights(sd1, sd2)
    assert len(copied) == 42
    for k in ref:
        assert torch.equal(moved[k], ref[k]), k
    # partial leading-slice copy: (512,3,3,3) -> first 256 output channels
    k = "inc.conv.conv_op.0.weight"
    assert torch.equal(moved[k], sd1[k][:256])
    fx = torch.load(os.path.join(golden_dir, "primitives.pt"))
    for key, s in fx["transfer_1_to_2"].items():
        f = moved[key].flatten()
        got = f if "full" in s else f[s["idx"]]
        assert torch.equal(got, s["full"] if "full" in s else s["vals"]), key
    pu = U.ProgressiveUNet(3, 1)
    again = pu.transfer_weights(sd1, sd2, 2)
    assert all(torch.equal(again[k], moved[k]) for k in moved)


def test_no_cpu_fallback():
    m = U.PGUNet1(3, 1)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.randn(1, 3, 32, 32))
    with pytest.raises(RuntimeError, match="CUDA"):
        U.DoubleConv(64, 64)(torch.randn(1, 64, 16, 16))
    with pytest.raises(RuntimeError, match="CUDA"):
        U.UncertaintyGuidedLoss("cpu").generate_uncertainty_map(torch.randn(1, 3, 64, 64), m, 32, 64)
    with pytest.raises(ValueError):
        U.DoubleConv(96, 64)           # channel counts the kernels do not cover fail loudly at construction
    with pytest.raises(ValueError):
        U.OutConv(64, 9)


def test_generic_criterion_path_matches_reference_arithmetic():
    """A loss_fn that is not the fused BCEWithLogitsLoss goes through the reference's own arithmetic
    (UG_unet.py:78-94); works on CPU tensors because no kernel is involved."""
    g = torch.Generator().manual_seed(0)
    z = torch.randn(2, 1, 8, 8, generator=g, requires_grad=True)
    y = torch.randint(0, 2, (2, 1, 8, 8), generator=g).float()
    A = torch.rand(2, 1, 8, 8, generator=g)
    ul = U.UncertaintyGuidedLoss("cpu")
    crit = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor([5.0]), reduction="none")
    final, base = ul.apply_uncertainty_weighted_loss(crit, z, y, A, 0.5)
    rf, rb = O.weighted_loss(z, y, A, 0.5, 5.0)
    assert abs(final.item() - rf.item()) < 1e-6 and abs(base - rb.item()) < 1e-6
    f1, b1 = ul.apply_uncertainty_weighted_loss(crit, z, y, None, 1.0)
    assert abs(f1.item() - b1) < 1e-6


def test_trainer_surface_and_metric_arithmetic():
    tr = U.UncertaintyGuidedProgressiveTrainer(3, 1, device="cpu", uncertainty_alpha=0.7)
    assert tr.stage_configs[1] == {"resolution": 32, "epochs_per_stage": 40, "lr": 3e-4}
    assert tr.stage_configs[4]["lr"] == 1e-4 and set(tr.models) == {1, 2, 3, 4}
    assert tr.current_model is tr.models[1] and tr.uncertainty_alpha == 0.7
    assert isinstance(tr.base_criterion, torch.nn.BCEWithLogitsLoss) and float(tr.base_criterion.pos_weight) == 5.0
    assert isinstance(tr.optimizer, torch.optim.RMSprop)
    d = tr.optimizer.defaults
    assert (d["lr"], d["weight_decay"], d["alpha"], d["eps"], d["momentum"], d["centered"]) == (3e-4, 1e-4, 0.99, 1e-8, 0, False)
    for key in ("train_loss", "val_loss", "train_dice", "val_dice", "uncertainty_weights_mean",
                "uncertainty_weights_std", "base_loss", "stage_transitions"):
        assert tr.history[key] == []
    # fused-kernel accumulators -> reference metrics (uncertainty_guided_trainer.py:90-123,177-181)
    g = torch.Generator().manual_seed(1)
    B, n = 3, 3 * 16
    z = torch.randn(B, 1, 4, 4, generator=g)
    y = torch.randint(0, 2, (B, 1, 4, 4), generator=g).float()
    A = torch.rand(B, 1, 4, 4, generator=g)
    l = O.bce_with_logits(z, y, 5.0)
    pred = (z > 0).float()
    acc = [(l * (1 + 0.7 * A)).sum(), l.sum(), A.sum(), (A * A).sum(), pred.ne(y).sum().float()]
    for b in range(B):
        acc += [(pred[b] * y[b]).sum(), pred[b].sum(), y[b].sum()]
    m = torch.stack([a.double() for a in acc])
    final, base, dice, accuracy, um, us = tr._metrics_from_device(m, n, B, True)
    assert abs(final - (l * (1 + 0.7 * A)).mean().item()) < 1e-6
    assert abs(base - l.mean().item()) < 1e-6
    assert abs(dice - O.dice_coefficient(pred.squeeze(1), y.squeeze(1)).item()) < 1e-6
    assert abs(accuracy - O.accuracy(pred.squeeze(1), y.squeeze(1))) < 1e-9
    assert abs(um - A.mean().item()) < 1e-6 and abs(us - A.std().item()) < 1e-5
    # reference helper methods keep their semantics
    assert torch.equal(tr.get_predictions(z), O.predictions(z))
    assert abs(tr.dice_coefficient(pred.squeeze(1), y.squeeze(1)).item() - dice) < 1e-6


def test_c_abi_library_exports_every_declared_symbol():
    from ug_pg_unet_b200 import _lib

    header = open(os.path.join(ROOT, "include", "ugp_b200.h")).read()
    declared = set(re.findall(r"\b(ugp_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared == set(_lib.EXPORTED), declared ^ set(_lib.EXPORTED)
    handle = _lib.lib()                      # loads (and builds if nvcc is present); no compute call
    for name in declared:
        assert hasattr(handle, name), name
    assert handle.ugp_version() >= 100
    assert handle.ugp_set_option(b"no_such_option", 1) < 0
    assert b"no_such_option" in handle.ugp_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ug_pg_unet_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "ug_oracle" not in src and "import oracle" not in src, fn


def test_herlev_model_surface_matches_reference_tables():
    """HerlevClassificationModel: state_dict keys/shapes, attribute names, encoder parameter set."""
    import ug_oracle as O
    import ug_pg_unet_b200.herlev as H

    for stage in (1, 2, 3, 4):
        m = H.HerlevClassificationModel(stage, 7)
        sd = m.state_dict()
        want = O.seeded_herlev_state_dict(stage, 7)
        assert set(sd) == set(want)
        for k, v in want.items():
            assert tuple(sd[k].shape) == tuple(v.shape), k
        m.load_state_dict(want, strict=True)
        assert isinstance(m.classifier[0], torch.nn.AdaptiveAvgPool2d) and isinstance(m.classifier[9], torch.nn.Linear)
        enc = {id(p) for p in m.encoder_parameters()}
        names = {k.split(".")[1] for k, p in m.named_parameters() if id(p) in enc}
        assert names == {"inc", *[f"down{i}" for i in range(5 - stage, 4)]}, (stage, names)
    with pytest.raises(ValueError):
        H.HerlevClassificationModel(5, 7)
    m = H.HerlevClassificationModel(1, 7)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 3, 32, 32))


def test_metrics_from_counts_matches_the_oracle():
    """evaluate.metrics_from_counts (host arithmetic on the kernel's confusion counts) vs
    oracle.seg_metrics (MoNuSegImprove/test_monuseg.py:264-297), including empty masks."""
    import torch
    import ug_oracle as O
    from ug_pg_unet_b200.evaluate import metrics_from_counts

    g = torch.Generator().manual_seed(0)
    cases = [((torch.rand(16, 16, generator=g) > 0.5).float(), (torch.rand(16, 16, generator=g) > 0.7).float()),
             (torch.zeros(8, 8), torch.zeros(8, 8)), (torch.ones(8, 8), torch.zeros(8, 8)),
             (torch.zeros(8, 8), torch.ones(8, 8))]
    for p, t in cases:
        want = O.seg_metrics(p, t)
        got = metrics_from_counts(float((p * t).sum()), float(p.sum()), float(t.sum()), float(p.numel()))
        for k, v in want.items():
            assert abs(got[k] - v) < 1e-12, k


def test_fused_optimizers_delegate_cpu_parameters_to_the_stock_implementation():
    """FusedAdam / FusedRMSprop subclass the torch optimisers: with parameters the kernels cannot take
    (CPU tensors here) `step()` is the stock update, and the state_dict layout is the stock one."""
    import torch
    from ug_pg_unet_b200.optim import FusedAdam, FusedRMSprop

    g = torch.Generator().manual_seed(0)
    for fused_cls, stock_cls, kw in ((FusedAdam, torch.optim.Adam, {"lr": 1e-3, "weight_decay": 1e-4}),
                                     (FusedRMSprop, torch.optim.RMSprop, {"lr": 1e-3, "weight_decay": 1e-4})):
        pa = [torch.randn(5, 3, generator=g).requires_grad_(True), torch.randn(7, generator=g).requires_grad_(True)]
        pb = [p.detach().clone().requires_grad_(True) for p in pa]
        oa, ob = fused_cls(pa, **kw), stock_cls(pb, **kw)
        for _ in range(3):
            for a, b in zip(pa, pb):
                gr = torch.randn(a.shape, generator=g)
                a.grad, b.grad = gr.clone(), gr.clone()
            oa.step()
            ob.step()
        for a, b in zip(pa, pb):
            torch.testing.assert_close(a, b, rtol=1e-6, atol=1e-7)
        sa, sb = oa.state_dict(), ob.state_dict()
        assert set(sa["state"][0]) == set(sb["state"][0])
        assert float(sa["state"][0]["step"]) == float(sb["state"][0]["step"]) == 3.0


def test_legacy_unet_surface_matches_reference_tables():
    """UNet / DownSample / UpSample (UG_unet.py:430-462, UG_unet_parts.py:31-41, 57-67): attribute names,
    the 136 state_dict keys / shapes of the reference (oracle table, itself loaded strict=True into the
    reference by tools/make_goldens.py) and a strict load of the oracle's seeded weights."""
    import torch
    import ug_oracle as O
    import ug_pg_unet_b200 as U

    net = U.UNet(3, 2)
    want = O.legacy_unet_shapes(3, 2)
    got = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    assert got == want and len(got) == 136
    for name in ("down_convolution_1", "down_convolution_4", "bottle_neck", "up_convolution_1", "up_convolution_4", "out"):
        assert hasattr(net, name)
    assert isinstance(net.up_convolution_2.up, torch.nn.ConvTranspose2d) and net.up_convolution_2.up.stride == (2, 2)
    assert isinstance(net.down_convolution_1.pool, torch.nn.MaxPool2d)
    net.load_state_dict(O.seeded_legacy_state_dict(3, 2, seed=1), strict=True)
    with pytest.raises(RuntimeError, match="CUDA"):
        net(torch.zeros(1, 3, 256, 256))                      # no CPU fallback
    with pytest.raises(ValueError):
        U.UpSample(96, 64)                                    # channel counts the GEMM tiles cannot take


def test_byte_inputs_are_validated_on_the_host():
    import torch
    import ug_pg_unet_b200 as U

    m = U.PGUNet1(3, 1)
    assert m.set_input_norm((0.485, 0.456, 0.406), (0.229, 0.224, 0.225)) is m
    assert m.input_mean == (0.485, 0.456, 0.406) and m.input_std == (0.229, 0.224, 0.225)
    assert "input_mean" not in m.state_dict()                # not part of the checkpoint format
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 32, 32, 3, dtype=torch.uint8))       # bytes on the CPU: still no fallback


def test_classification_metrics_from_confusion_match_the_reference_golden(golden_dir):
    """evaluate.classification_metrics_from_confusion (host arithmetic on the device-side confusion matrix of
    ugp_cls_eval) vs the fixture produced by the unmodified HerlevEvaluator.calculate_metrics."""
    import os
    import torch
    from ug_pg_unet_b200.evaluate import classification_metrics_from_confusion

    fx = torch.load(os.path.join(golden_dir, "herlev_evaluator.pt"))
    K = fx["K"]
    cm = torch.zeros(K, K, dtype=torch.int64)
    for t, p in zip(fx["y"].tolist(), fx["p"].tolist()):
        cm[t, p] += 1                                         # what the kernel accumulates: counts[label][pred]
    got = classification_metrics_from_confusion(cm)
    assert torch.equal(got["confusion_matrix"], fx["confusion_matrix"])
    assert abs(got["accuracy"] - fx["accuracy"]) < 1e-12 and got["total_samples"] == 200
    for k, rep in fx["per_class"].items():
        assert abs(got["precision"][k] - rep["precision"]) < 1e-12
        assert abs(got["recall"][k] - rep["recall"]) < 1e-12
        assert abs(got["f1"][k] - rep["f1-score"]) < 1e-12
        assert got["support"][k] == int(rep["support"])
    for name in ("macro_avg", "weighted_avg"):
        for m, key in (("precision", "precision"), ("recall", "recall"), ("f1", "f1-score")):
            assert abs(got[name][m] - fx[name][key]) < 1e-12
    assert {k: round(v, 12) for k, v in got["per_class_accuracy"].items()} == \
        {k: round(v, 12) for k, v in fx["per_class_accuracy"].items()}


def test_fused_first_layer_shape_rule_and_options():
    """Host-only entry points of the C ABI (no GPU touched): which first-layer shapes the fused kernels take
    (ugp_conv_first_supported / _wgrad_supported, include/ugp_b200.h) - everything else keeps the packer +
    GEMM path - and the tuning switches this round added are known to ugp_set_option."""
    from ug_pg_unet_b200 import _lib
    handle = _lib.lib()
    fwd, wg = handle.ugp_conv_first_supported, handle.ugp_conv_first_wgrad_supported
    assert fwd(3, 64, 16, 256, 256) == 1 and wg(3, 64, 16, 256, 256) == 1          # configs[1]
    assert fwd(3, 128, 16, 128, 128) == 1 and wg(3, 128, 16, 128, 128) == 0        # the stage-3 network: forward only
    assert fwd(3, 64, 16, 224, 224) == 1                                           # Herlev at the reference's 224 x 224
    assert fwd(3, 256, 16, 64, 64) == 0 and fwd(3, 512, 16, 32, 32) == 0           # stage-2 / stage-1 widths: packer + GEMM
    assert fwd(1, 64, 16, 256, 256) == 0 and fwd(4, 64, 16, 256, 256) == 0         # other channel counts
    assert fwd(3, 64, 1, 28, 28) == 0                                              # 784 pixels: not a multiple of 128
    for name in (b"pdl", b"conv_first_fused", b"conv_first_patch", b"wgrad_max_ctas", b"conv_small_kn"):
        assert handle.ugp_set_option(name, 1 if name != b"wgrad_max_ctas" else 0) == 0
    assert handle.ugp_set_option(b"conv_first_fused", 0) == 0 and fwd(3, 64, 16, 256, 256) == 0
    assert handle.ugp_set_option(b"conv_first_fused", 1) == 0 and fwd(3, 64, 16, 256, 256) == 1

"""GPU parity tests (run with `-m gpu` on a B200): the CUDA path through the C ABI against
(a) plain torch ops kernel by kernel, (b) the CPU oracle and the reference-generated golden fixtures
model by model, (c) size-independent properties at BASELINE.json's full sizes.

Tolerances.  Integer / index work (max-pool indices, BN step counters, masks / classes outside the
logit-noise band) is bit-exact.  Floating point: a kernel vs torch on identical bf16 inputs is within bf16
output rounding (rel 6e-3 on a tensor; 1e-5 for the fp32 head / loss kernels).  Whole models are compared
with the reference-generated fixtures against north_star's bars (logits / gradients rel 2e-2, loss 1e-3)
wherever the measured table (profiles/r02_parity_table.md) shows a bf16-storage pipeline can
meet them, and against the fixture's recorded bf16-storage floor (the fp32 oracle with tensors rounded to
bf16 where the kernels store bf16) where it cannot - see the comment above `_check_grad_rows`.
"""
import math
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


@pytest.fixture(scope="module")
def dev():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return torch.device("cuda", 0)


# --------------------------------------------------------------------------- kernel by kernel
@pytest.mark.parametrize("group", ["elementwise", "conv_fwd", "conv_dgrad", "conv_wgrad", "head_loss"])
def test_kernels_against_torch(dev, group):
    import gpu_check_ops

    assert getattr(gpu_check_ops, "group_" + group)(), f"kernel group {group} failed (see stdout)"


def test_halo_and_per_tap_convolutions_agree(dev):
    from ug_pg_unet_b200 import _lib, ops

    g = torch.Generator(device="cuda").manual_seed(0)
    try:
        for (B, H, W, c0, c1, cout) in [(1, 4, 128, 64, 0, 64), (2, 128, 128, 64, 64, 128), (1, 256, 256, 64, 0, 64)]:
            x0 = torch.randn(B, H, W, c0, device="cuda", generator=g).bfloat16()
            x1 = torch.randn(B, H, W, c1, device="cuda", generator=g).bfloat16() if c1 else None
            w = torch.randn(cout, c0 + c1, 3, 3, device="cuda", generator=g) / 30
            wf, _ = ops.pack_weights(w)
            outs = []
            for mode in (0, 1):
                _lib.set_option("conv_halo", mode)
                st = ops.new_stats(cout, "cuda")
                outs.append((ops.conv_gemm(x0, x1, wf, 9, cout, stats=st), st))
            assert rel(outs[1][0], outs[0][0]) < 1e-4          # same products, fp32 order differs per chunk
            assert rel(outs[1][1].sum(0), outs[0][1].sum(0)) < 1e-5
    finally:
        _lib.set_option("conv_halo", 1)


# --------------------------------------------------------------------------- model level
def _run_cuda_step(stage, nc, x, y, sd, sd_prev, dev):
    import ug_pg_unet_b200 as U

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
    return model, logits, A, final, base, ugl


# north_star: "logits and gradients within rel 2e-2 (bf16) ..., loss within 1e-3, bit-exact argmax masks".
# Measured (profiles/r02_parity_table.md): a bf16-STORAGE pipeline cannot meet 2e-2 everywhere
# on these random-init networks - the fp32 oracle with tensors merely rounded to bf16 where the kernels
# store bf16 (all arithmetic fp32) is already 6e-2 away on the stage-4 logits and 15-55 % away on most
# gradient tensors (random labels: the gradient is what is left of a sum of 10^6 sign-random terms), and
# a TF32-operand emulation still misses the 1e-3 alternative by 6x / 100x.  So every assertion below
# names north_star's bar where the table shows it is met and the fixture's recorded emulation floor where
# it is not; a kernel bug (wrong tap, wrong scale, dropped term) moves a tensor by O(1) and fails both.
NORTH_STAR_REL = 2e-2
NORTH_STAR_LOSS = 1e-3


def _check_grad_rows(rows, what):
    """Per gradient tensor: north_star's 2e-2, or - where the bf16-storage emulation itself is further
    away than that - within 1.3x of the emulation's distance AND pointing the same way as the reference
    gradient (cosine no worse than the emulation's minus 0.05, never below 0.8) AND of the right size."""
    for k, v in rows.items():
        if v["rel"] <= NORTH_STAR_REL:
            continue
        assert v["rel"] <= 1.3 * v["floor_rel"] + 1e-2, (what, k, v)
        assert v["cos"] >= max(0.8, v["floor_cos"] - 0.05), (what, k, v)
        assert 0.75 <= v["norm_ratio"] <= 1.3, (what, k, v)
        # nothing beyond 0.3 that the emulation does not itself show
        assert v["rel"] <= 0.3 or v["floor_rel"] >= 0.25, (what, k, v)


# fixtures whose logits meet north_star's 2e-2 outright (profiles/r02_parity_table.md); for the others the
# bf16-storage emulation itself is further away (stage 2: 2.3e-2, stage 4: 5.9e-2 / 6.1e-2) and the bar is
# that measured floor
LOGITS_MEET_2E2 = {"trainer_stage1_nc1", "trainer_stage1_nc2_config1", "trainer_stage3_nc1", "trainer_stage3_nc1_b16"}


@pytest.mark.parametrize("name", ["trainer_stage1_nc1", "trainer_stage1_nc2_config1", "trainer_stage2_nc1",
                                  "trainer_stage3_nc1", "trainer_stage4_nc1", "trainer_stage3_nc1_b16",
                                  "trainer_stage4_nc1_b16"])
def test_trainer_step_against_reference_goldens(dev, name):
    """Same seeded weights + inputs as the fixtures produced by the unmodified reference
    (tools/make_goldens.py; uncertainty_guided_trainer.py:143-184): stages 1-4 at small batches, BASELINE.json
    configs[0]'s shape, and configs[1] at its real batch size 16 (plus stage 3 at batch 16)."""
    import parity_table as P

    r = P.measure_trainer(name, dev)
    model, logits, A, final, base, ugl, fx = r.pop("_objects")
    print({k: v for k, v in r.items() if k != "grads"})
    if name in LOGITS_MEET_2E2:
        assert r["logits_rel"] <= NORTH_STAR_REL, r["logits_rel"]                     # north_star's bar, met
    else:
        # the bf16-storage emulation is already beyond 2e-2 here; the kernels must not be further than it
        assert r["logits_rel"] <= 1.2 * r["logits_floor"] + 2e-3, (r["logits_rel"], r["logits_floor"])
    # loss within 1e-3 (north_star) - or the emulation's own distance where that is already ~1e-3
    assert r["final_loss_rel"] <= max(NORTH_STAR_LOSS, 1.5 * r["final_loss_floor_rel"]), r
    assert r["base_loss_rel"] <= max(NORTH_STAR_LOSS, 1.5 * r["base_loss_floor_rel"]), r
    # masks: bit-exact on every pixel whose reference logit is not within bf16 noise of the threshold; the
    # excluded band is |z_ref| <= max(5 % of max|z_ref|, 6 sigma of the emulation's rms logit error) - its
    # pixel fraction is printed above and tabulated in DESIGN.md; overall flips no more frequent than the
    # emulation's
    assert r["mask_mismatch_safe"] == 0
    assert r["mask_mismatch"] <= 1.5 * r["mask_mismatch_floor"] + 1e-3
    if A is not None:
        assert r["uncertainty_max_abs"] < 5e-2
        assert float(A.min()) >= 0.0 and float(A.max()) <= 1.0
    assert r["bn_worst_rel"] < 5e-3
    sd_after = model.state_dict()
    for k, v in fx["bn"].items():
        if k.endswith("num_batches_tracked"):
            assert int(sd_after[k]) == int(v)                      # integer state: exact
    _check_grad_rows(r["grads"], name)
    # conv bias in front of train-mode BN: analytically zero gradient (rounding noise in the reference)
    for k, p in model.named_parameters():
        if k.endswith(("conv_op.0.bias", "conv_op.3.bias")):
            scale = model.get_parameter(k.replace(".bias", ".weight")).grad.abs().max().item()
            assert p.grad.abs().max().item() <= 2e-2 * scale + 1e-6, k
    if r["nc"] != 1:
        return
    # fused metrics against the reference's python-side metrics
    import ug_pg_unet_b200 as U
    tr = U.UncertaintyGuidedProgressiveTrainer.__new__(U.UncertaintyGuidedProgressiveTrainer)
    f, b, dice, acc, um, us = tr._metrics_from_device(ugl.last_metrics, logits.numel(), logits.shape[0], A is not None)
    assert abs(acc - fx["accuracy"]) <= r["mask_mismatch"] + 1e-6
    assert abs(dice - fx["dice"]) <= 4 * r["mask_mismatch"] + 1e-3
    assert abs(um - fx["uncertainty_weight_mean"]) < 5e-3 and abs(us - fx["uncertainty_weight_std"]) < 5e-3


@pytest.mark.parametrize("name", ["herlev_stage1_k7", "herlev_stage2_k7", "herlev_stage3_k7_b16", "herlev_stage4_k7_b16_256",
                                  "herlev_stage4_k7_b16_224"])
def test_herlev_step_against_reference_goldens(dev, name):
    """BASELINE.json configs[3] (7 classes, stage 4 at 256x256, batch 16), the reference's own stage-4 size
    224 and stage 3, against fixtures from the unmodified Herlev/train_herlev.py:216-296."""
    import parity_table as P

    r = P.measure_herlev(name, dev)
    tr, model, metrics, fx, x, y = r.pop("_objects")
    print({k: v for k, v in r.items() if k != "grads"})
    assert r["logits_rel"] <= NORTH_STAR_REL, r["logits_rel"]          # north_star's bar (measured ~3e-3)
    assert r["final_loss_rel"] <= NORTH_STAR_LOSS and r["base_loss_rel"] <= NORTH_STAR_LOSS, r
    # predicted classes: identical, except for a sample whose reference top-2 margin is inside the logit noise
    # (herlev_stage4_k7_b16_256 has two samples with margins of 2e-5 and 3e-5 on logits of rms 3e-2)
    assert r["argmax_flips"] <= 2 and r["argmax_flip_max_margin"] <= r["argmax_margin_band"], r
    assert abs(r["correct"] - r["correct_ref"]) <= r["argmax_flips"]
    assert r["unc_mean_abs"] < 5e-3 and r["unc_std_abs"] < 5e-3
    assert r["no_grad_ok"] and r["bn_worst_rel"] < 5e-3
    _check_grad_rows(r["grads"], name)
    if r["B"] > 4:
        return
    # feature extractor exposed the way Herlev/train_herlev.py:83-102 returns it, and one optimiser step
    stage = r["stage"]
    feats = model._extract_features(x)
    assert feats.shape == (fx["B"], 512, 32, 32) and feats.dtype == torch.float32
    tr.setup_optimizer_scheduler(stage)
    w0 = model.classifier[9].weight.detach().clone()
    m5 = tr.train_step(x, y, stage)
    assert m5.shape == (5,) and not torch.equal(w0, model.classifier[9].weight.detach())


def test_eval_mode_and_no_grad_inference(dev):
    import ug_oracle as O
    import ug_pg_unet_b200 as U

    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 3, 128, 128, generator=g)
    sd = O.seeded_state_dict(3, 3, 2, seed=9)
    model = U.PGUNet3(3, 2).to(dev)
    model.load_state_dict(sd)
    model.eval()
    with torch.no_grad():
        out = model(x.to(dev))
        out2 = model(x.to(dev))
    ref = O.pgunet_forward(sd, x, 3, training=False)
    with O.bf16_storage(True):
        emu = O.pgunet_forward(sd, x, 3, training=False)
    assert out.shape == ref.shape and out.dtype == torch.float32
    assert rel(out, ref) <= 1.5 * rel(emu, ref) + 2e-3
    assert torch.equal(out, out2)                                   # deterministic
    # running statistics untouched in eval mode
    assert int(model.state_dict()["inc.conv.conv_op.1.num_batches_tracked"]) == 0
    # backward through eval-mode BN is refused loudly rather than silently wrong
    xg = x.to(dev)
    with pytest.raises(RuntimeError, match="eval-mode"):
        model(xg).sum().backward()


def test_blocks_are_drop_in_for_fp32_nchw_tensors(dev):
    """DoubleConv / Down / Up / OutConv called the way Herlev/train_herlev.py:83-102 calls them."""
    import torch.nn.functional as F
    import ug_pg_unet_b200 as U

    torch.manual_seed(0)
    inc, down, up, outc = U.InConv(3, 64).to(dev), U.Down(64, 128).to(dev), U.Up(192, 64).to(dev), U.OutConv(64, 7).to(dev)
    x = torch.randn(2, 3, 64, 64, device=dev)
    x1 = inc(x)
    x2 = down(x1)
    x3 = up(x2, x1)
    y = outc(x3)
    assert x1.dtype == torch.float32 and x1.shape == (2, 64, 64, 64) and x1.is_contiguous()
    assert x2.shape == (2, 128, 32, 32) and x3.shape == (2, 64, 64, 64) and y.shape == (2, 7, 64, 64)
    # against stock torch modules sharing the same parameters
    def ref_dc(dc, t):
        return dc.conv_op(t)
    r1 = ref_dc(inc.conv, x)
    for m in (inc, down, up):
        for bn in [mm for mm in m.modules() if isinstance(mm, torch.nn.BatchNorm2d)]:
            bn.reset_running_stats()
    r2 = ref_dc(down.mpconv[1], F.max_pool2d(r1, 2))
    r3 = ref_dc(up.conv, torch.cat([r1, F.interpolate(r2, scale_factor=2, mode="bilinear", align_corners=True)], 1))
    ry = outc.conv(r3)
    assert rel(x1, r1) < 2e-2 and rel(x2, r2) < 3e-2 and rel(x3, r3) < 4e-2 and rel(y, ry) < 5e-2
    y.square().mean().backward()
    assert inc.conv.conv_op[0].weight.grad is not None and torch.isfinite(inc.conv.conv_op[0].weight.grad).all()


def test_loss_properties(dev):
    """final(alpha) = base + alpha*mean(l*A); alpha = 0 or stage 1 -> final == base
    (demo_uncertainty_guided.py:155-201, UG_unet.py:80-82)."""
    import ug_pg_unet_b200 as U

    g = torch.Generator(device="cuda").manual_seed(0)
    z = torch.randn(4, 1, 64, 64, device=dev, generator=g) * 2
    y = torch.randint(0, 2, (4, 1, 64, 64), device=dev, generator=g).float()
    A = torch.rand(4, 1, 64, 64, device=dev, generator=g)
    ul = U.UncertaintyGuidedLoss(dev)
    crit = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor([5.0], device=dev), reduction="none")
    f = [ul.apply_uncertainty_weighted_loss(crit, z, y, A, a)[0].item() for a in (0.0, 1.0, 2.0)]
    f_none, b_none = ul.apply_uncertainty_weighted_loss(crit, z, y, None, 1.0)
    assert abs(f[0] - b_none) < 1e-6 and abs(f_none.item() - b_none) < 1e-6
    assert abs((f[2] - f[1]) - (f[1] - f[0])) < 1e-5
    with pytest.raises(ValueError):
        ul.apply_uncertainty_weighted_loss(crit, z, y[:, :, :32], None, 1.0)


def test_cuda_graph_step_equals_eager_step(dev):
    import contextlib
    import io
    import ug_pg_unet_b200 as U

    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 3, 64, 64, generator=g).to(dev)
    y = torch.randint(0, 2, (2, 1, 64, 64), generator=g).float().to(dev)
    results = []
    for use_graph in (False, True):
        torch.manual_seed(0)
        with contextlib.redirect_stdout(io.StringIO()):
            tr = U.UncertaintyGuidedProgressiveTrainer(3, 1, device=dev, use_cuda_graph=use_graph, log_interval=0)
        tr.current_stage, tr.current_model = 2, tr.models[2]
        # a tiny learning rate keeps the two runs on the same trajectory (RMSprop's first steps are
        # sign-like, which would amplify fp32-atomic ordering noise into a different trajectory)
        tr.stage_configs[2]["lr"] = 1e-7
        tr.setup_optimizer(2)
        tr.current_model.train()
        tr.models[1].eval()
        ms = []
        for _ in range(6):                      # 3 eager warm-ups, capture, replays
            m, _, _ = tr.train_step(x, y, 2)
            ms.append(m.clone())
        results.append((ms, {k: v.clone() for k, v in tr.current_model.state_dict().items()}, tr))
    for a, b in zip(results[0][0], results[1][0]):
        assert rel(a[:2], b[:2]) < 5e-3
    sd_e, sd_g = results[0][1], results[1][1]
    assert int(sd_g["inc.conv.conv_op.1.num_batches_tracked"]) == 6      # BN buffers advance inside replays
    assert rel(sd_g["up2.conv.conv_op.4.running_var"], sd_e["up2.conv.conv_op.4.running_var"]) < 5e-3
    # a replay contains the optimizer step, which autograd's version counters do not see: the bf16
    # operand caches must be invalidated so the next eager use (e.g. validation) sees the new weights
    from ug_pg_unet_b200 import ops
    dc = results[1][2].models[2].up2.conv
    wf_now, _ = ops.pack_weights(dc.conv_op[3].weight.detach())
    assert torch.equal(dc._packs[1].get()[0], wf_now)
    # and the weights really moved during the replays (optimizer inside the graph)
    w4 = dc.conv_op[3].weight.detach().clone()
    results[1][2].train_step(x, y, 2)
    assert not torch.equal(w4, dc.conv_op[3].weight.detach())


# --------------------------------------------------------------------------- full-size properties
def test_full_size_properties_stage4_batch16(dev):
    """BASELINE.json configs[1] shape (batch 16, 256x256): the oracle is too slow here, so check
    size-independent properties — BatchNorm output statistics, pooling/ReLU invariants, conv
    linearity, finite gradients for every parameter."""
    import ug_pg_unet_b200 as U
    from ug_pg_unet_b200 import ops

    torch.manual_seed(0)
    model = U.PGUNet4(3, 1).to(dev)
    model.train()
    g = torch.Generator().manual_seed(11)
    x = torch.randn(16, 3, 256, 256, generator=g).to(dev)
    y = torch.randint(0, 2, (16, 1, 256, 256), generator=g).float().to(dev)
    feats = model.encode(x)
    a = feats[0].float()                         # post BN+ReLU activations of `inc`: [B,H,W,C]
    assert a.min().item() >= 0.0
    # BN(train) with default gamma=1, beta=0 then ReLU: per-channel mean of relu(N(0,1)) ~ 0.3989
    ch_mean = a.mean(dim=(0, 1, 2))
    assert (ch_mean - 0.3989).abs().max().item() < 0.05
    # pooled tensor feeding down1 equals the max of each 2x2 window of the skip tensor (bit-exact)
    pooled = ops.maxpool2x2(feats[0])
    ref_pool = torch.nn.functional.max_pool2d(feats[0].permute(0, 3, 1, 2).float(), 2).permute(0, 2, 3, 1)
    assert torch.equal(pooled.float(), ref_pool)
    logits = model(x)
    assert logits.shape == (16, 1, 256, 256) and torch.isfinite(logits).all()
    ul = U.UncertaintyGuidedLoss(dev)
    crit = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor([5.0], device=dev), reduction="none")
    loss, base = ul.apply_uncertainty_weighted_loss(crit, logits, y, None, 1.0)
    loss.backward()
    for k, p in model.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), k
    # conv linearity at full size: conv(x1 + x2) == conv(x1) + conv(x2) (no bias), to bf16 rounding
    gg = torch.Generator(device="cuda").manual_seed(2)
    x1 = torch.randn(16, 256, 256, 64, device=dev, generator=gg).bfloat16()
    x2 = torch.randn(16, 256, 256, 64, device=dev, generator=gg).bfloat16()
    w = torch.randn(64, 64, 3, 3, device=dev, generator=gg) / 24
    wf, _ = ops.pack_weights(w)
    s = (x1.float() + x2.float()).bfloat16()
    lhs = ops.conv_gemm(s, None, wf, 9, 64).float()
    rhs = ops.conv_gemm(x1, None, wf, 9, 64).float() + ops.conv_gemm(x2, None, wf, 9, 64).float()
    assert rel(lhs, rhs) < 8e-3


def test_fused_rmsprop_matches_torch_rmsprop(dev):
    """One-launch RMSprop vs torch.optim.RMSprop with the reference's configuration
    (uncertainty_guided_trainer.py:84-88), including state_dict interchange."""
    from ug_pg_unet_b200.optim import FusedRMSprop

    torch.manual_seed(0)
    shapes = [(64, 3, 3, 3), (64,), (128, 64, 3, 3), (1, 64, 1, 1), (1,), (7,)]
    pa = [torch.randn(s, device=dev, requires_grad=True) for s in shapes]
    pb = [p.detach().clone().requires_grad_(True) for p in pa]
    oa = FusedRMSprop(pa, lr=1e-4, weight_decay=1e-4)
    ob = torch.optim.RMSprop(pb, lr=1e-4, weight_decay=1e-4)
    assert isinstance(oa, torch.optim.RMSprop)
    for step in range(4):
        for x, y in zip(pa, pb):
            g = torch.randn_like(x)
            x.grad, y.grad = g.clone(), g.clone()
        oa.step()
        ob.step()
    for x, y in zip(pa, pb):
        torch.testing.assert_close(x, y, rtol=1e-6, atol=1e-7)
    sa, sb = oa.state_dict(), ob.state_dict()
    assert sa["param_groups"][0]["lr"] == sb["param_groups"][0]["lr"]
    for k in sb["state"]:
        torch.testing.assert_close(sa["state"][k]["square_avg"], sb["state"][k]["square_avg"], rtol=1e-5, atol=1e-8)
        assert float(sa["state"][k]["step"]) == float(sb["state"][k]["step"]) == 4.0
    ob.load_state_dict(sa)          # checkpoints interchange


def test_rmsprop_conv_packed_equals_unpack_rmsprop_pack(dev):
    """The one-launch optimiser for deferred conv gradients (packed gradient in, fp32 OIHW master +
    bf16 operand packs out) is bit-identical to unpack_wgrad -> rmsprop_multi -> pack_weights."""
    from ug_pg_unet_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(7)
    shapes = [(64, 64), (128, 64), (64, 128), (256, 512), (512, 256), (64, 1024)]
    W = [torch.randn(co, ci, 3, 3, device=dev, generator=g) * 0.05 for co, ci in shapes]
    V = [torch.rand(co, ci, 3, 3, device=dev, generator=g) * 1e-3 for co, ci in shapes]
    D = [torch.randn(9 * ci * co, device=dev, generator=g) for co, ci in shapes]
    hp = dict(lr=1e-3, alpha=0.99, eps=1e-8, weight_decay=1e-4)
    # unfused reference path
    W1, V1 = [w.clone() for w in W], [v.clone() for v in V]
    G1 = [ops.unpack_wgrad(d.clone(), co, ci) for d, (co, ci) in zip(D, shapes)]
    ops.rmsprop_multi(W1, G1, V1, hp["lr"], hp["alpha"], hp["eps"], hp["weight_decay"], 0.5)
    P1 = [ops.pack_weights(w) for w in W1]
    # fused path; the last tensor has no input-gradient pack
    W2, V2, D2 = [w.clone() for w in W], [v.clone() for v in V], [d.clone() for d in D]
    WF = [torch.empty(9, co, ci, dtype=torch.bfloat16, device=dev) for co, ci in shapes]
    WD = [torch.empty(9, ci, co, dtype=torch.bfloat16, device=dev) for co, ci in shapes]
    WD[-1] = None
    ops.rmsprop_conv_packed(W2, V2, D2, WF, WD, hp["lr"], hp["alpha"], hp["eps"], hp["weight_decay"], 0.5)
    for i in range(len(shapes)):
        assert torch.equal(W1[i], W2[i]) and torch.equal(V1[i], V2[i]), shapes[i]
        assert torch.equal(P1[i][0], WF[i]), shapes[i]
        if WD[i] is not None:
            assert torch.equal(P1[i][1], WD[i]), shapes[i]
        assert not D2[i].any()                                   # accumulator left zeroed
    assert not torch.equal(W[0], W2[0])


def test_deferred_packed_gradients_train_like_materialised_gradients(dev):
    """Trainer with packed (deferred) conv gradients + CUDA graph vs eager steps with OIHW `.grad`
    and the multi-tensor optimiser: same loss trajectory at the reference's learning rate, and the
    loss really falls on a repeated batch (a replay that kept stale bf16 weights would not move)."""
    import contextlib
    import io
    import ug_pg_unet_b200 as U
    from ug_pg_unet_b200 import engine

    g = torch.Generator().manual_seed(5)
    x = torch.randn(4, 3, 64, 64, generator=g).to(dev)
    y = torch.randint(0, 2, (4, 1, 64, 64), generator=g).float().to(dev)
    runs = []
    for deferred, use_graph in ((False, False), (True, False), (True, True)):
        torch.manual_seed(0)
        with contextlib.redirect_stdout(io.StringIO()):
            tr = U.UncertaintyGuidedProgressiveTrainer(3, 1, device=dev, use_cuda_graph=use_graph, log_interval=0)
        tr.defer_conv_grads = deferred
        tr.current_stage, tr.current_model = 2, tr.models[2]
        tr.setup_optimizer(2)
        assert any(pk.defer for pk in engine.conv_packs(tr.models[2])) == deferred
        assert not any(pk.defer for pk in engine.conv_packs(tr.models[1]))
        tr.current_model.train()
        tr.models[1].eval()
        losses = []
        for _ in range(8):
            m, _, _ = tr.train_step(x, y, 2)
            losses.append(float(m[1]) / y.numel())
        w = tr.models[2].up2.conv.conv_op[3].weight
        if deferred:
            assert w.grad is None                   # never materialised
            assert not tr._grad_arena.any()         # consumed and zeroed by the optimiser kernel
        else:
            assert w.grad is not None
        # state_dict of the optimiser keeps torch.optim.RMSprop's layout either way
        st = tr.optimizer.state_dict()["state"]
        assert len(st) == len(list(tr.models[2].parameters()))
        assert all(float(s["step"]) == 8.0 for s in st.values())
        runs.append((losses, w.detach().clone()))
    base = runs[0][0]
    assert base[-1] < 0.9 * base[0], base          # the repeated batch is being fitted
    # RMSprop's first steps are sign-like (|dw| = 10 lr), so fp32 reduction-order noise in near-zero
    # gradient elements makes the trajectories drift apart slowly: tight early, looser later
    for losses, _ in runs[1:]:
        for i, (a, b) in enumerate(zip(losses, base)):
            assert abs(a - b) <= (1.5e-2 if i < 3 else 0.15) * abs(b), (i, losses, base)
        assert losses[-1] < 0.9 * losses[0], losses
    for _, w in runs[1:]:
        assert rel(w, runs[0][1]) < 0.1


def test_fused_gap_mlp_head_against_stock_modules(dev):
    """ugp_gap_mlp_head_fwd / ugp_mlp_head_bwd (one launch each) against the reference's classifier
    nn.Sequential (Herlev/train_herlev.py:66-77) in stock torch: dropout inactive, and with explicit
    dropout uniforms against the same masks applied by hand."""
    import torch.nn.functional as F
    from ug_pg_unet_b200 import engine
    import ug_pg_unet_b200.herlev as H

    torch.manual_seed(3)
    for B, hw, K in ((16, 32, 7), (5, 28, 2), (64, 8, 7)):
        m = H.HerlevClassificationModel(stage=1, num_classes=K).to(dev)
        fc1, fc2, fc3 = m.classifier[3], m.classifier[6], m.classifier[9]
        x = (torch.randn(B, hw, hw, 512, device=dev) * 0.5 + 0.3).bfloat16().requires_grad_(True)
        for use_u in (False, True):
            u = torch.rand(B, 512 + 512 + 256, device=dev) if use_u else None
            ps = (0.5, 0.3, 0.2) if use_u else (0.0, 0.0, 0.0)
            for t in (x, *m.classifier.parameters()):
                t.grad = None
            out = engine.GapMlpHeadFn.apply(x, fc1.weight, fc1.bias, fc2.weight, fc2.bias, fc3.weight, fc3.bias, u, *ps)
            gy = torch.randn_like(out)
            (out * gy).sum().backward()
            got = [out.detach().clone(), x.grad.float().clone()] + [p.grad.clone() for p in m.classifier.parameters()]
            # stock: fp32 mean over pixels, then the Linear stack with the same masks
            for t in (x, *m.classifier.parameters()):
                t.grad = None
            xf = x.detach().float().requires_grad_(True)
            h = xf.mean(dim=(1, 2))
            def drop(t, lo, p):
                if u is None:
                    return t
                return t * (u[:, lo:lo + t.shape[1]] >= p).float() / (1.0 - p)
            h = F.relu(fc1(drop(h, 0, 0.5)))
            h = F.relu(fc2(drop(h, 512, 0.3)))
            ref = fc3(drop(h, 1024, 0.2))
            (ref * gy).sum().backward()
            want = [ref.detach(), xf.grad] + [p.grad for p in m.classifier.parameters()]
            assert rel(got[0], want[0]) < 1e-5, (B, use_u, rel(got[0], want[0]))
            assert rel(got[1], want[1]) < 6e-3                         # dx is stored in bf16
            for a, b in zip(got[2:], want[2:]):
                assert rel(a, b) < 1e-4, (B, use_u, a.shape, rel(a, b))
        # the module's forward routes through the fused kernel and is repeatable in eval mode
        m.eval()
        with torch.no_grad():
            img = torch.randn(2, 3, 32, 32, device=dev)
            assert torch.equal(m(img), m(img))
            os.environ["UGP_NO_FUSED_HEAD"] = "1"
            try:
                stock = m(img)
            finally:
                del os.environ["UGP_NO_FUSED_HEAD"]
            assert rel(m(img), stock) < 1e-5


@pytest.mark.parametrize("shape", [(2, 16, 16, 64), (1, 7, 5, 8), (1, 1, 1, 64), (1, 28, 28, 256), (1, 3, 112, 256),
                                   (2, 128, 128, 64), (1, 2, 300, 128)])
def test_separable_upsample_kernels_against_torch(dev, shape):
    """Bilinear x2 (align_corners=True, UG_unet_parts.py:76-79) and its adjoint: vertical / horizontal passes
    through shared memory, channel slabs when a row does not fit; odd, tiny, non-square and multi-slab sizes."""
    import torch.nn.functional as F
    from ug_pg_unet_b200 import ops

    B, h, w, C = shape
    g = torch.Generator(device="cuda").manual_seed(h * 1000 + w)
    x = torch.randn(B, h, w, C, device=dev, generator=g).bfloat16()
    up = ops.upsample2x(x)
    xr = x.permute(0, 3, 1, 2).float().requires_grad_(True)
    ref = F.interpolate(xr, scale_factor=2, mode="bilinear", align_corners=True)
    assert up.shape == (B, 2 * h, 2 * w, C)
    assert rel(up.permute(0, 3, 1, 2), ref) < 4e-3
    dU = torch.randn(B, 2 * h, 2 * w, C, device=dev, generator=g).bfloat16()
    ref.backward(dU.permute(0, 3, 1, 2).float())
    dx = ops.upsample2x_bwd(dU)
    assert rel(dx.permute(0, 3, 1, 2), xr.grad) < 4e-3
    # deterministic (no atomics) and exact on constants: a constant image stays constant
    assert torch.equal(ops.upsample2x_bwd(dU), dx)
    c = torch.full((B, h, w, C), 1.5, device=dev).bfloat16()
    assert torch.equal(ops.upsample2x(c), torch.full((B, 2 * h, 2 * w, C), 1.5, device=dev).bfloat16())


@pytest.mark.parametrize("shape", [(2, 16, 16, 64), (1, 32, 32, 512), (3, 64, 64, 128), (1, 6, 10, 64)])
def test_batchnorm_backward_kernels_against_torch(dev, shape):
    """The software-pipelined BatchNorm/ReLU(/max-pool) backward (plain and pooled variants, ragged grid
    tails) against autograd of F.batch_norm + relu + max_pool2d on the same bf16 tensors."""
    import torch.nn.functional as F
    from ug_pg_unet_b200 import ops

    B, H, W, C = shape
    g = torch.Generator(device="cuda").manual_seed(C + H)
    z = (torch.randn(B, H, W, C, device=dev, generator=g) * 1.3 + 0.2).bfloat16()
    gamma = torch.rand(C, device=dev, generator=g) + 0.5
    beta = torch.randn(C, device=dev, generator=g) * 0.2
    for pooled in (False, True):
        zr = z.permute(0, 3, 1, 2).float().requires_grad_(True)
        gm, bt = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
        a = F.relu(F.batch_norm(zr, None, None, gm, bt, training=True, eps=1e-5))
        dA = torch.randn(B, H, W, C, device=dev, generator=g).bfloat16()
        loss = (a * dA.permute(0, 3, 1, 2).float()).sum()
        dP = None
        if pooled:
            dP = torch.randn(B, H // 2, W // 2, C, device=dev, generator=g).bfloat16()
            # pool the bf16-rounded activation, as the forward kernel does (arg-max ties resolve identically)
            ar = a.bfloat16().float()
            loss = loss + (F.max_pool2d(a + (ar - a).detach(), 2) * dP.permute(0, 3, 1, 2).float()).sum()
        loss.backward()
        zf = z.float().reshape(-1, C)
        mean, var = zf.mean(0), zf.var(0, unbiased=False)
        rstd = (var + 1e-5).rsqrt()
        scale, shift = gamma * rstd, beta - mean * gamma * rstd
        small = torch.empty(3, C, device=dev)
        dz = ops.bn_relu_bwd(z, dA, dP, scale, shift, mean, rstd, gamma, small[:2], small[2])
        assert rel(dz.permute(0, 3, 1, 2), zr.grad) < 8e-3, (shape, pooled, rel(dz.permute(0, 3, 1, 2), zr.grad))
        assert rel(small[0], gm.grad) < 2e-3 and rel(small[1], bt.grad) < 2e-3, (shape, pooled)


@pytest.mark.parametrize("shape", [(2, 16, 16, 64, 1), (1, 8, 32, 128, 2), (3, 32, 32, 256, 1), (1, 5, 3, 64, 2)])
def test_fused_decoder_head_kernels_against_unfused(dev, shape):
    """ugp_bn_relu_apply_head / ugp_bn_relu_bwd_head (the OutConv of a decoder block riding in the block's
    BN/ReLU pass and BatchNorm backward) against the separate kernels: bn_relu_apply + outconv_fwd, and
    outconv_bwd -> bn_relu_bwd."""
    from ug_pg_unet_b200 import ops

    B, H, W, C, nc = shape
    g = torch.Generator(device="cuda").manual_seed(C + nc)
    z = (torch.randn(B, H, W, C, device=dev, generator=g) * 1.2 + 0.1).bfloat16()
    gamma = torch.rand(C, device=dev, generator=g) + 0.5
    zf = z.float().reshape(-1, C)
    mean, var = zf.mean(0), zf.var(0, unbiased=False)
    rstd = (var + 1e-5).rsqrt()
    scale, shift = gamma * rstd, torch.randn(C, device=dev, generator=g) * 0.1 - mean * gamma * rstd
    wh = torch.randn(nc, C, device=dev, generator=g) / 8
    bh = torch.randn(nc, device=dev, generator=g)
    a_f, lmap = ops.bn_relu_apply_head(z, scale, shift, wh, bh)
    a_u = ops.bn_relu_apply(z, scale, shift)
    assert torch.equal(a_f, a_u)
    want = torch.einsum("bhwc,kc->bkhw", torch.relu(z.float() * scale + shift), wh) + bh.view(1, nc, 1, 1)
    assert rel(lmap, want) < 1e-5
    assert rel(lmap, ops.outconv_fwd(a_u, wh, bh)) < 4e-3          # the unfused head reads the bf16 activation
    # backward
    dl = torch.randn(B, nc, H, W, device=dev, generator=g)
    for with_dA in (True, False):
        dA = torch.randn(B, H, W, C, device=dev, generator=g).bfloat16() if with_dA else None
        small = torch.empty(3, C, device=dev)
        dz, hg = ops.bn_relu_bwd_head(z, dA, scale, shift, mean, rstd, gamma, small[:2], small[2], dl, wh)
        # reference: materialise the head's input gradient in fp32, add, then the plain BN backward in torch
        gin = torch.einsum("bkhw,kc->bhwc", dl, wh) + (dA.float() if with_dA else 0.0)
        act = z.float() * scale + shift
        gm = torch.where(act > 0, gin, torch.zeros_like(gin)).reshape(-1, C)
        xhat = (zf - mean) * rstd
        dgamma, dbeta = (gm * xhat).sum(0), gm.sum(0)
        n = gm.shape[0]
        dz_ref = (gamma * rstd) * (gm - dbeta / n - xhat * dgamma / n)
        assert rel(dz, dz_ref.view(B, H, W, C)) < 6e-3, (shape, with_dA, rel(dz, dz_ref.view(B, H, W, C)))
        assert rel(small[0], dgamma) < 1e-3 and rel(small[1], dbeta) < 1e-3
        dwh = torch.einsum("bkhw,bhwc->kc", dl, torch.relu(act))
        assert rel(hg[: nc * C].view(nc, C), dwh) < 1e-4 and rel(hg[nc * C:], dl.sum((0, 2, 3))) < 1e-4


def test_fused_and_unfused_decoder_heads_train_alike(dev):
    """PGUNet3 forward + backward with the heads fused into the decoder blocks vs UGP_NO_FUSED_OUTCONV=1."""
    import ug_oracle as O
    import ug_pg_unet_b200 as U

    g = torch.Generator().manual_seed(17)
    x = torch.randn(2, 3, 128, 128, generator=g).to(dev)
    y = torch.randint(0, 2, (2, 1, 128, 128), generator=g).float().to(dev)
    sd = O.seeded_state_dict(3, 3, 1, seed=3)
    outs = []
    for env in ("0", "1"):
        os.environ["UGP_NO_FUSED_OUTCONV"] = env
        try:
            m = U.PGUNet3(3, 1).to(dev)
            m.load_state_dict(sd)
            m.train()
            logits = m(x)
            torch.nn.functional.binary_cross_entropy_with_logits(logits, y).backward()
            outs.append((logits.detach(), {k: p.grad.clone() for k, p in m.named_parameters()}))
        finally:
            del os.environ["UGP_NO_FUSED_OUTCONV"]
    assert rel(outs[0][0], outs[1][0]) < 5e-3
    for k in outs[0][1]:
        if k.endswith(("conv_op.0.bias", "conv_op.3.bias")):
            continue
        # same kernels below the heads; the two paths differ by the bf16 rounding of the head's input /
        # input gradient, which the deep layers amplify like any other bf16 perturbation
        assert rel(outs[0][1][k], outs[1][1][k]) < (2e-2 if k.startswith("outc") else 0.5), k


def test_no_kernel_writes_outside_its_output_buffers(dev):
    """Guard bands around every buffer the op wrappers hand to a kernel (tools/sanitize.py --guards): one
    small launch of every kernel family - per-tap / pair / halo convolutions, both weight-gradient kernels,
    transposed convolution, element-wise, reductions, optimisers, the fused heads - and no byte of a 4 KiB
    margin below or above an output may change.  (compute-sanitizer is closed on the GPU pool.)"""
    import subprocess
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sanitize.py"), "--guards"], capture_output=True,
                       text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0 and "sanitize.py --guards: OK" in r.stdout


def test_joint_flips_of_loader_bytes_are_bit_exact(dev):
    """ugp_flip_u8 against torch.flip (= TF.hflip / TF.vflip of monuseg_dataset.py:150-158) per sample."""
    from ug_pg_unet_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(4)
    img = torch.randint(0, 256, (8, 37, 53, 3), device=dev, dtype=torch.uint8, generator=g)
    msk = torch.randint(0, 2, (8, 1, 37, 53), device=dev, dtype=torch.uint8, generator=g)
    flags = torch.arange(8, device=dev, dtype=torch.uint8) % 4
    fi = ops.flip_u8(img, flags)
    fm = ops.flip_u8(msk.reshape(8, 37, 53, 1).contiguous(), flags).reshape(msk.shape)
    for b in range(8):
        dims_i = [d for d, bit in ((1, 1), (0, 2)) if int(flags[b]) & bit]          # [H, W, C]: W = dim 1, H = dim 0
        assert torch.equal(fi[b], torch.flip(img[b], dims_i) if dims_i else img[b])
        dims_m = [d for d, bit in ((2, 1), (1, 2)) if int(flags[b]) & bit]          # [1, H, W]
        assert torch.equal(fm[b], torch.flip(msk[b], dims_m) if dims_m else msk[b])
    a, m2, fl = ops.joint_flip_u8(img, msk, generator=g)
    assert a.shape == img.shape and m2.shape == msk.shape and fl.shape == (8,)


@pytest.mark.parametrize("shape", [(2, 16, 16, 64, 64), (2, 16, 16, 256, 256), (1, 32, 32, 128, 128), (1, 4, 128, 64, 64),
                                   (1, 4, 128, 128, 128), (2, 64, 64, 256, 256), (16, 128, 128, 64, 64)])
def test_dgrad_with_fused_batchnorm_backward_sums(dev, shape):
    """ugp_conv_dgrad_bnstats_bf16 + ugp_bn_relu_bwd_apply against ugp_conv_gemm_bf16 + ugp_bn_relu_bwd (the
    reduction pass as its own kernel), for the per-tap, pair and halo kernels: the input gradient bit for
    bit, the sums to fp32 summation order, dz / dgamma / dbeta / dbias alike."""
    from ug_pg_unet_b200 import ops

    B, H, W, c, cin = shape          # dz has c channels, the gradient (and z of the BN below) cin channels
    g = torch.Generator(device="cuda").manual_seed(sum(shape))
    dz2 = torch.randn(B, H, W, c, device=dev, generator=g).bfloat16()
    w = torch.randn(c, cin, 3, 3, device=dev, generator=g) / (3.0 * c ** 0.5)
    _, wd = ops.pack_weights(w)
    z1 = (torch.randn(B, H, W, cin, device=dev, generator=g) * 1.1 + 0.2).bfloat16()
    gamma = torch.rand(cin, device=dev, generator=g) + 0.5
    zf = z1.float().reshape(-1, cin)
    mean, var = zf.mean(0), zf.var(0, unbiased=False)
    rstd = (var + 1e-5).rsqrt()
    scale, shift = gamma * rstd, torch.randn(cin, device=dev, generator=g) * 0.1 - mean * gamma * rstd
    dA_ref = ops.conv_gemm(dz2, None, wd, 9, cin)
    small_r = torch.empty(3, cin, device=dev)
    dz_ref = ops.bn_relu_bwd(z1, dA_ref, None, scale, shift, mean, rstd, gamma, small_r[:2], small_r[2])
    dA, raw = ops.conv_dgrad_bnstats(dz2, wd, cin, z1, scale, shift)
    assert torch.equal(dA, dA_ref)
    m = (z1.float() * scale + shift) > 0
    gm = torch.where(m, dA_ref.float(), torch.zeros((), device=dev)).reshape(-1, cin)
    assert rel(raw[0], gm.sum(0)) < 1e-4 and rel(raw[1], (gm * zf).sum(0)) < 1e-4
    small = torch.empty(3, cin, device=dev)
    dz = ops.bn_relu_bwd_apply(z1, dA, raw, scale, shift, mean, rstd, gamma, small[:2], small[2])
    assert rel(dz, dz_ref) < 2e-3
    torch.testing.assert_close(small[0], small_r[0], rtol=2e-3, atol=1e-3 * float(small_r[0].abs().max()))
    torch.testing.assert_close(small[1], small_r[1], rtol=2e-3, atol=1e-3 * float(small_r[1].abs().max()))
    torch.testing.assert_close(small[2], small_r[2], rtol=2e-2, atol=2e-3 * float(small_r[1].abs().max()) + 1e-3)


def test_custom_criterion_runs_one_forward_per_step_and_graph_cache_keys(dev):
    """A base_criterion the fused loss kernel does not cover runs the reference's own arithmetic - ONE
    train-mode forward per step (BatchNorm counters advance once per step, reference :203-234) - and the
    step-graph cache keeps one captured graph per (shape, alpha, criterion) instead of evicting."""
    import contextlib
    import io
    import ug_pg_unet_b200 as U

    g = torch.Generator().manual_seed(2)
    batches = [(torch.randn(2, 3, 64, 64, generator=g), torch.randint(0, 2, (2, 1, 64, 64), generator=g).float())
               for _ in range(3)]
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        tr = U.UncertaintyGuidedProgressiveTrainer(3, 1, device=dev, log_interval=0)
        tr.current_stage, tr.current_model = 2, tr.models[2]
        tr.setup_optimizer(2)
        tr.base_criterion = lambda z, t: torch.nn.functional.binary_cross_entropy_with_logits(z, t, reduction="none")
        res = tr.train_epoch(batches, 2)
    assert all(math.isfinite(v) for v in res) and 0.0 <= res[2] <= 1.0 and 0.0 <= res[3] <= 1.0
    assert int(tr.models[2].inc.conv.conv_op[1].num_batches_tracked) == 3          # not 6
    # fused criterion again: graphs keyed by everything a capture bakes in
    with contextlib.redirect_stdout(io.StringIO()):
        tr.base_criterion = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor([5.0], device=dev), reduction="none")
        x, y = batches[0][0].to(dev), batches[0][1].to(dev)
        for _ in range(5):
            tr.train_step(x, y, 2)
        assert len(tr._graphs) == 1 and next(iter(tr._graphs.values()))["graph"] is not None
        tr.uncertainty_alpha = 0.5                       # baked into the capture as a float: a new key
        for _ in range(5):
            tr.train_step(x, y, 2)
        assert len(tr._graphs) == 2
        m_a = tr.train_step(x, y, 2)[0]
        tr.uncertainty_alpha = 1.0                       # back to the first graph, still captured
        k = len(tr._graphs)
        m_b = tr.train_step(x, y, 2)[0]
        assert len(tr._graphs) == k and not torch.equal(m_a[:1], m_b[:1])
        xr, yr = x[:1].contiguous(), y[:1].contiguous()  # ragged batch: own entry, nothing evicted
        tr.train_step(xr, yr, 2)
        assert len(tr._graphs) == 3 and sum(g_["graph"] is not None for g_ in tr._graphs.values()) == 2
        tr.close()
        assert not tr._graphs


def test_global_avg_pool_kernels(dev):
    from ug_pg_unet_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(3, 16, 16, 128, device=dev, generator=g).bfloat16()
    out = ops.global_avg_pool(x)
    torch.testing.assert_close(out, x.float().mean(dim=(1, 2)), rtol=1e-5, atol=1e-6)
    dy = torch.randn(3, 128, device=dev, generator=g)
    dx = ops.global_avg_pool_bwd(dy, 16, 16)
    want = (dy / 256.0).bfloat16()[:, None, None, :].expand(3, 16, 16, 128)
    assert torch.equal(dx, want)


# --------------------------------------------------------------------------- BASELINE.json configs as test cases
def test_config2_full_progressive_pipeline(dev, tmp_path):
    """configs[2]: stages 1 -> 4 (32 -> 64 -> 128 -> 256) with weight transfer between stages, cross-stage
    uncertainty weighting, validation and checkpoints — `train_progressive` end to end (1 epoch of 2
    batches per stage, batch 2)."""
    import contextlib
    import io
    import ug_pg_unet_b200 as U

    g = torch.Generator().manual_seed(3)
    batches = [(torch.randn(2, 3, 256, 256, generator=g), torch.randint(0, 2, (2, 1, 256, 256), generator=g).float())
               for _ in range(2)]

    class Loader(list):
        pass

    torch.manual_seed(0)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        tr = U.UncertaintyGuidedProgressiveTrainer(3, 1, device=dev, log_interval=1)
        for s in tr.stage_configs.values():
            s["epochs_per_stage"] = 1
        before = {s: {k: v.clone() for k, v in tr.models[s].state_dict().items()} for s in (1, 2, 3, 4)}
        tr.train_progressive(Loader(batches), Loader(batches[:1]), max_stages=4, save_dir=str(tmp_path))
    log = buf.getvalue()
    assert "Transferring weights from stage 3 to stage 4" in log and "Unc_mean" in log
    assert len(tr.history["train_loss"]) == 4 and all(math.isfinite(v) for v in tr.history["train_loss"])
    assert tr.history["stage_transitions"] == [0, 1, 2, 3]
    assert tr.history["uncertainty_weights_mean"][0] == 0.0 and tr.history["uncertainty_weights_mean"][3] > 0.0
    for s in (1, 2, 3, 4):       # every stage trained: weights moved, BN counters advanced by its 2 steps
        sd = tr.models[s].state_dict()
        assert not torch.equal(sd["up1.conv.conv_op.0.weight"], before[s]["up1.conv.conv_op.0.weight"])
        # the scalar BN step counter transfers with the weights (same key, same shape), so it accumulates
        assert int(sd["inc.conv.conv_op.1.num_batches_tracked"]) == 2 * s
    # checkpoints: reference layout, loadable into a fresh model
    ck = torch.load(tmp_path / "ug_pgunet_stage4_best.pth", map_location="cpu", weights_only=False)
    assert {"stage", "epoch", "model_state_dict", "optimizer_state_dict", "val_dice", "history"} <= set(ck)
    m = U.PGUNet4(3, 1)
    m.load_state_dict(ck["model_state_dict"], strict=True)
    assert len(ck["optimizer_state_dict"]["state"]) == len(list(m.parameters()))


def test_config3_herlev_stage4_256_seven_classes(dev):
    """configs[3]: Herlev-shaped 7-class classification, stage 4 at 256x256 with the stage-3 model supplying
    the entropy weights; two optimiser steps, then validation."""
    import contextlib
    import io
    import ug_pg_unet_b200.herlev as H

    cfg = {"device": dev, "epochs_per_stage": 1, "num_classes": 7, "uncertainty_alpha": 1.0, "log_interval": 0,
           "stage4_resolution": 256}
    torch.manual_seed(1)
    with contextlib.redirect_stdout(io.StringIO()):
        tr = H.HerlevTrainer(cfg)
        tr.setup_optimizer_scheduler(4)
        tr.transfer_weights(3, 4)
    assert tr.stage_configs[4]["resolution"] == 256
    g = torch.Generator().manual_seed(9)
    x = torch.randn(4, 3, 256, 256, generator=g)
    y = torch.randint(0, 7, (4,), generator=g)
    with contextlib.redirect_stdout(io.StringIO()):
        loss, base, acc, um, us = tr.train_epoch([(x, y), (x, y)], 4)
        vl, vb, va, vum, vus = tr.validate_epoch([(x, y)], 4)
    assert all(math.isfinite(v) for v in (loss, base, vl, vb)) and 0.0 <= acc <= 100.0
    assert 1.0 <= um <= 2.0 and 1.0 <= vum <= 2.0          # weights 1 + alpha*u with u in [0, 1]
    m = tr.models[4]
    assert m.unet.down3.mpconv[1].conv_op[0].weight.grad is not None
    assert m.unet.down4.mpconv[1].conv_op[0].weight.grad is None and m.unet.up1.conv.conv_op[0].weight.grad is None
    assert int(m.unet.inc.conv.conv_op[1].num_batches_tracked) == 2


def test_config5_batch64_step(dev):
    """configs[4] per-GPU shape: batch 64 at stage 4, 256x256 — eager warm-ups, graph capture, replays."""
    import contextlib
    import io
    import ug_pg_unet_b200 as U

    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        tr = U.UncertaintyGuidedProgressiveTrainer(3, 1, device=dev, log_interval=0)
    tr.current_stage, tr.current_model = 4, tr.models[4]
    tr.setup_optimizer(4)
    tr.current_model.train()
    tr.models[3].eval()
    g = torch.Generator().manual_seed(64)
    x = torch.randn(64, 3, 256, 256, generator=g).to(dev)
    y = (torch.rand(64, 1, 256, 256, generator=g) < 0.2).float().to(dev)      # sparse positives, SURVEY §8d
    # RMSprop's first (sign-like) updates overshoot: the loss goes 2.68 -> 4.78 -> 3.37 -> ... -> 2.67 (step
    # 6) -> 2.48 (step 10) in every mode (eager / graph / single stream), so the fit is judged at step 10
    ms = [tr.train_step(x, y, 4)[0] for _ in range(10)]
    torch.cuda.synchronize()
    f = [tr._metrics_from_device(m, y.numel(), 64, True) for m in ms]
    assert all(math.isfinite(v[0]) and math.isfinite(v[1]) for v in f)
    assert f[-1][0] < f[0][0] - 0.1                # the repeated batch is being fitted
    assert 0.0 < f[-1][4] < 1.0                    # uncertainty mean
    assert torch.cuda.max_memory_allocated() < 40e9


# --------------------------------------------------------------------------- ConvTranspose2d / legacy UNet
@pytest.mark.parametrize("shape", [(2, 16, 16, 128, 64), (1, 32, 32, 256, 128), (1, 16, 16, 1024, 512),
                                   (1, 128, 128, 128, 64), (3, 64, 64, 512, 256)])
def test_conv_transpose2x2_kernels_against_torch(dev, shape):
    """ugp_conv_transpose2x2_{fwd,dgrad,wgrad}, the tconv (un)pack kernels and ugp_channel_sum vs
    F.conv_transpose2d + autograd on identical bf16-rounded inputs (bf16 output rounding: rel <= 6e-3;
    the fp32 weight / bias gradients: 2e-3)."""
    import torch.nn.functional as F
    from ug_pg_unet_b200 import ops

    B, H, W, cin, cout = shape
    g = torch.Generator(device="cuda").manual_seed(cin + H)
    x = torch.randn(B, H, W, cin, device="cuda", generator=g).bfloat16()
    w = (torch.randn(cin, cout, 2, 2, device="cuda", generator=g) / math.sqrt(cin)).bfloat16().float()
    b = torch.randn(cout, device="cuda", generator=g)
    dout = torch.randn(B, 2 * H, 2 * W, cout, device="cuda", generator=g).bfloat16()
    wf, wd = ops.pack_weights_tconv(w)
    # packs are exact re-layouts of the bf16-representable weights
    assert torch.equal(wf.float().view(2, 2, cout, cin), w.permute(2, 3, 1, 0))
    assert torch.equal(wd.float().view(cin, 2, 2, cout), w.permute(0, 2, 3, 1))
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    ref = F.conv_transpose2d(xr, wr, br, stride=2)
    ref.backward(dout.float().permute(0, 3, 1, 2))
    out = ops.conv_transpose2x2_fwd(x, wf, cout, bias=b)
    assert out.shape == (B, 2 * H, 2 * W, cout)
    assert rel(out.permute(0, 3, 1, 2), ref) < 6e-3
    dx = ops.conv_transpose2x2_dgrad(dout, wd, cin)
    assert rel(dx.permute(0, 3, 1, 2), xr.grad) < 6e-3
    dwp = torch.zeros(4 * cin * cout, device="cuda")
    ops.conv_transpose2x2_wgrad(x, dout, dwp)
    dw = ops.unpack_wgrad_tconv(dwp, cin, cout)
    assert rel(dw, wr.grad) < 2e-3
    assert float(dwp.abs().max()) == 0.0                      # consumed and re-zeroed
    ops.conv_transpose2x2_wgrad(x, dout, dwp)
    dw2 = ops.unpack_wgrad_tconv(dwp, cin, cout, grad=dw.clone(), accumulate=True)
    assert rel(dw2, 2 * wr.grad) < 2e-3
    db = ops.channel_sum(dout)
    assert rel(db, br.grad) < 2e-3


def test_upsample_block_against_reference_golden(dev, golden_dir):
    """UpSample (ConvTranspose2d -> cat([up, skip]) -> DoubleConv) called with fp32 NCHW tensors like the
    reference's legacy UNet calls it, against the fixture produced by the unmodified UG_unet_parts.UpSample."""
    import ug_oracle as O
    import ug_pg_unet_b200 as U

    p = torch.load(os.path.join(golden_dir, "legacy_unet.pt"))["prim"]
    sd_up = {k[len("up_convolution_4."):]: v for k, v in O.seeded_legacy_state_dict(3, 1, seed=5).items()
             if k.startswith("up_convolution_4.")}
    blk = U.UpSample(128, 64).to(dev)
    blk.load_state_dict(sd_up, strict=True)
    blk.train()
    x1 = p["x1"].to(dev).requires_grad_(True)
    x2 = p["x2"].to(dev).requires_grad_(True)
    y = blk(x1, x2)
    assert y.dtype == torch.float32 and tuple(y.shape) == tuple(p["out"].shape)
    (y * p["gy"].to(dev)).sum().backward()
    # bf16-storage floor of the oracle for the same block
    with O.bf16_storage(True):
        e1 = p["x1"].clone().requires_grad_(True)
        e2 = p["x2"].clone().requires_grad_(True)
        ew = {k: v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v
              for k, v in sd_up.items()}
        t = O.conv_transpose2x2(e1, ew["up.weight"], ew["up.bias"])
        ey = O.double_conv(torch.cat([t, e2], 1), ew, "conv", True)
        (ey * p["gy"]).sum().backward()
    assert rel(y, p["out"]) <= 1.5 * rel(ey, p["out"]) + 2e-3
    assert rel(x1.grad, p["dx1"]) <= 1.5 * rel(e1.grad, p["dx1"]) + 2e-2
    assert rel(x2.grad, p["dx2"]) <= 1.5 * rel(e2.grad, p["dx2"]) + 2e-2
    for k, prm in blk.named_parameters():
        if k.endswith(("conv_op.0.bias", "conv_op.3.bias")):
            continue                                            # analytically zero in front of train-mode BN
        assert rel(prm.grad, p["grads"][k]) <= 1.5 * rel(ew[k].grad, p["grads"][k]) + 2e-2, k


def test_legacy_unet_against_reference_golden(dev, golden_dir):
    """The legacy UNet (UG_unet.py:430-462: DownSample x4, 512->1024 bottleneck, ConvTranspose2d up path)
    against the fixture produced by the unmodified reference, same seeded weights and input."""
    import torch.nn.functional as F
    import ug_oracle as O
    import ug_pg_unet_b200 as U

    fx = torch.load(os.path.join(golden_dir, "legacy_unet.pt"))
    sd = O.seeded_legacy_state_dict(3, 1, seed=6)
    net = U.UNet(3, 1).to(dev)
    net.load_state_dict(sd, strict=True)
    assert list(net.state_dict().keys()) == list(O.legacy_unet_shapes(3, 1).keys()) or \
        set(net.state_dict().keys()) == set(O.legacy_unet_shapes(3, 1).keys())
    net.train()
    x, y = fx["x"].to(dev), fx["y"].float().to(dev)
    logits = net(x)
    loss = F.binary_cross_entropy_with_logits(logits, y)
    loss.backward()
    torch.cuda.synchronize()
    esd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v)
           for k, v in sd.items()}
    with O.bf16_storage(True):
        elog = O.legacy_unet_forward(esd, fx["x"], True)
        eloss = F.binary_cross_entropy_with_logits(elog, fx["y"].float())
        eloss.backward()
    floor = rel(elog, fx["logits"])
    assert rel(logits, fx["logits"]) <= 1.5 * floor + 2e-3, (rel(logits, fx["logits"]), floor)
    assert abs(loss.item() - fx["loss"]) <= max(1e-3 * fx["loss"], 2.5 * abs(eloss.item() - fx["loss"])) + 2e-4
    zr = fx["logits"]
    safe = zr.abs() > 0.05 * zr.abs().max()
    assert torch.equal((logits.cpu() > 0)[safe], (zr > 0)[safe])

    def sample_rel(t, s):
        f = t.detach().flatten().cpu().double()
        if "full" in s:
            r = s["full"].double()
            return ((f - r).norm() / (r.norm() + 1e-30)).item()
        r = s["vals"].double()
        return ((f[s["idx"]] - r).norm() / (r.norm() + 1e-30)).item()

    for k, prm in net.named_parameters():
        if k.endswith(("conv_op.0.bias", "conv_op.3.bias")):
            continue
        fl = sample_rel(esd[k].grad, fx["grads"][k])
        got = sample_rel(prm.grad, fx["grads"][k])
        assert got <= 1.5 * fl + 2e-2, (k, got, fl)
    net.eval()
    with torch.no_grad():
        ev = net(x)
    # eval-mode logits use the running statistics updated by ONE train step from the seeded buffers
    esd2 = {k: v.detach() for k, v in sd.items()}
    bn = {}
    O.legacy_unet_forward(esd2, fx["x"], True, bn)
    esd2.update(bn)
    with O.bf16_storage(True):
        eev = O.legacy_unet_forward(esd2, fx["x"], False)
    assert rel(ev, fx["eval_logits"]) <= 1.5 * rel(eev, fx["eval_logits"]) + 5e-3


# --------------------------------------------------------------------------- uint8 input path / evaluators
def test_byte_input_packer_matches_float_packer(dev):
    """ugp_im2col_first_u8 (uint8 HWC bytes, /255 + Normalize + resize folded in) against the fp32 NCHW
    packer fed with the tensor the reference's loaders would build (ToTensor, Normalize)."""
    from ug_pg_unet_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(5)
    img = torch.randint(0, 256, (3, 96, 80, 3), device="cuda", generator=g, dtype=torch.uint8)
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    for m, s in ((None, None), (mean, std)):
        x = img.permute(0, 3, 1, 2).float() / 255.0
        if m is not None:
            x = (x - torch.tensor(m, device="cuda").view(1, 3, 1, 1)) / torch.tensor(s, device="cuda").view(1, 3, 1, 1)
        for (H, W) in ((96, 80), (64, 64)):
            a = ops.im2col_first_u8(img, H, W, m, s).float()
            b = ops.im2col_first(x.contiguous(), H, W).float()
            # same arithmetic up to one fp32 rounding before the bf16 store: at most 1 bf16 ulp, rarely
            assert (a - b).abs().max().item() <= 2.0 ** -7 * b.abs().max().item()
            assert (a != b).float().mean().item() < 0.02
            assert torch.equal(a[..., 27:], torch.zeros_like(a[..., 27:]))


def test_byte_targets_and_byte_images_train_like_float_tensors(dev):
    """The fused loss with uint8 masks is bit-identical to the float-mask kernel, and a model fed uint8
    HWC bytes produces the logits / gradients of the same model fed the loader's float tensor."""
    import ug_oracle as O
    import ug_pg_unet_b200 as U
    from ug_pg_unet_b200 import ops

    g = torch.Generator().manual_seed(9)
    z = (torch.randn(2, 1, 64, 64, generator=g) * 2).to(dev)
    t8 = torch.randint(0, 2, (2, 1, 64, 64), generator=g, dtype=torch.uint8).to(dev)
    A = torch.rand(2, 1, 64, 64, generator=g).to(dev)
    o1, d1 = ops.bce_uncertainty_loss(z, t8.float(), A, 1.0, 5.0, True, 1.0 / z.numel())
    o2, d2 = ops.bce_uncertainty_loss(z, t8, A, 1.0, 5.0, True, 1.0 / z.numel())
    assert torch.equal(d1, d2) and torch.allclose(o1, o2, rtol=1e-12, atol=0)

    img = torch.randint(0, 256, (2, 64, 64, 3), generator=g, dtype=torch.uint8).to(dev)
    model = U.PGUNet2(3, 1).to(dev)
    model.load_state_dict(O.seeded_state_dict(2, 3, 1, seed=8))
    model.train()
    outs = []
    for x in (img, (img.permute(0, 3, 1, 2).float() / 255.0).contiguous()):
        model.zero_grad(set_to_none=True)
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.reset_running_stats()
        logits = model(x)
        ugl = U.UncertaintyGuidedLoss(dev)
        crit = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor([5.0], device=dev), reduction="none")
        loss, _ = ugl.apply_uncertainty_weighted_loss(crit, logits, t8 if x.dtype == torch.uint8 else t8.float(), None, 1.0)
        loss.backward()
        outs.append((logits.detach().clone(), loss.item(), model.up2.conv.conv_op[3].weight.grad.clone()))
    assert rel(outs[0][0], outs[1][0]) < 2e-2 and abs(outs[0][1] - outs[1][1]) < 2e-3 * abs(outs[1][1])
    assert rel(outs[0][2], outs[1][2]) < 1e-1
    # resize folded into the packer: 64 x 64 bytes -> the 32 x 32 stage-1 model
    prev = U.PGUNet1(3, 1).to(dev)
    prev.load_state_dict(O.seeded_state_dict(1, 3, 1, seed=1))
    A8 = U.UncertaintyGuidedLoss(dev).generate_uncertainty_map(img, prev, 32, 64)
    Af = U.UncertaintyGuidedLoss(dev).generate_uncertainty_map((img.permute(0, 3, 1, 2).float() / 255.0).contiguous(),
                                                               prev, 32, 64)
    assert (A8 - Af).abs().max().item() < 5e-2 and rel(A8, Af) < 2e-2


def test_trainer_step_accepts_loader_bytes(dev):
    """UncertaintyGuidedProgressiveTrainer.train_step with uint8 HWC images + uint8 masks (eager and
    CUDA-graph replays) tracks the same trainer fed float tensors."""
    import ug_oracle as O
    import ug_pg_unet_b200 as U

    g = torch.Generator().manual_seed(12)
    img = torch.randint(0, 256, (2, 64, 64, 3), generator=g, dtype=torch.uint8).to(dev)
    y8 = torch.randint(0, 2, (2, 1, 64, 64), generator=g, dtype=torch.uint8).to(dev)
    losses = []
    for as_bytes in (True, False):
        tr = U.UncertaintyGuidedProgressiveTrainer(3, 1, device=dev, uncertainty_alpha=1.0)
        tr.models[1].load_state_dict(O.seeded_state_dict(1, 3, 1, seed=1))
        tr.models[2].load_state_dict(O.seeded_state_dict(2, 3, 1, seed=2))
        tr.current_stage, tr.current_model = 2, tr.models[2]
        tr.setup_optimizer(2)
        tr.current_model.train()
        tr.models[1].eval()
        x = img if as_bytes else (img.permute(0, 3, 1, 2).float() / 255.0).contiguous()
        y = y8 if as_bytes else y8.float()
        ls = []
        for _ in range(6):                       # 3 eager warm-ups, capture, 2 replays
            m, out, unc = tr.train_step(x, y, 2)
            ls.append(tr._metrics_from_device(m, out.numel(), out.shape[0], unc is not None)[0])
        losses.append(ls)
    # same trajectory: identical first step, then within the drift RMSprop's sign-like first updates
    # give two bf16 runs (tools/check_noise.py)
    assert abs(losses[0][0] - losses[1][0]) <= 1e-3 * abs(losses[1][0])
    for a, b in zip(*losses):
        assert abs(a - b) <= 5e-2 * abs(b) + 1e-3, (losses)
    assert losses[0][-1] < losses[0][0]          # it trains


def test_segmentation_evaluator_against_reference_golden(dev, golden_dir):
    """MoNuSegEvaluator mirror (uint8 bytes and float tensors) vs the fixture produced by the unmodified
    MoNuSegImprove/test_monuseg.py arithmetic: masks bit-exact away from the decision boundary, metrics
    within the effect of the flipped pixels."""
    import ug_oracle as O
    import ug_pg_unet_b200 as U

    fx = torch.load(os.path.join(golden_dir, "evaluator.pt"))
    model = U.PGUNet2(3, 1).to(dev)
    model.load_state_dict(O.seeded_state_dict(2, 3, 1, seed=8))
    ev = U.MoNuSegEvaluator(model, device=dev)
    zr = fx["logits"]
    safe = zr.abs() > 0.05 * zr.abs().max()
    for images in (fx["images"], (fx["images"].permute(0, 3, 1, 2).float() / 255.0).contiguous()):
        mask, conf = ev.predict_batch(images)
        assert mask.dtype == torch.uint8 and tuple(mask.shape) == tuple(fx["masks"].shape)
        assert torch.equal(mask.cpu()[safe], fx["masks"][safe])
        flips = (mask.cpu() != fx["masks"]).float().flatten(1).mean(1)
        for i in range(4):
            assert abs(conf[i].item() - fx["confidence"][i]) < 2e-2
        avg, std = ev.evaluate_batches([(images[:2], fx["gts"][:2]), (images[2:], fx["gts"][2:].float())])
        for k in fx["avg"]:
            assert abs(avg[k] - fx["avg"][k]) <= 4 * flips.max().item() + 1e-3, (k, avg[k], fx["avg"][k])
            assert abs(std[k] - fx["std"][k]) <= 4 * flips.max().item() + 1e-3
    # the counting kernel itself is exact: feed it the reference's logits
    from ug_pg_unet_b200 import ops
    counts, m2 = ops.seg_confusion(zr.to(dev).contiguous(), fx["gts"].to(dev).view(4, -1), want_mask=True)
    assert torch.equal(m2.cpu(), fx["masks"])
    got = U.metrics_from_counts(counts[:, 0], counts[:, 1], counts[:, 2], 64.0 * 64.0)
    for i, m in enumerate(fx["per_image"]):
        for k, v in m.items():
            assert abs(got[k][i].item() - v) < 1e-6
    up, _ = ev.predict_batch(fx["images"], out_size=(100, 90))
    assert tuple(up.shape[-2:]) == (100, 90)


def test_classification_evaluator_kernels(dev):
    """ugp_cls_eval vs torch softmax / argmax / a host confusion matrix; HerlevEvaluator on uint8 bytes
    equals the same model on the ImageNet-normalised float tensor."""
    import ug_pg_unet_b200 as U
    from ug_pg_unet_b200 import ops

    g = torch.Generator().manual_seed(3)
    z = (torch.randn(37, 7, generator=g) * 3).to(dev)
    y = torch.randint(0, 7, (37,), generator=g).to(dev)
    probs, preds, cm = ops.cls_eval(z, y)
    assert torch.equal(preds, z.argmax(1))
    assert torch.allclose(probs, torch.softmax(z, 1), rtol=1e-5, atol=1e-7)
    ref = torch.zeros(7, 7, dtype=torch.int64)
    for a, b in zip(y.tolist(), z.argmax(1).tolist()):
        ref[a, b] += 1
    assert torch.equal(cm.cpu(), ref)
    _, _, cm = ops.cls_eval(z, y, cm)
    assert torch.equal(cm.cpu(), 2 * ref)

    torch.manual_seed(0)
    model = U.HerlevClassificationModel(stage=2, num_classes=7).to(dev)
    ev = U.HerlevEvaluator(model, device=dev)
    img = torch.randint(0, 256, (5, 64, 64, 3), generator=g, dtype=torch.uint8)
    x = img.permute(0, 3, 1, 2).float() / 255.0
    x = (x - torch.tensor(U.evaluate.IMAGENET_MEAN).view(1, 3, 1, 1)) / torch.tensor(U.evaluate.IMAGENET_STD).view(1, 3, 1, 1)
    p8, pr8, c8 = ev.predict_batch(img)
    pf, prf, cf = ev.predict_batch(x.contiguous())
    assert rel(pr8, prf) < 2e-2 and pr8.shape == (5, 7) and torch.allclose(pr8.sum(1), torch.ones(5, device=dev), atol=1e-5)
    labels = torch.randint(0, 7, (5,), generator=g)
    res = ev.evaluate_batches([(img[:3], labels[:3]), (img[3:], labels[3:])])
    assert res["total_samples"] == 5 and int(res["confusion_matrix"].sum()) == 5
    assert abs(res["accuracy"] - (p8.cpu() == labels).double().mean().item()) < 1e-9


# --------------------------------------------------------------------------- general image sizes
@pytest.mark.parametrize("shape", [(2, 56, 56, 64, 0, 128), (1, 28, 28, 128, 64, 256), (1, 112, 112, 64, 0, 64),
                                   (2, 24, 40, 64, 0, 64), (1, 224, 224, 64, 0, 64), (1, 14, 14, 256, 0, 512),
                                   (3, 12, 20, 128, 128, 128), (1, 48, 48, 64, 0, 256)])
def test_convolutions_at_general_sizes_against_torch(dev, shape):
    """Forward (+ BN statistics), input gradient and weight gradient of the 3x3 convolution at sizes that
    do not tile into 128-pixel runs (the reference's Herlev stage 4 runs 224 -> 112 -> 56 -> 28; H != W):
    boxes with BW*BH <= 128 pixels, masked accumulator rows, 4-D stores."""
    import torch.nn.functional as F
    from ug_pg_unet_b200 import ops

    B, H, W, c0, c1, cout = shape
    cin = c0 + c1
    g = torch.Generator(device="cuda").manual_seed(H * W + cout)
    x0 = torch.randn(B, H, W, c0, device="cuda", generator=g).bfloat16()
    x1 = torch.randn(B, H, W, c1, device="cuda", generator=g).bfloat16() if c1 else None
    w = (torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (3.0 * math.sqrt(cin))).bfloat16().float()
    bias = torch.randn(cout, device="cuda", generator=g)
    dz = torch.randn(B, H, W, cout, device="cuda", generator=g).bfloat16()
    wf, wd = ops.pack_weights(w)
    xin = x0.float().permute(0, 3, 1, 2) if x1 is None else torch.cat([x0, x1], 3).float().permute(0, 3, 1, 2)
    xin = xin.contiguous().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    ref = F.conv2d(xin, wr, bias, padding=1)
    ref.backward(dz.float().permute(0, 3, 1, 2))
    stats = ops.new_stats(cout, "cuda")
    out = ops.conv_gemm(x0, x1, wf, 9, cout, bias=bias, stats=stats)
    assert rel(out.permute(0, 3, 1, 2), ref) < 6e-3
    o = out.float().reshape(-1, cout)
    s = stats.sum(0)
    assert rel(s[0], o.sum(0)) < 1e-4 and rel(s[1], (o * o).sum(0)) < 1e-4      # masked rows contribute nothing
    if c1:
        d0, d1 = ops.conv_gemm(dz, None, wd, 9, cin, split_n=c0)
        dx = torch.cat([d0, d1], 3)
    else:
        dx = ops.conv_gemm(dz, None, wd, 9, cin)
    assert rel(dx.permute(0, 3, 1, 2), xin.grad) < 6e-3
    dwp = torch.zeros(9 * cin * cout, device="cuda")
    ops.conv_wgrad(x0, x1, dz, 9, dwp)
    assert rel(ops.unpack_wgrad(dwp, cout, cin), wr.grad) < 2e-3
    # eval-mode epilogue (affine + ReLU) on the same tiles
    sc = torch.rand(cout, device="cuda", generator=g) + 0.5
    sh = torch.randn(cout, device="cuda", generator=g) * 0.2
    oe = ops.conv_gemm(x0, x1, wf, 9, cout, bias=bias, scale=sc, shift=sh, relu=True)
    assert rel(oe.permute(0, 3, 1, 2), F.relu(ref.detach() * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1))) < 6e-3


@pytest.mark.parametrize("size", [(224, 224), (96, 160)])
def test_pgunet_at_general_sizes_against_oracle(dev, size):
    """A whole PGUNet4 training step at 224 x 224 (14 x 14 bottleneck) and at H != W against the oracle."""
    import ug_oracle as O
    import ug_pg_unet_b200 as U

    H, W = size
    g = torch.Generator().manual_seed(H)
    x = torch.randn(1, 3, H, W, generator=g)
    y = torch.randint(0, 2, (1, 1, H, W), generator=g).float()
    sd = O.seeded_state_dict(4, 3, 1, seed=4)
    model = U.PGUNet4(3, 1).to(dev)
    model.load_state_dict(sd)
    model.train()
    logits = model(x.to(dev))
    loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, y.to(dev))
    loss.backward()

    def oracle(bf):
        esd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v)
               for k, v in sd.items()}
        with O.bf16_storage(bf):
            z = O.pgunet_forward(esd, x, 4, training=True)
            l = torch.nn.functional.binary_cross_entropy_with_logits(z, y)
            l.backward()
        return z.detach(), l.item(), esd

    zr, lr_, rsd = oracle(False)
    ze, le, esd = oracle(True)
    assert tuple(logits.shape) == (1, 1, H, W)
    assert rel(logits, zr) <= 1.5 * rel(ze, zr) + 2e-3, (rel(logits, zr), rel(ze, zr))
    assert abs(loss.item() - lr_) <= max(1e-3 * lr_, 2.5 * abs(le - lr_)) + 2e-4
    for k, prm in model.named_parameters():
        if k.endswith(("conv_op.0.bias", "conv_op.3.bias")):
            continue
        fl = rel(esd[k].grad, rsd[k].grad)
        assert rel(prm.grad, rsd[k].grad) <= 1.5 * fl + 2e-2, (k, rel(prm.grad, rsd[k].grad), fl)


def test_herlev_stage4_at_the_reference_resolution_224(dev):
    """Herlev/train_herlev.py:138 trains stage 4 at 224 x 224 (encoder maps 224 / 112 / 56 / 28)."""
    import contextlib
    import io
    import ug_pg_unet_b200.herlev as Hm

    cfg = {"device": dev, "epochs_per_stage": 1, "num_classes": 7, "uncertainty_alpha": 1.0, "log_interval": 0}
    torch.manual_seed(1)
    with contextlib.redirect_stdout(io.StringIO()):
        tr = Hm.HerlevTrainer(cfg)
        tr.setup_optimizer_scheduler(4)
        tr.transfer_weights(3, 4)
    assert tr.stage_configs[4]["resolution"] == 224
    g = torch.Generator().manual_seed(9)
    x = torch.randn(4, 3, 224, 224, generator=g)
    y = torch.randint(0, 7, (4,), generator=g)
    with contextlib.redirect_stdout(io.StringIO()):
        loss, base, acc, um, us = tr.train_epoch([(x, y), (x, y)], 4)
        vl, vb, va, vum, vus = tr.validate_epoch([(x, y)], 4)
    assert all(math.isfinite(v) for v in (loss, base, vl, vb)) and 0.0 <= acc <= 100.0
    assert tr.models[4].unet.down3.mpconv[1].conv_op[0].weight.grad is not None


def test_fused_adam_matches_torch_adam(dev):
    """FusedAdam (ugp_adam_multi, the Herlev trainer's optimiser) vs torch.optim.Adam: same updates, same
    state_dict layout, checkpoints interchangeable."""
    from ug_pg_unet_b200.optim import FusedAdam

    g = torch.Generator().manual_seed(4)
    shapes = [(512, 256, 3, 3), (64,), (7, 256), (1,), (100003,)]
    pa = [torch.randn(s, generator=g).to(dev).requires_grad_(True) for s in shapes]
    pb = [p.detach().clone().requires_grad_(True) for p in pa]
    oa = FusedAdam(pa, lr=1e-3, weight_decay=1e-4)
    ob = torch.optim.Adam(pb, lr=1e-3, weight_decay=1e-4)
    for step in range(4):
        for a, b in zip(pa, pb):
            gr = torch.randn(a.shape, generator=g).to(dev)
            a.grad, b.grad = gr.clone(), gr.clone()
        oa.step()
        ob.step()
        if step == 1:                       # checkpoint round trip through the stock optimiser's format
            import copy
            oa.load_state_dict(copy.deepcopy(ob.state_dict()))    # state_dict() hands out the live tensors
    for a, b in zip(pa, pb):
        assert rel(a, b) < 1e-6
    sa, sb = oa.state_dict(), ob.state_dict()
    assert sa["param_groups"][0]["lr"] == sb["param_groups"][0]["lr"]
    for k in sb["state"]:
        assert set(sa["state"][k]) == set(sb["state"][k]) == {"step", "exp_avg", "exp_avg_sq"}
        assert float(sa["state"][k]["step"]) == float(sb["state"][k]["step"]) == 4.0
        assert rel(sa["state"][k]["exp_avg_sq"], sb["state"][k]["exp_avg_sq"]) < 1e-6


def test_evaluator_graph_replay_equals_eager(dev):
    """MoNuSegEvaluator.predict_batch through its CUDA graph (third call on) returns what the eager passes
    return, and follows new inputs and in-place weight updates."""
    import ug_oracle as O
    import ug_pg_unet_b200 as U

    model = U.PGUNet2(3, 1).to(dev)
    model.load_state_dict(O.seeded_state_dict(2, 3, 1, seed=8))
    ev_g = U.MoNuSegEvaluator(model, device=dev, use_cuda_graph=True)
    ev_e = U.MoNuSegEvaluator(model, device=dev, use_cuda_graph=False)
    g = torch.Generator().manual_seed(2)
    for i in range(5):
        img = torch.randint(0, 256, (2, 64, 64, 3), generator=g, dtype=torch.uint8)
        if i == 4:
            model.load_state_dict(O.seeded_state_dict(2, 3, 1, seed=9))      # in place: same storages
        mg, cg = ev_g.predict_batch(img)
        me, ce = ev_e.predict_batch(img)
        assert torch.equal(mg, me) and torch.equal(cg, ce), i


def test_herlev_graph_step_tracks_eager_step(dev):
    """HerlevTrainer.train_step: CUDA-graph replays (FusedAdam with its device-side step count inside the
    graph) follow the eager trajectory, and the optimiser's state_dict reports the true step count."""
    import contextlib
    import io
    import ug_pg_unet_b200.herlev as Hm

    g = torch.Generator().manual_seed(21)
    x = torch.randn(4, 3, 64, 64, generator=g).to(dev)
    y = torch.randint(0, 7, (4,), generator=g).to(dev)
    traj = []
    for use_graph in (False, True):
        cfg = {"device": dev, "epochs_per_stage": 1, "num_classes": 7, "uncertainty_alpha": 1.0, "log_interval": 0,
               "use_cuda_graph": use_graph}
        torch.manual_seed(3)
        with contextlib.redirect_stdout(io.StringIO()):
            tr = Hm.HerlevTrainer(cfg)
            tr.setup_optimizer_scheduler(2)
        model = tr.models[2]
        model.train()
        for m in model.classifier:
            if isinstance(m, torch.nn.Dropout):
                m.eval()                                  # same trajectory in both modes
        tr.models[1].eval()
        ms = torch.stack([tr.train_step(x, y, 2) for _ in range(8)]).cpu()
        sd = tr.optimizers[2].state_dict()
        assert {float(s["step"]) for s in sd["state"].values()} == {8.0}
        traj.append((ms, model.classifier[9].weight.detach().clone()))
    (me, we), (mg, wg) = traj
    assert abs(me[0, 0] - mg[0, 0]) <= 1e-3 * abs(me[0, 0])           # identical first step
    assert rel(mg[:, 0], me[:, 0]) < 5e-2                             # bf16 / atomic-order drift afterwards
    assert rel(wg, we) < 5e-2 and mg[-1, 0] < mg[0, 0]


@pytest.mark.gpu
def test_early_optimizer_matches_the_default_order(dev):
    """UGP_EARLY_OPTIMIZER: the packed-weight RMSprop launched from inside backward (side stream, after the
    last weight gradient and the last dgrad are queued) ends at the same weights as the default order - up to
    the run-to-run floor of the atomically accumulated gradients, measured here with two default runs."""
    import contextlib
    import io
    import ug_pg_unet_b200 as U

    g = torch.Generator().manual_seed(5)
    x = torch.randn(4, 3, 64, 64, generator=g).to(dev)
    y = torch.randint(0, 2, (4, 1, 64, 64), generator=g).float().to(dev)

    def run(early):
        torch.manual_seed(0)
        with contextlib.redirect_stdout(io.StringIO()):
            tr = U.UncertaintyGuidedProgressiveTrainer(3, 1, device=dev, log_interval=0)
            tr.current_stage, tr.current_model = 2, tr.models[2]
            tr.setup_optimizer(2)
            tr.early_optimizer = early
            losses = [float(tr.train_step(x, y, 2)[0][0]) for _ in range(6)]      # eager warm-ups, capture, replays
        w = torch.cat([p.detach().float().reshape(-1) for p in tr.current_model.parameters()])
        pending = [pk.grad_pending for pk in tr._arena_packs]
        tr.close()
        return w, losses, pending

    w_a, l_a, _ = run(False)
    w_b, l_b, _ = run(False)
    w_c, l_c, _ = run(False)
    w_e, l_e, pend = run(True)
    assert not any(pend)
    # run-to-run floor of the default order: the largest distance among three runs (two runs alone can
    # land close to each other by luck; six chaotic steps of a random-init net amplify fp32-atomic noise)
    dist = lambda u, v: ((u - v).norm() / u.norm()).item()
    floor = max(dist(w_a, w_b), dist(w_a, w_c), dist(w_b, w_c))
    err = min(dist(w_e, w_a), dist(w_e, w_b), dist(w_e, w_c))
    assert err <= 3 * floor + 1e-3, (err, floor)
    drift = max(max(abs(a - b), abs(a - c), abs(b - c)) / abs(a) for a, b, c in zip(l_a, l_b, l_c))
    assert all(min(abs(a - e), abs(b - e), abs(c - e)) <= (3 * drift + 5e-3) * abs(a)
               for a, b, c, e in zip(l_a, l_b, l_c, l_e)), (l_a, l_b, l_c, l_e)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 64, 64, 64, 64), (1, 32, 32, 32, 32), (16, 256, 256, 256, 256),
                                   (3, 224, 224, 224, 224), (4, 256, 256, 128, 128), (5, 200, 120, 96, 64),
                                   (8, 16, 16, 16, 16), (2, 32, 128, 32, 128), (3, 128, 64, 128, 64)])
@pytest.mark.parametrize("u8", [False, True])
@pytest.mark.parametrize("cout", [64, 128])
def test_fused_first_layer_is_bit_identical_to_packer_plus_gemm(dev, shape, u8, cout):
    """ugp_conv_first_fwd / ugp_conv_first_wgrad (first convolution with its 64-wide rows built in shared memory,
    UG_unet_parts.py:10 + the resize of uncertainty_guided_trainer.py:208) against the path they replace -
    ugp_im2col_first[_u8] + ugp_conv_gemm_bf16 / ugp_conv_wgrad_bf16 on the packed rows: the stored bf16 output,
    the BatchNorm statistics and the eval-mode folded output bit for bit; the weight gradient up to the
    order of the fp32 atomics.  fp32 inputs at the layer's resolution with power-of-two rows take the TMA
    image-patch producers (16 .. 256 wide, H != W), everything else the per-row gathers."""
    from ug_pg_unet_b200 import ops

    B, Hin, Win, H, W = shape
    g = torch.Generator().manual_seed(B * 1000 + H)
    if u8:
        x = torch.randint(0, 256, (B, Hin, Win, 3), generator=g, dtype=torch.uint8).to(dev)
        mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    else:
        x = torch.randn(B, 3, Hin, Win, generator=g).to(dev)
        mean = std = None
    w = (torch.randn(cout, 3, 3, 3, generator=g) * 0.2).to(dev)
    bias = torch.randn(cout, generator=g).to(dev)
    wf, _ = ops.pack_weights(w, need_dgrad=False)
    rows = ops.first_rows(x, H, W, mean, std, cout=cout)
    assert isinstance(rows, ops.FirstRows), "shape must be covered by the fused kernels"
    col = rows.materialize()
    # train mode: raw output + per-CTA statistics
    st_a, st_b = ops.new_stats(cout, dev), ops.new_stats(cout, dev)
    z_a = ops.conv_gemm(rows, None, wf, 1, cout, bias=bias, stats=st_a)
    z_b = ops.conv_gemm(col, None, wf, 1, cout, bias=bias, stats=st_b)
    assert torch.equal(z_a, z_b)
    sa, sb = st_a.double().sum(0), st_b.double().sum(0)
    assert torch.allclose(sa, sb, rtol=1e-5, atol=1e-3), (sa - sb).abs().max()
    st_c = ops.new_stats(cout, dev)                       # run-to-run identical statistics (shared slots commute)
    ops.conv_gemm(rows, None, wf, 1, cout, bias=bias, stats=st_c)
    assert torch.equal(st_a, st_c)
    # eval mode: folded BatchNorm + ReLU in the epilogue
    sc, sh = torch.rand(cout, generator=g).to(dev) + 0.5, torch.randn(cout, generator=g).to(dev)
    a_a = ops.conv_gemm(rows, None, wf, 1, cout, bias=bias, scale=sc, shift=sh, relu=True)
    a_b = ops.conv_gemm(col, None, wf, 1, cout, bias=bias, scale=sc, shift=sh, relu=True)
    assert torch.equal(a_a, a_b)
    # weight gradient (cout = 128 falls back to the packed rows inside ops.conv_wgrad)
    dz = torch.randn(B, H, W, cout, generator=g).to(dev).to(torch.bfloat16)
    dw_a = torch.zeros(64 * cout, device=dev)
    dw_b = torch.zeros(64 * cout, device=dev)
    ops.conv_wgrad(rows, None, dz, 1, dw_a)
    ops.conv_wgrad(col, None, dz, 1, dw_b)
    ref = col.double().reshape(-1, 64).t() @ dz.double().reshape(-1, cout)          # [k][cout]
    err_a = (dw_a.double().view(64, cout) - ref).abs().max().item()
    err_b = (dw_b.double().view(64, cout) - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-6
    assert err_a <= max(2 * err_b, 1e-5 * scale), (err_a, err_b, scale)
    # accumulation: a second launch adds
    ops.conv_wgrad(rows, None, dz, 1, dw_a)
    assert (dw_a.double().view(64, cout) - 2 * ref).abs().max().item() <= max(4 * err_b, 2e-5 * scale)


@pytest.mark.gpu
def test_prefetch_ring_delivers_every_batch_in_order(dev):
    """trainer._prefetch: host batches copied on the copy stream into a ring of reused device buffers
    (slot reuse ordered by GPU events after the consuming step) arrive bit-exact and in order - more batches
    than slots, a slow consumer on the main stream, a ragged last batch and a dtype change."""
    import contextlib
    import io
    import ug_pg_unet_b200 as U

    with contextlib.redirect_stdout(io.StringIO()):
        tr = U.UncertaintyGuidedProgressiveTrainer(3, 1, device=dev, log_interval=0)
    g = torch.Generator().manual_seed(11)
    batches = [(torch.randn(4, 3, 32, 32, generator=g).pin_memory(), torch.randint(0, 2, (4, 1, 32, 32), generator=g).float().pin_memory())
               for _ in range(8)]
    batches.append((torch.randn(2, 3, 32, 32, generator=g), torch.randint(0, 2, (2, 1, 32, 32), generator=g).float()))   # ragged
    batches.append((torch.randint(0, 256, (4, 32, 32, 3), generator=g, dtype=torch.uint8),
                    torch.randint(0, 2, (4, 1, 32, 32), generator=g, dtype=torch.uint8)))                              # loader bytes
    batches += batches[:3]
    got = []
    busy = torch.empty(64 << 20, device=dev)
    for d, t in tr._prefetch(batches):
        assert d.device.type == "cuda" and t.device.type == "cuda"
        for _ in range(3):
            busy.add_(1.0)                       # the consumer keeps the main stream busy: copies run ahead of it
        got.append((d.clone(), t.clone()))       # what the step would read, on the main stream
    torch.cuda.synchronize()
    assert len(got) == len(batches)
    for (d, t), (hd, ht) in zip(got, batches):
        assert d.dtype == hd.dtype and torch.equal(d.cpu(), hd) and torch.equal(t.cpu(), ht)
    assert sum(s is not None for s in tr._stage_ring) == 3
    tr.close()

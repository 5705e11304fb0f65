"""Bring-up checks of the C-ABI kernels against plain torch on the GPU.

Each group runs in its own subprocess (a trapped kernel poisons the CUDA
context), results go to stdout and gpurun_out/check_ops.log.

    python tools/gpu_check_ops.py            # all groups
    python tools/gpu_check_ops.py conv_fwd   # one group, in-process
"""
from __future__ import annotations

import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

GROUPS = ["elementwise", "conv_fwd", "conv_dgrad", "conv_wgrad", "head_loss"]


def rel_err(a, b):
    import torch

    a = a.double()
    b = b.double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item(), (a - b).abs().max().item()


def report(name, ok, detail=""):
    print(f"[{'PASS' if ok else 'FAIL'}] {name} {detail}", flush=True)
    return ok


def to_nchw(x_nhwc):
    return x_nhwc.permute(0, 3, 1, 2).float()


def group_conv_fwd():
    import torch
    import torch.nn.functional as F
    from ug_pg_unet_b200 import ops

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(0)
    ok = True
    cases = [  # B, H, W, c0, c1, cout
        (2, 16, 16, 64, 0, 64),
        (1, 32, 32, 128, 0, 128),
        (2, 32, 32, 128, 64, 256),
        (1, 64, 64, 64, 0, 512),
        (1, 128, 128, 64, 0, 64),
        (1, 256, 256, 64, 64, 64),
        (3, 16, 16, 512, 512, 256),
    ]
    for (B, H, W, c0, c1, cout) in cases:
        x0 = torch.randn(B, H, W, c0, device="cuda", generator=g).bfloat16()
        x1 = torch.randn(B, H, W, c1, device="cuda", generator=g).bfloat16() if c1 else None
        cin = c0 + c1
        w = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (3.0 * cin ** 0.5)
        bias = torch.randn(cout, device="cuda", generator=g)
        wf, wd = ops.pack_weights(w)
        stats = ops.new_stats(cout, "cuda")
        out = ops.conv_gemm(x0, x1, wf, 9, cout, bias=bias, stats=stats)
        torch.cuda.synchronize()
        xin = to_nchw(x0) if x1 is None else torch.cat([to_nchw(x0), to_nchw(x1)], 1)
        ref = F.conv2d(xin, w.bfloat16().float(), bias, padding=1)
        e, m = rel_err(to_nchw(out), ref)
        o64 = out.double().reshape(-1, cout)
        tot = stats.sum(0)                      # [P][2][C] per-CTA partials -> [2][C]
        se, _ = rel_err(tot[0], o64.sum(0))
        s2e, _ = rel_err(tot[1], (o64 * o64).sum(0))
        ok &= report(f"conv_fwd B{B} {H}x{W} {c0}+{c1}->{cout}", e < 6e-3 and se < 1e-4 and s2e < 1e-4,
                     f"rel={e:.2e} max={m:.2e} stats_rel={se:.1e}/{s2e:.1e}")
    # eval-mode epilogue: affine + relu
    B, H, W, c0, cout = 2, 32, 32, 64, 128
    x0 = torch.randn(B, H, W, c0, device="cuda", generator=g).bfloat16()
    w = torch.randn(cout, c0, 3, 3, device="cuda", generator=g) / 24.0
    bias = torch.randn(cout, device="cuda", generator=g)
    sc = torch.rand(cout, device="cuda", generator=g) + 0.5
    sh = torch.randn(cout, device="cuda", generator=g)
    wf, _ = ops.pack_weights(w)
    out = ops.conv_gemm(x0, None, wf, 9, cout, bias=bias, scale=sc, shift=sh, relu=True)
    ref = F.relu(F.conv2d(to_nchw(x0), w.bfloat16().float(), bias, padding=1) * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1))
    e, m = rel_err(to_nchw(out), ref)
    ok &= report("conv_fwd affine+relu epilogue", e < 6e-3, f"rel={e:.2e} max={m:.2e}")
    # first layer through im2col (taps = 1)
    for (B, R, Rin) in [(2, 32, 32), (1, 256, 256), (2, 64, 128)]:
        x = torch.randn(B, 3, Rin, Rin, device="cuda", generator=g)
        w = torch.randn(64, 3, 3, 3, device="cuda", generator=g) / 5.0
        bias = torch.randn(64, device="cuda", generator=g)
        wf, _ = ops.pack_weights(w)
        col = ops.im2col_first(x, R, R)
        out = ops.conv_gemm(col, None, wf, 1, 64, bias=bias)
        xr = x if R == Rin else F.interpolate(x, size=(R, R), mode="bilinear", align_corners=True)
        ref = F.conv2d(xr.bfloat16().float(), w.bfloat16().float(), bias, padding=1)
        e, m = rel_err(to_nchw(out), ref)
        ok &= report(f"first conv (im2col) B{B} {Rin}->{R}", e < 8e-3, f"rel={e:.2e} max={m:.2e}")
    return ok


def group_conv_dgrad():
    import torch
    import torch.nn.functional as F
    from ug_pg_unet_b200 import ops

    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(1)
    ok = True
    for (B, H, W, cin, cout, split) in [(2, 16, 16, 64, 64, None), (1, 32, 32, 256, 128, None),
                                        (1, 64, 64, 128, 64, 64), (2, 32, 32, 1024, 256, 512),
                                        (1, 128, 128, 64, 128, None)]:
        dz = torch.randn(B, H, W, cout, device="cuda", generator=g).bfloat16()
        w = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (3.0 * cout ** 0.5)
        _, wd = ops.pack_weights(w)
        res = ops.conv_gemm(dz, None, wd, 9, cin, split_n=split)
        ref = F.conv_transpose2d(to_nchw(dz), w.bfloat16().float(), padding=1)
        if split:
            got = torch.cat([to_nchw(res[0]), to_nchw(res[1])], 1)
        else:
            got = to_nchw(res)
        e, m = rel_err(got, ref)
        ok &= report(f"conv_dgrad B{B} {H}x{W} cin{cin} cout{cout} split{split}", e < 6e-3, f"rel={e:.2e} max={m:.2e}")
    return ok


def group_conv_wgrad():
    import torch
    from ug_pg_unet_b200 import ops

    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(2)
    ok = True
    for (B, H, W, c0, c1, cout) in [(2, 16, 16, 64, 0, 64), (1, 32, 32, 128, 0, 128), (2, 64, 64, 64, 64, 64),
                                    (1, 32, 32, 256, 256, 256), (1, 128, 128, 64, 0, 128), (2, 16, 16, 512, 0, 512),
                                    (1, 256, 256, 64, 0, 64)]:
        x0 = torch.randn(B, H, W, c0, device="cuda", generator=g).bfloat16()
        x1 = torch.randn(B, H, W, c1, device="cuda", generator=g).bfloat16() if c1 else None
        dz = torch.randn(B, H, W, cout, device="cuda", generator=g).bfloat16()
        cin = c0 + c1
        dwp = torch.zeros(9, cin, cout, device="cuda")
        ops.conv_wgrad(x0, x1, dz, 9, dwp)
        grad = ops.unpack_wgrad(dwp, cout, cin)
        xin = to_nchw(x0) if x1 is None else torch.cat([to_nchw(x0), to_nchw(x1)], 1)
        ref = torch.nn.grad.conv2d_weight(xin, (cout, cin, 3, 3), to_nchw(dz), padding=1)
        e, m = rel_err(grad, ref)
        ok &= report(f"conv_wgrad B{B} {H}x{W} {c0}+{c1}->{cout}", e < 2e-3, f"rel={e:.2e} max={m:.2e}")
    # first layer (im2col'ed input, taps = 1)
    B, R = 2, 64
    x = torch.randn(B, 3, R, R, device="cuda", generator=g)
    dz = torch.randn(B, R, R, 64, device="cuda", generator=g).bfloat16()
    col = ops.im2col_first(x)
    dwp = torch.zeros(64, 64, device="cuda")
    ops.conv_wgrad(col, None, dz, 1, dwp)
    grad = ops.unpack_wgrad(dwp, 64, 3)
    ref = torch.nn.grad.conv2d_weight(x.bfloat16().float(), (64, 3, 3, 3), to_nchw(dz), padding=1)
    e, m = rel_err(grad, ref)
    ok &= report("conv_wgrad first layer", e < 2e-3, f"rel={e:.2e} max={m:.2e}")
    return ok


def group_elementwise():
    import torch
    import torch.nn.functional as F
    from ug_pg_unet_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(3)
    ok = True
    B, H, W, Cn = 2, 32, 32, 128
    z = torch.randn(B, H, W, Cn, device="cuda", generator=g).bfloat16()
    sc = torch.rand(Cn, device="cuda", generator=g) + 0.5
    sh = torch.randn(Cn, device="cuda", generator=g) * 0.3
    a, p = ops.bn_relu_apply(z, sc, sh, pool=True)
    a2 = ops.bn_relu_apply(z, sc, sh)
    ref = F.relu(z.float() * sc + sh).bfloat16()
    # fma in the kernel vs mul+add in torch: allow one bf16 ulp
    ok &= report("bn_relu_apply", torch.equal(a, a2) and (a.float() - ref.float()).abs().max().item() <= 2 ** -7 * ref.float().abs().max().item())
    refp, refi = F.max_pool2d(to_nchw(ref), 2, return_indices=True)
    ok &= report("bn_relu_apply fused pool", torch.equal(to_nchw(p), F.max_pool2d(to_nchw(a), 2)))
    pv, pi = ops.maxpool2x2(ref, return_indices=True)
    ok &= report("maxpool2x2 values", torch.equal(to_nchw(pv), refp))
    ok &= report("maxpool2x2 indices (bit-exact, first max wins)",
                 torch.equal(pi.permute(0, 3, 1, 2).long(), refi))
    dP = torch.randn(B, H // 2, W // 2, Cn, device="cuda", generator=g).bfloat16()
    xa = to_nchw(ref).requires_grad_(True)
    F.max_pool2d(xa, 2).backward(to_nchw(dP))
    dxp = ops.maxpool2x2_bwd(ref, dP)
    ok &= report("maxpool2x2_bwd", torch.equal(to_nchw(dxp), xa.grad))
    # upsample
    x = torch.randn(B, 16, 16, Cn, device="cuda", generator=g).bfloat16()
    up = ops.upsample2x(x)
    xr = to_nchw(x).requires_grad_(True)
    refu = F.interpolate(xr, scale_factor=2, mode="bilinear", align_corners=True)
    e, m = rel_err(to_nchw(up), refu.detach())
    ok &= report("upsample2x", e < 4e-3, f"rel={e:.2e} max={m:.2e}")
    dU = torch.randn(B, 32, 32, Cn, device="cuda", generator=g).bfloat16()
    refu.backward(to_nchw(dU))
    dx = ops.upsample2x_bwd(dU)
    e, m = rel_err(to_nchw(dx), xr.grad)
    ok &= report("upsample2x_bwd", e < 4e-3, f"rel={e:.2e} max={m:.2e}")
    # BN finalize
    n = B * H * W
    zz = z.double().reshape(-1, Cn)
    stats = torch.stack([zz.sum(0), (zz * zz).sum(0)]).float().unsqueeze(0).contiguous()   # one partial slot
    gamma = torch.rand(Cn, device="cuda", generator=g) + 0.5
    beta = torch.randn(Cn, device="cuda", generator=g)
    rm = torch.zeros(Cn, device="cuda")
    rv = torch.ones(Cn, device="cuda")
    nbt = torch.zeros((), dtype=torch.int64, device="cuda")
    scale = torch.empty(Cn, device="cuda"); shift = torch.empty(Cn, device="cuda")
    mean = torch.empty(Cn, device="cuda"); rstd = torch.empty(Cn, device="cuda")
    ops.bn_finalize(stats, n, gamma, beta, 1e-5, 0.1, rm, rv, nbt, scale, shift, mean, rstd)
    bn = torch.nn.BatchNorm2d(Cn).cuda()
    bn.weight.data.copy_(gamma); bn.bias.data.copy_(beta)
    zin = to_nchw(z).requires_grad_(True)
    y = F.relu(bn(zin))
    refy = y.detach()
    got = to_nchw(ops.bn_relu_apply(z, scale, shift))
    e, m = rel_err(got, refy)
    ok &= report("bn_finalize + apply vs nn.BatchNorm2d(train)", e < 4e-3, f"rel={e:.2e} max={m:.2e}")
    e1, _ = rel_err(rm, bn.running_mean); e2, _ = rel_err(rv, bn.running_var)
    ok &= report("bn running stats", e1 < 1e-4 and e2 < 1e-4 and int(nbt) == 1 and float(stats.abs().max()) == 0.0,
                 f"{e1:.1e} {e2:.1e} nbt={int(nbt)}")
    # BN + ReLU backward (full-res grad only, pooled grad only, both)
    dA = torch.randn(B, H, W, Cn, device="cuda", generator=g).bfloat16()
    for mode in ("full", "pool", "both"):
        zin.grad = None; bn.weight.grad = None; bn.bias.grad = None
        y = F.relu(bn(zin))
        loss_terms = []
        if mode in ("full", "both"):
            loss_terms.append((y * to_nchw(dA)).sum())
        if mode in ("pool", "both"):
            loss_terms.append((F.max_pool2d(y, 2) * to_nchw(dP)).sum())
        sum(loss_terms).backward()
        sums = torch.empty(2, Cn, device="cuda")
        dbias = torch.zeros(Cn, device="cuda")
        dz = ops.bn_relu_bwd(z, dA if mode != "pool" else None, dP if mode != "full" else None, scale, shift,
                             mean, rstd, gamma, sums, dbias)
        e, m = rel_err(to_nchw(dz), zin.grad)
        eg, _ = rel_err(sums[0], bn.weight.grad); eb, _ = rel_err(sums[1], bn.bias.grad)
        # the fused pool arg-max is taken on bf16-rounded activations: a handful of
        # near-ties may route differently from the fp32 reference
        tol = 4e-3 if mode == "full" else 3e-2
        ok &= report(f"bn_relu_bwd[{mode}]", e < tol and eg < tol and eb < tol,
                     f"dz rel={e:.2e} dgamma={eg:.1e} dbeta={eb:.1e} |dbias|max={dbias.abs().max().item():.1e}")
    # weight pack round trip
    w = torch.randn(128, 64, 3, 3, device="cuda", generator=g)
    wf, wd = ops.pack_weights(w)
    ok &= report("pack_weights fwd", torch.equal(wf, w.bfloat16().permute(2, 3, 0, 1).reshape(9, 128, 64)))
    ok &= report("pack_weights dgrad", torch.equal(wd, w.bfloat16().flip(2, 3).permute(2, 3, 1, 0).reshape(9, 64, 128)))
    # rmsprop
    p0 = torch.randn(100003, device="cuda", generator=g); gr = torch.randn(100003, device="cuda", generator=g)
    pt = p0.clone().requires_grad_(True); pt.grad = gr.clone()
    opt = torch.optim.RMSprop([pt], lr=1e-3, weight_decay=1e-4)
    mine = p0.clone(); v = torch.zeros_like(mine)
    for _ in range(3):
        opt.step()
        ops.rmsprop_step(mine, gr, v, 1e-3, 0.99, 1e-8, 1e-4)
    e, m = rel_err(mine, pt.detach())
    ok &= report("rmsprop_step x3", e < 1e-6, f"rel={e:.2e}")
    return ok


def group_head_loss():
    import torch
    import torch.nn.functional as F
    from ug_pg_unet_b200 import ops

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(4)
    ok = True
    B, R, nc = 2, 64, 2
    feats = [torch.randn(B, R >> s, R >> s, c, device="cuda", generator=g).bfloat16() for s, c in ((3, 256), (2, 128), (1, 64), (0, 64))]
    ws = [torch.randn(nc, f.shape[3], device="cuda", generator=g) / 8 for f in feats]
    bs = [torch.randn(nc, device="cuda", generator=g) for _ in feats]
    maps = [ops.outconv_fwd(f, w, b) for f, w, b in zip(feats, ws, bs)]
    logits = ops.head_combine_fwd(maps)
    fr = [to_nchw(f).requires_grad_(True) for f in feats]
    wr = [w.clone().requires_grad_(True) for w in ws]
    br = [b.clone().requires_grad_(True) for b in bs]
    ref = None
    for s, (f, w, b) in enumerate(zip(fr, wr, br)):
        l = F.conv2d(f, w.view(nc, -1, 1, 1), b)
        if l.shape[-1] != R:
            l = F.interpolate(l, scale_factor=R // l.shape[-1], mode="bilinear", align_corners=True)
        ref = l if ref is None else ref + l
    e, m = rel_err(logits, ref.detach())
    ok &= report("outconv + head_combine fwd", e < 1e-5, f"rel={e:.2e} max={m:.2e}")
    dlog = torch.randn(B, nc, R, R, device="cuda", generator=g)
    ref.backward(dlog)
    for s, f in enumerate(feats):
        dl = ops.head_combine_bwd(dlog, f.shape[1])
        dx = torch.empty_like(f)
        dw = torch.zeros_like(ws[s]); db = torch.zeros_like(bs[s])
        ops.outconv_bwd(f, ws[s], dl, dx, False, dw, db)
        e1, _ = rel_err(to_nchw(dx), fr[s].grad)
        e2, _ = rel_err(dw, wr[s].grad)
        e3, _ = rel_err(db, br[s].grad)
        ok &= report(f"head bwd scale {s}", e1 < 4e-3 and e2 < 1e-4 and e3 < 1e-4, f"dx={e1:.1e} dw={e2:.1e} db={e3:.1e}")
    # uncertainty map
    zp = torch.randn(B, 1, 32, 32, device="cuda", generator=g) * 3
    A = ops.uncertainty_margin(zp, 64, 64)
    P = F.interpolate(torch.sigmoid(zp), size=(64, 64), mode="bilinear", align_corners=True)
    refA = 1.0 - 2.0 * torch.abs(P - 0.5)
    ok &= report("uncertainty_margin", (A - refA).abs().max().item() < 1e-5, f"max={(A - refA).abs().max().item():.1e}")
    # fused BCE loss
    z = (torch.randn(B, 1, 64, 64, device="cuda", generator=g) * 2).requires_grad_(True)
    t = torch.randint(0, 2, (B, 1, 64, 64), device="cuda", generator=g).float()
    crit = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor([5.0], device="cuda"), reduction="none")
    pl = crit(z, t)
    final = (pl * (1.0 + 0.7 * refA)).mean()
    final.backward()
    out, dlogits = ops.bce_uncertainty_loss(z.detach(), t, refA, 0.7, 5.0, True, 1.0 / z.numel())
    N = z.numel()
    pred = (torch.sigmoid(z.detach()) > 0.5).float()
    checks = [abs(out[0].item() / N - final.item()) < 1e-5, abs(out[1].item() / N - pl.mean().item()) < 1e-5,
              abs(out[2].item() / N - refA.mean().item()) < 1e-6, out[4].item() == pred.ne(t).sum().item()]
    e, m = rel_err(dlogits, z.grad)
    inter = (pred * t).flatten(1).sum(1)
    dice_ref = ((2 * inter + 1) / (pred.flatten(1).sum(1) + t.flatten(1).sum(1) + 1)).mean().item()
    o = out[5:].view(B, 3)
    dice = ((2 * o[:, 0] + 1) / (o[:, 1] + o[:, 2] + 1)).mean().item()
    ok &= report("bce_uncertainty_loss", all(checks) and e < 1e-5 and abs(dice - dice_ref) < 1e-6, f"checks={checks} dlogits rel={e:.1e}")
    # classifier variant
    lg = (torch.randn(16, 7, device="cuda", generator=g) * 2).requires_grad_(True)
    lab = torch.randint(0, 7, (16,), device="cuda", generator=g)
    prev = torch.randn(16, 7, device="cuda", generator=g)
    u = ops.softmax_entropy(prev)
    pr = F.softmax(prev, 1)
    uref = -(pr * torch.log(pr + 1e-8)).sum(1) / torch.log(torch.tensor(7.0))
    ok &= report("softmax_entropy", (u - uref).abs().max().item() < 1e-5)
    lossr = (F.cross_entropy(lg, lab, reduction="none") * (1 + 0.5 * uref)).mean()
    lossr.backward()
    out, dl = ops.ce_weighted_loss(lg.detach(), lab, u, 0.5, True, 1.0 / 16)
    e, _ = rel_err(dl, lg.grad)
    ok &= report("ce_weighted_loss", abs(out[0].item() / 16 - lossr.item()) < 1e-5 and e < 1e-5, f"dl rel={e:.1e}")
    return ok


def main():
    if len(sys.argv) > 1:
        fn = globals()["group_" + sys.argv[1]]
        ok = fn()
        sys.exit(0 if ok else 1)
    os.makedirs("gpurun_out", exist_ok=True)
    summary = []
    with open("gpurun_out/check_ops.log", "w") as log:
        for grp in GROUPS:
            t0 = time.time()
            try:
                r = subprocess.run([sys.executable, __file__, grp], capture_output=True, text=True, timeout=600)
                out, rc = r.stdout + r.stderr[-4000:], r.returncode
            except subprocess.TimeoutExpired as e:
                out, rc = f"TIMEOUT {e}", -9
            msg = f"===== {grp}: rc={rc} ({time.time() - t0:.1f}s)\n{out}\n"
            print(msg, flush=True)
            log.write(msg)
            summary.append((grp, rc))
    print("SUMMARY", summary)
    sys.exit(0 if all(rc == 0 for _, rc in summary) else 1)


if __name__ == "__main__":
    main()

"""The reference's trainer step in STOCK PyTorch on the GPU (cuDNN / cuBLAS / ATen kernels) - the "kernel to
beat" of SURVEY.md §2.1 / BASELINE.md §5.3, timed by bench.py beside the sm_100a path (`stock_gpu` key).

The reference's Python files cannot travel to the GPU box, so the step is restated here with stock modules
only (nn.Conv2d / nn.BatchNorm2d / nn.ReLU containers exactly as UG_unet_parts.py:5-19 builds them,
F.max_pool2d, F.interpolate(bilinear, align_corners=True), torch.cat, nn.BCEWithLogitsLoss, optim.RMSprop),
following uncertainty_guided_trainer.py:143-184, 203-218 and UG_unet.py:19-94, 282-304.  Nothing of this
repository's kernels, engine or oracle is on this path.  Two modes: "tf32" (fp32 tensors, cuDNN TF32
convolutions = PyTorch's default on Ampere+) and "bf16" (torch.autocast(bfloat16) + channels_last).  The
reference's per-step `.item()` host syncs are left out (they only slow it down); one sync ends the timed region.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class _DoubleConv(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv_op = nn.Sequential(nn.Conv2d(cin, cout, 3, padding=1), nn.BatchNorm2d(cout), nn.ReLU(inplace=True),
                                     nn.Conv2d(cout, cout, 3, padding=1), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))

    def forward(self, x):
        return self.conv_op(x)


class StockPGUNet(nn.Module):
    """PGUNet{stage}(3, nc) of UG_unet.py:178-304 with stock modules."""

    def __init__(self, stage: int, nc: int = 1):
        super().__init__()
        enc = {1: [512, 512], 2: [256, 512, 512], 3: [128, 256, 512, 512], 4: [64, 128, 256, 512, 512]}[stage]
        up_out = [256, 128, 64, 64]
        self.stage = stage
        self.inc = _DoubleConv(3, enc[0])
        self.downs = nn.ModuleList(_DoubleConv(enc[i], enc[i + 1]) for i in range(len(enc) - 1))
        ups, heads, prev = [], [], enc[-1]
        for j in range(stage):
            ups.append(_DoubleConv(enc[stage - 1 - j] + prev, up_out[j]))
            prev = up_out[j]
            heads.append(nn.Conv2d(prev, nc, 1))
        self.ups = nn.ModuleList(ups)
        self.heads = nn.ModuleList(heads if stage > 1 else heads[-1:])

    def forward(self, x):
        feats = [self.inc(x)]
        for d in self.downs:
            feats.append(d(F.max_pool2d(feats[-1], 2)))
        cur, dec = feats[-1], []
        for j, u in enumerate(self.ups):
            up = F.interpolate(cur, scale_factor=2, mode="bilinear", align_corners=True)
            cur = u(torch.cat([feats[self.stage - 1 - j], up], dim=1))
            dec.append(cur)
        if self.stage == 1:
            return self.heads[0](dec[-1])
        out = None
        for h, d in zip(self.heads, dec):
            l = h(d)
            f = x.shape[-1] // l.shape[-1]
            if f > 1:
                l = F.interpolate(l, scale_factor=f, mode="bilinear", align_corners=True)
            out = l if out is None else out + l
        return out


def time_stock_step(mode: str, batch: int, steps: int, warmup: int, device, res: int = 256, stage: int = 4):
    """-> (images/s, ms/step) of the stage-`stage` trainer step (train forward, previous-stage uncertainty
    forward, uncertainty-weighted BCE, backward, RMSprop) in stock PyTorch, `mode` in {"tf32", "bf16"}."""
    assert mode in ("tf32", "bf16")
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    try:
        torch.manual_seed(0)
        fmt = torch.channels_last if mode == "bf16" else torch.contiguous_format
        model = StockPGUNet(stage).to(device).to(memory_format=fmt).train()
        prev = StockPGUNet(stage - 1).to(device).to(memory_format=fmt).eval()
        opt = torch.optim.RMSprop(model.parameters(), lr=1e-4, weight_decay=1e-4)
        crit = nn.BCEWithLogitsLoss(pos_weight=torch.tensor([5.0], device=device), reduction="none")
        g = torch.Generator().manual_seed(1234)
        x = torch.randn(batch, 3, res, res, generator=g).to(device).contiguous(memory_format=fmt)
        y = torch.randint(0, 2, (batch, 1, res, res), generator=g).float().to(device)
        cast = torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16"))

        def step():
            opt.zero_grad(set_to_none=True)
            with cast:
                out = model(x)
                with torch.no_grad():
                    xp = F.interpolate(x, size=(res // 2, res // 2), mode="bilinear", align_corners=True)
                    p = torch.sigmoid(prev(xp).float())
                    p = F.interpolate(p, size=(res, res), mode="bilinear", align_corners=True)
                    A = 1.0 - 2.0 * torch.abs(p - 0.5)
            loss = torch.mean(crit(out.float(), y) * (1.0 + A).detach())
            loss.backward()
            opt.step()
            return loss

        for _ in range(max(warmup, 2)):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return batch / (ms / 1e3), ms, float(loss)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old

"""Inference / evaluator timing: PGUNet4 eval forward + on-device thresholding and
confidence through `MoNuSegEvaluator.predict_batch` at batch 1 and 16, eager launches vs CUDA-graph replay,
fp32 NCHW inputs vs uint8 HWC loader bytes; device-resident and end to end from pinned host buffers
(H2D of the image, D2H of the uint8 mask + confidence inside the timed region).

    python tools/t_infer.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ug_pg_unet_b200 as U

GF_PER_IMG = 61.8445     # PGUNet4 forward, 256 x 256


def timed(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return a.elapsed_time(e) / n


def main():
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    model = U.PGUNet4(3, 1).to(dev)
    for graph in (False, True):
        ev = U.MoNuSegEvaluator(model, device=dev, use_cuda_graph=graph)
        for B in (1, 16):
            g = torch.Generator().manual_seed(B)
            img8 = torch.randint(0, 256, (B, 256, 256, 3), generator=g, dtype=torch.uint8)
            imgf = (img8.permute(0, 3, 1, 2).float() / 255.0).contiguous()
            for name, host in (("fp32 NCHW", imgf), ("uint8 HWC", img8)):
                d = host.to(dev)
                hp = host.pin_memory()
                ms = timed(lambda: ev.predict_batch(d), 30)

                def e2e():
                    mask, conf = ev.predict_batch(hp.to(dev, non_blocking=True))
                    return mask.cpu(), conf.cpu()

                ms2 = timed(e2e, 30)
                print(f"{'graph' if graph else 'eager'}  batch {B:2d}  {name}: {ms:6.3f} ms ({B / ms * 1e3:7.0f} images/s, "
                      f"{B * GF_PER_IMG / ms:6.0f} TFLOP/s)   host->host {ms2:6.3f} ms ({B / ms2 * 1e3:7.0f} images/s, "
                      f"{host.numel() * host.element_size() / 1e6:.2f} MB in)", flush=True)


if __name__ == "__main__":
    main()

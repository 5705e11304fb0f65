"""ug_pg_unet_b200 — the UG-PG-UNet hot path on hand-written sm_100a kernels.

Public surface (mirrors the reference's UG_unet_parts.py / UG_unet.py /
uncertainty_guided_trainer.py):

    DoubleConv, InConv, Down, Up, OutConv, DownSample, UpSample
    PGUNet1..4, ProgressiveUNet, UncertaintyGuidedLoss, UNet (legacy, ConvTranspose2d up path)
    UncertaintyGuidedProgressiveTrainer
    HerlevClassificationModel, HerlevTrainer          (Herlev/train_herlev.py)
    MoNuSegEvaluator, HerlevEvaluator                 (MoNuSegImprove/test_monuseg.py, Herlev/test_herlev.py)

Everything executes through libugp_b200.so (C ABI in include/ugp_b200.h);
there is no CPU or library fallback.
"""
from .unet_parts import DoubleConv, Down, DownSample, InConv, OutConv, Up, UpSample
from .unet import (PGUNet1, PGUNet2, PGUNet3, PGUNet4, ProgressiveUNet, STAGE_MODELS, UNet,
                   UncertaintyGuidedLoss, transfer_state)
from .trainer import UncertaintyGuidedProgressiveTrainer
from .herlev import HerlevClassificationModel, HerlevTrainer
from .evaluate import HerlevEvaluator, MoNuSegEvaluator, classification_metrics_from_confusion, metrics_from_counts

__all__ = [
    "DoubleConv", "InConv", "Down", "Up", "OutConv", "DownSample", "UpSample", "UNet",
    "PGUNet1", "PGUNet2", "PGUNet3", "PGUNet4", "ProgressiveUNet", "STAGE_MODELS",
    "UncertaintyGuidedLoss", "transfer_state", "UncertaintyGuidedProgressiveTrainer",
    "HerlevClassificationModel", "HerlevTrainer", "MoNuSegEvaluator", "HerlevEvaluator", "metrics_from_counts",
    "classification_metrics_from_confusion",
]

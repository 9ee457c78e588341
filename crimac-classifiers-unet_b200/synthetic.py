"""Synthetic workloads of the benchmark configurations: echogram patches in dB and label maps.

Kept in the package so that the measured path (bench.py's B200 arm, the examples) does not import anything from
oracle/ - the oracle is test infrastructure.  The generators are deterministic in `seed`.
"""
import torch
import torch.nn.functional as F

IGNORE_INDEX = -100  # nn.CrossEntropyLoss default (reference pipeline_train_predict/pipeline.py:138)


def synthetic_echogram(batch, channels, height, width, seed=0, device="cpu"):
    """configs[0]/[1] input: x = clip(10*log10(sv + 1e-10), -75, 0) with sv = 10**U(-9,-2) - what remove_nan_inf +
    db_with_limits (batch/data_transforms/) produce from volume backscatter."""
    g = torch.Generator().manual_seed(seed)
    sv = 10.0 ** (torch.rand((batch, channels, height, width), generator=g) * 7.0 - 9.0)
    return torch.clamp(10.0 * torch.log10(sv + 1e-10), -75.0, 0.0).to(device)


def synthetic_labels(batch, height, width, seed=1, device="cpu"):
    """configs[1] labels: blobs of class 1 / 2 on background 0 (~90/5/5 %), 2 % of the pixels ignored (-100)."""
    g = torch.Generator().manual_seed(seed)
    coarse = torch.rand((batch, 1, max(height // 16, 1), max(width // 16, 1)), generator=g)
    field = F.interpolate(coarse, size=(height, width), mode="bilinear", align_corners=False)[:, 0]
    lab = torch.zeros((batch, height, width), dtype=torch.long)
    lab[field > 0.80] = 1
    lab[field < 0.20] = 2
    lab[torch.rand((batch, height, width), generator=g) < 0.02] = IGNORE_INDEX
    return lab.to(device)

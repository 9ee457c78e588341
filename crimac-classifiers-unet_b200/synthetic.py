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


def structured_batch(batch, height, width, seed=0, device="cpu"):
    """A LEARNABLE 4-frequency workload (SURVEY.md section 7.2's "trained-like" parity runs need a net that has learned
    something): the class blobs of `synthetic_labels` are imprinted on the data - class 1 raises the two low
    frequencies, class 2 the two high ones - and everything stays inside the dB range [-75, 0]."""
    y = synthetic_labels(batch, height, width, seed=seed)
    g = torch.Generator().manual_seed(seed + 1000)
    x = -60.0 + 6.0 * torch.randn((batch, 4, height, width), generator=g)
    x[:, 0:2] += 18.0 * (y == 1).unsqueeze(1)
    x[:, 2:4] += 18.0 * (y == 2).unsqueeze(1)
    return torch.clamp(x, -75.0, 0.0).to(device), y.to(device)


def synthetic_survey_pings(freqs, n_range, p0, p1, seed=0, device="cpu", nan_fraction=1e-3, block=20000):
    """BASELINE configs[3] survey, pings [p0, p1): sv (F, R, p1-p0) fp32 = 10**U(-9,-2) with 0.1 % NaNs.  The values of a
    ping depend only on (seed, ping index) - generated in fixed blocks of `block` pings - so that any rank loading any
    window of the ONE shared survey sees the same data (inference shards by ping range, save_predict.py:160-171)."""
    out = torch.empty((freqs, n_range, p1 - p0), dtype=torch.float32, device=device)
    b0 = p0 // block
    while b0 * block < p1:
        lo, hi = b0 * block, (b0 + 1) * block
        g = torch.Generator(device=device).manual_seed(seed * 1000003 + b0)
        blk = torch.pow(10.0, torch.rand((freqs, n_range, block), device=device, generator=g) * 7.0 - 9.0)
        blk[torch.rand((freqs, n_range, block), device=device, generator=g) < nan_fraction] = float("nan")
        s, e = max(lo, p0), min(hi, p1)
        out[:, :, s - p0:e - p0] = blk[:, :, s - lo:e - lo]
        b0 += 1
    return out


def synthetic_seabed(p0, p1, device="cpu"):
    """configs[3] seabed index per ping: 200 + 20*sin(2*pi*p/5000), int32."""
    import math
    p = torch.arange(p0, p1, device=device, dtype=torch.float64)
    return (200 + 20 * torch.sin(2 * math.pi * p / 5000)).to(torch.int32)

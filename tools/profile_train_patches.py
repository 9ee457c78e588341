#!/usr/bin/env python
"""Workload for timing / profiling the on-device training-sample path alone: a synthetic 4-frequency survey resident
in HBM, batches of 32 crops of 256x256.  Prints kernel-only and feeder-level times (CUDA events, L2 flushed by the
working set: 32 crops read ~36 MB and write ~50 MB per batch, the survey is 480 MB)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

ge.load_package()
from crimac_unet_b200 import engine as E  # noqa: E402
from crimac_unet_b200 import train_patches as TP  # noqa: E402

dev = torch.device("cuda:0")
F, NP, R, n, patch = 4, 60000, 500, 32, 256
g = torch.Generator(device=dev).manual_seed(0)
sv = 10.0 ** (torch.rand((F, NP, R), device=dev, generator=g) * 7 - 9)
labels = torch.zeros((NP, R), device=dev)
rng = np.random.default_rng(0)
for _ in range(400):
    cy, cx = int(rng.integers(0, NP)), int(rng.integers(0, R))
    labels[max(0, cy - 60):cy + 60, max(0, cx - 40):cx + 40] = float(rng.choice([27, 1]))
feeder = TP.SurveyPatchFeeder(sv, labels, n, (patch, patch), seed=3)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for _ in range(3):
    feeder.next_batch()
cen, fl = feeder.draw()
cen_d, fl_d = torch.from_numpy(cen).to(dev), torch.from_numpy(fl).to(dev)
x = torch.empty((n, F, patch, patch), device=dev)
y = torch.empty((n, patch, patch), dtype=torch.int64, device=dev)
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
t0.record()
for i in range(steps):
    E.train_patches(sv, labels, cen_d, fl_d, (patch, patch), seed=i, out=x, labels_out=y)
t1.record()
torch.cuda.synchronize()
k_ms = t0.elapsed_time(t1) / steps
t0.record()
for i in range(steps):
    feeder.next_batch()
t1.record()
torch.cuda.synchronize()
f_ms = t0.elapsed_time(t1) / steps
# algorithmic bytes per crop: sv in + x out (F*ph*pw*8), label in (4), code out + in (16), label out (8)
gb = n * (F * patch * patch * 8 + patch * patch * 28) / 1e9
print(f"train_patches batch {n} of {F}x{patch}x{patch}: kernels {k_ms:.3f} ms ({n / k_ms * 1e3:.0f} patches/s, "
      f"{gb / k_ms * 1e3:.0f} GB/s algorithmic), through the feeder {f_ms:.3f} ms")

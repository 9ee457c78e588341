"""Training samples straight from a survey resident in HBM.

The reference feeds training from CPU DataLoader workers: Dataset.__getitem__ (batch/dataset.py:75-108) crops one
sample at a time (get_crop_zarr, :358-407), augments it in numpy (batch/data_augmentation/add_noise.py, flip_x_axis.py),
refines the labels with a 7x7 binary closing (batch/label_transforms/refine_label_boundary.py), re-indexes them
(convert_label_indexing.py) and applies the dB transform (batch/data_transforms/*.py); pipeline.py:161-164 then copies
the batch to the GPU.  Here the survey's sv and annotation arrays stay on the device in the zarr store's own
[frequency][ping][range] order and a whole batch is produced by two kernel launches (engine.train_patches ->
crimac_train_patches), so the only per-step host->device traffic is 9 bytes per sample (centre + two coin flips).

Which samples to draw is the samplers' business (batch/samplers/*.py, out of scope): `SurveyPatchFeeder` takes any
callable returning crop centres and defaults to uniform centres over the survey.
"""
import numpy as np
import torch

from . import engine as _engine


def uniform_centres(n, n_range, n_pings, rng):
    """(n,2) int32 centres (range, ping) uniform over the survey (a stand-in for the reference's samplers)."""
    return np.stack([rng.integers(0, n_range, n), rng.integers(0, n_pings, n)], axis=1).astype(np.int32)


class SurveyPatchFeeder:
    """Draws (x, labels) training batches on the device.

    sv: fp32 (F, P, R) device tensor; labels: fp32 (P, R) raw annotation categories (0 background, 27 sandeel, 1 other,
    other species > 0, NaN / -100 no data).  The two per-sample coin flips of the reference's augmentation
    (add_noise.py:25, flip_x_axis.py:22) are drawn on the host from `rng`; the per-element noise field comes from the
    kernel's counter-based generator keyed by (seed, step)."""

    def __init__(self, sv, labels, batch_size, patch_hw=(256, 256), seed=0, centre_sampler=None, augment=True,
                 scaled=False, thr_freq=None, thr=(1e-7, 1e-4)):
        if sv.dim() != 3 or labels.dim() != 2 or tuple(labels.shape) != tuple(sv.shape[1:]):
            raise ValueError("sv must be (F, P, R) and labels (P, R)")
        if not sv.is_cuda:
            raise ValueError("the survey must be resident on the GPU (there is no CPU path)")
        self.sv, self.labels = sv.contiguous(), labels.contiguous()
        self.batch_size, self.patch_hw = int(batch_size), tuple(patch_hw)
        self.rng = np.random.default_rng(seed)
        self.seed, self.step = int(seed), 0
        self.centre_sampler = centre_sampler
        self.augment, self.scaled, self.thr_freq, self.thr = bool(augment), bool(scaled), thr_freq, tuple(thr)
        dev = sv.device
        F = sv.shape[0]
        n, (ph, pw) = self.batch_size, self.patch_hw
        # two batches in flight: the previous step may still be reading its inputs when the next batch is produced
        self._x = [torch.empty((n, F, ph, pw), dtype=torch.float32, device=dev) for _ in range(2)]
        self._y = [torch.empty((n, ph, pw), dtype=torch.int64, device=dev) for _ in range(2)]
        self._host = [torch.empty((n, 3), dtype=torch.int32).pin_memory() for _ in range(2)]
        self._centres = [torch.empty((n, 2), dtype=torch.int32, device=dev) for _ in range(2)]
        self._flags = [torch.empty((n,), dtype=torch.uint8, device=dev) for _ in range(2)]
        self._staged = [torch.empty((n, 3), dtype=torch.int32, device=dev) for _ in range(2)]
        self._copied = [None, None]   # event after the slot's host->device copy: the pinned buffer may be rewritten

    def draw(self):
        """Host-side decisions of one batch: centres (n,2) int32 and flags (n) uint8 (bit0 noise, bit1 flip)."""
        n = self.batch_size
        _, P, R = self.sv.shape
        if self.centre_sampler is not None:
            centres = np.asarray(self.centre_sampler(n, self.rng), dtype=np.int32).reshape(n, 2)
        else:
            centres = uniform_centres(n, R, P, self.rng)
        if self.augment:
            flags = (self.rng.integers(0, 2, n) | (self.rng.integers(0, 2, n) << 1)).astype(np.uint8)
        else:
            flags = np.zeros(n, dtype=np.uint8)
        return centres, flags

    def next_batch(self, centres=None, flags=None):
        """Returns device tensors (x fp32 (n,F,ph,pw), labels int64 (n,ph,pw)) valid until the call after next."""
        slot = self.step & 1
        if centres is None:
            centres, flags = self.draw()
        h = self._host[slot]
        if self._copied[slot] is not None:
            self._copied[slot].synchronize()   # the host may run several steps ahead of the device
        h[:, :2] = torch.from_numpy(np.ascontiguousarray(centres, dtype=np.int32))
        h[:, 2] = torch.from_numpy(np.ascontiguousarray(flags).astype(np.int32))
        self._staged[slot].copy_(h, non_blocking=True)
        self._copied[slot] = torch.cuda.Event()
        self._copied[slot].record()
        self._centres[slot].copy_(self._staged[slot][:, :2])
        self._flags[slot].copy_(self._staged[slot][:, 2])
        _engine.train_patches(self.sv, self.labels, self._centres[slot], self._flags[slot], self.patch_hw,
                              seed=(self.seed << 32) ^ self.step, thr_freq=self.thr_freq, thr=self.thr,
                              scaled=self.scaled, out=self._x[slot], labels_out=self._y[slot])
        self.step += 1
        return self._x[slot], self._y[slot]

    def __iter__(self):
        while True:
            yield self.next_batch()

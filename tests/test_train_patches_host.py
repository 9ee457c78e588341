"""CPU tests of the training-sample path.

* the oracle restatement (oracle/pipeline_oracle.py: get_crop_zarr ... train_patch_item) against the fixture made by
  RUNNING the reference's own functions (oracle/make_golden.py:golden_train_pipeline -> golden/pipeline_train.npz);
* the product's kernel bodies (csrc/train_patch_core.h — the very functions the CUDA kernels call) executed on the
  host by tests/host/train_patch_hostcheck.cc, against the same fixture, against scipy's binary_closing and against
  the distribution of the reference's add_noise.  No product code path runs on the CPU: this is a test harness.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import pipeline_oracle as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
vp = ctypes.c_void_p


@pytest.fixture(scope="module")
def hostlib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("tphost") / "libtphost.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tests", "host", "train_patch_hostcheck.cc")],
                   check=True)
    return ctypes.CDLL(so)


def noise_field(lib, seed, chan, ph, pw):
    out = np.zeros((ph, pw), np.float32)
    lib.tp_host_noise_field(ctypes.c_uint64(seed), chan, ph, pw, vp(out.ctypes.data))
    return out


def host_train_patches(lib, sv, labels, centres, flags, mult, patch, seed=0, thr_freq=None, scaled=0, border=0):
    F, NP, R = sv.shape
    n = len(centres)
    x = np.zeros((n, F, patch, patch), np.float32)
    y = np.zeros((n, patch, patch), np.int64)
    sv, labels = np.ascontiguousarray(sv, np.float32), np.ascontiguousarray(labels, np.float32)
    centres, flags = np.ascontiguousarray(centres, np.int32), np.ascontiguousarray(flags, np.uint8)
    mp = vp(0) if mult is None else vp(np.ascontiguousarray(mult, np.float32).ctypes.data)
    keep = mult  # noqa: F841  (keeps the array alive during the call)
    rc = lib.tp_host_train_patches(vp(sv.ctypes.data), vp(labels.ctypes.data), F, NP, R, vp(centres.ctypes.data),
                                   vp(flags.ctypes.data), mp, ctypes.c_uint64(seed), n, patch, patch,
                                   F - 1 if thr_freq is None else thr_freq, ctypes.c_double(1e-7), ctypes.c_double(1e-4),
                                   scaled, border, vp(x.ctypes.data), vp(y.ctypes.data))
    assert rc == 0
    return x, y


def load_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "pipeline_train.npz"))
    flags = (g["noise_on"].astype(np.uint8) | (g["flip"].astype(np.uint8) << 1)).astype(np.uint8)
    return g, flags


def test_oracle_matches_the_reference_made_fixture(golden_dir):
    g, _ = load_fixture(golden_dir)
    patch = tuple(int(v) for v in g["patch"])
    refined = rescued = 0
    for i, c in enumerate(g["centres"]):
        for key, kw in (("data", {}), ("data_scaled", {"scaled": True}), ("data_border", {"border_zero": True})):
            d, l = P.train_patch_item(g["sv"], g["labels"], c, g["noise_on"][i], g["flip"][i], g["mult"][i], patch, **kw)
            assert np.array_equal(l, g["out_labels"][i].astype(np.int64))
            assert np.array_equal(np.isnan(d), np.isnan(g[key][i]))
            assert np.allclose(d, g[key][i], rtol=0, atol=1e-5, equal_nan=True)
        # the fixture exercises both effects of refine_label_boundary: schools samples rejected by the closing and
        # below-threshold samples rescued by it
        data, raw = P.get_crop_zarr(g["sv"], g["labels"], c, patch)
        if g["noise_on"][i]:
            data = data * g["mult"][i]
        if g["flip"][i]:
            data, raw = np.flip(data, 2), np.flip(raw, 1)
        fish = (raw == 27) | (raw == 1)
        thr = (data[-1] > 1e-7) & (data[-1] < 1e-4)
        out = g["out_labels"][i]
        refined += int((fish & (out == -100)).sum())
        rescued += int((fish & ~thr & (out != -100)).sum())
    assert refined > 50 and rescued > 50, (refined, rescued)


def test_kernel_bodies_on_the_host_match_the_reference_made_fixture(hostlib, golden_dir):
    g, flags = load_fixture(golden_dir)
    patch = int(g["patch"][0])
    for key, kw in (("data", {}), ("data_scaled", {"scaled": 1}), ("data_border", {"border": 1})):
        x, y = host_train_patches(hostlib, g["sv"], g["labels"], g["centres"], flags, g["mult"], patch, **kw)
        assert np.array_equal(y, g["out_labels"].astype(np.int64))          # labels: bit-exact
        assert np.array_equal(np.isnan(x), np.isnan(g[key]))
        assert np.nanmax(np.abs(x - g[key])) <= 1e-4                        # dB values: fp32 log10f vs numpy float64


def test_kernel_bodies_on_the_host_match_the_oracle_on_random_crops(hostlib):
    rng = np.random.default_rng(21)
    F, NP, R, patch, n = 2, 300, 200, 96, 6
    sv = (10.0 ** rng.uniform(-9, -2, size=(F, NP, R))).astype(np.float32)
    labels = np.zeros((NP, R), np.float32)
    for _ in range(12):
        cy, cx = rng.integers(0, NP), rng.integers(0, R)
        labels[max(0, cy - 20):cy + 20, max(0, cx - 15):cx + 15] = rng.choice([27, 1, 5])
    sv[F - 1][labels > 0] = (10.0 ** rng.uniform(-7.6, -3.6, size=int((labels > 0).sum()))).astype(np.float32)
    labels[rng.random(labels.shape) < 0.002] = np.nan
    sv[0][rng.random((NP, R)) < 0.002] = np.nan
    centres = np.stack([rng.integers(-20, R + 20, n), rng.integers(-20, NP + 20, n)], 1).astype(np.int32)
    flags = rng.integers(0, 4, n).astype(np.uint8)
    mult = np.stack([P.noise_multiplier_field((F, patch, patch), rng) for _ in range(n)])
    x, y = host_train_patches(hostlib, sv, labels, centres, flags, mult, patch)
    for i in range(n):
        d, l = P.train_patch_item(sv, labels, centres[i], flags[i] & 1, flags[i] & 2, mult[i], (patch, patch))
        assert np.array_equal(y[i], l), i
        assert np.allclose(x[i], d, rtol=0, atol=1e-4, equal_nan=True), i


def test_kernel_bodies_on_the_host_edge_cases(hostlib):
    """Survey smaller than the patch, crops hanging over every edge or entirely outside, a one-sample survey, an
    all-fish survey whose threshold channel is entirely inside / entirely outside the window."""
    rng = np.random.default_rng(33)
    for (F, NP, R, patch) in ((1, 20, 50, 64), (3, 70, 17, 32), (2, 1, 1, 32), (2, 40, 40, 32)):
        sv = (10.0 ** rng.uniform(-9, -2, size=(F, NP, R))).astype(np.float32)
        labels = rng.choice([0, 0, 27, 1, 9, np.nan, -100, -1], size=(NP, R)).astype(np.float32)
        centres = np.array([[0, 0], [R - 1, NP - 1], [R // 2, NP // 2], [-patch, NP // 2], [R // 2, NP + patch],
                            [R + 3 * patch, -3 * patch], [patch // 2 - 1, patch // 2 - 1]], np.int32)
        n = len(centres)
        flags = rng.integers(0, 4, n).astype(np.uint8)
        mult = np.stack([P.noise_multiplier_field((F, patch, patch), rng) for _ in range(n)])
        x, y = host_train_patches(hostlib, sv, labels, centres, flags, mult, patch)
        for i in range(n):
            d, l = P.train_patch_item(sv, labels, centres[i], flags[i] & 1, flags[i] & 2, mult[i], (patch, patch))
            assert np.array_equal(y[i], l), (F, NP, R, patch, i)
            assert np.allclose(x[i], d, rtol=0, atol=1e-4, equal_nan=True), (F, NP, R, patch, i)
    for lo_hi, expect_fish in (((-6.5, -4.5), True), ((-3.5, -2.5), False)):
        sv = (10.0 ** rng.uniform(*lo_hi, size=(2, 64, 64))).astype(np.float32)
        labels = np.full((64, 64), 27, np.float32)
        c = np.array([[31, 31]], np.int32)
        x, y = host_train_patches(hostlib, sv, labels, c, np.zeros(1, np.uint8), None, 64)
        d, l = P.train_patch_item(sv, labels, c[0], 0, 0, None, (64, 64))
        assert np.array_equal(y[0], l)
        # scipy's closing erodes with border_value 0: even a fully passing school loses its 3-sample rim
        assert bool((y[0, 3:-3, 3:-3] == 1).all()) == expect_fish and bool((y[0, 0] == -100).all())
        assert bool((y == -100).all()) == (not expect_fish)


def test_bit_row_closing_matches_scipy(hostlib):
    from scipy.ndimage import binary_closing
    rng = np.random.default_rng(4)
    for t in range(60):
        H, W = 32 * rng.integers(1, 4), 32 * rng.integers(1, 4)
        y0, x0 = rng.integers(0, H - 1), rng.integers(0, W - 1)
        y1, x1 = rng.integers(y0 + 1, H + 1), rng.integers(x0 + 1, W + 1)
        m = np.zeros((H, W), np.uint8)
        m[y0:y1, x0:x1] = rng.random((y1 - y0, x1 - x0)) < rng.uniform(0.05, 0.95)   # the mask is 0 outside the box
        out = np.zeros((H, W), np.uint8)
        hostlib.tp_host_closing(vp(m.ctypes.data), int(H), int(W), int(y0), int(y1), int(x0), int(x1), vp(out.ctypes.data))
        ref = np.zeros((H, W), bool)
        ref[y0:y1, x0:x1] = binary_closing(m[y0:y1, x0:x1].astype(bool), structure=P.DISC7)
        assert np.array_equal(out.astype(bool), ref), (t, H, W, y0, y1, x0, x1)
        assert np.array_equal(P.binary_closing(m[y0:y1, x0:x1].astype(bool)), ref[y0:y1, x0:x1])


def test_counter_based_noise_has_the_distribution_of_add_noise(hostlib):
    """add_noise.py:28-38: 5 % of the samples change; half of those are multiplied by U(1,10), half by U(0,1)."""
    m = noise_field(hostlib, 12345, 3, 640, 640).astype(np.float64).ravel()
    changed = m != 1.0
    assert abs(changed.mean() - 0.05) < 0.002
    up, down = m[changed & (m > 1.0)], m[changed & (m < 1.0)]
    assert len(up) + len(down) == changed.sum()
    assert abs(len(up) / changed.sum() - 0.5) < 0.02
    assert 1.0 < up.min() and up.max() < 10.0 and abs(up.mean() - 5.5) < 0.1 and abs(up.std() - 9 / 12 ** 0.5) < 0.1
    assert 0.0 < down.min() and abs(down.mean() - 0.5) < 0.02 and abs(down.std() - 1 / 12 ** 0.5) < 0.02
    # the two samples that share one generator call (rows py and py + 8) are independent
    f = noise_field(hostlib, 12345, 3, 640, 640) != 1.0
    rows = np.arange(640)
    base = rows[(rows >> 3) & 1 == 0]
    both = (f[base] & f[base + 8]).mean()
    assert abs(both - 0.0025) < 0.0006
    # a pure function of (seed, channel, row, ping): reproducible, and different for another seed / channel
    assert np.array_equal(noise_field(hostlib, 1, 0, 64, 64), noise_field(hostlib, 1, 0, 64, 64))
    assert (noise_field(hostlib, 1, 0, 64, 64) != noise_field(hostlib, 2, 0, 64, 64)).mean() > 0.05
    assert (noise_field(hostlib, 1, 0, 64, 64) != noise_field(hostlib, 1, 1, 64, 64)).mean() > 0.05


def test_feeder_argument_validation_needs_no_gpu(pkg):
    import importlib
    import torch
    tp = importlib.import_module("crimac_unet_b200.train_patches")
    with pytest.raises(ValueError):
        tp.SurveyPatchFeeder(torch.zeros(2, 10, 10), torch.zeros(10, 11), 4)
    with pytest.raises(ValueError):      # a CPU survey is refused: there is no CPU path
        tp.SurveyPatchFeeder(torch.zeros(2, 10, 10), torch.zeros(10, 10), 4)
    c = tp.uniform_centres(100, 50, 70, np.random.default_rng(0))
    assert c.shape == (100, 2) and c.dtype == np.int32 and c[:, 0].max() < 50 and c[:, 1].max() < 70

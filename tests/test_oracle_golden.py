"""The oracle (oracle/*.py) against the golden vectors produced by RUNNING THE REFERENCE (oracle/make_golden.py).
CPU only; this is what pins the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import pipeline_oracle as P
from oracle import unet_oracle as O


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _state_from(g):
    return {k[len("state/"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("state/")}


def _seeded_d5_state():
    """Re-create the depth-5 weights of unet_d5.npz from their seed (the 124 MB state is not committed)."""
    import importlib
    import __graft_entry__ as ge
    ge.load_package()
    M = importlib.import_module("crimac_unet_b200.models.unet")
    torch.manual_seed(0)
    return O.trained_like_state(M.UNet_Baseline(3, 4, depth=5).state_dict(), seed=0)


def test_unet_depth2_eval_and_train_match_reference(golden_dir):
    g = _load(golden_dir, "unet_d2.npz")
    sd = _state_from(g)
    assert len(sd) == 2 * 14 + 16 + 2
    x, y = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
    logits = O.unet_forward(sd, x, train=False)
    assert np.allclose(logits.numpy(), g["eval_logits"], rtol=0, atol=1e-5)
    lg, loss, grads, stats = O.train_step(sd, x, y)
    assert np.allclose(lg.numpy(), g["train_logits"], rtol=0, atol=1e-4)
    assert abs(loss.item() - float(g["loss"])) < 1e-5
    for k in g.files:
        if k.startswith("grad/"):
            name = k[5:]
            ref = g[k]
            pre_bn_bias = name.endswith(".bias") and any(s in name for s in ("main.0", "main.3", "conv1", "conv2"))
            if pre_bn_bias:   # mathematically zero (BN removes it); both sides are rounding noise
                assert np.abs(grads[name].numpy()).max() < 1e-6 and np.abs(ref).max() < 1e-6
            else:
                assert np.linalg.norm(grads[name].numpy() - ref) <= 1e-3 * np.linalg.norm(ref), name
        if k.startswith("stat/"):
            assert np.allclose(stats[k[5:]].numpy(), g[k], rtol=1e-4, atol=1e-5), k


def test_unet_depth5_matches_reference(golden_dir):
    g = _load(golden_dir, "unet_d5.npz")
    sd = _seeded_d5_state()
    chk = np.array([float(v.double().abs().sum()) for v in sd.values()])
    if not np.allclose(chk, g["state_checksums"], rtol=1e-9):
        pytest.skip("seeded weights differ from the fixture's (different torch build): depth-5 golden not applicable")
    x, y = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
    assert np.allclose(O.unet_forward(sd, x, train=False).numpy(), g["eval_logits"], rtol=0, atol=1e-5)
    lg, loss, grads, _ = O.train_step(sd, x, y)
    assert np.allclose(lg.numpy(), g["train_logits"], rtol=0, atol=2e-4)
    assert abs(loss.item() - float(g["loss"])) < 1e-5
    n = 0
    for k in g.files:
        if k.startswith("grad/") and not k.endswith("main.0.bias"):
            ref = g[k]
            assert np.linalg.norm(grads[k[5:]].numpy() - ref) <= 2e-3 * np.linalg.norm(ref) + 1e-9, k
            n += 1
    assert n > 30


def test_weighted_ce_edge_cases():
    lg = torch.randn(2, 3, 4, 4)
    y = torch.full((2, 4, 4), -100)
    assert torch.isnan(O.weighted_ce(lg, y))            # all ignored -> NaN, as nn.CrossEntropyLoss
    y[0, 0, 0] = 2
    ref = torch.nn.CrossEntropyLoss(weight=torch.tensor(O.CLASS_WEIGHTS))(lg, y)
    assert torch.allclose(O.weighted_ce(lg, y), ref, atol=1e-6)


def test_pipeline_matches_reference(golden_dir):
    g = _load(golden_dir, "pipeline_small.npz")
    sv, labels, seabed = g["sv"], g["labels"], g["seabed"]
    patch, overlap = tuple(int(v) for v in g["patch"]), int(g["overlap"])
    F_, NP, R = sv.shape
    splits = P.get_data_split([[0, NP]], 300)
    assert np.array_equal(splits, g["splits"])
    for ci, (s, e) in enumerate(splits):
        grid = P.get_data_grid(int(s), int(e), 0, P.end_range_from_seabed(R, seabed[s:e]), patch, overlap)
        assert np.array_equal(grid, g[f"chunk{ci}/grid"])
        d0, d1 = P.preload_extents(grid, NP, patch[1])
        assert [d0, d1] == list(g[f"chunk{ci}/preload"])
        sv_pre = np.swapaxes(sv[:, d0:d1, :], 1, 2)
        out = np.zeros((2, R, e - s))
        for i, c in enumerate(grid):
            d, l = P.patch_item(sv_pre, d0, labels[s:e, :].T, int(s), c, seabed, R, NP, patch, overlap)
            assert np.array_equal(l, g[f"chunk{ci}/labels"][i])
            assert np.allclose(d, g[f"chunk{ci}/data"][i], rtol=0, atol=1e-4)
            probs = np.stack([np.full(patch, 0.1 * k + 0.001 * i) + 1e-4 * np.arange(patch[1])[None, :] for k in range(3)])
            P.fill_out_array(out, probs, l, c, int(s))
        assert np.allclose(out, g[f"chunk{ci}/stitched"], rtol=0, atol=1e-6)


def test_reference_grid_geometry_of_config4():
    """1 M pings / 20 000 -> 50 chunks; R = 256 -> 2 rows x 93 columns = 186 patches per chunk."""
    splits = P.get_data_split([[0, 1_000_000]], 20000)
    assert len(splits) == 50 and all(e - s == 20000 for s, e in splits)
    grid = P.get_data_grid(0, 20000, 0, 256, (256, 256), 20)
    assert len(grid) == 186 and list(grid[0]) == [107, 107] and list(grid[1]) == [107, 323]
    assert sorted(set(grid[:, 0])) == [107, 323]


@pytest.mark.parametrize("variant,kw", [("add", dict(up_mode="transpose", merge_mode="add")),
                                        ("upsample", dict(up_mode="upsample", merge_mode="concat"))])
def test_oracle_decoder_variants_against_reference_golden(golden_dir, variant, kw):
    """oracle.unet_forward / train_step with up_mode / merge_mode against outputs of the reference's own classes
    (oracle/make_golden_variants.py, models/unet.py:47-56,113-118,131-134)."""
    import importlib
    import __graft_entry__ as ge
    ge.load_package()
    Mm = importlib.import_module("crimac_unet_b200.models.unet")
    g = np.load(os.path.join(golden_dir, f"unet_{variant}_d3.npz"))
    torch.manual_seed(7)
    m = Mm.UNet_Baseline(3, 4, depth=3, **kw)
    sd0 = O.trained_like_state({k: v.detach().clone() for k, v in m.state_dict().items()}, seed=1, head_gain=2.0)
    assert np.allclose(np.array([float(v.double().abs().sum()) for v in sd0.values()]), g["state_checksum"], rtol=1e-6)
    x, y = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
    with torch.no_grad():
        ev = O.unet_forward(sd0, x, **kw)
    assert (ev - torch.from_numpy(g["eval_logits"])).abs().max().item() < 1e-4
    tl, loss, grads, stats = O.train_step(sd0, x, y, **kw)
    assert (tl - torch.from_numpy(g["train_logits"])).abs().max().item() < 1e-3
    assert abs(float(loss) - float(g["loss"])) < 1e-5
    for k in g.files:
        if k.startswith("grad/") and not (k.endswith(".bias") and any(t in k for t in ("main.0", "main.3", "conv1", "conv2"))):
            a, b = grads[k[5:]], torch.from_numpy(g[k])
            assert ((a - b).norm() / (b.norm() + 1e-30)).item() < 2e-2, k


def test_oracle_metadata_channels_against_reference_golden(golden_dir):
    """oracle.meta_channels against the `meta` arrays of the reference's get_crop_memmap (batch/dataset.py:296-349) run on a
    fake echogram (oracle/make_golden_meta.py): crops inside, across every edge and outside the data, all / some
    channel kinds, and the re-centring branch of an echogram shallower than the window."""
    g = np.load(os.path.join(golden_dir, "meta_channels.npz"))
    n = 0
    for tag, ci, mtag, c, window, n_range, cfg, vec, ref in P.iter_meta_golden(g):
        got = P.meta_channels(c, window, n_range, cfg, **vec)
        assert got.shape == ref.shape, (tag, ci, mtag)
        assert np.array_equal(got, ref), (tag, ci, mtag)
        n += 1
    assert n == 20

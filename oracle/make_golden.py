"""Generates tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE from /root/reference (build container only —
the reference does not travel to the GPU box) and cross-checks oracle/ against it.  Usage: python oracle/make_golden.py

Vectors:
  unet_d2.npz        reference UNet_Baseline(3, 4, depth=2): full state_dict, input, eval logits, train logits / loss /
                     all gradients / updated BN buffers (models/unet.py + pipeline.py:135-138,176-177)
  unet_d5.npz        reference UNet_Baseline(3, 4) (the production depth): weights regenerated from seed 0 (checksums
                     stored), eval + train logits, loss, a subset of gradients
  unet_late_d2.npz   reference UNet_LateMetInject(3, 4, 2, depth=2) on the body weights of unet_d2.npz (only the head and
                     metadata-MLP tensors are stored): eval + train logits, loss, gradients
  pipeline_train.npz reference get_crop_zarr + add_noise + flip_x_axis + refine_label_boundary + convert_label_indexing +
                     remove_nan_inf + db_with_limits(_scaled) + set_data_border_value on scripted random decisions
  pipeline_small.npz reference DatasetGriddedReader (preload branch) + remove_nan_inf + db_with_limits +
                     mask_label_seabed + mask_label_overlap + fill_out_array, driven by a fake in-memory zarr reader
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/crimac_unet"
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)


def _stub_missing_modules():
    """xarray / dask / numcodecs / matplotlib are absent here and only touched at import time by the reference."""
    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Anything:
        def __getattr__(self, k):
            return _Anything()

        def __call__(self, *a, **k):
            return _Anything()

    if "xarray" not in sys.modules:
        stub("xarray", Dataset=_Anything(), DataArray=_Anything(), where=_Anything(), open_zarr=_Anything(),
             zeros_like=_Anything(), ufuncs=_Anything())
    if "dask" not in sys.modules:
        d = stub("dask")
        d.config = _Anything()
    if "numcodecs" not in sys.modules:
        stub("numcodecs", Blosc=_Anything())
    if "matplotlib" not in sys.modules:
        mp = stub("matplotlib", use=lambda *a, **k: None)
        mp.pyplot = stub("matplotlib.pyplot")
        mp.colors = stub("matplotlib.colors", LinearSegmentedColormap=_Anything(), ListedColormap=_Anything())
        mp.patches = stub("matplotlib.patches")


def load_reference_unet():
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_unet", os.path.join(REF, "models", "unet.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def checksum(t):
    return float(t.double().abs().sum())


def golden_unet(ref, depth, hw, batch, fname, store_state, grad_subset=None):
    from oracle import unet_oracle as O
    torch.manual_seed(0)
    model = ref.UNet_Baseline(3, 4, depth=depth)
    sd = O.trained_like_state(model.state_dict(), seed=0)
    model.load_state_dict(sd)
    x = O.synthetic_echogram(batch, 4, hw, hw, seed=10 + depth)
    y = O.synthetic_labels(batch, hw, hw, seed=20 + depth)
    out = {"x": x.numpy(), "y": y.numpy()}
    # eval
    model.eval()
    with torch.no_grad():
        eval_logits = model(x)
    out["eval_logits"] = eval_logits.numpy()
    chk = (O.unet_forward(sd, x, train=False) - eval_logits).abs().max().item()
    print(f"{fname}: oracle vs reference eval logits max|d| = {chk:.3g}")
    assert chk < 1e-4
    # train step exactly as pipeline.py:167-177
    model.train()
    crit = torch.nn.CrossEntropyLoss(weight=torch.tensor([10.0, 300.0, 250.0]))
    model.zero_grad()
    logits = model(x)
    loss = crit(logits, y)
    loss.backward()
    out["train_logits"] = logits.detach().numpy()
    out["loss"] = np.array(loss.item(), dtype=np.float64)
    o_logits, o_loss, o_g, o_stats = O.train_step(sd, x, y)
    print(f"{fname}: oracle vs reference train logits {(o_logits - logits).abs().max().item():.3g} loss "
          f"{abs(o_loss.item() - loss.item()):.3g}")
    after = model.state_dict()
    for name, p in model.named_parameters():
        if grad_subset is None or any(s in name for s in grad_subset):
            out["grad/" + name] = p.grad.numpy()
        d = (o_g[name] - p.grad).norm().item() / (p.grad.norm().item() + 1e-30)
        pre_bn_bias = name.endswith(".bias") and any(s in name for s in ("main.0", "main.3", "conv1", "conv2"))
        assert d < 1e-3 or pre_bn_bias, (name, d)
    for k, v in o_stats.items():
        out["stat/" + k] = after[k].numpy()
        assert torch.allclose(v.float(), after[k].float(), rtol=1e-4, atol=1e-5), k
    if store_state:
        for k, v in sd.items():
            out["state/" + k] = v.numpy()
    else:
        out["state_checksums"] = np.array([checksum(v) for v in sd.values()])
    np.savez_compressed(os.path.join(OUT, fname), **out)


def golden_late_meta_inject(ref):
    """tests/golden/unet_late_d2.npz: reference UNet_LateMetInject(3, 4, 2, depth=2) (models/unet.py:346-391) with the
    body weights of unet_d2.npz, a seeded 65-input head and metadata MLP: eval logits, train logits / loss / gradients."""
    g = np.load(os.path.join(OUT, "unet_d2.npz"))
    body = {k[6:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("state/") and not k.startswith("state/conv_final")}
    torch.manual_seed(7)
    model = ref.UNet_LateMetInject(3, 4, 2, depth=2)
    sd = model.state_dict()
    sd.update(body)
    sd["conv_final.weight"] = sd["conv_final.weight"] * 4.0          # a confident head, as in the other fixtures
    model.load_state_dict(sd)
    x, y = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
    gen = torch.Generator().manual_seed(8)
    meta = torch.rand((x.shape[0], 2, x.shape[2], x.shape[3]), generator=gen) * 2 - 0.5
    out = {"meta": meta.numpy()}
    for k, v in sd.items():
        if k.startswith("conv_final") or k.startswith("post_processing_weights"):
            out["state/" + k] = v.numpy()
    model.eval()
    with torch.no_grad():
        out["eval_logits"] = model(x, meta).numpy()
    # the decomposition the product uses, checked here against the reference in fp32 torch:
    # conv1x1_65(cat(f, m)) == conv1x1_64(f; W[:, :64], b) + W[:, 64] * m
    model.eval()
    with torch.no_grad():
        feats = x
        enc = []
        for blk in model.down_convs:
            feats, bp = blk(feats)
            enc.append(bp)
        for i, blk in enumerate(model.up_convs):
            feats = blk(enc[-(i + 2)], feats)
        w = model.conv_final.weight
        split = torch.nn.functional.conv2d(feats, w[:, :64], model.conv_final.bias) + \
            w[:, 64:, 0, 0].reshape(1, 3, 1, 1) * model.post_processing_weights(meta)
        d = (split - torch.from_numpy(out["eval_logits"])).abs().max().item()
    print(f"unet_late_d2.npz: head decomposition vs reference eval logits max|d| = {d:.3g}")
    assert d < 1e-5
    model.train()
    crit = torch.nn.CrossEntropyLoss(weight=torch.tensor([10.0, 300.0, 250.0]))
    model.zero_grad()
    logits = model(x, meta)
    loss = crit(logits, y)
    loss.backward()
    out["train_logits"] = logits.detach().numpy()
    out["loss"] = np.array(loss.item(), dtype=np.float64)
    for name, p_ in model.named_parameters():
        if name.startswith(("conv_final", "post_processing_weights")) or name in (
                "up_convs.0.bn2.weight", "up_convs.0.bn2.bias", "up_convs.0.conv2.weight", "down_convs.0.main.0.weight",
                "down_convs.1.main.3.weight", "up_convs.0.upconv.weight"):
            out["grad/" + name] = p_.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "unet_late_d2.npz"), **out)


class FakeZarrReader:
    """In-memory stand-in for data/data_reader.py:DataReaderZarr exposing only what DatasetGriddedReader, the
    transforms and fill_out_array touch."""
    data_format = "zarr"

    def __init__(self, sv, labels, seabed_idx):
        self.sv = sv              # (F, pings, range)
        self.labels = labels      # (pings, range)
        self.seabed_idx = seabed_idx
        self.shape = (sv.shape[1], sv.shape[2])
        self.name = "fake"
        self.time_vector = np.arange(sv.shape[1])
        self.range_vector = np.arange(sv.shape[2])

    def get_data_slice(self, idx_ping, n_pings, idx_range=None, n_range=None, frequencies=None, drop_na=False,
                       return_numpy=True):
        r0 = 0 if idx_range is None else idx_range
        r1 = self.shape[1] if n_range is None else r0 + n_range
        return self.sv[:, idx_ping:idx_ping + n_pings, r0:r1].copy()

    def get_label_slice(self, idx_ping, n_pings, idx_range=None, n_range=None, return_numpy=True, **kw):
        r0 = 0 if idx_range is None else idx_range
        r1 = self.shape[1] if n_range is None else r0 + n_range
        return self.labels[idx_ping:idx_ping + n_pings, r0:r1].copy()

    def get_seabed(self, idx_ping, n_pings=1, idx_range=None, n_range=None, return_numpy=True):
        s = self.seabed_idx[idx_ping:idx_ping + n_pings]
        if return_numpy:
            return s.astype(int)

        class _V:
            def __init__(self, v):
                self.values = v

            def max(self):
                return _V(self.values.max())
        return _V(s)

    def get_seabed_mask(self, idx_ping, n_pings, idx_range=None, n_range=None, return_numpy=False, seabed_pad=0):
        idx_range = 0 if idx_range is None else idx_range
        hi = self.shape[1] if n_range is None else idx_range + n_range
        rng = np.arange(idx_range, min(hi, self.shape[1]))[None, :]
        m = (rng >= self.seabed_idx[idx_ping:idx_ping + n_pings, None]).astype(float)   # (pings, range)
        if seabed_pad != 0:
            mp = np.zeros_like(m)
            mp[:, seabed_pad:] = m[:, :-seabed_pad]
            return mp
        return m


def golden_pipeline():
    from oracle import pipeline_oracle as P
    _stub_missing_modules()
    sys.path.insert(0, REF)
    from batch.dataset import DatasetGriddedReader
    from batch.data_transforms.remove_nan_inf import remove_nan_inf
    from batch.data_transforms.db_with_limits import db_with_limits
    from batch.label_transforms.mask_label_seabed import mask_label_seabed
    from batch.label_transforms.mask_label_overlap import mask_label_overlap
    from utils.combine_functions import CombineFunctions
    from utils.preload_data_split import get_data_split
    from pipeline_train_predict.save_predict import fill_out_array

    rng = np.random.default_rng(5)
    F, NP, R = 2, 700, 96
    patch, overlap, preload = [64, 64], 8, 300
    sv = (10.0 ** rng.uniform(-9, -2, size=(F, NP, R))).astype(np.float32)
    sv[0, 100:104, 10:30] = np.nan
    sv[1, 250:252, 40:50] = np.inf
    sv[0, 411, 5] = -np.inf
    labels = np.zeros((NP, R), dtype=np.float32)
    labels[120:140, 20:30] = 1
    labels[300:320, 60:90] = 2          # partly below the seabed: fish labels win over the seabed mask
    seabed = (70 + 10 * np.sin(2 * np.pi * np.arange(NP) / 180.0)).astype(int)
    reader = FakeZarrReader(sv, labels, seabed)

    splits = get_data_split([[0, NP]], preload)
    assert np.array_equal(splits, P.get_data_split([[0, NP]], preload))
    out = {"sv": sv, "labels": labels, "seabed": seabed, "patch": np.array(patch), "overlap": np.array(overlap),
           "splits": splits}
    for ci, (s, e) in enumerate(splits):
        ds = DatasetGriddedReader(reader, patch, [18, 38], grid_start=int(s), grid_end=int(e), patch_overlap=overlap,
                                  data_preload=True,
                                  label_transform_function=CombineFunctions([mask_label_seabed(),
                                                                             mask_label_overlap(overlap=overlap)]),
                                  data_transform_function=CombineFunctions([remove_nan_inf, db_with_limits]))
        assert ds.data_preload
        grid = np.array(ds.data_grid)
        er = P.end_range_from_seabed(R, seabed[s:e])
        ogrid = P.get_data_grid(int(s), int(e), 0, er, patch, overlap)
        assert np.array_equal(grid, ogrid), (grid[:4], ogrid[:4])
        d0, d1 = P.preload_extents(ogrid, NP, patch[1])
        assert (d0, d1) == (ds.data_preload_start, ds.data_preload_end)
        out_array = np.zeros([2, R, e - s])
        o_out = np.zeros([2, R, e - s])
        datas, labs = [], []
        sv_pre = np.swapaxes(sv[:, d0:d1, :], 1, 2)
        lab_chunk = labels[s:e, :].T
        for i in range(len(ds)):
            item = ds[i]
            datas.append(item["data"])
            labs.append(item["labels"])
            c = item["center_coordinates"]
            probs = np.stack([np.full(patch, 0.1 * k + 0.001 * i) + 1e-4 * np.arange(patch[1])[None, :] for k in range(3)])
            fill_out_array(out_array, probs, item["labels"], c, int(s))
            od, ol = P.patch_item(sv_pre, d0, lab_chunk, int(s), c, seabed, R, NP, tuple(patch), overlap)
            assert np.array_equal(ol, item["labels"]), (ci, i)
            assert np.allclose(od, item["data"], rtol=0, atol=1e-4), (ci, i, np.abs(od - item["data"]).max())
            P.fill_out_array(o_out, probs, ol, c, int(s))
        assert np.array_equal(out_array, o_out)
        out[f"chunk{ci}/grid"] = grid
        out[f"chunk{ci}/preload"] = np.array([d0, d1])
        out[f"chunk{ci}/data"] = np.stack(datas).astype(np.float32)
        out[f"chunk{ci}/labels"] = np.stack(labs).astype(np.int16)
        out[f"chunk{ci}/stitched"] = out_array.astype(np.float32)
        print(f"pipeline chunk {ci}: pings [{s},{e}) patches={len(ds)} written={(out_array[0] != 0).mean() * 100:.1f}%")
    np.savez_compressed(os.path.join(OUT, "pipeline_small.npz"), **out)


class _ScriptedRandom:
    """Stands in for numpy's global generator while the reference's add_noise / flip_x_axis run, so that their random
    decisions are known: randint -> the scripted coin, binomial / uniform -> the scripted fields."""

    def __init__(self):
        self.coins, self.binomials, self.uniforms = [], [], []

    def __enter__(self):
        self.saved = (np.random.randint, np.random.binomial, np.random.uniform)
        np.random.randint = lambda *a, **k: self.coins.pop(0)
        np.random.binomial = lambda *a, **k: self.binomials.pop(0)
        np.random.uniform = lambda *a, **k: self.uniforms.pop(0)
        return self

    def __exit__(self, *exc):
        np.random.randint, np.random.binomial, np.random.uniform = self.saved


def train_survey(seed=11, F=3, NP=150, R=110):
    """Small synthetic survey for the training-sample path: schools of sandeel (27), other (1) and an unused species
    (12) whose threshold-frequency response straddles refine_label_boundary's (1e-7, 1e-4) window, NaN / inf holes."""
    rng = np.random.default_rng(seed)
    sv = (10.0 ** rng.uniform(-9, -2, size=(F, NP, R))).astype(np.float32)
    labels = np.zeros((NP, R), dtype=np.float32)
    yy, xx = np.meshgrid(np.arange(NP), np.arange(R), indexing="ij")
    for (cy, cx, ry, rx, val) in [(30, 40, 14, 20, 27), (70, 75, 18, 12, 1), (110, 30, 10, 16, 12), (5, 100, 9, 9, 27),
                                  (140, 8, 12, 10, 1)]:
        blob = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1
        labels[blob] = val
        # inside a school the last frequency is mostly inside the window, with speckle outside it
        inside = 10.0 ** rng.uniform(-6.8, -4.2, size=blob.sum())
        speckle = rng.random(blob.sum())
        inside = np.where(speckle < 0.25, 10.0 ** rng.uniform(-9, -7.2, size=blob.sum()), inside)
        inside = np.where(speckle > 0.9, 10.0 ** rng.uniform(-3.9, -2, size=blob.sum()), inside)
        sv[F - 1][blob] = inside.astype(np.float32)
    sv[0, 60:63, 20:50] = np.nan
    sv[F - 1, 28:31, 35:45] = np.nan
    sv[1, 100, 60:64] = np.inf
    labels[90:94, 90:100] = np.nan
    labels[20:24, 100:104] = -1
    return sv, labels


def golden_train_pipeline():
    """tests/golden/pipeline_train.npz: the reference's Dataset.__getitem__ composition (dataset.py:75-108 with
    train.py:56-60's functions) on scripted random decisions."""
    from oracle import pipeline_oracle as P
    _stub_missing_modules()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from batch.dataset import get_crop_zarr
    from batch.data_augmentation.add_noise import add_noise
    from batch.data_augmentation.flip_x_axis import flip_x_axis
    from batch.label_transforms.refine_label_boundary import refine_label_boundary
    from batch.label_transforms.convert_label_indexing import convert_label_indexing
    from batch.data_transforms.remove_nan_inf import remove_nan_inf
    from batch.data_transforms.db_with_limits import db_with_limits, db_with_limits_scaled
    from batch.data_transforms.set_data_border_value import set_data_border_value

    sv, labels = train_survey()
    F, NP, R = sv.shape
    freqs = [18, 38, 200]
    reader = FakeZarrReader(sv, labels, np.zeros(NP, dtype=int))
    patch = [64, 64]
    centres = np.array([[40, 30], [75, 70], [30, 110], [100, 5], [8, 140], [55, 75], [105, 149], [400, 400],
                        [31, 31], [78, 118]])
    rng = np.random.default_rng(3)
    noise_on = np.array([1, 0, 1, 1, 0, 1, 0, 1, 0, 1])
    flip = np.array([0, 1, 1, 0, 0, 1, 1, 0, 0, 1])
    refine = refine_label_boundary(frequencies=freqs, threshold_freq=freqs[-1])
    out = {"sv": sv, "labels": labels, "centres": centres.astype(np.int32), "noise_on": noise_on.astype(np.uint8),
           "flip": flip.astype(np.uint8), "patch": np.array(patch)}
    mults, datas, labs, datas_scaled, datas_border = [], [], [], [], []
    for i, c in enumerate(centres):
        shape = (F, patch[0], patch[1])
        change = rng.binomial(1, 0.05, shape)
        inc = rng.binomial(1, 0.5, shape)
        u_hi = rng.uniform(1, 10, shape).astype(np.float32).astype(np.float64)
        u_lo = rng.uniform(0, 1, shape).astype(np.float32).astype(np.float64)
        mult = np.where(change == 1, np.where(inc == 1, u_hi, u_lo), 1.0).astype(np.float32)
        data, lab = get_crop_zarr(reader, list(c), patch, freqs)
        with _ScriptedRandom() as sr:
            sr.coins = [int(noise_on[i])]
            sr.binomials = [change, inc]
            sr.uniforms = [u_hi, u_lo]
            data, lab, _ = add_noise(data, lab, reader)
            sr.coins = [int(flip[i])]
            data, lab, _ = flip_x_axis(data, lab, reader)
        data, lab, _, _ = refine(data, lab, list(c), reader)
        data, lab, _, _ = convert_label_indexing(data, lab, list(c), reader)
        data, lab, _, _ = remove_nan_inf(data, lab, reader, freqs)
        with np.errstate(invalid="ignore"):
            d_plain, _, _, _ = db_with_limits(data.copy(), lab, reader, freqs)
            d_scaled, _, _, _ = db_with_limits_scaled(data.copy(), lab, reader, freqs)
        d_border, _, _, _ = set_data_border_value(d_plain.copy(), lab, reader, freqs)
        mults.append(mult)
        datas.append(d_plain.astype(np.float32))
        datas_scaled.append(d_scaled.astype(np.float32))
        datas_border.append(d_border.astype(np.float32))
        labs.append(lab.astype(np.int64))
        # the oracle restatement must agree with the reference on the same decisions
        od, ol = P.train_patch_item(sv, labels, c, noise_on[i], flip[i], mult, tuple(patch))
        assert np.array_equal(ol, labs[-1]), (i, (ol != labs[-1]).sum())
        assert np.array_equal(np.isnan(od), np.isnan(datas[-1])) and np.allclose(od, datas[-1], atol=1e-5, equal_nan=True), i
        od, _ = P.train_patch_item(sv, labels, c, noise_on[i], flip[i], mult, tuple(patch), scaled=True)
        assert np.allclose(od, datas_scaled[-1], atol=1e-6, equal_nan=True), i
        od, _ = P.train_patch_item(sv, labels, c, noise_on[i], flip[i], mult, tuple(patch), border_zero=True)
        assert np.allclose(od, datas_border[-1], atol=1e-5, equal_nan=True), i
        vals, cnt = np.unique(labs[-1], return_counts=True)
        print(f"train crop {i} centre {c.tolist()} noise {noise_on[i]} flip {flip[i]} labels {dict(zip(vals.tolist(), cnt.tolist()))}")
    out["mult"] = np.stack(mults)
    out["data"] = np.stack(datas)
    out["data_scaled"] = np.stack(datas_scaled)
    out["data_border"] = np.stack(datas_border)
    out["out_labels"] = np.stack(labs).astype(np.int16)
    np.savez_compressed(os.path.join(OUT, "pipeline_train.npz"), **out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    ref = load_reference_unet()
    golden_unet(ref, depth=2, hw=32, batch=2, fname="unet_d2.npz", store_state=True)
    golden_unet(ref, depth=5, hw=64, batch=2, fname="unet_d5.npz", store_state=False,
                grad_subset=("conv_final", "main.1.", "main.4.", "bn1", "bn2", "down_convs.0.main.0", "upconv.bias"))
    golden_late_meta_inject(ref)
    golden_pipeline()
    golden_train_pipeline()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")

"""Generates tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE from /root/reference (build container only —
the reference does not travel to the GPU box) and cross-checks oracle/ against it.  Usage: python oracle/make_golden.py

Vectors:
  unet_d2.npz        reference UNet_Baseline(3, 4, depth=2): full state_dict, input, eval logits, train logits / loss /
                     all gradients / updated BN buffers (models/unet.py + pipeline.py:135-138,176-177)
  unet_d5.npz        reference UNet_Baseline(3, 4) (the production depth): weights regenerated from seed 0 (checksums
                     stored), eval + train logits, loss, a subset of gradients
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

    def get_data_slice(self, idx_ping, n_pings, frequencies=None, drop_na=False, return_numpy=True):
        return self.sv[:, idx_ping:idx_ping + n_pings, :].copy()

    def get_label_slice(self, idx_ping, n_pings, return_numpy=True, **kw):
        return self.labels[idx_ping:idx_ping + n_pings, :].copy()

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


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    ref = load_reference_unet()
    golden_unet(ref, depth=2, hw=32, batch=2, fname="unet_d2.npz", store_state=True)
    golden_unet(ref, depth=5, hw=64, batch=2, fname="unet_d5.npz", store_state=False,
                grad_subset=("conv_final", "main.1.", "main.4.", "bn1", "bn2", "down_convs.0.main.0", "upconv.bias"))
    golden_pipeline()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")

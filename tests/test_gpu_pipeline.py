"""-m gpu: preprocessing (patch gather + sv->dB) and overlap stitching kernels, and the sliding-window driver,
against the golden vectors produced by the reference's own dataset / fill_out_array code and against the oracle."""
import importlib
import os

import numpy as np
import pytest
import torch

from oracle import pipeline_oracle as P
from oracle import unet_oracle as O

pytestmark = pytest.mark.gpu
dev = torch.device("cuda:0")


@pytest.fixture(scope="module")
def E(pkg):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return importlib.import_module("crimac_unet_b200.engine")


def test_preprocess_and_stitch_against_reference_golden(E, golden_dir):
    g = np.load(os.path.join(golden_dir, "pipeline_small.npz"))
    sv, labels, seabed = g["sv"], g["labels"], g["seabed"]
    patch, overlap = tuple(int(v) for v in g["patch"]), int(g["overlap"])
    F_, NP, R = sv.shape
    for ci, (s, e) in enumerate(g["splits"]):
        s, e = int(s), int(e)
        grid = g[f"chunk{ci}/grid"]
        d0, d1 = (int(v) for v in g[f"chunk{ci}/preload"])
        sv_dev = torch.from_numpy(np.ascontiguousarray(np.swapaxes(sv[:, d0:d1, :], 1, 2))).to(dev)   # (F, R, P)
        centres = torch.from_numpy(grid.astype(np.int32)).to(dev)
        x, nan_mask = E.preprocess(sv_dev, d0, centres, patch)
        ref = torch.from_numpy(g[f"chunk{ci}/data"]).to(dev)
        assert (x - ref).abs().max().item() <= 1e-4          # fp32 log10: a few ulp between numpy and CUDA
        ref_lab = g[f"chunk{ci}/labels"]
        # stitch with the same fake probabilities the golden used
        n = len(grid)
        probs = torch.stack([torch.stack([torch.full(patch, 0.1 * k + 0.001 * i) + 1e-4 * torch.arange(patch[1])[None, :]
                                          for k in range(3)]) for i in range(n)]).float().to(dev)
        out = torch.zeros((2, R, e - s), dtype=torch.float16, device=dev)
        lab_chunk = torch.from_numpy(np.ascontiguousarray(labels[s:e, :].T).astype(np.int16)).to(dev)
        sb = torch.from_numpy(seabed[s:e].astype(np.int32)).to(dev)
        E.stitch(probs, centres, nan_mask, out, s, overlap, labels=lab_chunk, seabed=sb, seabed_pad=10)
        ref_out = g[f"chunk{ci}/stitched"]
        assert np.array_equal(out.cpu().numpy() != 0, ref_out != 0)          # exactly the same pixels are written
        assert np.array_equal(out.cpu().numpy(), ref_out.astype(np.float16))  # and with the same (fp16-cast) values
        # the written set is exactly "label not in {-70,-50,-100}" of the reference's label patches
        keep = sum(int(((l != -70) & (l != -50) & (l != -100)).sum()) for l in ref_lab)
        assert keep == int((ref_out[0] != 0).sum())


def test_preprocess_edge_cases(E):
    F_, R, Pn = 3, 40, 50
    sv = torch.full((F_, R, Pn), 1e-3, device=dev)
    sv[0, 5, 5] = float("nan")
    sv[1, 6, 6] = float("inf")
    sv[2, 7, 7] = 5.0          # > 0 dB -> clipped to 0
    sv[2, 8, 8] = 0.0          # -100 dB -> clipped to -75
    centres = torch.tensor([[15, 15], [-100, -100], [39, 49]], dtype=torch.int32, device=dev)
    x, nan = E.preprocess(sv, 0, centres, (32, 32))
    off = 15 - 16 + 1
    assert x[0, 0, 5 - off, 5 - off].item() == -75.0 and nan[0, 5 - off, 5 - off].item() == 1
    assert x[0, 1, 6 - off, 6 - off].item() == -75.0 and nan[0, 6 - off, 6 - off].item() == 0   # only frequency 0 flags
    assert x[0, 2, 7 - off, 7 - off].item() == 0.0 and x[0, 2, 8 - off, 8 - off].item() == -75.0
    assert abs(x[0, 0, 0, 0].item() + 30.0) < 1e-4
    assert torch.all(x[1] == -75.0) and nan[1].sum().item() == 0        # patch entirely outside the data
    assert torch.all(x[2][:, 17:, :] == -75.0)                           # rows below the last range bin


def test_sliding_window_driver_matches_oracle_pipeline(E, pkg):
    """One small survey end to end: product (preprocess -> tcgen05 forward -> stitch) vs oracle
    (numpy gather/transform/masks -> fp32 torch forward -> fill_out_array), same weights."""
    Mm = importlib.import_module("crimac_unet_b200.models.unet")
    Pr = importlib.import_module("crimac_unet_b200.predict")
    torch.manual_seed(0)
    rng = np.random.default_rng(0)
    Fq, NP, R, patch, ov, preload = 4, 900, 128, (64, 64), 8, 400
    sv = (10.0 ** rng.uniform(-9, -2, size=(Fq, R, NP))).astype(np.float32)
    sv[0, 30:34, 100:120] = np.nan
    seabed = (100 + 10 * np.sin(np.arange(NP) / 40.0)).astype(np.int32)
    m = Mm.UNet_Baseline(3, Fq, depth=3)
    m.load_state_dict(O.trained_like_state(m.state_dict(), 0, head_gain=2.0))
    m = m.to(dev).eval()
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    pred = Pr.SurveyPredictor(m, patch, ov, preload, batch_size=8)
    sv_dev = torch.from_numpy(sv).to(dev)
    sb_dev = torch.from_numpy(seabed).to(dev)

    def load(d0, d1, s, e):
        return sv_dev[:, :, d0:d1].contiguous(), None, sb_dev[s:e].contiguous()

    got = {(s, e): o.float().cpu().numpy() for s, e, o in
           pred.predict_survey(load, NP, R, seabed_max_of=lambda s, e: int(seabed[s:e].max()))}
    assert [k for k in got] == [tuple(int(v) for v in r) for r in P.get_data_split([[0, NP]], preload)]
    for (s, e), o in got.items():
        grid = P.get_data_grid(s, e, 0, P.end_range_from_seabed(R, seabed[s:e]), patch, ov)
        d0, d1 = P.preload_extents(grid, NP, patch[1])
        ref = np.zeros((2, R, e - s))
        for c in grid:
            d, l = P.patch_item(sv[:, :, d0:d1], d0, np.zeros((R, e - s)), s, c, seabed, R, NP, patch, ov)
            with torch.no_grad():
                p = O.softmax_probs(O.unet_forward(sd, torch.from_numpy(d)[None])).numpy()[0]
            P.fill_out_array(ref, p, l, c, s)
        assert np.array_equal(o != 0, ref.astype(np.float16) != 0) or np.mean((o != 0) != (ref != 0)) < 1e-4
        assert np.abs(o - ref).max() <= 2.5e-2                         # 2e-2 (bf16) + fp16 output cast


def test_sharding_covers_the_survey_once(pkg):
    Pr = importlib.import_module("crimac_unet_b200.predict")
    chunks = Pr.split_pings(0, 100000, 20000)
    seen = []
    for r in range(4):
        seen += Pr.shard_chunks(chunks, 4, r)
    assert seen == chunks

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


@pytest.mark.parametrize("Fq", [4, 6, 1])
def test_direct_preprocessing_feeds_the_first_conv_bit_identically(E, pkg, Fq):
    """north_star: the preprocessing kernel feeds the first conv directly.  crimac_preprocess_staged writes the bf16
    hi/lo NHWC operand of the tensor-core first conv; the result must equal - bit for bit - the two-step form
    (crimac_preprocess -> fp32 NCHW patches -> the first conv's own split pass), incl. out-of-data patches, NaN / inf
    samples and ragged edges, for the one-plane (<= 4 frequencies) and the two-plane (5..8) layout."""
    Mm = importlib.import_module("crimac_unet_b200.models.unet")
    torch.manual_seed(Fq)
    R, Pn, patch = 100, 300, (64, 64)
    g = torch.Generator(device=dev).manual_seed(1)
    sv = torch.pow(10.0, torch.rand((Fq, R, Pn), device=dev, generator=g) * 7.0 - 9.0)
    sv[0, 10:14, 20:40] = float("nan")
    sv[Fq - 1, 50, 60] = float("inf")
    sv[0, 70, 70] = -1.0                                           # negative sv -> log10 of a negative number -> NaN -> clipped like the reference
    centres = torch.tensor([[31, 31], [31, 95], [95, 250], [-200, -200], [99, 299], [40, 5]], dtype=torch.int32, device=dev)
    m = Mm.UNet_Baseline(3, Fq, depth=3)
    m.load_state_dict(O.trained_like_state(m.state_dict(), 0, head_gain=2.0))
    m = m.to(dev).eval()
    x, nan_a = E.preprocess(sv, 7, centres, patch)
    with torch.no_grad():
        a = m.predict_proba(x)
        b, nan_b = m.predict_proba_patches(sv, 7, centres, patch)
        c, _ = m.predict_proba_patches(sv, 7, centres[:2].contiguous(), patch)      # smaller batch in the same context
    assert torch.equal(nan_a, nan_b)
    assert torch.equal(torch.nan_to_num(a, nan=-1.0), torch.nan_to_num(b, nan=-1.0))
    assert torch.equal(torch.nan_to_num(b[:2], nan=-1.0), torch.nan_to_num(c, nan=-1.0))


def test_full_size_chunk_of_config3_against_the_oracle_pipeline(E, pkg):
    """BASELINE configs[3] at FULL size for one preload chunk: pings [60000, 80000) of the synthetic
    1 M-ping x 256-range survey (4 frequencies, 0.1 % NaNs, seabed 200 + 20 sin), preload_n_pings = 20000, 256x256
    patches, overlap 20 -> 186 patches incl. the second patch row that is mostly below the data
    (batch/samplers/gridded.py:40-47), neighbouring chunks' pings as context (batch/dataset.py:176-184).  Product =
    SurveyPredictor (direct preprocessing -> tcgen05 forward -> stitch); oracle = numpy gather / transforms / label
    masks per patch (oracle.patch_item) -> fp32 torch forward -> fill_out_array (save_predict.py:41-65).
    Bounds: identical written-pixel set, values within 2.5e-2 (2e-2 bf16 + fp16 output)."""
    Mm = importlib.import_module("crimac_unet_b200.models.unet")
    Pr = importlib.import_module("crimac_unet_b200.predict")
    S = importlib.import_module("crimac_unet_b200.synthetic")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    NP, R, patch, ov, preload = 1_000_000, 256, (256, 256), 20, 20000
    s, e = 60000, 80000
    torch.manual_seed(0)
    m = Mm.UNet_Baseline(3, 4)
    m.load_state_dict(O.trained_like_state(m.state_dict(), 0, head_gain=2.0))
    m = m.to(dev).eval()
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    pred = Pr.SurveyPredictor(m, patch, ov, preload, batch_size=62)
    seabed_all = S.synthetic_seabed(0, NP).numpy()
    grid, (d0, d1) = pred.chunk_geometry(s, e, R, NP, seabed_max=int(seabed_all[s:e].max()))
    ref_grid = P.get_data_grid(s, e, 0, P.end_range_from_seabed(R, seabed_all[s:e]), patch, ov)
    assert np.array_equal(grid, ref_grid) and len(grid) == 186
    assert (d0, d1) == tuple(P.preload_extents(ref_grid, NP, patch[1]))
    sv_dev = S.synthetic_survey_pings(4, R, d0, d1, seed=5, device=dev)
    sb_dev = torch.from_numpy(seabed_all[s:e].astype(np.int32)).to(dev)
    got = pred.predict_chunk(sv_dev, d0, grid, s, e, seabed=sb_dev).float().cpu().numpy()
    # ---- oracle
    sv_np = sv_dev.cpu().numpy()
    ref = np.zeros((2, R, e - s))
    lab0 = np.zeros((R, e - s))
    items = [P.patch_item(sv_np, d0, lab0, s, c, seabed_all, R, NP, patch, ov) for c in grid]
    with torch.no_grad():
        for i in range(0, len(items), 31):
            xb = torch.from_numpy(np.stack([d for d, _ in items[i:i + 31]])).to(dev)
            pb = O.softmax_probs(O.unet_forward(sd, xb)).cpu().numpy()
            for j, p in enumerate(pb):
                P.fill_out_array(ref, p, items[i + j][1], grid[i + j], s)
    written, ref_written = got[0] != 0, ref[0] != 0
    frac = ref_written.mean()
    print(f"configs[3] chunk [{s},{e}): {len(grid)} patches, {frac:.4f} of the output pixels written; max |d| {np.abs(got - ref).max():.4f}")
    assert np.array_equal(written, ref_written)
    assert 0.7 < frac < 0.9                                            # rows below seabed + 10 and NaN pixels stay 0
    assert np.abs(got - ref).max() <= 2.5e-2


def test_metadata_channels_kernel_against_reference_golden(E, golden_dir):
    """crimac_meta_channels against the reference's get_crop_memmap `meta` arrays (batch/dataset.py:296-349; fixture
    made by oracle/make_golden_meta.py), written straight into planes [F, F+M) of a network input tensor."""
    g = np.load(os.path.join(golden_dir, "meta_channels.npz"))
    worst = 0.0
    for tag, ci, mtag, c, window, n_range, cfg, vec, ref in P.iter_meta_golden(g):
        M_, F_ = ref.shape[0], 4
        x = torch.full((2, F_ + M_, window[0], window[1]), 7.0, device=dev)
        centres = torch.tensor([list(c), list(c)], dtype=torch.int32, device=dev)
        wrote = E.meta_channels(x, F_, centres, cfg, portion_year=vec["portion_year"],
                                portion_of_day=torch.from_numpy(vec["portion_of_day"]).to(dev),
                                time_diff=torch.from_numpy(vec["time_diff"]).to(dev),
                                seabed=torch.from_numpy(vec["seabed"]).to(dev), n_range=n_range)
        assert wrote == M_
        assert torch.all(x[:, :F_] == 7.0)                       # the frequency planes are not touched
        want = torch.from_numpy(ref).float().to(dev)             # what the model sees: batch['data'].float()
        assert torch.equal(x[0, F_:], x[1, F_:])
        worst = max(worst, ((x[0, F_:] - want).abs() / want.abs().clamp_min(1.0)).max().item())
    print(f"metadata channels: worst relative deviation {worst:.2e}")
    assert worst <= 2e-7                                         # double arithmetic on both sides, one fp32 rounding


def test_stitching_in_the_head_epilogue_equals_the_three_step_form(E, pkg):
    """SurveyPredictor's default path (preprocessing -> first conv operand, forward, stitch in the last conv's epilogue:
    crimac_preprocess_staged + crimac_forward_infer_stitch) against the three separate steps (crimac_preprocess,
    predict_proba, crimac_stitch) on a survey with NaNs, a seabed, chunk labels carrying every mask code and patches
    hanging over all four edges: the stitched fp16 arrays must be bit-identical."""
    Mm = importlib.import_module("crimac_unet_b200.models.unet")
    Pr = importlib.import_module("crimac_unet_b200.predict")
    torch.manual_seed(0)
    rng = np.random.default_rng(1)
    Fq, NP, R, patch, ov, preload = 4, 700, 150, (64, 64), 8, 300
    sv = (10.0 ** rng.uniform(-9, -2, size=(Fq, R, NP))).astype(np.float32)
    sv[0, 40:44, 90:130] = np.nan
    seabed = (110 + 15 * np.sin(np.arange(NP) / 30.0)).astype(np.int32)
    labels = rng.choice(np.array([0, 0, 0, 1, 2, -100, -70, -50], dtype=np.int16), size=(R, NP))
    m = Mm.UNet_Baseline(3, Fq, depth=3)
    m.load_state_dict(O.trained_like_state(m.state_dict(), 0, head_gain=2.0))
    m = m.to(dev).eval()
    sv_dev, sb_dev, lab_dev = torch.from_numpy(sv).to(dev), torch.from_numpy(seabed).to(dev), torch.from_numpy(labels).to(dev)

    def load(d0, d1, s, e):
        return sv_dev[:, :, d0:d1].contiguous(), lab_dev[:, s:e].contiguous(), sb_dev[s:e].contiguous()

    outs = {}
    for direct in (True, False):
        pred = Pr.SurveyPredictor(m, patch, ov, preload, batch_size=7, direct=direct)
        outs[direct] = {(s, e): o.clone() for s, e, o in
                        pred.predict_survey(load, NP, R, seabed_max_of=lambda s, e: int(seabed[s:e].max()))}
    assert outs[True].keys() == outs[False].keys() and len(outs[True]) == 3
    for k in outs[True]:
        assert torch.equal(outs[True][k], outs[False][k])
        assert (outs[True][k] != 0).float().mean().item() > 0.2


def test_survey_edge_cases_short_and_shallow(E, pkg):
    """Ragged surveys (the reference's loop has no special cases for them, save_predict.py:160-171): fewer pings than one
    patch width, fewer range bins than one patch height, a last chunk shorter than preload_n_pings, no seabed / labels.
    Product (fused path) against the oracle pipeline: same written-pixel set, values within 2.5e-2."""
    Mm = importlib.import_module("crimac_unet_b200.models.unet")
    Pr = importlib.import_module("crimac_unet_b200.predict")
    torch.manual_seed(0)
    m = Mm.UNet_Baseline(3, 4, depth=3)
    m.load_state_dict(O.trained_like_state(m.state_dict(), 0, head_gain=2.0))
    m = m.to(dev).eval()
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    patch, ov = (64, 64), 8
    for NP, R, preload in ((50, 40, 100), (230, 70, 100)):
        rng = np.random.default_rng(NP)
        sv = (10.0 ** rng.uniform(-9, -2, size=(4, R, NP))).astype(np.float32)
        sv_dev = torch.from_numpy(sv).to(dev)
        pred = Pr.SurveyPredictor(m, patch, ov, preload, batch_size=5)
        got = {(s, e): o.float().cpu().numpy() for s, e, o in
               pred.predict_survey(lambda d0, d1, s, e: sv_dev[:, :, d0:d1].contiguous(), NP, R)}
        assert [k for k in got] == [tuple(int(v) for v in r) for r in P.get_data_split([[0, NP]], preload)]
        for (s, e), o in got.items():
            grid = P.get_data_grid(s, e, 0, R, patch, ov)
            d0, d1 = P.preload_extents(grid, NP, patch[1])
            ref = np.zeros((2, R, e - s))
            for c in grid:
                d = P.gather_data(sv[:, :, d0:d1], c, d0, patch)
                l = P.mask_overlap(P.gather_labels(np.zeros((R, e - s)), c, s, patch), ov)
                d, l = P.data_transform(d, l)
                with torch.no_grad():
                    p = O.softmax_probs(O.unet_forward(sd, torch.from_numpy(d.astype(np.float32))[None])).numpy()[0]
                P.fill_out_array(ref, p, l.astype(np.int16), c, s)
            assert np.array_equal(o != 0, ref.astype(np.float16) != 0)
            assert (o != 0).all()                                     # no seabed, no NaNs: every pixel of the chunk is written
            assert np.abs(o - ref).max() <= 2.5e-2

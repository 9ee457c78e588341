"""CPU-side checks: the C-ABI library builds for sm_100a, loads and exports every declared symbol (no compute calls),
argument validation, the nn.Module surface the reference's callers rely on, and the host-side logic (patch grid,
ping-range sharding, data-parallel gradient exchange with gloo at world_size 2)."""
import ctypes
import importlib
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import pipeline_oracle as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol(pkg):
    lib = importlib.import_module("crimac_unet_b200.lib").load()
    header = open(os.path.join(ROOT, "include", "crimac_b200.h")).read()
    syms = sorted(set(re.findall(r"\b(crimac_[a-z0-9_]+)\s*\(", header)))
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), s


def test_library_contains_blackwell_sass(pkg):
    """tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG in the built cubin (B200_PROFILING.md)."""
    so = os.path.join(ROOT, "crimac-classifiers-unet_b200", "libcrimac_b200.so")
    out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    if not out:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out
    for mnem in ("UTCHMMA", "LDTM", "UTMALDG"):
        assert mnem in out, mnem


def test_workspace_and_argument_validation(pkg):
    E = importlib.import_module("crimac_unet_b200.engine")
    L = importlib.import_module("crimac_unet_b200.lib")
    train = E.workspace_bytes(4, 3, 5, 64, 32, 256, 256, 1)
    infer = E.workspace_bytes(4, 3, 5, 64, 32, 256, 256, 0)
    assert 4e9 < train < 9e9 and infer < train / 2
    det = E.workspace_bytes(4, 3, 5, 64, 32, 256, 256, 1, deterministic=True)
    assert 0.5e9 < det - train < 1.5e9          # per-split slabs of the fixed-order weight-gradient reduction
    with pytest.raises(L.CrimacError, match="start_filts"):
        E.workspace_bytes(4, 3, 5, 32, 1, 256, 256, 0)
    with pytest.raises(L.CrimacError, match="multiples"):
        E.workspace_bytes(4, 3, 5, 64, 1, 250, 256, 0)
    with pytest.raises(L.CrimacError, match="in_channels"):
        E.workspace_bytes(13, 3, 5, 64, 1, 256, 256, 0)          # 4 frequencies + 7 metadata channels = 11 is the most the reference builds
    with pytest.raises(L.CrimacError, match="incompatible"):
        E.workspace_bytes(4, 3, 5, 64, 1, 256, 256, 0, up_mode="upsample", merge_mode="add")
    assert E.workspace_bytes(4, 3, 5, 64, 8, 256, 256, 1, merge_mode="add") < E.workspace_bytes(4, 3, 5, 64, 8, 256, 256, 1)
    lib = L.load()
    cfg = E._Config(4, 3, 5, 64, 1, 256, 256, 1, 0, 0, 0)
    assert lib.crimac_state_count(ctypes.byref(cfg)) == 136
    assert lib.crimac_grad_count(ctypes.byref(cfg)) == 82


def test_module_surface_matches_reference_contract(pkg, golden_dir):
    M = importlib.import_module("crimac_unet_b200.models.unet")
    for name in ("conv3x3", "upconv2x2", "conv1x1", "DownConv", "UpConv", "MetaPostProcessing", "UNet", "UNet_Baseline",
                 "UNet_LateMetInject"):
        assert hasattr(M, name)
    m = M.UNet_Baseline(3, 4)
    sd = m.state_dict()
    assert len(sd) == 136 and sum(v.numel() for v in sd.values()) == 31056021
    assert sum(p.numel() for p in m.parameters()) == 31044227 and len(list(m.parameters())) == 82
    # key names / shapes of a reference-generated state_dict (depth 2 golden) load strictly
    g = np.load(os.path.join(golden_dir, "unet_d2.npz"))
    ref_sd = {k[6:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("state/")}
    m2 = M.UNet_Baseline(3, 4, depth=2)
    assert list(m2.state_dict().keys()) == list(ref_sd.keys())
    m2.load_state_dict(ref_sd, strict=True)
    for a in ("in_channels", "start_filts", "depth", "n_classes", "meta_in_channels", "up_mode", "merge_mode",
              "down_convs", "up_convs", "conv_final", "valid", "pad", "fow", "dim", "type", "stride", "increase_fow"):
        assert hasattr(m, a), a
    assert m._state_tensors()[5] is m.down_convs[0].main[1].running_var
    with pytest.raises(ValueError):
        M.UNet(up_mode="nearest")
    with pytest.raises(ValueError):
        M.UNet(merge_mode="mul")
    with pytest.raises(ValueError):
        M.UNet(up_mode="upsample", merge_mode="add")
    # no CPU fallback on the hot path
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 4, 32, 32))
    # the late-metadata variant keeps the reference's 142-entry state_dict and has no CPU fallback either
    lm = M.UNet_LateMetInject(3, 4, 2, depth=2)
    keys = list(lm.state_dict().keys())
    assert keys[-8:] == ["conv_final.weight", "conv_final.bias"] + [
        f"post_processing_weights.main.{i}.{t}" for i in (0, 2, 4) for t in ("weight", "bias")]
    assert lm.state_dict()["conv_final.weight"].shape == (3, 65, 1, 1)
    assert len(M.UNet_LateMetInject(3, 4, 2).state_dict()) == 142
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        lm(torch.zeros(1, 4, 16, 16), torch.zeros(1, 2, 16, 16))
    with pytest.raises(RuntimeError, match="UNet_LateMetInject"):      # a wide head on the baseline class is refused
        M.UNet_Baseline(3, 4, 2, late_meta_inject=True, depth=2)._check_supported(torch.zeros(1, 4, 16, 16))


def test_late_meta_inject_composition_against_reference_golden(pkg, golden_dir):
    """UNet_LateMetInject = native 64-channel network + metadata term (models/unet.py docstring).  Here the native calls
    are replaced by the fp32 oracle (as the CHECKER: this test pins the module's host logic - the head split, the
    autograd routing of W[:, :64] / W[:, 64:], the metadata MLP - against the reference-made fixture on the CPU; the
    GPU test runs the same fixture through the real library)."""
    M = importlib.import_module("crimac_unet_b200.models.unet")
    from oracle import unet_oracle as O
    g2 = np.load(os.path.join(golden_dir, "unet_d2.npz"))
    gl = np.load(os.path.join(golden_dir, "unet_late_d2.npz"))
    sd = {k[6:]: torch.from_numpy(g2[k]) for k in g2.files if k.startswith("state/") and "conv_final" not in k}
    sd.update({k[6:]: torch.from_numpy(gl[k]) for k in gl.files if k.startswith("state/")})
    m = M.UNet_LateMetInject(3, 4, 2, depth=2)
    m.load_state_dict(sd, strict=True)
    x, y, meta = torch.from_numpy(g2["x"]), torch.from_numpy(g2["y"]), torch.from_numpy(gl["meta"])
    body_names = [n for n, _ in m.named_parameters() if n.startswith(("down_convs", "up_convs"))]

    def oracle_state(head_w, params=None):
        st = {k: v for k, v in m.state_dict(keep_vars=True).items() if not k.startswith(("conv_final", "post_"))}
        if params is not None:
            st.update(dict(zip(body_names, params[:-2])))
        st["conv_final.weight"], st["conv_final.bias"] = head_w, m.conv_final.bias
        return st

    m._infer = lambda xx, softmax: O.unet_forward(oracle_state(m._head_weight()), xx, train=False)
    m._train_forward = lambda xx, params: O.unet_forward(oracle_state(params[-2], params), xx, train=True)
    m.eval()
    with torch.no_grad():
        assert np.allclose(m(x, meta).numpy(), gl["eval_logits"], rtol=0, atol=1e-4)
    assert m._head_weight() is m._head_weight()                 # cached while the parameter is unchanged
    m.train()
    logits = m(x, meta)
    loss = torch.nn.functional.cross_entropy(logits, y, weight=torch.tensor(O.CLASS_WEIGHTS))
    loss.backward()
    assert np.allclose(logits.detach().numpy(), gl["train_logits"], rtol=0, atol=2e-4)
    assert abs(loss.item() - float(gl["loss"])) < 1e-5
    named = dict(m.named_parameters())
    n = 0
    for k in gl.files:
        if k.startswith("grad/"):
            ref = gl[k]
            assert np.linalg.norm(named[k[5:]].grad.numpy() - ref) <= 1e-3 * np.linalg.norm(ref) + 1e-9, k
            n += 1
    assert n >= 12 and named["conv_final.weight"].grad.shape == (3, 65, 1, 1)


def test_patch_grid_and_sharding_host_logic(pkg):
    Pr = importlib.import_module("crimac_unet_b200.predict")
    assert Pr.split_pings(0, 1_000_000, 20000) == [tuple(r) for r in P.get_data_split([[0, 1_000_000]], 20000)]
    assert Pr.split_pings(7, 1003, 300) == [tuple(int(v) for v in r) for r in P.get_data_split([[7, 1003]], 300)]
    for (s, e, er, patch, ov) in ((0, 20000, 256, (256, 256), 20), (233, 466, 96, (64, 64), 8), (40, 50, 30, (64, 64), 0)):
        assert np.array_equal(Pr.patch_grid(s, e, er, patch, ov), P.get_data_grid(s, e, 0, er, patch, ov))
        g = Pr.patch_grid(s, e, er, patch, ov)
        assert Pr.preload_window(g, 100000, patch[1]) == P.preload_extents(g, 100000, patch[1])
    chunks = Pr.split_pings(0, 1_000_000, 20000)
    sizes = [len(Pr.shard_chunks(chunks, 8, r)) for r in range(8)]
    assert sizes == [7, 7, 6, 6, 6, 6, 6, 6]
    joined = sum((Pr.shard_chunks(chunks, 8, r) for r in range(8)), [])
    assert joined == chunks
    assert Pr.shard_chunks(chunks[:3], 8, 5) == []   # ragged: more ranks than chunks


def test_gradient_buckets_tile_the_arena(pkg):
    """Host mirror of the native backward's bucket ranges (trainer.gradient_buckets / csrc/net_api.cu close_bucket): the
    buckets tile the padded arena exactly, start on 16-byte boundaries (the exchange kernel moves float4) and the big
    tensors are in the buckets that close first."""
    T = importlib.import_module("crimac_unet_b200.trainer")
    M = importlib.import_module("crimac_unet_b200.models.unet")
    for depth, kw in ((5, {}), (3, dict(merge_mode="add")), (2, {}), (4, dict(up_mode="upsample"))):
        m = M.UNet_Baseline(3, 4, depth=depth, **kw)
        total = sum(p.numel() for p in m.parameters())
        b = T.gradient_buckets(m)
        assert len(b) == (3 if depth >= 3 else 2)
        spans = sorted((o, o + n) for _, o, n in b)
        assert spans[0][0] == 0 and spans[-1][1] == (total + 1023) // 1024 * 1024
        assert all(spans[i][1] == spans[i + 1][0] for i in range(len(spans) - 1))
        assert all(o % 4 == 0 and n % 4 == 0 for _, o, n in b)
    m = M.UNet_Baseline(3, 4)
    b = dict((name, n) for name, _, n in T.gradient_buckets(m))
    assert b["decoder+head"] > 12_000_000 and b["deep encoder"] > 17_000_000 and b["shallow encoder"] < 1_300_000


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[3])
rank, world = int(sys.argv[1]), int(sys.argv[2])
os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=sys.argv[4], RANK=str(rank), WORLD_SIZE=str(world))
dist.init_process_group("gloo", rank=rank, world_size=world)
import __graft_entry__ as ge
ge.load_package()
from crimac_unet_b200.trainer import reduce_gradients
from crimac_unet_b200.predict import shard_chunks, split_pings
g = torch.arange(10, dtype=torch.float32) * (rank + 1)
scale = reduce_gradients(g, world)
assert torch.allclose(g * scale, torch.arange(10, dtype=torch.float32) * 1.5), g      # mean over the 2 replicas
mine = shard_chunks(split_pings(0, 1000, 100), world, rank)
lens = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
dist.all_gather(lens, torch.tensor([len(mine)]))
assert sum(int(t) for t in lens) == 10
dist.destroy_process_group()
print("ok", rank)
"""


def test_data_parallel_exchange_gloo_world2(pkg, tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), str(r), "2", ROOT, port], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) prints ONE JSON line with the keys of the
    bench contract; it needs no GPU and no /root/reference."""
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--batch", "2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "patches/s" and d["higher_is_better"] is True
    for k in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    # "reference" = the unmodified reference module from baseline/_ref (present wherever build() ran next to
    # /root/reference), "port" = the oracle restatement when that copy is absent
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["value"] > 0
    assert d["config"]["same_batch_as_b200_arm"] is True
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_bench_reference_arm_inference_mode():
    """BASELINE.json configs[0] (the reference's CPU-runnable inference case) through the same arm."""
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--mode", "infer", "--steps", "1",
                        "--warmup", "1", "--batch", "2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][0])
    assert d["impl"] == "reference" and "inference" in d["metric"] and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["e2e"]["value"] == d["value"]


def test_bench_flop_accounting_matches_the_survey_numbers():
    """bench.py's generic FLOP counter reproduces SURVEY.md section 8(d) / BASELINE.md exactly (4x256x256 and 6x512x512)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("crimac_bench", os.path.join(root, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    assert abs(b.unet_gflop(4, 256, False) - 96.42704896) < 1e-6 and abs(b.GFLOP_INFER - 96.42704896) < 1e-9
    assert abs(b.unet_gflop(4, 256, True) - 288.979156992) < 1e-6 and abs(b.GFLOP_TRAIN - 288.979156992) < 1e-9
    assert abs(b.unet_gflop(6, 512, False) - 386.312175616) < 1e-6
    assert abs(b.unet_gflop(6, 512, True) - 1157.12458752) < 1e-6

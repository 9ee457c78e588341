"""-m gpu: the U-Net hot path through the nn.Module surface / C-ABI against the oracle.

Stated tolerances (bf16 storage + bf16 tensor-core operands, fp32 accumulation; the reference is fp32):
  * class probabilities: max |dp| <= 2e-2, argmax agreement >= 99.9 % on pixels whose oracle top-2 probability gap
    exceeds 2*tol (all-pixel agreement is reported and bounded at 99 %: a random-init net has ~2 % near-ties);
  * train-mode logits <= 0.15 absolute vs fp32 (<= 0.05 vs the bf16-emulated oracle), loss <= 2e-3 relative;
  * gradients: (a) the backward pass at the native forward state (teacher-forced fp32 autograd): cosine >= 0.999 and
    relative L2 <= 2e-2 for every tensor with a non-trivial gradient, random-init and trained net
    (test_backward_at_the_native_forward_state); (b) against fp32 autograd of the fp32 forward the distance is set by
    the gradient's sensitivity to bf16 FORWARD rounding (55 % relative L2 at the bottleneck with bf16 storage emulated on
    the CPU, <= 1.1 % from rounding the gradient tensors): head / last block <= 2e-2 resp. 5e-3, every tensor cosine >=
    0.8 and inside the emulated-storage noise ball (test_train_step_vs_oracle, ..._inside_the_bf16_sensitivity);
    (c) 24 optimisation steps track the reference's own loss curve within 2 % (measured 0.1 %; golden train_curve.npz).
    Every backward kernel is separately pinned at op level (tests/test_gpu_ops.py, tests/test_gpu_elementwise.py);
  * conv biases that precede a BatchNorm have a mathematically zero gradient: |g| <= 1e-4 * max|dW| of the layer.
"""
import importlib
import os

import numpy as np
import pytest
import torch

from oracle import unet_oracle as O

pytestmark = pytest.mark.gpu
PROB_TOL = 2e-2


@pytest.fixture(scope="module")
def M(pkg):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return importlib.import_module("crimac_unet_b200.models.unet")


dev = torch.device("cuda:0")


def _state(m):
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def _rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def _cos(a, b):
    return (a.double().flatten() @ b.double().flatten() / (a.double().norm() * b.double().norm() + 1e-30)).item()


def _pre_bn_bias(name):
    return name.endswith(".bias") and any(s in name for s in ("main.0", "main.3", "conv1", "conv2"))


def _populate_bn(m, x):
    """default init, BN running stats populated by one train-mode pass with momentum 1."""
    sd = _state(m)
    stats = {}
    with torch.no_grad():
        O.unet_forward(sd, x, train=True, new_stats=stats)
    for k, v in stats.items():
        if "num_batches" not in k:
            # momentum 1.0 <=> running = batch statistic: undo the 0.9/0.1 blend the oracle recorded
            sd[k] = (v - 0.9 * sd[k]) / 0.1
    m.load_state_dict(sd)


def test_golden_depth2_from_reference(M, golden_dir):
    g = np.load(os.path.join(golden_dir, "unet_d2.npz"))
    sd = {k[6:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("state/")}
    m = M.UNet_Baseline(3, 4, depth=2)
    m.load_state_dict(sd)
    m = m.to(dev)
    x, y = torch.from_numpy(g["x"]).to(dev), torch.from_numpy(g["y"]).to(dev)
    m.eval()
    with torch.no_grad():
        lg = m(x)
    ref_lg = torch.from_numpy(g["eval_logits"])                  # this fixture has a x8 head: logits span +-12
    assert (lg.cpu() - ref_lg).abs().max().item() < 2e-2 * ref_lg.abs().max().item()
    m.train()
    loss = m.train_step_fused(x, y, torch.tensor(O.CLASS_WEIGHTS, device=dev))
    assert abs(loss.item() - float(g["loss"])) < 2e-3 * float(g["loss"])
    for k in ("conv_final.weight", "conv_final.bias", "up_convs.0.bn2.weight", "up_convs.0.bn2.bias"):
        assert _rel(dict(m.named_parameters())[k].grad.cpu(), torch.from_numpy(g["grad/" + k])) < 2e-2, k
    for k in g.files:
        if k.startswith("stat/") and "num_batches" not in k:
            assert _rel(m.state_dict()[k[5:]].cpu(), torch.from_numpy(g[k])) < 1e-2, k
        if k.startswith("stat/") and "num_batches" in k:
            assert int(m.state_dict()[k[5:]]) == int(g[k])


def test_late_meta_inject_variant_against_reference_golden(M, golden_dir):
    """UNet_LateMetInject (reference unet.py:346-391) with its 64-channel part on the native
    path, against outputs of the reference class itself (oracle/make_golden.py:golden_late_meta_inject)."""
    g2 = np.load(os.path.join(golden_dir, "unet_d2.npz"))
    gl = np.load(os.path.join(golden_dir, "unet_late_d2.npz"))
    sd = {k[6:]: torch.from_numpy(g2[k]) for k in g2.files if k.startswith("state/") and "conv_final" not in k}
    sd.update({k[6:]: torch.from_numpy(gl[k]) for k in gl.files if k.startswith("state/")})
    m = M.UNet_LateMetInject(3, 4, 2, depth=2)
    m.load_state_dict(sd, strict=True)
    m = m.to(dev)
    x, y, meta = (torch.from_numpy(a).to(dev) for a in (g2["x"], g2["y"], gl["meta"]))
    m.eval()
    with torch.no_grad():
        lg = m(x, meta)
        lg2 = m(x, meta)
    ref = torch.from_numpy(gl["eval_logits"])
    assert (lg.cpu() - ref).abs().max().item() < 2e-2 * ref.abs().max().item()
    assert torch.equal(lg, lg2)
    m.train()
    logits = m(x, meta)
    loss = torch.nn.functional.cross_entropy(logits, y, weight=torch.tensor(O.CLASS_WEIGHTS, device=dev))
    loss.backward()
    ref_t = torch.from_numpy(gl["train_logits"])
    assert (logits.detach().cpu() - ref_t).abs().max().item() < 2e-2 * ref_t.abs().max().item()
    assert abs(loss.item() - float(gl["loss"])) < 2e-3 * float(gl["loss"])
    named = dict(m.named_parameters())
    assert named["conv_final.weight"].grad.shape == (3, 65, 1, 1)
    # gradients upstream of any bf16 gradient tensor: head (both halves), metadata MLP, last BatchNorm
    for k in gl.files:
        if k.startswith("grad/") and k[5:].startswith(("conv_final", "post_processing_weights", "up_convs.0.bn2")):
            assert _rel(named[k[5:]].grad.cpu(), torch.from_numpy(gl[k])) < 2e-2, k
    for k in ("up_convs.0.conv2.weight", "down_convs.1.main.3.weight", "up_convs.0.upconv.weight"):
        gref = torch.from_numpy(gl["grad/" + k])
        cos = torch.nn.functional.cosine_similarity(named[k].grad.cpu().flatten(), gref.flatten(), dim=0).item()
        assert cos > 0.9, (k, cos)
    for n_, p_ in m.named_parameters():
        assert p_.grad is not None and torch.isfinite(p_.grad).all(), n_


@pytest.mark.parametrize("B,H,W,in_ch", [(4, 256, 256, 4), (2, 64, 96, 6)])
def test_inference_probabilities_vs_fp32_oracle(M, B, H, W, in_ch):
    torch.manual_seed(0)
    m = M.UNet_Baseline(3, in_ch)
    x = O.synthetic_echogram(B, in_ch, H, W, seed=0)
    _populate_bn(m, x)
    m = m.to(dev).eval()
    x = x.to(dev)
    with torch.no_grad():
        ref = O.softmax_probs(O.unet_forward(_state(m), x))
        got = m.predict_proba(x)
        got_logits = m(x)
    assert (got.sum(1) - 1).abs().max().item() < 1e-5
    assert torch.allclose(torch.softmax(got_logits, 1), got, atol=1e-5)      # fused softmax == softmax of the logits
    dp = (got - ref).abs().max().item()
    agree = (got.argmax(1) == ref.argmax(1)).float().mean().item()
    top2 = ref.topk(2, 1).values
    conf = (top2[:, 0] - top2[:, 1]) > 2 * PROB_TOL
    agree_conf = (got.argmax(1) == ref.argmax(1))[conf].float().mean().item()
    print(f"max|dp|={dp:.4f} argmax agreement all={agree:.5f} confident={agree_conf:.5f} ({conf.float().mean().item():.3f} of pixels)")
    assert dp <= PROB_TOL
    assert agree_conf >= 0.999
    assert agree >= 0.99


@pytest.mark.parametrize("depth,B,H,W,in_ch", [(5, 2, 128, 128, 4), (3, 2, 40, 72, 6)])
def test_fp32_validation_mode_matches_oracle_to_1e4(M, depth, B, H, W, in_ch):
    """north_star: per-pixel class probabilities within 1e-4 in an fp32 validation mode.  forward_fp32 is an independent
    plain-fp32 CUDA implementation (no bf16, no tensor cores, unfolded BatchNorm) of the same eval forward."""
    torch.manual_seed(0)
    m = M.UNet_Baseline(3, in_ch, depth=depth)
    x = O.synthetic_echogram(B, in_ch, H, W, seed=2)
    _populate_bn(m, x)
    m = m.to(dev).eval()
    x = x.to(dev)
    with torch.no_grad():
        ref_logits = O.unet_forward(_state(m), x)
        ref = O.softmax_probs(ref_logits)
        got_logits = m.forward_fp32(x)
        got = m.forward_fp32(x, softmax=True)
        fast = m.predict_proba(x)
    dp = (got - ref).abs().max().item()
    dl = (got_logits - ref_logits).abs().max().item()
    agree = (got.argmax(1) == ref.argmax(1)).float().mean().item()
    print(f"fp32 validation mode: max|dp|={dp:.2e} max|dlogit|={dl:.2e} argmax agreement {agree:.6f}; bf16 path vs fp32 mode max|dp|={(fast - got).abs().max().item():.4f}")
    assert dp <= 1e-4
    assert dl <= 1e-3
    assert agree >= 0.999
    assert (fast - got).abs().max().item() <= PROB_TOL      # the production path against the validation mode


def test_oracle_noise_floor_cpu_against_cuda_fp32(M):
    """no reference test pins results at this boundary, so the oracle's own noise is bounded instead - the
    same fp32 restatement evaluated by torch on the host cores (oneDNN) and on the device (cuDNN, TF32 off) on the same
    seeded input.  The 1e-4 budget of the fp32 validation mode must sit well above this floor."""
    torch.manual_seed(0)
    m = M.UNet_Baseline(3, 4)
    x = O.synthetic_echogram(2, 4, 128, 128, seed=2)
    _populate_bn(m, x)
    m.eval()
    with torch.no_grad():
        cpu_logits = O.unet_forward(_state(m), x)
        m = m.to(dev)
        gpu_logits = O.unet_forward(_state(m), x.to(dev)).cpu()
    dp = (O.softmax_probs(cpu_logits) - O.softmax_probs(gpu_logits)).abs().max().item()
    dl = (cpu_logits - gpu_logits).abs().max().item()
    print(f"oracle noise floor, torch CPU fp32 vs torch CUDA fp32: max|dp|={dp:.2e} max|dlogit|={dl:.2e}")
    assert dp <= 2e-5 and dl <= 2e-4


def test_inference_is_per_patch_and_deterministic(M):
    torch.manual_seed(1)
    m = M.UNet_Baseline(3, 4).to(dev).eval()
    x = O.synthetic_echogram(3, 4, 64, 64, seed=5, device=dev)
    with torch.no_grad():
        a = m.predict_proba(x)
        b = m.predict_proba(x)
        c = m.predict_proba(x[1:2])
    assert torch.equal(a, b)
    assert torch.equal(a[1:2], c)          # eval mode: patches are independent, bit for bit


@pytest.mark.parametrize("depth,B,H,W", [(2, 4, 64, 64), (5, 4, 128, 128)])
def test_train_step_vs_oracle(M, depth, B, H, W):
    torch.manual_seed(0)
    m = M.UNet_Baseline(3, 4, depth=depth).to(dev).train()
    st0 = _state(m)
    x = O.synthetic_echogram(B, 4, H, W, seed=3, device=dev)
    y = O.synthetic_labels(B, H, W, seed=4, device=dev)
    ref_logits, ref_loss, ref_g, ref_stats = O.train_step(st0, x, y)
    emu_logits, emu_loss, emu_g, _ = O.train_step(st0, x, y, quant=True)
    # (1) drop-in autograd path: forward -> nn.CrossEntropyLoss -> backward, as pipeline.py:171-177
    out = m(x)
    loss = torch.nn.CrossEntropyLoss(weight=torch.tensor(O.CLASS_WEIGHTS, device=dev))(out, y)
    loss.backward()
    d_ref, d_emu = (out - ref_logits).abs().max().item(), (out - emu_logits).abs().max().item()
    print(f"depth {depth}: train logits max|d| vs fp32 oracle {d_ref:.4f}, vs bf16-emulated oracle {d_emu:.4f}")
    assert d_ref < 0.15 and d_emu < 0.05
    assert abs(loss.item() - ref_loss.item()) < 2e-3 * abs(ref_loss.item())
    worst_emu, worst_cos = 0.0, 1.0
    for name, p in m.named_parameters():
        if _pre_bn_bias(name):
            wname = name[:-4] + "weight"
            assert p.grad.abs().max().item() <= 1e-4 * ref_g[wname].abs().max().item() + 1e-7, name
            continue
        r_emu, floor = _rel(p.grad, emu_g[name]), _rel(emu_g[name], ref_g[name])
        worst_emu = max(worst_emu, r_emu)
        worst_cos = min(worst_cos, _cos(p.grad, ref_g[name]))
        # closer to the storage-emulating oracle than that oracle is to fp32 (the problem amplifies ANY perturbation,
        # incl. fp32 summation order, by the same factor: see DESIGN.md section 4)
        # "floor" is how far bf16 STORAGE alone moves this gradient (fp32 arithmetic, identical rounding points); the
        # kernels must stay inside that noise ball, both around the emulated and around the fp32 oracle
        assert r_emu <= max(2e-2, 1.0 * floor), (name, r_emu, floor)
        assert _rel(p.grad, ref_g[name]) <= max(2e-2, 1.5 * floor), (name, _rel(p.grad, ref_g[name]), floor)
        assert _cos(p.grad, ref_g[name]) >= 0.8, (name, _cos(p.grad, ref_g[name]))
        assert 0.8 <= (p.grad.norm() / ref_g[name].norm()).item() <= 1.25, name
    print(f"depth {depth}: worst rel-L2 vs bf16-emulated oracle {worst_emu:.4g}; worst cosine vs fp32 oracle {worst_cos:.4f}")
    for k in ("conv_final.weight", "conv_final.bias", f"up_convs.{depth - 2}.bn2.weight"):
        assert _rel(dict(m.named_parameters())[k].grad, ref_g[k]) < 5e-3, k      # before any bf16 activation gradient
    sd = m.state_dict()
    for k, v in ref_stats.items():
        if "num_batches" in k:
            assert int(sd[k]) == int(v)
        else:
            assert _rel(sd[k], v) < 1e-2, k
    # (2) fused path (forward + CE + backward in one C call) gives the same numbers as the autograd path
    torch.manual_seed(0)
    m2 = M.UNet_Baseline(3, 4, depth=depth).to(dev).train()
    l2 = m2.train_step_fused(x, y, torch.tensor(O.CLASS_WEIGHTS, device=dev))
    assert abs(l2.item() - loss.item()) < 1e-5 * abs(loss.item()) + 1e-6
    for (n1, p1), (n2, p2) in zip(m.named_parameters(), m2.named_parameters()):
        if not _pre_bn_bias(n1):
            # the two paths differ only in the last bit of dlogits (torch's CE backward vs ours); every bf16 gradient
            # tensor re-rounds that difference, which reaches ~5e-3 at the first layer of a random-init net
            assert _rel(p2.grad, p1.grad) < 2e-2, n1


def test_loss_edge_cases(M):
    torch.manual_seed(0)
    m = M.UNet_Baseline(3, 4, depth=2).to(dev).train()
    x = O.synthetic_echogram(2, 4, 32, 32, seed=1, device=dev)
    cw = torch.tensor(O.CLASS_WEIGHTS, device=dev)
    y = torch.full((2, 32, 32), -100, device=dev)
    assert torch.isnan(m.train_step_fused(x, y, cw))               # every pixel ignored -> NaN, as the reference loss
    y[:, :16] = 1
    st0 = _state(m)
    got = m.train_step_fused(x, y, cw)
    ref = O.train_step(st0, x, y)[1]
    assert abs(got.item() - ref.item()) < 2e-3 * abs(ref.item())


def test_backward_is_linear_in_the_logit_gradient_full_size(M):
    """Size-independent property at BASELINE.json's full training size (batch 32 of 4x256x256): the backward pass is
    linear in dlogits, and the loss is finite."""
    torch.manual_seed(0)
    m = M.UNet_Baseline(3, 4).to(dev).train()
    x = O.synthetic_echogram(32, 4, 256, 256, seed=7, device=dev)
    out = m(x)
    g1 = torch.randn_like(out) * 1e-6
    (ga,) = torch.autograd.grad(out, [m.up_convs[3].conv1.weight], g1, retain_graph=True)
    (gb,) = torch.autograd.grad(out, [m.up_convs[3].conv1.weight], 4.0 * g1)
    assert torch.isfinite(out).all()
    assert _rel(gb, 4.0 * ga) < 2e-2          # bf16 rounding of the gradient tensors scales exactly with powers of 2
    y = O.synthetic_labels(32, 256, 256, seed=8, device=dev)
    loss = m.train_step_fused(x, y, torch.tensor(O.CLASS_WEIGHTS, device=dev))
    assert torch.isfinite(loss) and 0.5 < loss.item() < 3.0


def test_checkpoint_round_trip_and_weight_updates_are_seen(M, tmp_path):
    torch.manual_seed(0)
    m = M.UNet_Baseline(3, 4, depth=2).to(dev).eval()
    x = O.synthetic_echogram(1, 4, 32, 32, seed=2, device=dev)
    with torch.no_grad():
        a = m(x)
        torch.save(m.state_dict(), tmp_path / "best.pt")
        m.conv_final.bias.add_(1.0)              # in-place update must trigger a re-pack of the eval operands
        b = m(x)
        assert torch.allclose(b, a + 1.0, atol=1e-5)
        m2 = M.UNet_Baseline(3, 4, depth=2).to(dev).eval()
        m2.load_state_dict(torch.load(tmp_path / "best.pt", map_location=dev))
        assert torch.equal(m2(x), a)


def test_trainer_reduces_the_loss(M, pkg):
    T = importlib.import_module("crimac_unet_b200.trainer")
    torch.manual_seed(0)
    m = M.UNet_Baseline(3, 4, depth=3).to(dev).train()
    tr = T.Trainer(m, lr=0.01, momentum=0.9)
    x = O.synthetic_echogram(4, 4, 64, 64, seed=11, device=dev)
    y = (x[:, 0] > -40).long()                  # a learnable target: class = loud pixels at 18 kHz
    losses = [tr.step(x, y).item() for _ in range(30)]
    assert losses[-1] < 0.5 * losses[0], losses


def _structured_batch(B, H, W, seed, dev):
    """Echogram patches whose class blobs are imprinted on the data (so that there is something to learn): class 1
    raises the two low frequencies, class 2 the two high ones; everything stays inside the dB range [-75, 0]."""
    y = O.synthetic_labels(B, H, W, seed=seed)
    g = torch.Generator().manual_seed(seed + 1000)
    x = -60.0 + 6.0 * torch.randn((B, 4, H, W), generator=g)
    x[:, 0:2] += 18.0 * (y == 1).unsqueeze(1)
    x[:, 2:4] += 18.0 * (y == 2).unsqueeze(1)
    return torch.clamp(x, -75.0, 0.0).to(dev), y.to(dev)


@pytest.fixture(scope="module")
def trained(M, pkg):
    """A net trained for 120 steps on structured batches by the data-parallel trainer (fused train step + SGD, fed from
    HOST batches through the double-buffered copy pipeline).  Shared by the argmax-agreement and the gradient tests."""
    import importlib
    T = importlib.import_module("crimac_unet_b200.trainer")
    torch.manual_seed(0)
    m = M.UNet_Baseline(3, 4).to(dev).train()
    tr = T.Trainer(m, lr=0.005, momentum=0.95, lr_step=0)
    host = []
    for i in range(8):
        x, y = _structured_batch(8, 128, 128, seed=10 + i, dev="cpu")
        host.append((x.pin_memory(), y.pin_memory()))
    losses = [l.item() for l in tr.fit_host(host[i % 8] for i in range(120))]
    return m, tr, losses


def test_training_converges_and_trained_net_meets_argmax_tolerance(M, pkg, trained):
    """End to end: the trainer drives the loss down, and on the resulting confident net the bf16 path agrees with the
    fp32 oracle on >= 99.9 % of ALL pixels (north_star), probabilities within 2e-2."""
    import importlib
    P = importlib.import_module("crimac_unet_b200.predict")
    m, _, losses = trained
    first, last = sum(losses[:5]) / 5, sum(losses[-5:]) / 5
    print(f"loss {first:.4f} -> {last:.4f} over {len(losses)} steps")
    assert all(l == l for l in losses)                  # no NaN
    assert last < 0.25 * first
    m.eval()
    x, y = _structured_batch(8, 128, 128, seed=99, dev=dev)
    with torch.no_grad():
        ref = O.softmax_probs(O.unet_forward(_state(m), x))
        got = m.predict_proba(x)
        val = m.forward_fp32(x, softmax=True)
    agree = (got.argmax(1) == ref.argmax(1)).float().mean().item()
    acc = (ref.argmax(1) == y)[y >= 0].float().mean().item()
    dp = (got - ref).abs().max().item()
    print(f"trained net: max|dp|={dp:.4f} argmax agreement (all pixels)={agree:.5f}; oracle accuracy on the labels {acc:.4f}; fp32 mode max|dp|={(val - ref).abs().max().item():.2e}")
    assert acc > 0.9                                    # the net has really learned the blobs
    assert dp <= PROB_TOL
    assert agree >= 0.999
    assert (val - ref).abs().max().item() <= 1e-4
    # host-batch inference pipeline == direct call
    outs = [o.clone() for o in P.predict_host_batches(m, [x.cpu().pin_memory()] * 3)]
    assert len(outs) == 3 and all(torch.equal(o, got[:, 1:3].half().cpu()) for o in outs)
    m.train()


GRAD_COS, GRAD_REL = 0.999, 2e-2


def _check_tf_gradients(got, ref, label):
    """Every gradient tensor against the teacher-forced fp32 gradients: cosine >= 0.999 and relative L2 <= 2e-2.  Two kinds
    of tensors have a (near-)zero true gradient and are compared differently: conv biases in front of a BatchNorm (exactly
    zero: absolute bound), and the up-convolution biases - a constant added to the up-sampled half only shifts the next
    conv's output by a per-channel constant away from the image border, which that conv's BatchNorm removes, so their
    gradient is a border effect of order 1e-6, a near-cancelling sum of bf16-rounded values: cosine >= 0.98, relative
    L2 <= 0.2 (measured worst: cosine 0.9898 / 0.143 at 6x512x512, batch 2)."""
    rows = []
    for name, g in got.items():
        if _pre_bn_bias(name):
            assert g.abs().max().item() <= 1e-4 * ref[name[:-4] + "weight"].abs().max().item() + 1e-7, name
            continue
        c, r = _cos(g, ref[name]), _rel(g, ref[name])
        if ".upconv." in name and name.endswith(".bias"):
            assert c >= 0.98 and r <= 0.2, (name, c, r)
            continue
        rows.append((name, c, r))
    rows.sort(key=lambda t: -t[2])
    print(f"{label}: {len(rows)} gradient tensors at the native forward state: worst cosine {min(r[1] for r in rows):.6f}, "
          f"worst rel-L2 {rows[0][2]:.4f} ({rows[0][0]}); median rel-L2 {rows[len(rows) // 2][2]:.4f}")
    for name, c, r in rows:
        assert c >= GRAD_COS and r <= GRAD_REL, (name, c, r)


def _teacher_forced_gradients(m, eng, E, x, y, nb):
    """fp32 autograd of the reference network (oracle restatement of models/unet.py:327-343 + pipeline.py:135-138,176)
    evaluated AT THE NATIVE FORWARD STATE: every tensor the native forward stored (conv outputs before BatchNorm,
    activations, ConvTranspose outputs) replaces the torch value with a straight-through substitution
    t + (native - t).detach(), so the graph - ReLU masks, BatchNorm statistics, max-pool arg-max, softmax - is the one the
    native backward differentiates, and the gradient arithmetic is torch fp32."""
    import torch.nn.functional as F
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    names = [n for n, _ in m.named_parameters()]
    leaf = {k: (v.requires_grad_(True) if k in names else v) for k, v in sd.items()}
    D = m.depth

    def nchw(t):
        return t.float().permute(0, 3, 1, 2)

    def tf(t, native):
        return t + (nchw(native) - t).detach()

    def block(t, p_w, p_b, p_bn, index, first=False, act_destroyed=False):
        w = leaf[p_w] if first else O._qw(leaf[p_w], True)          # bf16 tensor-core operand, fp32 master gradient
        raw = tf(F.conv2d(t, w, leaf[p_b], padding=1), E.saved_tensor(eng, index, 0, nb))
        yb = F.batch_norm(raw, None, None, leaf[p_bn + ".weight"], leaf[p_bn + ".bias"], True, 0.1, 1e-5)
        a = torch.relu(yb)
        if act_destroyed:   # merge_mode "add": the decoder added into this activation in place; re-derive the stored value
            return a + (a.detach().bfloat16().float() - a).detach()
        return tf(a, E.saved_tensor(eng, index, 1, nb))

    t, skips = x, []
    for i in range(D):
        p = f"down_convs.{i}.main."
        t = block(t, p + "0.weight", p + "0.bias", p + "1", 2 * i, first=(i == 0))
        t = block(t, p + "3.weight", p + "3.bias", p + "4", 2 * i + 1, act_destroyed=(m.merge_mode == "add" and i < D - 1))
        skips.append(t)
        if i < D - 1:
            t = F.max_pool2d(t, 2, 2)
    for j in range(D - 1):
        p = f"up_convs.{j}."
        if m.up_mode == "transpose":
            up = F.conv_transpose2d(t, O._qw(leaf[p + "upconv.weight"], True), leaf[p + "upconv.bias"], stride=2)
        else:   # the native path applies the 1x1 conv BEFORE the bilinear 2x (the two commute): same graph order here
            low = F.conv2d(t, O._qw(leaf[p + "upconv.1.weight"], True), leaf[p + "upconv.1.bias"])
            up = F.interpolate(low, scale_factor=2, mode="bilinear", align_corners=False)
        if m.merge_mode == "concat":
            t = torch.cat((tf(up, E.saved_tensor(eng, j, 3, nb)), skips[-(j + 2)]), 1)
        else:   # "add": the native buffer holds up + skip (rounded once)
            t = tf(up + skips[-(j + 2)], E.saved_tensor(eng, j, 3, nb))
        t = block(t, p + "conv1.weight", p + "conv1.bias", p + "bn1", 2 * D + 2 * j)
        t = block(t, p + "conv2.weight", p + "conv2.bias", p + "bn2", 2 * D + 2 * j + 1)
    logits = F.conv2d(t, leaf["conv_final.weight"], leaf["conv_final.bias"])
    loss = O.weighted_ce(logits, y)
    grads = torch.autograd.grad(loss, [leaf[k] for k in names])
    return loss.detach(), dict(zip(names, grads))


@pytest.mark.parametrize("which", ["random_init", "trained"])
def test_backward_at_the_native_forward_state(M, pkg, trained, which):
    """The backward pass, isolated from forward rounding: all 82 gradient tensors of one native train step against fp32
    autograd evaluated at the SAME forward state (the native stored activations, see _teacher_forced_gradients).
    Stated tolerance: cosine >= 0.999 and relative L2 <= 2e-2 for every tensor with a non-trivial
    gradient; the two families whose true gradient is (near) zero - conv biases in front of a BatchNorm and the
    up-convolution biases - are bounded as _check_tf_gradients explains.

    Why not simply against fp32 autograd of the fp32 forward: measured on the CPU with the oracle's storage emulation
    (fp32 arithmetic, bf16 rounding of the stored tensors only), rounding the FORWARD activations alone moves the
    bottleneck gradients of this network by 55 % relative L2 (cosine 0.84) at every stage of training, while rounding
    the gradient tensors alone moves them by <= 1.1 % - the gradient of a BatchNorm U-Net is that sensitive to its
    forward state.  That forward-state distance (logits within 0.05 of fp32) is bounded by the forward tests; what the
    backward kernels add on top is what this test bounds."""
    E = importlib.import_module("crimac_unet_b200.engine")
    if which == "trained":
        m = trained[0]
    else:
        torch.manual_seed(3)
        m = M.UNet_Baseline(3, 4).to(dev)
    m.train()
    st0 = _state(m)
    x, y = _structured_batch(4, 128, 128, seed=321, dev=dev)
    loss = m.train_step_fused(x, y, torch.tensor(O.CLASS_WEIGHTS, device=dev))
    eng = m._engine_for(x, train=True)
    got = {n: p.grad.detach().clone() for n, p in m.named_parameters()}
    m.load_state_dict(st0)                              # undo the running-statistics update (the fixture is shared)
    ref_loss, ref_g = _teacher_forced_gradients(m, eng, E, x, y, x.shape[0])
    assert abs(loss.item() - ref_loss.item()) < 1e-4 * abs(ref_loss.item())
    _check_tf_gradients(got, ref_g, which)


def test_whole_network_gradient_vs_fp32_reference_is_inside_the_bf16_sensitivity(M, trained):
    """Against fp32 autograd of the fp32 forward (the reference's own numbers) the distance is dominated by the
    sensitivity of the gradient to bf16 forward rounding (see the test above): tensors downstream of at most a few
    bf16 layers agree tightly, the bottleneck ones only in direction.  Bounds (trained net, structured batch):
    head <= 5e-3, last decoder block <= 5e-2 relative; every tensor cosine >= 0.8 and norm ratio in [0.75, 1.33];
    measured worst cosine 0.855 (down_convs.4.main.1.bias) = the value the CPU emulation of bf16 storage alone gives."""
    m, _, _ = trained
    m.train()
    st0 = _state(m)
    x, y = _structured_batch(8, 128, 128, seed=321, dev=dev)
    ref_logits, ref_loss, ref_g, _ = O.train_step(st0, x, y)
    loss = m.train_step_fused(x, y, torch.tensor(O.CLASS_WEIGHTS, device=dev))
    got = {n: p.grad.detach().clone() for n, p in m.named_parameters()}
    m.load_state_dict(st0)
    assert abs(loss.item() - ref_loss.item()) < 2e-3 * abs(ref_loss.item())
    worst = (1.0, None)
    for name, g in got.items():
        if _pre_bn_bias(name):
            continue
        c, r = _cos(g, ref_g[name]), _rel(g, ref_g[name])
        if c < worst[0]:
            worst = (c, name)
        if name.startswith("conv_final"):
            assert r <= 5e-3, (name, r)
        if name.startswith(("up_convs.3.bn2", "up_convs.3.conv2")):
            assert r <= 5e-2, (name, r)
        assert c >= 0.8 and 0.75 <= (g.norm() / ref_g[name].norm()).item() <= 1.33, (name, c)
    print(f"trained net vs fp32 reference gradients: worst cosine {worst[0]:.4f} ({worst[1]})")


def test_training_curve_and_validation_step_match_the_reference_golden(M, pkg, golden_dir):
    """The optimisation loop end to end against the REFERENCE ITSELF: tests/golden/train_curve.npz holds the losses of 24
    steps of the unmodified reference module trained as pipeline.py:144-190 does (SGD 0.005 / momentum 0.95,
    ExponentialLR(0.5) every 8 batches, weighted CE) and its validation step (pipeline.py:249-270) - generated by
    oracle/make_golden_curve.py.  The native Trainer runs the same schedule on the same seeded batches from the same
    initial weights.  Tolerances: first loss 2e-3 relative (same weights, forward parity); every later loss within 2 %,
    mean deviation <= 0.5 % (measured on B200: max 0.1 %, mean 0.02 %); learning rates exact; validation loss within 2 %
    and sandeel probabilities within 5e-3 on average of the reference's (measured 0.04 % and 4e-4)."""
    T = importlib.import_module("crimac_unet_b200.trainer")
    S = importlib.import_module("crimac_unet_b200.synthetic")
    g = np.load(os.path.join(golden_dir, "train_curve.npz"))
    steps, lr_step, B, size = int(g["steps"]), int(g["lr_step"]), int(g["batch"]), int(g["size"])
    torch.manual_seed(0)
    m = M.UNet_Baseline(3, 4)
    chk = np.array([float(v.detach().abs().sum()) for v in m.state_dict().values()])
    assert np.allclose(chk, g["init_checksum"], rtol=1e-6), "initial weights differ from the reference run's"
    m = m.to(dev).train()
    tr = T.Trainer(m, lr=0.005, momentum=0.95, lr_reduction=0.5, lr_step=lr_step)
    batches = [S.structured_batch(B, size, size, seed=40 + i, device=dev) for i in range(4)]
    losses, lrs = [], []
    for i in range(steps):
        lrs.append(tr.lr)
        losses.append(tr.step(*batches[i % 4]).item())
    losses, ref = np.array(losses), g["losses"]
    dev_rel = np.abs(losses - ref) / ref
    print("native losses   ", np.round(losses, 4).tolist())
    print("reference losses", np.round(ref, 4).tolist())
    print(f"relative deviation: first {dev_rel[0]:.2e}, max {dev_rel.max():.3f}, mean {dev_rel.mean():.4f}")
    assert np.allclose(np.array(lrs), g["lrs"], rtol=1e-12)
    assert dev_rel[0] <= 2e-3 and dev_rel.max() <= 0.02 and dev_rel.mean() <= 0.005
    # validation step on the reference's label codes (int16, as the dataset emits them)
    m.eval()
    xv, _ = S.structured_batch(B, size, size, seed=77, device=dev)
    yv = torch.from_numpy(g["val_labels"]).to(dev)
    cw = torch.tensor(O.CLASS_WEIGHTS, device=dev)
    vloss, prob, logits = m.validate_batch(xv, yv, cw)
    # exactness of the fused step on THIS net's logits (label remap, loss, sandeel probability)
    lab = yv.long().clone()
    for v in (-70, -30, -100, -10):
        lab[lab == v] = -100
    lab[lab == -50] = 0
    assert abs(vloss.item() - torch.nn.functional.cross_entropy(logits, lab, weight=cw).item()) < 1e-5 * vloss.item()
    assert (prob - torch.softmax(logits, 1)[:, 1]).abs().max().item() < 1e-6
    # and against the reference's own validation numbers after ITS 24 steps (two slightly different training runs)
    dprob = (prob.cpu() - torch.from_numpy(g["val_sandeel_prob"])).abs()
    print(f"validation: loss {vloss.item():.4f} (reference {float(g['val_loss']):.4f}); sandeel probability mean |d| {dprob.mean().item():.4f}, max {dprob.max().item():.3f}")
    assert abs(vloss.item() - float(g["val_loss"])) <= 0.02 * float(g["val_loss"])
    assert dprob.mean().item() <= 5e-3


def test_eval_after_native_training_step_sees_the_new_weights(M, pkg):
    """The reference's loop validates every log_step batches (pipeline.py:182): eval, train step, eval on ONE model.
    The native SGD kernel and bn_finalize write parameters / running statistics through raw pointers; the eval path
    must re-pack its bf16 weights and re-fold BatchNorm after every such step."""
    import importlib
    T = importlib.import_module("crimac_unet_b200.trainer")
    torch.manual_seed(0)
    m = M.UNet_Baseline(3, 4, depth=3).to(dev)
    tr = T.Trainer(m, lr=0.05, momentum=0.9)
    x, y = _structured_batch(4, 64, 64, seed=5, dev=dev)
    for use_graph_steps in (1, 3):                      # eager first step, then CUDA-graph replays
        m.eval()
        with torch.no_grad():
            before = m.predict_proba(x)
            assert (before - O.softmax_probs(O.unet_forward(_state(m), x))).abs().max().item() <= PROB_TOL
        m.train()
        for _ in range(use_graph_steps):
            tr.step(x, y)
        m.eval()
        with torch.no_grad():
            after = m.predict_proba(x)
            ref = O.softmax_probs(O.unet_forward(_state(m), x))
        assert (after - ref).abs().max().item() <= PROB_TOL
        assert (after - before).abs().max().item() > 10 * PROB_TOL      # the step really changed the predictions


def test_autograd_path_guards_and_grad_reseating(M):
    torch.manual_seed(0)
    m = M.UNet_Baseline(3, 4, depth=2).to(dev).train()
    x = O.synthetic_echogram(2, 4, 32, 32, seed=1, device=dev)
    y = O.synthetic_labels(2, 32, 32, seed=2, device=dev)
    cw = torch.tensor(O.CLASS_WEIGHTS, device=dev)
    # two train-mode forwards before a backward: the first one's saved activations are gone -> loud error, not wrong grads
    out1 = m(x)
    out2 = m(x)
    with pytest.raises(RuntimeError, match="saved activations"):
        out1.sum().backward()
    out2.sum().backward()                                # the latest forward is fine
    # optimizer.zero_grad(set_to_none=True) drops the arena views: train_step_fused re-seats them
    m.train_step_fused(x, y, cw)
    g0 = m.conv_final.weight.grad.clone()
    torch.optim.SGD(m.parameters(), lr=0.1).zero_grad(set_to_none=True)
    assert m.conv_final.weight.grad is None
    m2 = m.train_step_fused(x, y, cw)
    assert m.conv_final.weight.grad is not None and m.conv_final.weight.grad.data_ptr() >= m._grad_arena.data_ptr()
    assert torch.isfinite(m2) and m.conv_final.weight.grad.abs().sum().item() > 0
    assert g0.shape == m.conv_final.weight.grad.shape
    # the parameter-holder blocks have no torch path of their own
    with pytest.raises(RuntimeError, match="parameter holder"):
        m.down_convs[0](x)
    # train-mode BatchNorm over a single value per channel: PyTorch raises, so does the library
    m1 = M.UNet_Baseline(3, 4, depth=5).to(dev).train()
    with pytest.raises(RuntimeError, match="more than 1 value per channel"):
        m1(O.synthetic_echogram(1, 4, 16, 16, seed=3, device=dev))


def test_full_size_config2_batch32_parity_and_invariances(M):
    """BASELINE.json configs[1] at its FULL size (batch 32 of 4x256x256): loss, head gradient, BatchNorm running
    statistics and class probabilities against the fp32 oracle (run with torch on the GPU), plus the size-independent
    properties the path offers: eval results are per-patch (any sub-batch reproduces its slice bit for bit) and the
    train-mode forward is deterministic."""
    torch.manual_seed(0)
    m = M.UNet_Baseline(3, 4).to(dev).train()
    st0 = _state(m)
    x = O.synthetic_echogram(32, 4, 256, 256, seed=0, device=dev)
    y = O.synthetic_labels(32, 256, 256, seed=1, device=dev)
    ref_logits, ref_loss, ref_g, ref_stats = O.train_step(st0, x, y)
    cw = torch.tensor(O.CLASS_WEIGHTS, device=dev)
    loss = m.train_step_fused(x, y, cw)
    assert abs(loss.item() - ref_loss.item()) < 2e-3 * abs(ref_loss.item())
    named = dict(m.named_parameters())
    for k in ("conv_final.weight", "conv_final.bias", "up_convs.3.bn2.weight", "up_convs.3.bn2.bias"):
        assert _rel(named[k].grad, ref_g[k]) < 5e-3, k
    for name, p in m.named_parameters():
        if not _pre_bn_bias(name):
            assert _cos(p.grad, ref_g[name]) >= 0.8, (name, _cos(p.grad, ref_g[name]))
    sd = m.state_dict()
    for k, v in ref_stats.items():
        if "num_batches" not in k:
            assert _rel(sd[k], v) < 1e-2, k
    del ref_logits, ref_g
    # train-mode forward is deterministic (no atomics on the forward path)
    torch.manual_seed(0)
    m2 = M.UNet_Baseline(3, 4).to(dev).train()
    with torch.no_grad():
        a = m2(x)
    torch.manual_seed(0)
    m3 = M.UNet_Baseline(3, 4).to(dev).train()
    with torch.no_grad():
        b = m3(x)
    assert torch.equal(a, b)
    # eval: probabilities vs oracle on the full batch; sub-batches reproduce their slice exactly
    m.eval()
    with torch.no_grad():
        ref = O.softmax_probs(O.unet_forward(_state(m), x))
        got = m.predict_proba(x)
        parts = torch.cat([m.predict_proba(x[i:i + 8]) for i in range(0, 32, 8)])
    dp = (got - ref).abs().max().item()
    top2 = ref.topk(2, 1).values
    conf = (top2[:, 0] - top2[:, 1]) > 2 * PROB_TOL
    agree_conf = (got.argmax(1) == ref.argmax(1))[conf].float().mean().item()
    print(f"full size: loss {loss.item():.5f} (oracle {ref_loss.item():.5f}); max|dp|={dp:.4f}; argmax agreement on confident pixels {agree_conf:.5f}")
    assert dp <= PROB_TOL and agree_conf >= 0.999
    assert torch.equal(got, parts)


def test_config4_shapes_six_frequencies_512x512(M):
    """BASELINE configs[4] shapes (6 frequencies, 512x512 patches; batch 2 here, the bench runs 64): tile counts, 64-bit
    indexing and split-K factors differ from the 256x256 case.  Eval probabilities and one train step (loss, head and
    last-block gradients, BatchNorm running statistics) against the fp32 oracle, and the backward pass at the native
    forward state for all 82 tensors."""
    E = importlib.import_module("crimac_unet_b200.engine")
    torch.manual_seed(0)
    m = M.UNet_Baseline(3, 6)
    x = O.synthetic_echogram(2, 6, 512, 512, seed=4)
    _populate_bn(m, x)
    m = m.to(dev).eval()
    x = x.to(dev)
    y = O.synthetic_labels(2, 512, 512, seed=5, device=dev)
    with torch.no_grad():
        ref = O.softmax_probs(O.unet_forward(_state(m), x))
        got = m.predict_proba(x)
    dp = (got - ref).abs().max().item()
    top2 = ref.topk(2, 1).values
    conf = (top2[:, 0] - top2[:, 1]) > 2 * PROB_TOL
    agree_conf = (got.argmax(1) == ref.argmax(1))[conf].float().mean().item()
    assert dp <= PROB_TOL and agree_conf >= 0.999
    m.train()
    st0 = _state(m)
    ref_logits, ref_loss, ref_g, ref_stats = O.train_step(st0, x, y)
    loss = m.train_step_fused(x, y, torch.tensor(O.CLASS_WEIGHTS, device=dev))
    assert abs(loss.item() - ref_loss.item()) < 2e-3 * abs(ref_loss.item())
    named = dict(m.named_parameters())
    for k, tol in (("conv_final.weight", 5e-3), ("conv_final.bias", 5e-3), ("up_convs.3.bn2.weight", 2e-2), ("up_convs.3.bn2.bias", 2e-2)):
        assert _rel(named[k].grad, ref_g[k]) < tol, (k, _rel(named[k].grad, ref_g[k]))
    sd = m.state_dict()
    for k, v in ref_stats.items():
        if "num_batches" not in k:
            assert _rel(sd[k], v) < 1e-2, k
    got_g = {n: p.grad.detach().clone() for n, p in m.named_parameters()}
    m.load_state_dict(st0)
    eng = m._engine_for(x, train=True)
    tf_loss, tf_g = _teacher_forced_gradients(m, eng, E, x, y, 2)
    print(f"6x512x512: max|dp| {dp:.4f}; loss {loss.item():.5f} (oracle {ref_loss.item():.5f})")
    _check_tf_gradients(got_g, tf_g, "6x512x512")


@pytest.mark.parametrize("variant", ["add", "upsample"])
def test_decoder_variants_against_reference_golden(M, golden_dir, variant):
    """the two non-default decoders of UNet.__init__ (reference unet.py:200-251) on the
    native path - merge_mode="add" (ConvTranspose epilogue adds into the skip activation in place) and
    up_mode="upsample" (conv1x1 as a one-tap tensor-core GEMM at the low resolution, then a bilinear 2x kernel) -
    against outputs of the reference classes themselves (oracle/make_golden_variants.py): eval logits, train loss,
    BatchNorm buffers, a subset of gradients; and the backward pass at the native forward state for all tensors."""
    E = importlib.import_module("crimac_unet_b200.engine")
    kw = dict(up_mode="transpose", merge_mode="add") if variant == "add" else dict(up_mode="upsample", merge_mode="concat")
    g = np.load(os.path.join(golden_dir, f"unet_{variant}_d3.npz"))
    torch.manual_seed(7)
    m = M.UNet_Baseline(3, 4, depth=3, **kw)
    assert list(m.state_dict().keys()) == [str(k) for k in g["state_keys"]]            # same state_dict schema as the reference
    sd0 = O.trained_like_state({k: v.detach().clone() for k, v in m.state_dict().items()}, seed=1, head_gain=2.0)
    chk = np.array([float(v.double().abs().sum()) for v in sd0.values()])
    assert np.allclose(chk, g["state_checksum"], rtol=1e-6)
    m.load_state_dict(sd0)
    m = m.to(dev)
    x, y = torch.from_numpy(g["x"]).to(dev), torch.from_numpy(g["y"]).to(dev)
    m.eval()
    with torch.no_grad():
        ev = m(x)
        pr = m.predict_proba(x)
    ref_ev = torch.from_numpy(g["eval_logits"])
    assert (ev.cpu() - ref_ev).abs().max().item() < 2e-2 * ref_ev.abs().max().item()
    assert (pr.cpu() - torch.softmax(ref_ev, 1)).abs().max().item() <= PROB_TOL
    m.train()
    st0 = _state(m)
    loss = m.train_step_fused(x, y, torch.tensor(O.CLASS_WEIGHTS, device=dev))
    assert abs(loss.item() - float(g["loss"])) < 2e-3 * float(g["loss"])
    named = dict(m.named_parameters())
    for k in g.files:
        if k.startswith("grad/conv_final"):
            assert _rel(named[k[5:]].grad.cpu(), torch.from_numpy(g[k])) < 2e-2, k
        elif k.startswith("grad/") and not _pre_bn_bias(k[5:]):
            assert _cos(named[k[5:]].grad.cpu(), torch.from_numpy(g[k])) > 0.9, (k, _cos(named[k[5:]].grad.cpu(), torch.from_numpy(g[k])))
        if k.startswith("stat/") and "num_batches" not in k:
            assert _rel(m.state_dict()[k[5:]].cpu(), torch.from_numpy(g[k])) < 1e-2, k
    got = {n: p.grad.detach().clone() for n, p in m.named_parameters()}
    eng = m._engine_for(x, train=True)
    m.load_state_dict(st0)
    tf_loss, tf_g = _teacher_forced_gradients(m, eng, E, x, y, x.shape[0])
    print(f"{variant}: eval max|dlogit| {(ev.cpu() - ref_ev).abs().max().item():.4f}")
    _check_tf_gradients(got, tf_g, variant)


@pytest.mark.parametrize("B,H,W", [(1, 32, 96), (3, 80, 48), (4, 16, 32)])
def test_rectangular_ragged_patches_and_batch_one(M, B, H, W):
    """Shapes off the beaten path: a single patch per batch, non-square patches whose sides are not multiples of the
    16 x 8 / 8 x 16 output tiles, the smallest legal patch of depth 5 (one row of two values per channel at the bottom;
    four of them per batch: with ONE, BatchNorm normalises over two values, running_var reaches 0 and the fp32 oracle
    and its own bf16-rounded form already differ by 0.06 in probability - a single such patch is run in eval mode).
    Eval probabilities, train loss, BatchNorm buffers against the fp32 oracle; every gradient tensor against fp32
    autograd at the native forward state."""
    E = importlib.import_module("crimac_unet_b200.engine")
    torch.manual_seed(B * H + W)
    m = M.UNet_Baseline(3, 4)
    x = O.synthetic_echogram(B, 4, H, W, seed=H)
    _populate_bn(m, x)
    m = m.to(dev).eval()
    x = x.to(dev)
    y = O.synthetic_labels(B, H, W, seed=W, device=dev)
    with torch.no_grad():
        ref = O.softmax_probs(O.unet_forward(_state(m), x))
        got = m.predict_proba(x)
        one = m.predict_proba(x[:1].contiguous())
    assert (got - ref).abs().max().item() <= PROB_TOL
    assert (one - ref[:1]).abs().max().item() <= PROB_TOL
    m.train()
    st0 = _state(m)
    ref_logits, ref_loss, ref_g, ref_stats = O.train_step(st0, x, y)
    loss = m.train_step_fused(x, y, torch.tensor(O.CLASS_WEIGHTS, device=dev))
    assert abs(loss.item() - ref_loss.item()) < 3e-3 * abs(ref_loss.item())
    sd = m.state_dict()
    for k, v in ref_stats.items():
        if "num_batches" not in k:
            assert _rel(sd[k], v) < 2e-2, k
    got_g = {n: p.grad.detach().clone() for n, p in m.named_parameters()}
    m.load_state_dict(st0)
    tf_loss, tf_g = _teacher_forced_gradients(m, m._engine_for(x, train=True), E, x, y, B)
    _check_tf_gradients(got_g, tf_g, f"{B}x4x{H}x{W}")


def test_deterministic_mode_is_bit_reproducible(M):
    """The reference's fix_seeds (utils/general.py:120-128) sets torch.backends.cudnn.deterministic = True.  The native
    path honours the same switch: split-K weight-gradient partial tiles are stored per split and summed in a fixed order
    instead of red.global.add, so two runs of the same steps are bit-identical (every other reduction of the step - BN
    statistics, BN / head / bias gradient partials - is per-CTA partial rows summed in a fixed order already)."""
    cw = torch.tensor(O.CLASS_WEIGHTS, device=dev)
    x, y = _structured_batch(4, 128, 128, seed=31, dev=dev)
    old = torch.backends.cudnn.deterministic
    try:
        torch.backends.cudnn.deterministic = True
        runs = []
        for _ in range(2):
            torch.manual_seed(5)
            m = M.UNet_Baseline(3, 4).to(dev).train()
            losses = [m.train_step_fused(x, y, cw).item() for _ in range(3)]
            runs.append((losses, m._grad_arena.clone(), {k: v.clone() for k, v in m.state_dict().items()}))
        assert runs[0][0] == runs[1][0]
        assert torch.equal(runs[0][1], runs[1][1])
        assert all(torch.equal(runs[0][2][k], runs[1][2][k]) for k in runs[0][2])
    finally:
        torch.backends.cudnn.deterministic = old
    # same numbers as the default (atomic) mode up to fp32 summation order
    torch.manual_seed(5)
    m = M.UNet_Baseline(3, 4).to(dev).train()
    for _ in range(3):
        m.train_step_fused(x, y, cw)
    assert _rel(m._grad_arena, runs[0][1]) < 1e-4


@pytest.mark.parametrize("in_ch", [1, 3, 8, 11])
def test_first_conv_channel_counts(M, in_ch):
    """Edge input-channel counts of the first layer (tensor-core path: one 16-byte chunk per tap up to 4 channels, two
    chunks - and two weight-gradient accumulator groups - up to 8; fp32 CUDA-core kernels for 9..12 = 4 frequencies +
    all 7 metadata channels, pipeline.py:392,413-425): forward, loss and the first conv's weight gradient against the
    fp32 oracle on a shallow net (the gradient passes through one bf16 layer only)."""
    torch.manual_seed(in_ch)
    m = M.UNet_Baseline(3, in_ch, depth=2).to(dev).train()
    st0 = _state(m)
    x = O.synthetic_echogram(2, in_ch, 48, 80, seed=7, device=dev)     # ragged tiles: 48 x 80 is not a multiple of 8 x 16 x 2
    y = O.synthetic_labels(2, 48, 80, seed=8, device=dev)
    ref_logits, ref_loss, ref_g, _ = O.train_step(st0, x, y)
    loss = m.train_step_fused(x, y, torch.tensor(O.CLASS_WEIGHTS, device=dev))
    assert abs(loss.item() - ref_loss.item()) < 2e-3 * abs(ref_loss.item())
    g = dict(m.named_parameters())["down_convs.0.main.0.weight"].grad
    r = _rel(g, ref_g["down_convs.0.main.0.weight"])
    c = _cos(g, ref_g["down_convs.0.main.0.weight"])
    print(f"in_ch {in_ch}: first-conv weight gradient rel-L2 {r:.4f}, cosine {c:.5f}")
    # (a mapping error - tap order, hi/lo rows, channel padding - gives a cosine near 0; the residual is the bf16-storage
    # noise of the layers above, ~0.1 relative on a random-init net, see test_train_step_vs_oracle)
    assert c >= 0.97 and r <= 0.3
    m.eval()
    with torch.no_grad():
        ref = O.softmax_probs(O.unet_forward(_state(m), x))
        got = m.predict_proba(x)
        val = m.forward_fp32(x, softmax=True)
    assert (got - ref).abs().max().item() <= PROB_TOL
    assert (val - ref).abs().max().item() <= 1e-4


@pytest.mark.parametrize("switches", ["CRIMAC_FC_CUDACORE"])
def test_non_default_kernel_variants_pass_the_same_parity_tests(switches):
    """The environment switches of INTEGRATION.md section 4 select alternative kernels.  They are read when the library
    creates a context, so the parity tests are re-run in a child process with the switch set.  Only the CUDA-core first
    conv is kept under test; the other switches are measurement aids whose status is stated in INTEGRATION.md."""
    import subprocess
    import sys
    env = dict(os.environ)
    for s_ in switches.split():
        env[s_] = "1"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_unet.py"), "-x", "-q", "-m", "gpu",
                        "-k", "golden_depth2 or train_step_vs_oracle or trainer_reduces_the_loss or first_conv_channel_counts"],
                       env=env, cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]

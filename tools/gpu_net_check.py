"""Network-level parity + quick timing on the B200 box (development probe; the formal versions live in tests/)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "crimac-classifiers-unet_b200"))
import models.unet as M  # noqa: E402
from oracle import unet_oracle as O  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")


def make_model(in_ch=4, ncls=3, depth=5, seed=0, trained_like=True):
    torch.manual_seed(seed)
    m = M.UNet_Baseline(ncls, in_ch, depth=depth)
    if trained_like:
        m.load_state_dict(O.trained_like_state(m.state_dict(), seed))
    return m.to(dev)


def state_of(m):
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-30)).item()


def sec_infer(B=4, H=256, W=256):
    m = make_model().eval()
    x = O.synthetic_echogram(B, 4, H, W, seed=0, device=dev)
    with torch.no_grad():
        ref_logits = O.unet_forward(state_of(m), x, train=False)
        ref_p = O.softmax_probs(ref_logits)
        got_logits = m(x)
        got_p = m.predict_proba(x)
    torch.cuda.synchronize()
    print(f"[infer] logits max|d|={(got_logits - ref_logits).abs().max().item():.4g} (ref range "
          f"{ref_logits.min().item():.3g}..{ref_logits.max().item():.3g})")
    dp = (got_p - ref_p).abs()
    agree = (got_p.argmax(1) == ref_p.argmax(1)).float().mean().item()
    top2 = ref_p.topk(2, 1).values
    conf = (top2[:, 0] - top2[:, 1]) > 0.04
    agree_conf = (got_p.argmax(1) == ref_p.argmax(1))[conf].float().mean().item()
    print(f"[infer] probs max|dp|={dp.max().item():.4g} mean={dp.mean().item():.4g} argmax agreement={agree * 100:.3f}% "
          f"(on pixels with top-2 gap>0.04: {agree_conf * 100:.3f}%, {conf.float().mean().item() * 100:.1f}% of pixels)")
    print(f"[infer] softmax sums to 1: {(got_p.sum(1) - 1).abs().max().item():.3g}")


def sec_train(B=4, H=128, W=128, depth=5):
    m = make_model(depth=depth, trained_like=False).train()
    st0 = state_of(m)
    x = O.synthetic_echogram(B, 4, H, W, seed=3, device=dev)
    y = O.synthetic_labels(B, H, W, seed=4, device=dev)
    ref_logits, ref_loss, ref_g, ref_stats = O.train_step(st0, x, y)
    crit = torch.nn.CrossEntropyLoss(weight=torch.tensor(O.CLASS_WEIGHTS, device=dev))
    out = m(x)
    loss = crit(out, y)
    loss.backward()
    torch.cuda.synchronize()
    print(f"[train d={depth}] logits max|d|={(out - ref_logits).abs().max().item():.4g}  loss {loss.item():.6f} vs ref "
          f"{ref_loss.item():.6f}")
    worst = 0.0
    for name, p in m.named_parameters():
        r = rel(p.grad, ref_g[name])
        gmax = ref_g[name].abs().max().item()
        flag = ""
        is_pre_bn_bias = name.endswith(".bias") and ("main.0" in name or "main.3" in name or "conv1" in name or "conv2" in name)
        if not is_pre_bn_bias:
            worst = max(worst, r)
        if r > 0.05 and not is_pre_bn_bias:
            flag = "  <-- BAD"
        print(f"   grad {name:38s} rel_l2={r:.4g} ref_max={gmax:.3g} got_max={p.grad.abs().max().item():.3g}{flag}")
    print(f"[train d={depth}] worst rel_l2 over non-pre-BN-bias params = {worst:.4g}")
    sd = m.state_dict()
    e = max(rel(sd[k], v) for k, v in ref_stats.items() if "num_batches" not in k)
    nbt = all(int(sd[k]) == int(v) for k, v in ref_stats.items() if "num_batches" in k)
    print(f"[train d={depth}] running stats worst rel={e:.4g}  num_batches_tracked ok={nbt}")
    # fused step on the same inputs
    m2 = make_model(depth=depth, trained_like=False).train()
    l2 = m2.train_step_fused(x, y, torch.tensor(O.CLASS_WEIGHTS, device=dev))
    torch.cuda.synchronize()
    w2 = max(rel(p.grad, ref_g[n]) for n, p in m2.named_parameters()
             if not (n.endswith(".bias") and ("main.0" in n or "main.3" in n or "conv1" in n or "conv2" in n)))
    print(f"[fused d={depth}] loss {l2.item():.6f} vs ref {ref_loss.item():.6f}; worst grad rel_l2={w2:.4g}")


def timeit(fn, warm=3, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def sec_time():
    m = make_model().eval()
    x = O.synthetic_echogram(16, 4, 256, 256, seed=0, device=dev)
    with torch.no_grad():
        ms = timeit(lambda: m.predict_proba(x))
    print(f"[time] infer B=16: {ms:.3f} ms -> {16 / ms * 1e3:.1f} patches/s = {16 / ms * 96.427:.1f} TFLOP/s")
    x = O.synthetic_echogram(32, 4, 256, 256, seed=0, device=dev)
    with torch.no_grad():
        ms = timeit(lambda: m.predict_proba(x))
    print(f"[time] infer B=32: {ms:.3f} ms -> {32 / ms * 1e3:.1f} patches/s = {32 / ms * 96.427:.1f} TFLOP/s")
    m.train()
    y = O.synthetic_labels(32, 256, 256, seed=1, device=dev)
    cw = torch.tensor(O.CLASS_WEIGHTS, device=dev)
    ms = timeit(lambda: m.train_step_fused(x, y, cw))
    print(f"[time] train fused B=32: {ms:.3f} ms -> {32 / ms * 1e3:.1f} patches/s = {32 / ms * 288.979:.1f} TFLOP/s")


if __name__ == "__main__":
    for name in sys.argv[1:]:
        t0 = time.time()
        try:
            if name == "infer":
                sec_infer()
            elif name == "train2":
                sec_train(depth=2, H=64, W=64)
            elif name == "train5":
                sec_train(depth=5)
            elif name == "time":
                sec_time()
        except Exception as e:  # noqa
            import traceback
            traceback.print_exc()
            print(f"[EXC] {name}: {type(e).__name__}: {e}")
        print(f"--- {name} done in {time.time() - t0:.1f}s", flush=True)

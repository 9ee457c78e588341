"""ORACLE (test infrastructure, not product code): fp32 restatement of the reference U-Net hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this file; the
product path (crimac-classifiers-unet_b200/) never does.

The reference is pure PyTorch: its arithmetic lives in the third-party ATen operators (torch==1.7.1 pinned in
crimac_unet/requirements.txt:47; torch 2.11.0 here), so this restatement is written against the same operators in
functional form, driven by a plain ``state_dict`` — it shares no module classes with the product.  Every function
cites the reference lines it follows (paths relative to /root/reference/crimac_unet/).

Pinning: the reference ships no tests, golden vectors or checkpoints, so the oracle is pinned
against OUTPUTS OF THE REFERENCE ITSELF: oracle/make_golden.py imports the unmodified reference module from
/root/reference in the build container, runs it on seeded inputs and commits the results under tests/golden/;
tests/test_oracle_golden.py checks this file against those vectors.
"""
import torch
import torch.nn.functional as F

BN_EPS = 1e-5        # nn.BatchNorm2d default, models/unet.py:78
BN_MOMENTUM = 0.1    # nn.BatchNorm2d default
CLASS_WEIGHTS = (10.0, 300.0, 250.0)  # pipeline_train_predict/pipeline.py:135
IGNORE_INDEX = -100  # nn.CrossEntropyLoss default, pipeline.py:138


def depth_of(state):
    """Number of encoder blocks in a UNet state_dict."""
    d = 0
    while f"down_convs.{d}.main.0.weight" in state:
        d += 1
    return d


def _bn(x, state, prefix, train, new_stats):
    """nn.BatchNorm2d forward (models/unet.py:78,81,121-122). In train mode uses batch statistics and records the
    running-stat update PyTorch performs (momentum 0.1, unbiased variance) in ``new_stats``."""
    w, b = state[prefix + ".weight"], state[prefix + ".bias"]
    rm, rv = state[prefix + ".running_mean"], state[prefix + ".running_var"]
    if not train:
        return F.batch_norm(x, rm, rv, w, b, False, BN_MOMENTUM, BN_EPS)
    n = x.numel() // x.shape[1]
    mean = x.mean((0, 2, 3))
    var = x.var((0, 2, 3), unbiased=False)
    if new_stats is not None:
        new_stats[prefix + ".running_mean"] = ((1 - BN_MOMENTUM) * rm + BN_MOMENTUM * mean).detach()
        new_stats[prefix + ".running_var"] = ((1 - BN_MOMENTUM) * rv + BN_MOMENTUM * var * n / max(n - 1, 1)).detach()
        new_stats[prefix + ".num_batches_tracked"] = state[prefix + ".num_batches_tracked"] + 1
    xh = (x - mean[None, :, None, None]) * torch.rsqrt(var + BN_EPS)[None, :, None, None]
    return xh * w[None, :, None, None] + b[None, :, None, None]


class _RoundBoth(torch.autograd.Function):
    """bf16 storage emulation: rounds the value on the way forward and the gradient on the way back."""

    @staticmethod
    def forward(ctx, t):
        return t.to(torch.bfloat16).to(t.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


def _q(t, quant):
    return _RoundBoth.apply(t) if quant else t


def _qw(w, quant):
    """bf16 operand copy of a weight; the gradient flows to the fp32 master weight unchanged."""
    return w + (w.to(torch.bfloat16).to(w.dtype) - w).detach() if quant else w


def unet_forward(state, x, train=False, new_stats=None, quant=False, up_mode="transpose", merge_mode="concat"):
    """UNet_Baseline.forward (models/unet.py:327-343): encoder blocks DownConv.forward (:88-93), decoder blocks
    UpConv.forward (:124-137, concat order = upsampled first), 1x1 head (:342). Returns raw logits.
    up_mode "upsample" = nn.Upsample(bilinear, x2) then conv1x1 (:50-56, state keys upconv.1.*); merge_mode "add" =
    from_up + from_down (:133-134).

    quant=True is the MIXED-PRECISION EMULATION of the product's storage format (not the reference): every tensor the
    sm_100a path keeps in HBM as bf16 (conv outputs before BN in train mode, activations, ConvTranspose outputs,
    tensor-core weight operands, and the matching gradients) is rounded to bf16 at the same point, with all
    arithmetic in fp32.  The first conv, BN statistics, the head and the loss stay fp32, as in the product."""
    depth = depth_of(state)
    skips = []

    def block(t, wkey, bkey, bnkey, first=False):
        w = state[wkey] if first else _qw(state[wkey], quant)
        raw = F.conv2d(t, w, state[bkey], padding=1)                                  # conv3x3, :35-44
        if train:
            raw = _q(raw, quant)   # train mode stores the pre-BN output; eval folds BN into the fp32 epilogue
        return _q(F.relu(_bn(raw, state, bnkey, train, new_stats)), quant)

    for i in range(depth):
        p = f"down_convs.{i}.main."
        x = block(x, p + "0.weight", p + "0.bias", p + "1", first=(i == 0))
        x = block(x, p + "3.weight", p + "3.bias", p + "4")
        skips.append(x)                                                               # before_pool, :90
        if i < depth - 1:
            x = F.max_pool2d(x, 2, 2)                                                 # :86,92
    for j in range(depth - 1):
        p = f"up_convs.{j}."
        if up_mode == "transpose":
            up = F.conv_transpose2d(x, _qw(state[p + "upconv.weight"], quant), state[p + "upconv.bias"], stride=2)  # :47-49,130
        else:
            up = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)                             # :53
            up = F.conv2d(up, _qw(state[p + "upconv.1.weight"], quant), state[p + "upconv.1.bias"])                 # :54
        if merge_mode == "concat":
            x = torch.cat((_q(up, quant), skips[-(j + 2)]), 1)                        # :132, :336
        else:
            x = _q(up + skips[-(j + 2)], quant)                                       # :134
        x = block(x, p + "conv1.weight", p + "conv1.bias", p + "bn1")
        x = block(x, p + "conv2.weight", p + "conv2.bias", p + "bn2")
    return F.conv2d(x, state["conv_final.weight"], state["conv_final.bias"])          # conv1x1, :59-60,342


def softmax_probs(logits):
    """F.softmax(outputs, dim=1) (pipeline.py:218)."""
    return F.softmax(logits, dim=1)


def weighted_ce(logits, labels, class_weights=CLASS_WEIGHTS, ignore_index=IGNORE_INDEX):
    """nn.CrossEntropyLoss(weight=[10,300,250]) with the default ignore_index and 'mean' reduction
    (pipeline.py:135-138, call :176), written out: sum_i w[y_i]*(lse(z_i) - z_i[y_i]) / sum_i w[y_i]."""
    w = torch.as_tensor(class_weights, dtype=logits.dtype, device=logits.device)
    z = logits.permute(0, 2, 3, 1).reshape(-1, logits.shape[1])
    y = labels.reshape(-1)
    keep = y != ignore_index
    z, y = z[keep], y[keep]
    nll = torch.logsumexp(z, 1) - z.gather(1, y[:, None])[:, 0]
    wy = w[y]
    return (wy * nll).sum() / wy.sum()


def train_step(state, x, labels, class_weights=CLASS_WEIGHTS, quant=False, up_mode="transpose", merge_mode="concat"):
    """pipeline.py:167-177: model.train(); outputs = model(x); loss = criterion(outputs, labels); loss.backward().
    Returns (logits, loss, grads by parameter name, updated BN buffers)."""
    names = [k for k, v in state.items() if v.dtype.is_floating_point and "running_" not in k]
    leaf = {k: (v.detach().clone().requires_grad_(True) if k in names else v) for k, v in state.items()}
    new_stats = {}
    logits = unet_forward(leaf, x, train=True, new_stats=new_stats, quant=quant, up_mode=up_mode, merge_mode=merge_mode)
    loss = weighted_ce(logits, labels, class_weights)
    grads = torch.autograd.grad(loss, [leaf[k] for k in names])
    return logits.detach(), loss.detach(), dict(zip(names, grads)), new_stats


def param_names(state):
    """Names in UNet_Baseline.parameters() order."""
    depth = depth_of(state)
    out = []
    for i in range(depth):
        for m in ("0", "1", "3", "4"):
            out += [f"down_convs.{i}.main.{m}.weight", f"down_convs.{i}.main.{m}.bias"]
    for j in range(depth - 1):
        for m in ("upconv", "conv1", "conv2", "bn1", "bn2"):
            out += [f"up_convs.{j}.{m}.weight", f"up_convs.{j}.{m}.bias"]
    return out + ["conv_final.weight", "conv_final.bias"]


# ------------------------------------------------------------------------------------------ synthetic workloads
def synthetic_echogram(batch, channels, height, width, seed=0, device="cpu"):
    """SURVEY.md §8d config 1/2 input: x = clip(10*log10(sv + 1e-10), -75, 0), sv = 10**U(-9,-2)."""
    g = torch.Generator().manual_seed(seed)
    sv = 10.0 ** (torch.rand((batch, channels, height, width), generator=g) * 7.0 - 9.0)
    return torch.clamp(10.0 * torch.log10(sv + 1e-10), -75.0, 0.0).to(device)


def synthetic_labels(batch, height, width, seed=1, device="cpu"):
    """SURVEY.md §8d config 2 labels: blobs of class 1 / 2 on background 0 (~90/5/5 %), 2 % of pixels = -100."""
    g = torch.Generator().manual_seed(seed)
    coarse = torch.rand((batch, 1, max(height // 16, 1), max(width // 16, 1)), generator=g)
    field = F.interpolate(coarse, size=(height, width), mode="bilinear", align_corners=False)[:, 0]
    lab = torch.zeros((batch, height, width), dtype=torch.long)
    lab[field > 0.80] = 1
    lab[field < 0.20] = 2
    lab[torch.rand((batch, height, width), generator=g) < 0.02] = IGNORE_INDEX
    return lab.to(device)


def trained_like_state(state, seed=0, head_gain=8.0):
    """Makes a randomly initialised state_dict behave like a trained checkpoint for parity runs:
    BN running statistics away from (0,1), non-trivial affine BN parameters and a confident head."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, v in state.items():
        v = v.clone()
        if k.endswith("running_mean"):
            v = 0.1 * torch.randn(v.shape, generator=g)
        elif k.endswith("running_var"):
            v = 0.5 + torch.rand(v.shape, generator=g)
        elif (".main.1." in k or ".main.4." in k or ".bn1." in k or ".bn2." in k) and k.endswith("weight"):
            v = 0.75 + 0.5 * torch.rand(v.shape, generator=g)
        elif (".main.1." in k or ".main.4." in k or ".bn1." in k or ".bn2." in k) and k.endswith("bias"):
            v = 0.1 * torch.randn(v.shape, generator=g)
        elif k == "conv_final.weight":
            v = v * head_gain
        out[k] = v
    return out

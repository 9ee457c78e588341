"""-m gpu: every tensor-core kernel in isolation, through the op-level C-ABI, against the oracle's operators
(torch fp32 on the same bf16-rounded operands, TF32 off).  Tolerances: outputs are stored as bf16, so the bound is
half a bf16 ulp of the value range (2^-8 relative) unless stated."""
import ctypes
import importlib

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env(pkg):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    L = importlib.import_module("crimac_unet_b200.lib")
    return L, L.load(), torch.device("cuda:0")


def _igemm(env, mode, x, w, n_total, scale=None, shift=None, relu=0, out=None, out_pitch=0, convt_cout=0, pool=None,
           stats=None, head=None, block_n=0, H=None, W=None):
    L, lib, dev = env
    NB, Hx, Wx, pitch = x.shape
    hw, hb, hout, ncls, sm = (None, None, None, 0, 0) if head is None else head
    rc = lib.crimac_op_igemm(mode, L.ptr(x), NB, H or Hx, W or Wx, pitch, pitch, L.ptr(w), n_total, L.ptr(scale),
                             L.ptr(shift), relu, L.ptr(out), out_pitch, convt_cout, L.ptr(pool),
                             0 if pool is None else pool.shape[-1], L.ptr(stats), L.ptr(hw), L.ptr(hb), L.ptr(hout),
                             ncls, sm, block_n, L.stream_ptr())
    L.check(rc, "crimac_op_igemm")
    torch.cuda.synchronize()


def _nchw(t):
    return t.float().permute(0, 3, 1, 2)


def _close(got, ref, rel=2 ** -8):
    tol = rel * max(ref.abs().max().item(), 1.0)
    err = (got.float() - ref.float()).abs().max().item()
    assert err <= tol, f"max abs err {err} > {tol}"


@pytest.mark.parametrize("NB,H,W,Cin,Cout,bn", [(2, 32, 32, 64, 64, 0), (1, 16, 16, 128, 256, 0),
                                                (2, 32, 48, 128, 128, 64), (3, 16, 16, 256, 512, 0),
                                                (1, 24, 40, 64, 128, 0)])   # last: ragged tiles (H,W not multiples of the 8x16 tile)
def test_conv3x3_bn_relu_pool_and_train_epilogue(env, NB, H, W, Cin, Cout, bn):
    dev = env[2]
    torch.manual_seed(NB * 1000 + Cin)
    x = torch.randn(NB, H, W, Cin, device=dev).bfloat16()
    w = (torch.randn(Cout, Cin, 3, 3, device=dev) / (3 * Cin ** 0.5)).bfloat16()
    wp = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
    scale, shift = torch.rand(Cout, device=dev) + 0.5, torch.randn(Cout, device=dev)
    out = torch.zeros(NB, H, W, Cout, device=dev, dtype=torch.bfloat16)
    pool = torch.zeros(NB, H // 2, W // 2, Cout, device=dev, dtype=torch.bfloat16)
    _igemm(env, 0, x, wp, Cout, scale, shift, 1, out, Cout, pool=pool, block_n=bn)
    conv = F.conv2d(_nchw(x), w.float(), padding=1)
    ref = torch.relu(conv * scale[None, :, None, None] + shift[None, :, None, None])
    _close(_nchw(out), ref)
    assert torch.equal(_nchw(pool), F.max_pool2d(_nchw(out), 2))       # pool of the stored values: bit exact
    # train-mode epilogue: raw + bias, per-tile channel sums of the stored values
    stats = torch.zeros(256, 2, Cout, device=dev)       # one partial row per CTA (<= number of SMs rows are written)
    raw = torch.zeros_like(out)
    _igemm(env, 0, x, wp, Cout, None, shift, 0, raw, Cout, stats=stats, block_n=bn)
    _close(_nchw(raw), conv + shift[None, :, None, None])
    s = stats.double().sum(0)
    rr = raw.double()
    assert torch.allclose(s[0], rr.sum((0, 1, 2)), rtol=1e-4, atol=1e-2)
    assert torch.allclose(s[1], (rr * rr).sum((0, 1, 2)), rtol=1e-4, atol=1e-2)


def test_conv3x3_fused_head_softmax(env):
    dev = env[2]
    torch.manual_seed(1)
    NB, H, W = 2, 32, 32
    x = torch.randn(NB, H, W, 64, device=dev).bfloat16()
    w = (torch.randn(64, 64, 3, 3, device=dev) / 24).bfloat16()
    wp = w.permute(0, 2, 3, 1).reshape(64, 576).contiguous()
    scale, shift = torch.rand(64, device=dev) + 0.5, torch.randn(64, device=dev) * 0.1
    hw, hb = torch.randn(3, 64, device=dev) * 0.2, torch.randn(3, device=dev)
    for softmax in (1, 0):
        outp = torch.zeros(NB, 3, H, W, device=dev)
        _igemm(env, 0, x, wp, 64, scale, shift, 1, None, 0, head=(hw, hb, outp, 3, softmax))
        act = torch.relu(F.conv2d(_nchw(x), w.float(), padding=1) * scale[None, :, None, None] + shift[None, :, None, None])
        logits = F.conv2d(act, hw[:, :, None, None], hb)
        ref = torch.softmax(logits, 1) if softmax else logits
        assert (outp - ref).abs().max().item() < (1e-3 if softmax else 5e-3)


def test_conv_transpose_forward_scatter_and_backward_data(env):
    dev = env[2]
    torch.manual_seed(2)
    NB, H, W, Cin, Cout = 2, 16, 16, 128, 64
    x = torch.randn(NB, H, W, Cin, device=dev).bfloat16()
    w = (torch.randn(Cin, Cout, 2, 2, device=dev) / Cin ** 0.5).bfloat16()
    b = torch.randn(Cout, device=dev)
    wp = w.permute(2, 3, 1, 0).reshape(4 * Cout, Cin).contiguous()
    cat = torch.full((NB, 2 * H, 2 * W, 2 * Cout), 7.0, device=dev, dtype=torch.bfloat16)
    _igemm(env, 1, x, wp, 4 * Cout, None, b, 0, cat, 2 * Cout, convt_cout=Cout)
    _close(_nchw(cat[..., :Cout]), F.conv_transpose2d(_nchw(x), w.float(), b, stride=2))
    assert torch.all(cat[..., Cout:] == 7.0)            # the skip half of the concat buffer is not touched
    dy = torch.randn(NB, 2 * H, 2 * W, Cout, device=dev).bfloat16()
    wd = w.permute(0, 2, 3, 1).reshape(Cin, 4 * Cout).contiguous()
    dx = torch.zeros(NB, H, W, Cin, device=dev, dtype=torch.bfloat16)
    _igemm(env, 2, dy, wd, Cin, None, None, 0, dx, Cin, H=H, W=W)
    xr = _nchw(x).clone().requires_grad_(True)
    F.conv_transpose2d(xr, w.float(), b, stride=2).backward(_nchw(dy))
    _close(_nchw(dx), xr.grad)
    # the product path: backward-data straight from the FORWARD-packed weights (MN-major B operand), no second packing
    dx2 = torch.zeros_like(dx)
    _igemm(env, 5, dy, wp, Cin, None, None, 0, dx2, Cin, H=H, W=W)
    _close(_nchw(dx2), xr.grad)


@pytest.mark.parametrize("NB,H,W,Cin,Cout", [(2, 32, 32, 128, 64), (1, 16, 16, 256, 512), (2, 24, 40, 64, 128)])
def test_conv3x3_backward_data(env, NB, H, W, Cin, Cout):
    dev = env[2]
    torch.manual_seed(3)
    w = (torch.randn(Cout, Cin, 3, 3, device=dev) / (3 * Cin ** 0.5)).bfloat16()
    dy = torch.randn(NB, H, W, Cout, device=dev).bfloat16()
    wd = w.flip(2, 3).permute(1, 2, 3, 0).reshape(Cin, 9 * Cout).contiguous()
    dx = torch.zeros(NB, H, W, Cin, device=dev, dtype=torch.bfloat16)
    _igemm(env, 0, dy, wd, Cin, None, None, 0, dx, Cin)
    xr = torch.randn(NB, Cin, H, W, device=dev, requires_grad=True)
    F.conv2d(xr, w.float(), padding=1).backward(_nchw(dy))
    _close(_nchw(dx), xr.grad)
    # the product path: the FORWARD-packed weights [Cout][tap][Cin] read as an MN-major operand, taps rotated in-kernel
    wf = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
    dx2 = torch.zeros_like(dx)
    _igemm(env, 4, dy, wf, Cin, None, None, 0, dx2, Cin)
    _close(_nchw(dx2), xr.grad)


@pytest.mark.parametrize("NB,H,W,Cin,Cout,splits", [(2, 32, 32, 64, 64, 1), (2, 32, 32, 128, 128, 4),
                                                    (2, 16, 16, 256, 128, 0), (1, 32, 32, 128, 64, 3),
                                                    (1, 20, 24, 64, 128, 2)])
def test_weight_gradient_conv3x3(env, NB, H, W, Cin, Cout, splits):
    L, lib, dev = env
    torch.manual_seed(4)
    x = torch.randn(NB, H, W, Cin, device=dev).bfloat16()
    dy = torch.randn(NB, H, W, Cout, device=dev).bfloat16()
    scratch = torch.empty(9 * Cout * Cin, device=dev)
    dw = torch.zeros(Cout, Cin, 3, 3, device=dev)
    L.check(lib.crimac_op_wgrad(0, L.ptr(dy), Cout, Cout, L.ptr(x), Cin, Cin, NB, H, W, L.ptr(scratch), L.ptr(dw),
                                splits, 0, L.stream_ptr()), "crimac_op_wgrad")
    torch.cuda.synchronize()
    wr = torch.zeros(Cout, Cin, 3, 3, device=dev, requires_grad=True)
    F.conv2d(_nchw(x), wr, padding=1).backward(_nchw(dy))
    # fp32 accumulation of exact bf16 products: only summation-order noise
    assert (dw - wr.grad).abs().max().item() <= 1e-4 * (NB * H * W) ** 0.5 * 4


@pytest.mark.parametrize("NB,H,W,Cin,Cout,splits", [(2, 32, 32, 64, 64, 1), (2, 32, 32, 128, 64, 4),
                                                    (2, 16, 16, 256, 128, 0), (1, 32, 32, 64, 192, 3),
                                                    (1, 20, 24, 64, 128, 2), (2, 64, 64, 64, 64, 0)])
def test_weight_gradient_conv3x3_all_taps_halo(env, NB, H, W, Cin, Cout, splits):
    """The production 3x3 weight-gradient kernel: nine taps per CTA from one halo tile, taps paired along M."""
    L, lib, dev = env
    torch.manual_seed(6)
    x = torch.randn(NB, H, W, Cin, device=dev).bfloat16()
    dy = torch.randn(NB, H, W, Cout, device=dev).bfloat16()
    scratch = torch.empty(9 * Cout * Cin, device=dev)
    dw = torch.zeros(Cout, Cin, 3, 3, device=dev)
    L.check(lib.crimac_op_wgrad_halo(L.ptr(dy), Cout, Cout, L.ptr(x), Cin, Cin, NB, H, W, L.ptr(scratch), L.ptr(dw),
                                     splits, L.stream_ptr()), "crimac_op_wgrad_halo")
    torch.cuda.synchronize()
    wr = torch.zeros(Cout, Cin, 3, 3, device=dev, requires_grad=True)
    F.conv2d(_nchw(x), wr, padding=1).backward(_nchw(dy))
    assert (dw - wr.grad).abs().max().item() <= 1e-4 * (NB * H * W) ** 0.5 * 4


def test_weight_gradient_conv_transpose(env):
    L, lib, dev = env
    torch.manual_seed(5)
    NB, H, W, Cin, Cout = 2, 16, 16, 128, 64
    x = torch.randn(NB, H, W, Cin, device=dev).bfloat16()
    dy = torch.randn(NB, 2 * H, 2 * W, Cout, device=dev).bfloat16()
    scratch = torch.empty(4 * Cout * Cin, device=dev)
    dw = torch.zeros(Cin, Cout, 2, 2, device=dev)
    L.check(lib.crimac_op_wgrad(1, L.ptr(x), Cin, Cin, L.ptr(dy), Cout, Cout, NB, H, W, L.ptr(scratch), L.ptr(dw), 2, 0,
                                L.stream_ptr()), "crimac_op_wgrad")
    torch.cuda.synchronize()
    wr = torch.zeros(Cin, Cout, 2, 2, device=dev, requires_grad=True)
    F.conv_transpose2d(_nchw(x), wr, stride=2).backward(_nchw(dy))
    assert (dw - wr.grad).abs().max().item() <= 1e-4 * (NB * H * W) ** 0.5 * 4


def test_invalid_arguments_are_reported_not_launched(env):
    L, lib, dev = env
    x = torch.zeros(1, 16, 16, 48, device=dev, dtype=torch.bfloat16)
    rc = lib.crimac_op_igemm(0, L.ptr(x), 1, 16, 16, 48, 48, L.ptr(x), 64, None, None, 0, L.ptr(x), 64, 0, None, 0, None,
                             None, None, None, 0, 0, 0, L.stream_ptr())
    assert rc == 1 and b"multiple of 64" in lib.crimac_last_error()

"""-m gpu: every HBM-bound kernel of the train step in isolation, through the op-level C-ABI (csrc/ops_api.cu), against
torch fp32 of the same operator on identical bf16-rounded inputs (TF32 off).

Tolerances: tensors the kernels store as bf16 are compared at half a bf16 ulp of the value range (2^-8 relative to
max|ref|); fp32 reductions at summation-order level (1e-4 relative to the norm); index / routing work (max-pool
arg-max, weight packing, gradient un-packing) is bit-exact.
Reference operators: nn.BatchNorm2d train mode + ReLU (reference models/unet.py:78-82,121-122,135-136),
nn.MaxPool2d(2,2) (:86,92), conv1x1 head (:342) + nn.CrossEntropyLoss(weight) (pipeline.py:135-138,176),
optim.SGD(momentum) + ExponentialLR stepped every lr_step batches (pipeline.py:156-157,178,188-189)."""
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
    lib = L.load()
    lib.crimac_op_scratch_bytes.restype = ctypes.c_size_t
    dev = torch.device("cuda:0")
    scratch = torch.zeros(lib.crimac_op_scratch_bytes(), dtype=torch.uint8, device=dev)
    return L, lib, dev, scratch


def _nchw(t):
    return t.float().permute(0, 3, 1, 2)


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def _close_bf16(got, ref, what=""):
    tol = 2 ** -8 * max(ref.abs().max().item(), 1e-3)
    err = (got.float() - ref.float()).abs().max().item()
    assert err <= tol, f"{what}: max abs err {err} > {tol}"


def _rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def _first_argmax_2x2(act_nhwc):
    """Index (row*2+col) of the FIRST maximal element of every 2x2 window in scan order (PyTorch's max-pool tie-break)."""
    a = act_nhwc.float()
    w = [a[:, 0::2, 0::2], a[:, 0::2, 1::2], a[:, 1::2, 0::2], a[:, 1::2, 1::2]]
    best, arg = w[0].clone(), torch.zeros_like(w[0], dtype=torch.int32)
    for i in (1, 2, 3):
        m = w[i] > best
        best = torch.where(m, w[i], best)
        arg = torch.where(m, torch.full_like(arg, i), arg)
    return arg


@pytest.mark.parametrize("N,H,W,C,rows", [(2, 32, 32, 64, 7), (3, 16, 24, 128, 148), (1, 8, 8, 1024, 3), (2, 20, 12, 256, 33)])
def test_batchnorm_train_forward_finalize_apply_pool(env, N, H, W, C, rows):
    L, lib, dev, _ = env
    torch.manual_seed(C + rows)
    raw = (torch.randn(N, H, W, C, device=dev) * 1.7 + 0.3).bfloat16()
    gamma, beta = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev) * 0.2
    gamma[::5] *= -1                                         # negative scales exist in trained nets
    rm, rv = torch.randn(C, device=dev) * 0.1, torch.rand(C, device=dev) + 0.5
    nbt = torch.tensor(41, dtype=torch.int64, device=dev)
    # partial rows as the conv kernels' EPI_STATS epilogue leaves them: [rows][2][C] fp32 sums over disjoint pixel sets
    flat = raw.float().reshape(-1, C)
    bounds = torch.linspace(0, flat.shape[0], rows + 1).long().tolist()
    part = torch.zeros(rows, 2, C, device=dev)
    for i in range(rows):
        seg = flat[bounds[i]:bounds[i + 1]]
        part[i, 0], part[i, 1] = seg.sum(0), (seg * seg).sum(0)
    scale, shift, mean, invstd = (torch.empty(C, device=dev) for _ in range(4))
    rm_ref, rv_ref = rm.clone(), rv.clone()
    ref = F.batch_norm(_nchw(raw), rm_ref, rv_ref, gamma, beta, training=True, momentum=0.1, eps=1e-5)
    L.check(lib.crimac_op_bn_finalize(L.ptr(part), rows, C, ctypes.c_double(N * H * W), L.ptr(gamma), L.ptr(beta),
                                      L.ptr(rm), L.ptr(rv), L.ptr(nbt), ctypes.c_float(0.1), ctypes.c_float(1e-5),
                                      L.ptr(scale), L.ptr(shift), L.ptr(mean), L.ptr(invstd), L.stream_ptr()),
            "crimac_op_bn_finalize")
    torch.cuda.synchronize()
    x64 = raw.double().reshape(-1, C)
    assert torch.allclose(mean.double(), x64.mean(0), rtol=1e-5, atol=1e-6)
    assert torch.allclose(invstd.double(), 1.0 / torch.sqrt(x64.var(0, unbiased=False) + 1e-5), rtol=1e-4)
    assert torch.allclose(rm, rm_ref, rtol=1e-5, atol=1e-6) and torch.allclose(rv, rv_ref, rtol=1e-4, atol=1e-6)
    assert int(nbt) == 42
    # apply + ReLU (+ pool + arg-max map)
    pooling = (H % 2 == 0 and W % 2 == 0)
    act = torch.zeros(N, H, W, C + 16, device=dev, dtype=torch.bfloat16)         # written through a pitched view
    pool = torch.zeros(N, H // 2, W // 2, C, device=dev, dtype=torch.bfloat16)
    arg = torch.zeros(N, H // 2, W // 2, C // 8, device=dev, dtype=torch.int16)
    L.check(lib.crimac_op_bn_apply(L.ptr(raw), C, N, H, W, C, L.ptr(scale), L.ptr(shift), L.ptr(act), C + 16,
                                   L.ptr(pool) if pooling else None, C, L.ptr(arg) if pooling else None, L.stream_ptr()),
            "crimac_op_bn_apply")
    torch.cuda.synchronize()
    _close_bf16(_nchw(act[..., :C]), torch.relu(ref), "bn_apply")
    assert torch.all(act[..., C:] == 0)                                          # nothing outside the view is touched
    if pooling:
        stored = act[..., :C]
        assert torch.equal(_nchw(pool), F.max_pool2d(_nchw(stored), 2))          # pool of the stored values: bit exact
        want = _first_argmax_2x2(stored)                                         # (N,H/2,W/2,C)
        got = torch.stack([(arg.int() & 0xFFFF) >> (2 * j) & 3 for j in range(8)], -1).reshape(N, H // 2, W // 2, C)
        assert torch.equal(got, want)


@pytest.mark.parametrize("N,H,W,C,gs", [(2, 32, 32, 64, None), (2, 16, 24, 128, 0.37), (4, 8, 8, 512, None), (1, 40, 24, 64, None)])
def test_batchnorm_relu_backward(env, N, H, W, C, gs):
    L, lib, dev, scratch = env
    torch.manual_seed(C + N)
    raw = (torch.randn(N, H, W, C, device=dev) * 1.3 - 0.2).bfloat16()
    dact = (torch.randn(N, H, W, C, device=dev) * 0.01).bfloat16()
    gamma = (torch.rand(C, device=dev) + 0.5).requires_grad_(True)
    beta = (torch.randn(C, device=dev) * 0.3).requires_grad_(True)
    x = _nchw(raw).clone().requires_grad_(True)
    y = F.batch_norm(x, None, None, gamma, beta, training=True, eps=1e-5)
    s = 1.0 if gs is None else gs
    torch.relu(y).backward(_nchw(dact) * s)
    x64 = raw.double().reshape(-1, C)
    mean = x64.mean(0).float()
    invstd = (1.0 / torch.sqrt(x64.var(0, unbiased=False) + 1e-5)).float()
    scale = (gamma.detach() * invstd).contiguous()
    shift = (beta.detach() - mean * scale).contiguous()
    draw = torch.zeros(N, H, W, C, device=dev, dtype=torch.bfloat16)
    dgamma, dbeta, dbias = (torch.full((C,), 7.0, device=dev) for _ in range(3))
    gsc = None if gs is None else torch.tensor([gs], device=dev)
    L.check(lib.crimac_op_bn_bwd(L.ptr(dact), C, L.ptr(raw), C, N, H, W, C, L.ptr(scale), L.ptr(shift), L.ptr(mean),
                                 L.ptr(invstd), L.ptr(draw), C, L.ptr(dgamma), L.ptr(dbeta), L.ptr(dbias), L.ptr(gsc),
                                 L.ptr(scratch), L.stream_ptr()), "crimac_op_bn_bwd")
    torch.cuda.synchronize()
    # pixels whose BatchNorm output is within rounding of 0 may take the other side of the ReLU: leave them out of the
    # elementwise comparison (their number is bounded below)
    sure = (y.detach().abs() > 1e-5)
    assert (~sure).float().mean().item() < 1e-4
    ref = x.grad
    tol = 2 ** -8 * ref.abs().max().item()
    err = ((_nchw(draw) - ref).abs() * sure).max().item()
    assert err <= tol, f"dRaw: max abs err {err} > {tol}"
    assert _rel(dgamma, gamma.grad) < 2e-4 and _rel(dbeta, beta.grad) < 2e-4
    assert torch.all(dbias == 0)                 # train-mode BatchNorm removes any conv bias: its gradient is exactly 0
    print(f"bn_bwd C={C}: dRaw max err {err:.3e} (tol {tol:.3e}); dgamma rel {_rel(dgamma, gamma.grad):.2e}, dbeta rel {_rel(dbeta, beta.grad):.2e}")


@pytest.mark.parametrize("N,H,W,C,skip", [(2, 32, 32, 64, True), (1, 16, 48, 256, True), (2, 8, 8, 512, False)])
def test_maxpool_backward_with_skip_add_is_bit_exact(env, N, H, W, C, skip):
    L, lib, dev, _ = env
    torch.manual_seed(C)
    raw = torch.randn(N, H, W, C, device=dev).bfloat16()
    one, zero = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    act = torch.zeros(N, H, W, C, device=dev, dtype=torch.bfloat16)
    pool = torch.zeros(N, H // 2, W // 2, C, device=dev, dtype=torch.bfloat16)
    arg = torch.zeros(N, H // 2, W // 2, C // 8, device=dev, dtype=torch.int16)
    L.check(lib.crimac_op_bn_apply(L.ptr(raw), C, N, H, W, C, L.ptr(one), L.ptr(zero), L.ptr(act), C, L.ptr(pool), C,
                                   L.ptr(arg), L.stream_ptr()), "crimac_op_bn_apply")     # relu(raw): half the windows tie at 0
    dpool = torch.randn(N, H // 2, W // 2, C, device=dev).bfloat16()
    cat = torch.randn(N, H, W, 2 * C, device=dev).bfloat16()                              # dskip = upper half of a concat gradient
    dskip = cat[..., C:]
    dact = torch.zeros(N, H, W, C, device=dev, dtype=torch.bfloat16)
    L.check(lib.crimac_op_pool_bwd_add(L.ptr(arg), L.ptr(dpool), C, L.ptr(dskip) if skip else None, 2 * C, L.ptr(dact), C,
                                       N, H, W, C, L.stream_ptr()), "crimac_op_pool_bwd_add")
    torch.cuda.synchronize()
    a = _nchw(act).clone().requires_grad_(True)
    F.max_pool2d(a, 2).backward(_nchw(dpool))
    ref = a.grad + (_nchw(dskip) if skip else 0.0)
    assert torch.equal(_nchw(dact), ref.bfloat16().float())
    assert (_nchw(act) == 0).float().mean().item() > 0.3       # the tie-break really was exercised


@pytest.mark.parametrize("N,H,W,C", [(2, 32, 32, 64), (1, 16, 48, 256), (3, 8, 8, 512)])
def test_batchnorm_backward_with_fused_maxpool_backward_and_skip_add(env, N, H, W, C):
    """What the encoder's second convs run: BatchNorm + ReLU backward whose incoming gradient dA = dSkip + unpool(dPool)
    (autograd of unet.py:86,92,132) is formed inside the kernel from the forward's arg-max map - against torch autograd of
    conv-output -> BatchNorm(train) -> ReLU -> {skip branch, MaxPool2d(2,2)} on the same bf16-rounded inputs."""
    L, lib, dev, scratch = env
    torch.manual_seed(C + H)
    raw = (torch.randn(N, H, W, C, device=dev) * 1.1 + 0.1).bfloat16()
    gamma = (torch.rand(C, device=dev) + 0.5).requires_grad_(True)
    beta = (torch.randn(C, device=dev) * 0.3).requires_grad_(True)
    x64 = raw.double().reshape(-1, C)
    mean = x64.mean(0).float()
    invstd = (1.0 / torch.sqrt(x64.var(0, unbiased=False) + 1e-5)).float()
    scale = (gamma.detach() * invstd).contiguous()
    shift = (beta.detach() - mean * scale).contiguous()
    act = torch.zeros(N, H, W, C, device=dev, dtype=torch.bfloat16)
    pool = torch.zeros(N, H // 2, W // 2, C, device=dev, dtype=torch.bfloat16)
    arg = torch.zeros(N, H // 2, W // 2, C // 8, device=dev, dtype=torch.int16)
    L.check(lib.crimac_op_bn_apply(L.ptr(raw), C, N, H, W, C, L.ptr(scale), L.ptr(shift), L.ptr(act), C, L.ptr(pool), C,
                                   L.ptr(arg), L.stream_ptr()), "crimac_op_bn_apply")
    dpool = (torch.randn(N, H // 2, W // 2, C, device=dev) * 0.02).bfloat16()
    dcat = (torch.randn(N, H, W, 2 * C, device=dev) * 0.01).bfloat16()
    dskip = dcat[..., C:]
    draw = torch.zeros(N, H, W, C, device=dev, dtype=torch.bfloat16)
    dgamma, dbeta, dbias = (torch.full((C,), 7.0, device=dev) for _ in range(3))
    L.check(lib.crimac_op_bn_bwd_pool(L.ptr(arg), L.ptr(dpool), C, L.ptr(dskip), 2 * C, L.ptr(raw), C, N, H, W, C,
                                      L.ptr(scale), L.ptr(shift), L.ptr(mean), L.ptr(invstd), L.ptr(draw), C, L.ptr(dgamma),
                                      L.ptr(dbeta), L.ptr(dbias), L.ptr(scratch), L.stream_ptr()), "crimac_op_bn_bwd_pool")
    torch.cuda.synchronize()
    # reference: the pool routes to the first maximum of the STORED (bf16) activation, as the forward kernel decided
    x = _nchw(raw).clone().requires_grad_(True)
    y = F.batch_norm(x, None, None, gamma, beta, training=True, eps=1e-5)
    a = torch.relu(y)
    a_st = a + (_nchw(act) - a).detach()
    (F.max_pool2d(a_st, 2) * _nchw(dpool)).sum().backward(retain_graph=True)
    (a_st * _nchw(dskip)).sum().backward()
    sure = (y.detach().abs() > 1e-5)
    tol = 2 ** -8 * x.grad.abs().max().item()
    err = ((_nchw(draw) - x.grad).abs() * sure).max().item()
    assert err <= tol, f"dRaw: max abs err {err} > {tol}"
    assert _rel(dgamma, gamma.grad) < 2e-4 and _rel(dbeta, beta.grad) < 2e-4
    assert torch.all(dbias == 0)


@pytest.mark.parametrize("ncls,N,H,W", [(3, 2, 32, 32), (2, 1, 24, 40), (8, 1, 16, 16)])
def test_head_cross_entropy_fused_forward_backward(env, ncls, N, H, W):
    L, lib, dev, scratch = env
    torch.manual_seed(ncls)
    act = torch.relu(torch.randn(N, H, W, 64, device=dev)).bfloat16()
    hw = (torch.randn(ncls, 64, device=dev) * 0.2).requires_grad_(True)
    hb = torch.randn(ncls, device=dev).requires_grad_(True)
    cw = torch.tensor([10.0, 300.0, 250.0, 1.0, 2.0, 3.0, 4.0, 5.0], device=dev)[:ncls].contiguous()
    y = torch.randint(0, ncls, (N, H, W), device=dev)
    y[torch.rand(N, H, W, device=dev) < 0.1] = -100
    a = _nchw(act).clone().requires_grad_(True)
    logits = F.conv2d(a, hw[:, :, None, None], hb)
    loss = F.cross_entropy(logits, y, weight=cw, ignore_index=-100)
    loss.backward()
    dact = torch.zeros(N, H, W, 64, device=dev, dtype=torch.bfloat16)
    dw, db, out3 = torch.zeros(ncls, 64, device=dev), torch.zeros(ncls, device=dev), torch.zeros(4, device=dev)
    L.check(lib.crimac_op_head_ce(L.ptr(act), 64, N, H, W, L.ptr(hw), L.ptr(hb), ncls, L.ptr(y), L.ptr(cw),
                                  ctypes.c_int64(-100), L.ptr(dact), 64, L.ptr(dw), L.ptr(db), L.ptr(out3), L.ptr(scratch),
                                  L.stream_ptr()), "crimac_op_head_ce")
    torch.cuda.synchronize()
    assert abs(out3[0].item() - loss.item()) < 1e-5 * abs(loss.item())
    assert abs(out3[1].item() * out3[2].item() - 1.0) < 1e-6
    assert abs(out3[2].item() - cw[y[y >= 0]].sum().item()) < 1e-5 * out3[2].item()
    _close_bf16(_nchw(dact) * out3[1], a.grad, "head dAct")      # the kernel leaves dAct un-normalised (x 1/sum_w later)
    assert _rel(dw, hw.grad) < 1e-4 and _rel(db, hb.grad) < 1e-4
    # the un-fused pair used by the autograd path
    lg = torch.zeros(N, ncls, H, W, device=dev)
    L.check(lib.crimac_op_head_fwd(L.ptr(act), 64, N, H, W, L.ptr(hw), L.ptr(hb), ncls, L.ptr(lg), L.stream_ptr()), "head_fwd")
    dl = torch.autograd.grad(F.cross_entropy(logits2 := lg.clone().requires_grad_(True), y, weight=cw), logits2)[0].contiguous()
    dact2, dw2, db2 = torch.zeros_like(dact), torch.zeros_like(dw), torch.zeros_like(db)
    L.check(lib.crimac_op_head_bwd(L.ptr(dl), None, L.ptr(act), 64, N, H, W, L.ptr(hw), ncls, L.ptr(dact2), 64, L.ptr(dw2),
                                   L.ptr(db2), L.ptr(scratch), L.stream_ptr()), "head_bwd")
    torch.cuda.synchronize()
    assert (lg - logits).abs().max().item() < 1e-4
    _close_bf16(_nchw(dact2), a.grad, "head_bwd dAct")
    assert _rel(dw2, hw.grad) < 1e-4 and _rel(db2, hb.grad) < 1e-4
    # edge cases of nn.CrossEntropyLoss: every pixel ignored -> NaN; a label outside [0, n_classes) is an error there
    # (device assert) and poisons the loss here
    y_all = torch.full_like(y, -100)
    L.check(lib.crimac_op_head_ce(L.ptr(act), 64, N, H, W, L.ptr(hw), L.ptr(hb), ncls, L.ptr(y_all), L.ptr(cw),
                                  ctypes.c_int64(-100), L.ptr(dact), 64, L.ptr(dw), L.ptr(db), L.ptr(out3), L.ptr(scratch),
                                  L.stream_ptr()), "crimac_op_head_ce")
    assert torch.isnan(out3[0])
    y_bad = y.clone()
    y_bad[0, 0, 0] = ncls
    L.check(lib.crimac_op_head_ce(L.ptr(act), 64, N, H, W, L.ptr(hw), L.ptr(hb), ncls, L.ptr(y_bad), L.ptr(cw),
                                  ctypes.c_int64(-100), L.ptr(dact), 64, L.ptr(dw), L.ptr(db), L.ptr(out3), L.ptr(scratch),
                                  L.stream_ptr()), "crimac_op_head_ce")
    assert torch.isnan(out3[0])


def test_convtranspose_bias_gradient_column_sum(env):
    L, lib, dev, scratch = env
    torch.manual_seed(9)
    cat = torch.randn(2, 32, 48, 256, device=dev).bfloat16()
    out = torch.zeros(128, device=dev)
    L.check(lib.crimac_op_colsum(L.ptr(cat), 256, 2, 32, 48, 128, L.ptr(out), L.ptr(scratch), L.stream_ptr()), "colsum")
    torch.cuda.synchronize()
    assert _rel(out, cat[..., :128].double().sum((0, 1, 2)).float()) < 1e-5


def test_weight_packing_is_bit_exact(env):
    L, lib, dev, _ = env
    torch.manual_seed(10)
    for cout, cin in ((64, 64), (128, 320), (8, 64)):
        w = torch.randn(cout, cin, 3, 3, device=dev)
        out = torch.zeros(cout, 9 * cin, device=dev, dtype=torch.bfloat16)
        L.check(lib.crimac_op_pack(0, L.ptr(w), cout, cin, L.ptr(out), L.stream_ptr()), "pack conv")
        torch.cuda.synchronize()
        assert torch.equal(out, w.permute(0, 2, 3, 1).reshape(cout, 9 * cin).bfloat16())
    for cin, cout in ((128, 64), (1024, 512)):
        w = torch.randn(cin, cout, 2, 2, device=dev)
        out = torch.zeros(4 * cout, cin, device=dev, dtype=torch.bfloat16)
        L.check(lib.crimac_op_pack(1, L.ptr(w), cout, cin, L.ptr(out), L.stream_ptr()), "pack convT")
        torch.cuda.synchronize()
        assert torch.equal(out, w.permute(2, 3, 1, 0).reshape(4 * cout, cin).bfloat16())


def test_weight_gradient_unpack_all_layers_one_launch(env):
    L, lib, dev, _ = env
    torch.manual_seed(11)
    shapes = [(9, 64 * 64), (4, 128 * 64), (9, 100), (9, 1024 + 300), (1, 5000)]       # ragged: not multiples of 256 / 1024
    scr = [torch.randn(t, mn, device=dev) for t, mn in shapes]
    want = [s.t().clone(memory_format=torch.contiguous_format) for s in scr]   # a COPY: the kernel zeroes scr
    dws = [torch.full((mn, t), 3.0, device=dev) for t, mn in shapes]
    n = len(shapes)
    a_s = (ctypes.c_void_p * n)(*[s.data_ptr() for s in scr])
    a_d = (ctypes.c_void_p * n)(*[d.data_ptr() for d in dws])
    a_mn = (ctypes.c_int64 * n)(*[mn for _, mn in shapes])
    a_t = (ctypes.c_int * n)(*[t for t, _ in shapes])
    L.check(lib.crimac_op_wgrad_unpack_all(n, a_s, a_d, a_mn, a_t, L.stream_ptr()), "unpack_all")
    torch.cuda.synchronize()
    for s, d, w in zip(scr, dws, want):
        assert torch.equal(d, w)
        assert torch.all(s == 0)                # the scratch is left zeroed for the next step's red.add accumulation


def test_sgd_momentum_kernel_and_lr_schedule_match_torch_optim(env, pkg):
    """crimac_sgd_step + Trainer's learning-rate rule against optim.SGD(momentum=0.95) + ExponentialLR(gamma) stepped
    every lr_step batches (pipeline.py:156-157,178,188-189), over 9 steps = 4 learning-rate changes."""
    L, lib, dev, _ = env
    E = importlib.import_module("crimac_unet_b200.engine")
    T = importlib.import_module("crimac_unet_b200.trainer")
    torch.manual_seed(12)
    n = 1_000_003
    p0 = torch.randn(n, device=dev)
    p_ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.SGD([p_ref], lr=0.005, momentum=0.95)
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.5)

    class _Holder(torch.nn.Module):             # the smallest thing Trainer accepts: parameters() + _native_mutated()
        def __init__(self):
            super().__init__()
            self.p = torch.nn.Parameter(p0.clone())

        def _native_mutated(self):
            pass

    tr = T.Trainer(_Holder(), lr=0.005, momentum=0.95, lr_reduction=0.5, lr_step=2, use_cuda_graph=False)
    for i in range(9):
        g = torch.randn(n, device=dev) * (1.0 + i)
        p_ref.grad = g.clone()
        opt.step()
        E.sgd_step(tr.flat_params, tr.flat_momentum, g, tr.lr, tr.momentum, 1.0)
        tr._advance()
        if (i + 1) % 2 == 0:
            sched.step()
        assert abs(tr.lr - opt.param_groups[0]["lr"]) < 1e-12, (i, tr.lr, opt.param_groups[0]["lr"])
        torch.cuda.synchronize()
        assert torch.allclose(tr.flat_params, p_ref.detach(), rtol=1e-5, atol=1e-6), i
    assert abs(tr.lr - 0.005 * 0.5 ** 4) < 1e-15
    vb = opt.state[p_ref]["momentum_buffer"]
    # summation-order level: the running sums reach ~30, so the absolute rounding error is ~30 * 2^-23 per step
    assert (tr.flat_momentum - vb).abs().max().item() <= 1e-5 * vb.abs().max().item()
    # the data-parallel mean is folded into the kernel: gscale = 1/world
    p1, v1 = p0.clone(), torch.zeros(n, device=dev)
    g = torch.randn(n, device=dev)
    E.sgd_step(p1, v1, g, 0.1, 0.9, 0.25)
    torch.cuda.synchronize()
    assert torch.allclose(p1, p0 - 0.1 * 0.25 * g, rtol=1e-6, atol=1e-7)


def test_validation_loss_label_remap_and_sandeel_probability(env):
    """crimac_eval_loss against the reference's validation step (pipeline.py:222-239 set_label_ignore_val, :264
    criterion, :269-270 softmax + SANDEEL channel), restated in torch."""
    L, lib, dev, scratch = env
    torch.manual_seed(13)
    N, H, W = 3, 40, 56
    logits = torch.randn(N, 3, H, W, device=dev) * 3
    codes = torch.tensor([0, 1, 2, -100, -70, -50, -30, -10], device=dev)
    y16 = codes[torch.randint(0, 8, (N, H, W), device=dev)].to(torch.int16)
    cw = torch.tensor([10.0, 300.0, 250.0], device=dev)
    ref_lab = y16.long().clone()
    for v in (-70, -30, -100, -10):
        ref_lab[ref_lab == v] = -100
    ref_lab[ref_lab == -50] = 0
    ref_loss = F.cross_entropy(logits, ref_lab, weight=cw)
    ref_prob = torch.softmax(logits, 1)[:, 1]
    for bits, lab in ((16, y16), (64, y16.long())):
        prob = torch.zeros(N, H, W, device=dev)
        lab_out = torch.zeros(N, H, W, dtype=torch.int64, device=dev)
        out3 = torch.zeros(4, device=dev)
        L.check(lib.crimac_eval_loss(L.ptr(logits), N, 3, H, W, L.ptr(lab), bits, L.ptr(cw), 1, L.ptr(prob),
                                     L.ptr(lab_out), L.ptr(out3), L.ptr(scratch), L.stream_ptr()), "crimac_eval_loss")
        torch.cuda.synchronize()
        assert torch.equal(lab_out, ref_lab)
        assert abs(out3[0].item() - ref_loss.item()) < 1e-5 * abs(ref_loss.item())
        assert (prob - ref_prob).abs().max().item() < 1e-6


@pytest.mark.parametrize("N,H,W,C", [(2, 16, 24, 64), (1, 1, 1, 8), (1, 5, 2, 128)])
def test_bilinear_upsample_forward_and_adjoint(env, N, H, W, C):
    """up_mode "upsample" (reference unet.py:50-56): nn.Upsample(mode="bilinear", scale_factor=2), align_corners=False."""
    L, lib, dev, _ = env
    torch.manual_seed(H * W + C)
    cat = torch.zeros(N, 2 * H, 2 * W, 2 * C, device=dev, dtype=torch.bfloat16)        # written through a concat view
    lo = torch.randn(N, H, W, C, device=dev).bfloat16()
    L.check(lib.crimac_op_upsample2x(L.ptr(lo), C, L.ptr(cat), 2 * C, N, H, W, C, 0, L.stream_ptr()), "upsample2x")
    torch.cuda.synchronize()
    x = _nchw(lo).clone().requires_grad_(True)
    ref = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
    _close_bf16(_nchw(cat[..., :C]), ref.detach(), "upsample")
    assert torch.all(cat[..., C:] == 0)
    g = torch.randn(N, 2 * H, 2 * W, 2 * C, device=dev).bfloat16()
    dlo = torch.zeros(N, H, W, C, device=dev, dtype=torch.bfloat16)
    L.check(lib.crimac_op_upsample2x(L.ptr(dlo), C, L.ptr(g), 2 * C, N, H, W, C, 1, L.stream_ptr()), "upsample2x bwd")
    torch.cuda.synchronize()
    ref.backward(_nchw(g[..., :C]))
    _close_bf16(_nchw(dlo), x.grad, "upsample adjoint")

"""Hardware probes run on the B200 box (tools/run_probe.sh): pin down UMMA descriptor / TMA swizzle semantics and
check the op-level kernels against torch on the GPU. Prints one line per check; exits 0 always (results are read
from the log)."""
import ctypes
import importlib.util
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "crimac-classifiers-unet_b200")
spec = importlib.util.spec_from_file_location("crimac_lib", os.path.join(PKG, "lib.py"))
L = importlib.util.module_from_spec(spec)
spec.loader.exec_module(L)
lib = L.load()
dev = torch.device("cuda:0")
torch.manual_seed(0)


def desc(off, lbo, sbo, base_offset=0, layout=2):
    d = (off >> 4) & 0x3FFF
    d |= ((lbo >> 4) & 0x3FFF) << 16
    d |= ((sbo >> 4) & 0x3FFF) << 32
    d |= 1 << 46
    d |= (base_offset & 7) << 49
    d |= (layout & 7) << 61
    return d


def idesc(M, N, amn, bmn):
    return (1 << 4) | (1 << 7) | (1 << 10) | (amn << 15) | (bmn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


def sw128(mat):
    """[R,64] bf16 -> flat int16 image in the 128B-swizzled order TMA writes / UMMA reads."""
    R = mat.shape[0]
    r = torch.arange(R)[:, None]
    c = torch.arange(64)[None, :]
    off = (r // 8) * 512 + (r % 8) * 64 + (((c // 8) ^ (r % 8)) * 8) + (c % 8)
    img = torch.zeros(((R + 7) // 8) * 512, dtype=torch.int16)
    img[off.flatten()] = mat.contiguous().view(torch.int16).flatten().cpu()
    return img


def run_umma(image, a_desc, b_desc, idsc, n_mma, a_step, b_step, N):
    img = image.to(dev)
    out = torch.full((128, N), float("nan"), device=dev)
    rc = lib.crimac_dbg_umma(L.ptr(img), ctypes.c_int(img.numel() * 2), ctypes.c_uint64(a_desc),
                             ctypes.c_uint64(b_desc), ctypes.c_uint32(idsc), n_mma, a_step, b_step, L.ptr(out), N,
                             L.stream_ptr())
    L.check(rc, "dbg_umma")
    torch.cuda.synchronize()
    return out.cpu()


def report(name, got, ref, tol=None):
    err = (got.float() - ref.float()).abs().max().item()
    scale = ref.float().abs().max().item()
    ok = err <= (tol if tol is not None else 1e-2 * max(scale, 1.0))
    print(f"[{'OK ' if ok else 'BAD'}] {name}: max_abs_err={err:.4g} ref_max={scale:.4g}", flush=True)
    return ok


def t_umma_kmajor():
    A = torch.randn(128, 64).bfloat16()
    B = torch.randn(64, 64).bfloat16()
    image = torch.cat([sw128(A), sw128(B)])
    got = run_umma(image, desc(0, 16, 1024), desc(16384, 16, 1024), idesc(128, 64, 0, 0), 4, 32, 32, 64)
    report("umma K-major SW128 (conv main loop layout)", got, A.float() @ B.float().T)


def t_umma_mnmajor():
    F = torch.randn(64, 128).bfloat16()  # [pixels(K)][m channels]
    T = torch.randn(64, 64).bfloat16()   # [pixels(K)][n channels]
    image = torch.cat([sw128(F[:, :64]), sw128(F[:, 64:]), sw128(T)])
    ref = F.float().T @ T.float()
    for lbo, sbo in ((8192, 1024),):
        got = run_umma(image, desc(0, lbo, sbo), desc(16384, lbo, sbo), idesc(128, 64, 1, 1), 4, 2048, 2048, 64)
        report(f"umma MN-major SW128 LBO={lbo} SBO={sbo} (wgrad layout)", got, ref)


def t_umma_unaligned():
    B = torch.randn(64, 64).bfloat16()
    # (c) plain row shift of a dense tile
    A = torch.randn(144, 64).bfloat16()
    for r0 in (8, 1, 3):
        for bo in sorted({0, r0 & 7}):
            image = torch.cat([sw128(A), sw128(B)])
            boff = sw128(A).numel() * 2
            got = run_umma(image, desc(r0 * 128, 16, 1024, bo), desc(boff, 16, 1024), idesc(128, 64, 0, 0), 4, 32,
                           32, 64)
            report(f"umma K-major start shifted by {r0} rows, base_offset={bo}", got,
                   A[r0:r0 + 128].float() @ B.float().T)
    # (d) halo tile: 18 x 10 pixels, output tile 16 rows x 8 pixels, SBO = 10*128
    Hh = torch.randn(184, 64).bfloat16()
    for dy, dx in ((0, 0), (0, 1), (1, 0), (1, 1), (2, 2)):
        rows = torch.tensor([(ty + dy) * 10 + tx + dx for ty in range(16) for tx in range(8)])
        ref = Hh[rows].float() @ B.float().T
        s = dy * 10 + dx
        for bo in sorted({0, s & 7}):
            image = torch.cat([sw128(Hh), sw128(B)])
            boff = sw128(Hh).numel() * 2
            got = run_umma(image, desc(s * 128, 16, 1280, bo), desc(boff, 16, 1024), idesc(128, 64, 0, 0), 4, 32, 32,
                           64)
            report(f"umma halo tile tap(dy={dy},dx={dx}) SBO=1280 base_offset={bo}", got, ref)


def t_tma():
    NB, H, W, C = 2, 16, 32, 128
    x = torch.randn(NB, H, W, C, device=dev).bfloat16()
    out = torch.zeros(128 * 64, dtype=torch.int16, device=dev)
    for (c0, x0, y0, n0) in ((0, 0, 0, 0), (64, -1, -1, 1), (64, 17, 9, 1)):
        rc = lib.crimac_dbg_tma_box(L.ptr(x), NB, H, W, C, C, 8, 0, 0, 0, c0, x0, y0, n0, L.ptr(out), L.stream_ptr())
        L.check(rc, "dbg_tma_box")
        torch.cuda.synchronize()
        exp = torch.zeros(8, 16, 64, dtype=torch.bfloat16)
        xc = x.cpu()
        for yy in range(8):
            for xx in range(16):
                y, xg = y0 + yy, x0 + xx
                if 0 <= y < H and 0 <= xg < W:
                    exp[yy, xx] = xc[n0, y, xg, c0:c0 + 64]
        ok = torch.equal(sw128(exp.reshape(128, 64)), out.cpu())
        print(f"[{'OK ' if ok else 'BAD'}] tma box at c0={c0} x0={x0} y0={y0} n={n0}: swizzle + zero fill", flush=True)
    # sub-sampled view
    rc = lib.crimac_dbg_tma_box(L.ptr(x), NB, H, W, C, C, 8, 1, 1, 0, 0, 0, 0, 0, L.ptr(out), L.stream_ptr())
    L.check(rc, "dbg_tma_box sub")
    torch.cuda.synchronize()
    exp = torch.zeros(8, 16, 64, dtype=torch.bfloat16)
    exp[:, :, :] = x.cpu()[0, 1::2, 0::2, :64][:8, :16]
    ok = torch.equal(sw128(exp.reshape(128, 64)), out.cpu())
    print(f"[{'OK ' if ok else 'BAD'}] tma sub-sampled (ky=1,kx=0) view", flush=True)


def igemm(mode, x, w, n_total, scale=None, shift=None, relu=0, out=None, out_pitch=0, convt_cout=0, pool=None,
          stats=None, head=None, block_n=0, H=None, W=None):
    NB, Hx, Wx, pitch = x.shape
    cin = pitch
    H = H or Hx
    W = W or Wx
    hw, hb, hout, ncls, sm = (None, None, None, 0, 0) if head is None else head
    rc = lib.crimac_op_igemm(mode, L.ptr(x), NB, H, W, cin, pitch, L.ptr(w), n_total, L.ptr(scale), L.ptr(shift), relu,
                             L.ptr(out), out_pitch, convt_cout, L.ptr(pool), 0 if pool is None else pool.shape[-1],
                             L.ptr(stats), L.ptr(hw), L.ptr(hb), L.ptr(hout), ncls, sm, block_n, L.stream_ptr())
    L.check(rc, "op_igemm")
    torch.cuda.synchronize()


def t_conv():
    import torch.nn.functional as F
    for (NB, H, W, Cin, Cout, bn) in ((2, 32, 32, 64, 64, 0), (1, 16, 16, 128, 256, 0), (2, 32, 48, 128, 128, 64),
                                      (3, 16, 16, 256, 512, 0)):
        x = torch.randn(NB, H, W, Cin, device=dev).bfloat16()
        w = (torch.randn(Cout, Cin, 3, 3, device=dev) / (3 * Cin ** 0.5)).bfloat16()
        wp = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
        scale = torch.rand(Cout, device=dev) + 0.5
        shift = torch.randn(Cout, device=dev)
        out = torch.zeros(NB, H, W, Cout, device=dev, dtype=torch.bfloat16)
        pool = torch.zeros(NB, H // 2, W // 2, Cout, device=dev, dtype=torch.bfloat16)
        igemm(0, x, wp, Cout, scale, shift, 1, out, Cout, pool=pool, block_n=bn)
        ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), padding=1)
        ref = torch.relu(ref * scale[None, :, None, None] + shift[None, :, None, None])
        tag = f"conv3x3 NB={NB} {H}x{W} {Cin}->{Cout} bn={bn}"
        report(tag + " (affine+relu)", out.float().permute(0, 3, 1, 2).cpu(), ref.cpu(), 3e-2)
        report(tag + " (fused 2x2 max-pool)", pool.float().permute(0, 3, 1, 2).cpu(),
               F.max_pool2d(ref.bfloat16().float(), 2).cpu(), 3e-2)
        # train-mode epilogue
        stats = torch.zeros(256, 2, Cout, device=dev)
        raw = torch.zeros_like(out)
        igemm(0, x, wp, Cout, None, shift, 0, raw, Cout, stats=stats, block_n=bn)
        refraw = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), padding=1) + shift[None, :, None, None]
        report(tag + " (raw+bias)", raw.float().permute(0, 3, 1, 2).cpu(), refraw.cpu(), 3e-2)
        s = stats.sum(0).cpu()
        rr = raw.float()
        report(tag + " (stats sum)", s[0], rr.sum((0, 1, 2)).cpu(), 1e-3 * NB * H * W)
        report(tag + " (stats sumsq)", s[1], (rr * rr).sum((0, 1, 2)).cpu(), 1e-3 * NB * H * W)


def t_head():
    import torch.nn.functional as F
    NB, H, W, Cin = 2, 32, 32, 64
    x = torch.randn(NB, H, W, Cin, device=dev).bfloat16()
    w = (torch.randn(64, Cin, 3, 3, device=dev) / (3 * Cin ** 0.5)).bfloat16()
    wp = w.permute(0, 2, 3, 1).reshape(64, 9 * Cin).contiguous()
    scale = torch.rand(64, device=dev) + 0.5
    shift = torch.randn(64, device=dev) * 0.1
    hw = torch.randn(3, 64, device=dev) * 0.2
    hb = torch.randn(3, device=dev)
    probs = torch.zeros(NB, 3, H, W, device=dev)
    igemm(0, x, wp, 64, scale, shift, 1, None, 0, head=(hw, hb, probs, 3, 1))
    act = torch.relu(F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), padding=1) * scale[None, :, None, None] +
                     shift[None, :, None, None])
    ref = torch.softmax(F.conv2d(act, hw[:, :, None, None], hb), 1)
    report("conv3x3 + fused 1x1 head + softmax", probs.cpu(), ref.cpu(), 2e-2)


def t_convt():
    import torch.nn.functional as F
    NB, H, W, Cin, Cout = 2, 16, 16, 128, 64
    x = torch.randn(NB, H, W, Cin, device=dev).bfloat16()
    w = (torch.randn(Cin, Cout, 2, 2, device=dev) / Cin ** 0.5).bfloat16()
    b = torch.randn(Cout, device=dev)
    wp = w.permute(2, 3, 1, 0).reshape(4 * Cout, Cin).contiguous()  # [(ky,kx,co)][ci]
    cat = torch.zeros(NB, 2 * H, 2 * W, 2 * Cout, device=dev, dtype=torch.bfloat16)
    igemm(1, x, wp, 4 * Cout, None, b.repeat(4), 0, cat, 2 * Cout, convt_cout=Cout)
    ref = F.conv_transpose2d(x.float().permute(0, 3, 1, 2), w.float(), b, stride=2)
    report("convT 2x2 forward, scatter into concat buffer", cat[..., :Cout].float().permute(0, 3, 1, 2).cpu(),
           ref.cpu(), 3e-2)
    report("convT forward leaves skip half untouched", cat[..., Cout:].float().cpu(),
           torch.zeros(NB, 2 * H, 2 * W, Cout), 0.0)
    # backward-data: dX[p][ci] = sum_{kk,co} dY[sub_kk(p)][co] W[ci][co][kk]
    dy = torch.randn(NB, 2 * H, 2 * W, Cout, device=dev).bfloat16()
    wd = w.permute(0, 2, 3, 1).reshape(Cin, 4 * Cout).contiguous()  # [ci][(ky,kx,co)]
    dx = torch.zeros(NB, H, W, Cin, device=dev, dtype=torch.bfloat16)
    igemm(2, dy, wd, Cin, None, None, 0, dx, Cin, H=H, W=W)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    F.conv_transpose2d(xr, w.float(), b, stride=2).backward(dy.float().permute(0, 3, 1, 2))
    report("convT 2x2 backward-data", dx.float().permute(0, 3, 1, 2).cpu(), xr.grad.cpu(), 5e-2)


def t_dgrad():
    import torch.nn.functional as F
    NB, H, W, Cin, Cout = 2, 32, 32, 128, 64
    w = (torch.randn(Cout, Cin, 3, 3, device=dev) / (3 * Cin ** 0.5)).bfloat16()
    dy = torch.randn(NB, H, W, Cout, device=dev).bfloat16()
    wd = w.flip(2, 3).permute(1, 2, 3, 0).reshape(Cin, 9 * Cout).contiguous()  # [ci][(2-ky,2-kx)][co]
    dx = torch.zeros(NB, H, W, Cin, device=dev, dtype=torch.bfloat16)
    igemm(0, dy, wd, Cin, None, None, 0, dx, Cin)
    xr = torch.randn(NB, Cin, H, W, device=dev, requires_grad=True)
    F.conv2d(xr, w.float(), padding=1).backward(dy.float().permute(0, 3, 1, 2))
    report("conv3x3 backward-data via rotated weights", dx.float().permute(0, 3, 1, 2).cpu(), xr.grad.cpu(), 5e-2)


def t_wgrad():
    import torch.nn.functional as F
    for (NB, H, W, Cin, Cout, splits) in ((2, 32, 32, 64, 64, 1), (2, 32, 32, 128, 128, 4), (2, 16, 16, 256, 128, 0),
                                          (1, 32, 32, 128, 64, 3)):
        x = torch.randn(NB, H, W, Cin, device=dev).bfloat16()
        dy = torch.randn(NB, H, W, Cout, device=dev).bfloat16()
        scratch = torch.empty(9 * Cout * Cin, device=dev)
        dw = torch.zeros(Cout, Cin, 3, 3, device=dev)
        rc = lib.crimac_op_wgrad(0, L.ptr(dy), Cout, Cout, L.ptr(x), Cin, Cin, NB, H, W, L.ptr(scratch), L.ptr(dw),
                                 splits, 0, L.stream_ptr())
        L.check(rc, "op_wgrad")
        torch.cuda.synchronize()
        wr = torch.zeros(Cout, Cin, 3, 3, device=dev, requires_grad=True)
        F.conv2d(x.float().permute(0, 3, 1, 2), wr, padding=1).backward(dy.float().permute(0, 3, 1, 2))
        report(f"wgrad conv3x3 NB={NB} {H}x{W} {Cin}->{Cout} splits={splits}", dw.cpu(), wr.grad.cpu(),
               2e-3 * (NB * H * W) ** 0.5)
    NB, H, W, Cin, Cout = 2, 16, 16, 128, 64
    x = torch.randn(NB, H, W, Cin, device=dev).bfloat16()
    dy = torch.randn(NB, 2 * H, 2 * W, Cout, device=dev).bfloat16()
    scratch = torch.empty(4 * Cout * Cin, device=dev)
    dw = torch.zeros(Cin, Cout, 2, 2, device=dev)
    rc = lib.crimac_op_wgrad(1, L.ptr(x), Cin, Cin, L.ptr(dy), Cout, Cout, NB, H, W, L.ptr(scratch), L.ptr(dw), 2, 0,
                             L.stream_ptr())
    L.check(rc, "op_wgrad convT")
    torch.cuda.synchronize()
    wr = torch.zeros(Cin, Cout, 2, 2, device=dev, requires_grad=True)
    F.conv_transpose2d(x.float().permute(0, 3, 1, 2), wr, stride=2).backward(dy.float().permute(0, 3, 1, 2))
    report("wgrad convT 2x2", dw.cpu(), wr.grad.cpu(), 2e-3 * (NB * H * W) ** 0.5)


TESTS = {k[2:]: v for k, v in list(globals().items()) if k.startswith("t_")}

if __name__ == "__main__":
    names = sys.argv[1:] or list(TESTS)
    for n in names:
        print(f"=== {n}", flush=True)
        try:
            TESTS[n]()
        except Exception as e:  # noqa
            print(f"[EXC] {n}: {type(e).__name__}: {e}", flush=True)

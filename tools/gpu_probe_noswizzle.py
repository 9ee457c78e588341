"""Probe (B200): UMMA shared-memory descriptors WITHOUT swizzle (layout 0), in the core-matrix layout a TMA box of
{8 channels (16 B), 16 pings, 8 rows} produces: 128 pixel rows x 16 B contiguous = 16 core matrices of 8 rows x 16 B.
Used to decide the first-conv redesign (im2col by TMA: one box per tap).  Prints one line per check."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gpu_probe import desc, idesc, run_umma, report, sw128
import torch

torch.manual_seed(1)


def img_kmajor(mat, kstride, gstride, total):
    """mat [R, K] bf16 -> int16 image: element (r,k) at byte (k//8)*kstride + (r//8)*gstride + (r%8)*16 + (k%8)*2."""
    R, K = mat.shape
    r = torch.arange(R)[:, None]
    k = torch.arange(K)[None, :]
    off = ((k // 8) * kstride + (r // 8) * gstride + (r % 8) * 16 + (k % 8) * 2) // 2
    img = torch.zeros(total // 2, dtype=torch.int16)
    img[off.flatten()] = mat.contiguous().view(torch.int16).flatten()
    return img


def t_kmajor():
    A = torch.randn(128, 32).bfloat16()   # 2 k-slices of 16 = 4 "tap boxes" of 8
    B = torch.randn(64, 32).bfloat16()
    ia = img_kmajor(A, 2048, 128, 4 * 2048)          # tap boxes 2 KB apart, core matrices 128 B apart
    ib = img_kmajor(B, 1024, 128, 4 * 1024)
    image = torch.cat([ia, ib])
    boff = ia.numel() * 2
    ref = A.float() @ B.float().T
    for name, (la, sa, lb, sb) in {"LBO=K-dir,SBO=MN-dir": (2048, 128, 1024, 128)}.items():
        got = run_umma(image, desc(0, la, sa, 0, 0), desc(boff, lb, sb, 0, 0), idesc(128, 64, 0, 0), 2, 4096, 2048, 64)
        report(f"umma K-major NO-swizzle {name}", got, ref)


def t_mnmajor_a():
    # A^T: M = 128 "k'" rows = 16 tap boxes x 8, K = 64 pixels; B = dRaw [64 px][64 co] SW128 MN-major
    At = torch.randn(64, 128).bfloat16()  # [px][m]
    T = torch.randn(64, 64).bfloat16()    # [px][co]
    px = torch.arange(64)[:, None]
    m = torch.arange(128)[None, :]
    off = ((m // 8) * 2048 + px * 16 + (m % 8) * 2) // 2
    ia = torch.zeros(16 * 2048 // 2, dtype=torch.int16)
    ia[off.flatten()] = At.contiguous().view(torch.int16).flatten()
    image = torch.cat([ia, sw128(T)])
    boff = ia.numel() * 2
    ref = At.float().T @ T.float()
    which = os.environ.get("PROBE_VARIANT", "0")
    variants = {"0": ("LBO=M-dir,SBO=K-dir", (2048, 128)), "1": ("LBO=K-dir,SBO=M-dir", (128, 2048))}
    for name, (la, sa) in [variants[which]]:
        got = run_umma(image, desc(0, la, sa, 0, 0), desc(boff, 8192, 1024), idesc(128, 64, 1, 1), 4, 256, 2048, 64)
        report(f"umma A MN-major NO-swizzle {name}, B MN-major SW128", got, ref)


if __name__ == "__main__":
    for t in ((t_kmajor, t_mnmajor_a) if os.environ.get("PROBE_VARIANT", "0") == "0" else (t_mnmajor_a,)):
        try:
            t()
        except Exception as e:  # noqa
            print(f"[ERR] {t.__name__}: {e}", flush=True)

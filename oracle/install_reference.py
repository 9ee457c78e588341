"""Places an UNMODIFIED copy of the reference's Python package under baseline/_ref/ (git-ignored, not gpurun-ignored:
it travels to the GPU box like the built .so, and never enters the history).

The reference (CRIMAC-classifiers-unet) has no setup.py / pyproject, so `pip install --target baseline/_ref` has
nothing to build; this script is the equivalent "install": a byte-for-byte copy of /root/reference/crimac_unet/**/*.py.
bench.py's `--impl reference` arm and its in-line `cpu_baseline` import `crimac_unet/models/unet.py` from there and
time the reference's own nn.Module on the host cores (kind: "reference"); when the copy is absent they fall back to
the oracle port (kind: "port").  Run by __graft_entry__.build() whenever /root/reference is present.
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/crimac_unet"
DST = os.path.join(ROOT, "baseline", "_ref", "crimac_unet")


def install(src=SRC, dst=DST):
    if not os.path.isdir(src):
        return None
    n = 0
    for base, dirs, files in os.walk(src):
        dirs[:] = [d for d in dirs if not d.startswith(".") and d != "__pycache__"]
        for f in files:
            if not f.endswith((".py", ".yaml", ".txt")):
                continue
            rel = os.path.relpath(os.path.join(base, f), src)
            out = os.path.join(dst, rel)
            os.makedirs(os.path.dirname(out), exist_ok=True)
            shutil.copyfile(os.path.join(base, f), out)
            n += 1
    with open(os.path.join(os.path.dirname(dst), "INSTALLED_FROM"), "w") as fh:
        fh.write(f"{src}: {n} files copied unmodified by oracle/install_reference.py\n")
    return dst


if __name__ == "__main__":
    print(install() or "reference not present: nothing installed", file=sys.stderr)

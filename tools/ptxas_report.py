#!/usr/bin/env python
"""nvcc -Xptxas -v for one csrc/*.cu: registers / spills / smem per kernel (CPU-only check before spending GPU time)."""
import os, re, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(HERE, "..", "crimac-classifiers-unet_b200")
src = sys.argv[1]
cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC",
       "--expt-relaxed-constexpr", "-Xptxas", "-v", "-I", os.path.join(PKG, "..", "include"), "-c",
       os.path.join(PKG, "csrc", src), "-o", "/dev/null"]
r = subprocess.run(cmd, capture_output=True, text=True)
if r.returncode:
    sys.stderr.write(r.stdout + r.stderr); sys.exit(1)
name = None
for ln in (r.stdout + r.stderr).splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", ln)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name).split("(")[0]
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", ln)
    if m: stack = m.groups()
    m = re.search(r"Used (\d+) registers", ln)
    if m and name:
        print(f"{name:60s} regs {m.group(1):>4s} stack {stack[0]:>4s} spill st/ld {stack[1]}/{stack[2]}")
        name = None

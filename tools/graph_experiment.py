"""Experiment: does capturing one train step (library launches + side stream + SGD) in a CUDA graph shorten the step?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from __graft_entry__ import load_package
from oracle import unet_oracle as O
load_package()
import crimac_unet_b200.models.unet as M
from crimac_unet_b200.trainer import Trainer
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = M.UNet_Baseline(3, 4).to(dev).train()
tr = Trainer(m)
x = O.synthetic_echogram(32, 4, 256, 256, seed=0, device=dev)
y = O.synthetic_labels(32, 256, 256, seed=1, device=dev)
def timed(fn, k=20):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
with torch.no_grad():
    for _ in range(3): tr.step(x, y)
    print("eager ms/step", timed(lambda: tr.step(x, y)))
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2): tr.step(x, y)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g):
            loss = tr.step(x, y)
        print("graph ms/step", timed(g.replay), "loss", loss.item())
        print("eager again", timed(lambda: tr.step(x, y)))
    except Exception as e:
        print("capture failed:", repr(e)[:500])

"""Generates tests/golden/train_curve.npz by TRAINING THE UNMODIFIED REFERENCE MODULE (from /root/reference, build
container only) exactly as pipeline_train_predict/pipeline.py:144-190 does: optim.SGD(lr 0.005, momentum 0.95),
ExponentialLR(gamma 0.5) stepped every lr_step batches, nn.CrossEntropyLoss(weight=[10,300,250]), model.train(),
zero_grad / forward / loss / backward / step.  Stored: the loss of every step and the validation loss + sandeel-class
probabilities of pipeline.py:249-270 (set_label_ignore_val, criterion, softmax[:, SANDEEL]) at the end.

The workload comes from crimac-classifiers-unet_b200/synthetic.py (structured_batch), seeded, so the GPU test can rebuild
the identical batches.  Usage: python oracle/make_golden_curve.py   (~3 minutes on 8 cores)
"""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "train_curve.npz")
STEPS, LR_STEP, BATCH, SIZE = 24, 8, 4, 64


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = _load("crimac_reference_unet", "/root/reference/crimac_unet/models/unet.py")
    S = _load("crimac_synthetic", os.path.join(ROOT, "crimac-classifiers-unet_b200", "synthetic.py"))
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    model = ref.UNet_Baseline(n_classes=3, in_channels=4)
    state0 = {k: v.detach().clone().numpy() for k, v in model.state_dict().items()}
    opt = torch.optim.SGD(model.parameters(), lr=0.005, momentum=0.95)
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.5)
    crit = torch.nn.CrossEntropyLoss(weight=torch.tensor([10.0, 300.0, 250.0]))
    batches = [S.structured_batch(BATCH, SIZE, SIZE, seed=40 + i) for i in range(4)]
    losses, lrs = [], []
    for i in range(STEPS):
        x, y = batches[i % 4]
        model.train()
        opt.zero_grad()
        loss = crit(model(x), y)
        loss.backward()
        opt.step()
        losses.append(float(loss))
        lrs.append(opt.param_groups[0]["lr"])
        if (i + 1) % LR_STEP == 0:
            sched.step()
        print(i, losses[-1], lrs[-1], flush=True)
    # validation step of pipeline.py:249-270 on a batch with every label code
    xv, yv = S.structured_batch(BATCH, SIZE, SIZE, seed=77)
    g = torch.Generator().manual_seed(5)
    codes = torch.tensor([-100, -70, -50, -30, -10])
    sel = torch.rand(yv.shape, generator=g) < 0.2
    yv = torch.where(sel, codes[torch.randint(0, 5, yv.shape, generator=g)], yv).to(torch.int16)
    model.eval()
    with torch.no_grad():
        out = model(xv)
        lab = yv.long().clone()
        for v in (-70, -30, -100, -10):
            lab[lab == v] = -100
        lab[lab == -50] = 0
        vloss = float(crit(out, lab))
        prob = torch.nn.functional.softmax(out, dim=1)[:, 1].numpy()
    np.savez_compressed(OUT, losses=np.array(losses), lrs=np.array(lrs), val_loss=vloss, val_labels=yv.numpy(),
                        val_sandeel_prob=prob.astype(np.float32), steps=STEPS, lr_step=LR_STEP, batch=BATCH, size=SIZE,
                        init_checksum=np.array([float(np.abs(v).sum()) for v in state0.values()]))
    print("wrote", OUT, "val_loss", vloss)


if __name__ == "__main__":
    main()

"""Generates tests/golden/unet_add_d3.npz and unet_upsample_d3.npz by RUNNING THE UNMODIFIED REFERENCE classes
(/root/reference/crimac_unet/models/unet.py, build container only) with the two non-default decoder variants of
UNet.__init__ (unet.py:200-251): merge_mode="add" and up_mode="upsample".  Stored per variant: checksums of the
(reproducible) state_dict, input, labels, eval logits, train logits, loss, a subset of gradients, updated BatchNorm buffers.  Cross-checks oracle/
unet_oracle.py (up_mode / merge_mode arguments) against the same outputs.  Usage: python oracle/make_golden_variants.py"""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    from oracle import unet_oracle as O
    ref = _load("crimac_reference_unet", "/root/reference/crimac_unet/models/unet.py")
    S = _load("crimac_synthetic", os.path.join(ROOT, "crimac-classifiers-unet_b200", "synthetic.py"))
    for name, kw in (("unet_add_d3", dict(up_mode="transpose", merge_mode="add")),
                     ("unet_upsample_d3", dict(up_mode="upsample", merge_mode="concat"))):
        torch.manual_seed(7)
        m = ref.UNet_Baseline(n_classes=3, in_channels=4, depth=3, **kw)
        sd0 = O.trained_like_state({k: v.detach().clone() for k, v in m.state_dict().items()}, seed=1, head_gain=2.0)
        m.load_state_dict(sd0)
        x, y = S.structured_batch(2, 48, 80, seed=21)
        m.eval()
        with torch.no_grad():
            ev = m(x)
        m.train()
        m.zero_grad()
        tl = m(x)
        loss = torch.nn.CrossEntropyLoss(weight=torch.tensor(O.CLASS_WEIGHTS))(tl, y)
        loss.backward()
        out = {"x": x.numpy(), "y": y.numpy(), "eval_logits": ev.numpy(), "train_logits": tl.detach().numpy(),
               "loss": float(loss)}
        # the weights are reproducible (default init under torch.manual_seed(7), then oracle.trained_like_state(seed 1)):
        # only checksums are stored, and the gradients of a subset of tensors (every kind of layer, shallow and deep)
        out["state_keys"] = np.array(list(sd0.keys()))
        out["state_checksum"] = np.array([float(v.double().abs().sum()) for v in sd0.values()])
        keep = ("down_convs.0.main.0.weight", "down_convs.1.main.3.weight", "down_convs.2.main.4.weight", "down_convs.2.main.4.bias",
                "up_convs.0.conv1.weight", "up_convs.0.bn1.weight", "up_convs.1.conv2.weight", "up_convs.1.bn2.bias",
                "conv_final.weight", "conv_final.bias")
        for k, p in m.named_parameters():
            if k in keep or ".upconv." in k:
                out["grad/" + k] = p.grad.numpy()
        for k, v in m.state_dict().items():
            if "running_" in k or "num_batches" in k:
                out["stat/" + k] = v.numpy()
        # ---- the oracle restatement against the reference's own numbers
        o_ev = O.unet_forward(sd0, x, **kw)
        o_tl, o_loss, o_g, _ = O.train_step(sd0, x, y, **kw)
        print(name, "oracle vs reference: eval", float((o_ev - ev).abs().max()), "train", float((o_tl - tl.detach()).abs().max()),
              "loss", abs(float(o_loss) - float(loss)),
              "worst grad rel", max(float((o_g[k] - p.grad).norm() / (p.grad.norm() + 1e-30)) for k, p in m.named_parameters()
                                    if not (k.endswith(".bias") and any(t in k for t in ("main.0", "main.3", "conv1", "conv2")))))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


if __name__ == "__main__":
    main()

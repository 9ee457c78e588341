#!/usr/bin/env python
"""Headline benchmark: U-Net 256^2 patches/s, training fwd + class-weighted CE + bwd (+ NCCL gradient all-reduce and
SGD step), batch 32 of 4x256x256 synthetic echogram patches per B200 (BASELINE.json configs[1] / configs[2]).

  python bench.py --gpus 1 --steps K --warmup W                # this repo's sm_100a path
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...                         # the reference's PyTorch CPU path (oracle port), host cores

Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement for the definition of every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GFLOP_TRAIN = 288.979156992   # per 4x256x256 patch, fwd+bwd (SURVEY.md §8d / BASELINE.md)
GFLOP_INFER = 96.42704896
METRIC = "U-Net 256x256 patches/s, train fwd+bwd (class-weighted CE), batch 32 per B200"
WORKLOAD = "configs[1]: UNet(4 freq -> 3 classes) training fwd+bwd, class-weighted CE, batch 32 of 4x256x256 synthetic patches per GPU"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="train", choices=["train", "infer", "survey", "feed"],
                    help="train = BASELINE configs[1]/[2] (headline); infer = forward+softmax; survey = configs[3] sliding window")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--in-ch", type=int, default=4, help="frequencies (configs[4] stress: 6)")
    ap.add_argument("--size", type=int, default=256, help="patch height = width (configs[4] stress: 512)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="headline numbers only: no sustained / infer / survey / parity sub-records")
    ap.add_argument("--profile-out", default=None, help="write the per-kernel CUDA-event breakdown of one step here")
    return ap.parse_args()


def unet_gflop(in_ch, size, train):
    """Algorithmic GFLOP per patch (2*MACs of convs, convT, head) - SURVEY.md App. A generalised to (in_ch, size)."""
    px = size * size
    fwd, first = 0.0, 2.0 * px * 64 * 9 * in_ch
    fwd += first
    ch = [64, 128, 256, 512, 1024]
    for l in range(5):
        p = px / 4 ** l
        if l > 0:
            fwd += 2.0 * p * ch[l] * 9 * ch[l - 1]
        fwd += 2.0 * p * ch[l] * 9 * ch[l]
    for l in range(3, -1, -1):
        p = px / 4 ** l
        fwd += 2.0 * (p / 4) * (4 * ch[l]) * ch[l + 1]      # ConvTranspose2d
        fwd += 2.0 * p * ch[l] * 9 * (2 * ch[l])            # conv1 on the concat
        fwd += 2.0 * p * ch[l] * 9 * ch[l]
    fwd += 2.0 * px * 3 * 64
    return (3 * fwd - first) / 1e9 if train else fwd / 1e9


def committed_traffic(mode):
    """DRAM bytes of the dominant kernel family for one step, from the committed ncu --set full capture (or None)."""
    p = os.path.join(ROOT, "profiles", "conv_igemm_traffic.json")
    if not os.path.exists(p):
        return None
    try:
        return json.load(open(p)).get(mode)
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tflops_burst": d["bf16_tflops"], "tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "hbm_gbs": d["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"tflops_burst": 1590.0, "tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------------------ CPU reference arm
REF_DIR = os.path.join(ROOT, "baseline", "_ref", "crimac_unet")


def reference_module():
    """The UNMODIFIED reference crimac_unet/models/unet.py, imported from the copy oracle/install_reference.py placed
    under baseline/_ref/ (git-ignored; travels to the GPU box).  None when the copy is absent."""
    path = os.path.join(REF_DIR, "models", "unet.py")
    if not os.path.exists(path):
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location("crimac_reference_unet", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _cpu_workload(batch, structured=False):
    load = __import__("__graft_entry__").load_package
    load()
    import crimac_unet_b200.synthetic as S
    if structured:
        return S.structured_batch(batch, 256, 256, seed=7)
    return S.synthetic_echogram(batch, 4, 256, 256, seed=0), S.synthetic_labels(batch, 256, 256, seed=1)


def cpu_train_patches_per_s(batch, reps, warm=1):
    """The reference train step on the host cores (fp32 torch CPU, all threads): model.train(); zero_grad; forward;
    nn.CrossEntropyLoss(weight); backward; optim.SGD(momentum).step() - pipeline.py:156,167-178.  kind "reference" =
    the reference's own nn.Module (baseline/_ref), kind "port" = the oracle restatement (same ATen operators)."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    x, y = _cpu_workload(batch)
    ref = reference_module()
    if ref is not None:
        model = ref.UNet_Baseline(n_classes=3, in_channels=4)
        crit = torch.nn.CrossEntropyLoss(weight=torch.tensor([10.0, 300.0, 250.0]))
        opt = torch.optim.SGD(model.parameters(), lr=0.005, momentum=0.95)

        def step():
            model.train()
            opt.zero_grad()
            loss = crit(model(x), y)
            loss.backward()
            opt.step()
            return loss.item()
        kind = "reference"
    else:
        from oracle import unet_oracle as O
        sys.path.insert(0, os.path.join(ROOT, "crimac-classifiers-unet_b200"))
        import models.unet as M
        state = {k: v.clone() for k, v in M.UNet_Baseline(3, 4).state_dict().items()}
        step = lambda: O.train_step(state, x, y)
        kind = "port"
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(reps):
        step()
    dt = (time.perf_counter() - t0) / reps
    return batch / dt, cores, dt, kind


def cpu_infer_patches_per_s(batch, reps, warm=1):
    """The reference inference call (model.eval(); no_grad; forward; F.softmax - pipeline.py:205-218) on the host cores -
    BASELINE.json configs[0], the reference's own CPU-runnable case."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    x, _ = _cpu_workload(batch)
    ref = reference_module()
    if ref is not None:
        model = ref.UNet_Baseline(n_classes=3, in_channels=4).eval()
        run = lambda: torch.nn.functional.softmax(model(x), dim=1)
        kind = "reference"
    else:
        from oracle import unet_oracle as O
        sys.path.insert(0, os.path.join(ROOT, "crimac-classifiers-unet_b200"))
        import models.unet as M
        state = {k: v.clone() for k, v in M.UNet_Baseline(3, 4).state_dict().items()}
        run = lambda: O.softmax_probs(O.unet_forward(state, x))
        kind = "port"
    with torch.no_grad():
        for _ in range(warm):
            run()
        t0 = time.perf_counter()
        for _ in range(reps):
            run()
    dt = (time.perf_counter() - t0) / reps
    return batch / dt, cores, dt, kind


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the path on this box's host cores, same metric / unit /
    workload as the B200 arm.  Batch = the workload's own (32, configs[1]) when K + W steps of it fit a few minutes
    (probed with one step), else the largest power-of-two fraction that does; the choice is stated in `sample`."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    infer = args.mode == "infer"
    fn = cpu_infer_patches_per_s if infer else cpu_train_patches_per_s
    warm = max(1, min(args.warmup, 2))
    budget_s = 420.0
    batch = args.batch
    _, _, probe, _ = fn(min(batch, 4), 1, warm=1)                       # seconds per step at batch <= 4
    per_patch = probe / min(batch, 4)
    while batch > 2 and per_patch * batch * (args.steps + warm) > budget_s:
        batch //= 2
    pps, cores, dt, kind = fn(batch, args.steps, warm=warm)
    what = "eval forward + softmax (pipeline.py:205-218)" if infer else "train step: forward, class-weighted CE, backward, SGD-momentum update (pipeline.py:156,167-178)"
    impl = "the reference's own nn.Module (unmodified copy under baseline/_ref)" if kind == "reference" else "oracle port of the reference (baseline/_ref absent)"
    sample = f"{batch} of the {args.batch} patches of a batch per step; {impl}; {what}; fp32 torch CPU, {cores} threads"
    line = {
        "impl": "reference",
        "metric": "U-Net 256x256 patches/s, inference (softmax probabilities)" if infer else METRIC,
        "value": pps, "unit": "patches/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": warm,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": ("UNet inference, batch %d of 4x256x256" % args.batch) if infer else WORKLOAD,
                   "sample": sample, "same_batch_as_b200_arm": batch == args.batch},
        "cpu_baseline": {"value": pps, "unit": "patches/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": pps, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, pw = [], None, set(), []
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            try:
                pw.append(float(f[2]))
            except ValueError:
                pass
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        pw.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                "power_w": pw[len(pw) // 2] if pw else None}


# ------------------------------------------------------------------------------------------------ sub-records
def infer_subrecord(model, dev, rank, world, timed, S):
    """BASELINE configs[0] shape on the B200 path: eval forward + fused softmax, batch 16 and 32 of 4x256x256, inputs
    resident in HBM, per GPU (every rank runs its own batch: weak scaling, value = all ranks)."""
    model.eval()
    out = {"metric": "U-Net 256x256 patches/s, inference (softmax probabilities)", "unit": "patches/s"}
    for b in (16, 32):
        xb = S.synthetic_echogram(b, 4, 256, 256, seed=300 + rank, device=dev)
        for _ in range(3):
            model.predict_proba(xb)
        ms = timed(lambda: model.predict_proba(xb), 20) / 20
        out[f"batch{b}"] = {"value": world * b / (ms * 1e-3), "ms_per_step": ms,
                            "tflops_per_gpu": b / (ms * 1e-3) * GFLOP_INFER / 1e3}
    return out


def survey_subrecord(model, dev, rank, world, E, n_pings=1_000_000, n_range=256, preload=20000):
    """BASELINE configs[3]: ONE synthetic 4-frequency survey of 1 M pings x 256 range bins, preload_n_pings = 20000 ->
    50 chunks of 186 patches, sharded over the ranks by contiguous ping range through the product's own
    SurveyPredictor.predict_survey / shard_chunks (save_predict.py:160-171 semantics; 7,7,6,6,6,6,6,6 chunks on 8 GPUs).
    Every rank generates the pings of its own chunks (+ context) from the shared (seed, ping) -> sv rule.  Timed per
    rank with CUDA events around its whole shard (chunk generation excluded: it stands for the zarr read), max over ranks."""
    import torch
    import torch.distributed as dist
    import crimac_unet_b200.synthetic as S
    from crimac_unet_b200.predict import SurveyPredictor, split_pings, shard_chunks
    model.eval()
    sp = SurveyPredictor(model, patch_hw=(256, 256), overlap=20, preload_n_pings=preload, batch_size=93)
    mine = shard_chunks(split_pings(0, n_pings, preload), world, rank)
    ev = []

    def load_chunk(d0, d1, s, e):
        sv = S.synthetic_survey_pings(4, n_range, d0, d1, seed=5, device=dev)
        seabed = S.synthetic_seabed(s, e, device=dev)
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record()
        ev.append([e0, None])
        return sv, None, seabed

    def run(limit=None):
        n_patch, written, n_px = 0, 0.0, 0
        for i, (s, e, out) in enumerate(sp.predict_survey(load_chunk, n_pings, n_range, rank=rank, world=world,
                                                          seabed_max_of=lambda s, e: 220)):
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            ev[-1][1] = e1
            n_patch += sp.last_chunk_patches
            if i == 0:
                written, n_px = float((out != 0).float().mean().item()), out.numel()
            if limit is not None and i + 1 >= limit:
                break
        return n_patch, written

    run(limit=1)                                    # warm-up: contexts, tensor maps, allocator
    ev.clear()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = E.launch_count()
    n_patch, written = run()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([ms, float(n_patch)], device=dev, dtype=torch.float64)
    tmax = t.clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        total_patches = t[1].item()
    else:
        total_patches = float(n_patch)
    return {"metric": "U-Net 256x256 patches/s, sliding-window survey inference (preprocess + forward + stitch)",
            "value": total_patches / (tmax[0].item() * 1e-3), "unit": "patches/s",
            "pings_per_s": n_pings / (tmax[0].item() * 1e-3), "patches": total_patches, "chunks_this_rank": len(mine),
            "ms_max_over_ranks": tmax[0].item(), "fraction_of_output_pixels_written": written,
            "gpu_launches_this_rank": E.launch_count() - l0,
            "workload": "configs[3]: ONE survey of %d pings x %d range bins, %d-ping chunks, sharded by ping range over %d GPU(s)" % (n_pings, n_range, preload, world)}


def reference_loop_subrecord(M, S, dev, B, C, size, steps):
    """The reference's training loop, statement for statement (pipeline.py:156-181: torch.optim.SGD, ExponentialLR,
    nn.CrossEntropyLoss(weight), model(inputs), loss.backward(), optimizer.step(), loss.item()), on the native module:
    what a user gets who swaps the model class and changes nothing else.  Host batches, as a DataLoader hands them over."""
    import torch
    from torch import nn, optim
    torch.manual_seed(1)
    model = M.UNet_Baseline(3, C).to(dev)
    optimizer = optim.SGD(model.parameters(), lr=0.005, momentum=0.95)
    scheduler = optim.lr_scheduler.ExponentialLR(optimizer, gamma=0.5)
    criterion = nn.CrossEntropyLoss(weight=torch.tensor([10.0, 300.0, 250.0], device=dev), ignore_index=-100)
    batch = {"data": S.synthetic_echogram(B, C, size, size, seed=300).pin_memory(),
             "labels": S.synthetic_labels(B, size, size, seed=301).pin_memory()}

    def run(k):
        for i in range(k):
            inputs_train = batch["data"].float().to(dev)
            labels_train = batch["labels"].long().to(dev)
            model.train()
            optimizer.zero_grad()
            outputs_train = model(inputs_train)
            loss_train = criterion(outputs_train, labels_train)
            loss_train.backward()
            optimizer.step()
            loss_train.item()
            if (i + 1) % 1000 == 0:
                scheduler.step()

    run(3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(steps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"value": B / (ms * 1e-3), "unit": "patches/s", "ms_per_step": ms, "steps": steps,
            "api": "pipeline.py:156-181 verbatim on crimac_unet_b200 UNet_Baseline: autograd Function over the C-ABI (eager "
                   "launches), torch CrossEntropyLoss and torch.optim.SGD outside the library; host batch -> device, loss.item() per step"}


def parity_subrecord(M, Trainer, S, dev, steps=60):
    """parity printed with the bench line.  A trained-like net (the native trainer runs `steps`
    optimisation steps on a structured workload; there is no checkpoint to download) is copied into the reference's
    own nn.Module (baseline/_ref; the oracle port if absent) on the CPU, and both sides run the SAME batch: eval
    probabilities (max |dp|, argmax agreement) and one train step's loss and 82 gradient tensors (worst cosine, worst
    relative L2; conv biases in front of a BatchNorm have a zero gradient and are left out)."""
    import torch
    torch.manual_seed(0)
    model = M.UNet_Baseline(3, 4).to(dev).train()
    tr = Trainer(model, lr=0.005, momentum=0.95, lr_step=0)
    batches = [S.structured_batch(8, 128, 128, seed=10 + i, device=dev) for i in range(8)]
    for i in range(steps):
        tr.step(*batches[i % 8])
    state = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    x, y = S.structured_batch(4, 128, 128, seed=777)
    ref = reference_module()
    cw = torch.tensor([10.0, 300.0, 250.0])
    if ref is not None:
        rm = ref.UNet_Baseline(n_classes=3, in_channels=4)
        rm.load_state_dict(state)
        rm.eval()
        with torch.no_grad():
            p_ref = torch.nn.functional.softmax(rm(x), dim=1)
        rm.train()
        rm.zero_grad()
        loss_ref = torch.nn.CrossEntropyLoss(weight=cw)(rm(x), y)
        loss_ref.backward()
        g_ref = {n: p.grad for n, p in rm.named_parameters()}
        kind = "reference"
    else:
        from oracle import unet_oracle as O
        with torch.no_grad():
            p_ref = O.softmax_probs(O.unet_forward(state, x))
        _, loss_ref, g_ref, _ = O.train_step(state, x, y)
        kind = "port"
    model.load_state_dict(state)
    model.eval()
    with torch.no_grad():
        p = model.predict_proba(x.to(dev)).cpu()
    model.train()
    model._set_native_opt(None)        # the Trainer above fused its SGD update into train_step_fused: gradients only here
    loss = model.train_step_fused(x.to(dev), y.to(dev), cw.to(dev)).item()
    worst_cos, rels, head_rel = 1.0, [], 0.0
    for n, prm in model.named_parameters():
        if n.endswith(".bias") and any(t in n for t in ("main.0", "main.3", "conv1", "conv2", "upconv")):
            continue        # (near-)zero true gradient: conv biases in front of a BatchNorm
        a, b = prm.grad.detach().cpu().double().flatten(), g_ref[n].double().flatten()
        worst_cos = min(worst_cos, float(a @ b / (a.norm() * b.norm() + 1e-30)))
        r = float((a - b).norm() / (b.norm() + 1e-30))
        rels.append(r)
        if n.startswith("conv_final"):
            head_rel = max(head_rel, r)
    rels.sort()
    return {"against": kind, "net": f"{steps} native SGD steps on the structured workload (trained-like)",
            "batch": "4 x 4x128x128", "max_abs_dp": float((p - p_ref).abs().max()),
            "argmax_agreement": float((p.argmax(1) == p_ref.argmax(1)).float().mean()),
            "loss": loss, "loss_ref": float(loss_ref), "loss_rel_err": abs(loss - float(loss_ref)) / abs(float(loss_ref)),
            "grad_head_rel_l2": head_rel, "grad_median_rel_l2": rels[len(rels) // 2], "grad_worst_rel_l2": rels[-1],
            "grad_worst_cosine": worst_cos,
            "grad_note": "gradients of the bf16 path against fp32 autograd of the FP32 forward: the distance of the deep tensors is the "
                         "sensitivity of this network's gradient to bf16 rounding of the forward activations (CPU emulation of bf16 "
                         "storage alone: 55 % / cosine 0.84), not of the backward kernels - at the same forward state every tensor is "
                         "within 0.9 % (cosine >= 0.9999) of fp32 autograd, and 24 optimisation steps track the reference's loss curve "
                         "within 0.1 % (tests/test_gpu_unet.py, DESIGN.md section 4)"}


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the U-Net hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    load_package()
    import crimac_unet_b200.engine as E
    import crimac_unet_b200.synthetic as O   # workload generators; nothing under oracle/ is imported by this arm
    import crimac_unet_b200.models.unet as M
    from crimac_unet_b200.trainer import Trainer

    B = args.batch
    C, S = args.in_ch, args.size
    torch.manual_seed(0)
    model = M.UNet_Baseline(3, C).to(dev)
    if args.mode == "survey":
        return run_survey(args, model, dev, rank, world, E)
    if args.mode == "feed":
        return run_feed(args, model, dev, rank, world, E)
    x = O.synthetic_echogram(B, C, S, S, seed=100 + rank, device=dev)
    y = O.synthetic_labels(B, S, S, seed=200 + rank, device=dev)
    standard = (C == 4 and S == 256)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.mode == "train":
        model.train()
        trainer = Trainer(model, lr=0.005, momentum=0.95)
        trainer.broadcast_parameters(0)
        step = lambda xx, yy: trainer.step(xx, yy)
        gflop = GFLOP_TRAIN if standard else unet_gflop(C, S, True)
    else:
        model.eval()
        step = lambda xx, yy: model.predict_proba(xx)
        gflop = GFLOP_INFER if standard else unet_gflop(C, S, False)

    def timed(fn, k):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        sync_all()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            step(x, y)
        # ---- device-resident timing (value)
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        ms_total = timed(lambda: step(x, y), args.steps)
        clocks = sampler.stop() if rank == 0 else None
        ms_step = ms_total / args.steps
        value = world * B / (ms_step * 1e-3)

        # ---- end to end through the module API with host buffers
        xh = x.cpu().pin_memory()
        yh = y.cpu().pin_memory()
        xd, yd = torch.empty_like(x), torch.empty_like(y)

        # the repo's host-facing API: double-buffered host->device copies on a copy stream overlap the compute of the
        # previous batch (Trainer.fit_host / predict.predict_host_batches); every step still moves its own inputs from
        # pinned host memory and reads its result back (loss.item() per step, as pipeline.py:181)
        from crimac_unet_b200.predict import predict_host_batches

        def e2e_run(k):
            if args.mode == "train":
                for loss in trainer.fit_host((xh, yh) for _ in range(k)):
                    loss.item()
            else:
                for out_h in predict_host_batches(model, (xh for _ in range(k))):
                    pass

        e2e_run(3)
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        e2e_run(args.steps)
        e1.record()
        sync_all()
        te = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        ms_e2e = te.item() / args.steps
        h2d = xh.numel() * 4 + (yh.numel() * 8 if args.mode == "train" else 0)
        d2h = 4 if args.mode == "train" else B * 2 * S * S * 2
        e2e_value = world * B / (ms_e2e * 1e-3)

        # ---- sustained figure: the same device-resident step repeated for >= 3 s (clocks settle under the power cap)
        sustained = None
        if args.mode == "train" and not args.quick:
            k_sus = max(args.steps, int(3000.0 / ms_step) + 1)
            ms_sus = timed(lambda: step(x, y), k_sus) / k_sus
            sustained = {"value": world * B / (ms_sus * 1e-3), "unit": "patches/s", "steps": k_sus,
                         "seconds": ms_sus * k_sus * 1e-3, "ms_per_step": ms_sus}

        # ---- per-kernel breakdown of one more step (CUDA events around every launch, on the launching stream)
        # (an EAGER step: the timed steps replay a CUDA graph of exactly these launches, which the library's launch
        # counter and per-launch events cannot see)
        torch.cuda.synchronize()
        E.profile_enable(True)
        l0 = E.launch_count()
        if args.mode == "train":
            trainer._launch_step(x, y)
        else:
            step(x, y)
        launches_per_step = E.launch_count() - l0
        recs = E.profile_read()
        E.profile_enable(False)
        launches = launches_per_step * args.steps

    fam = {}
    for name, ms, fl, by, ln in recs:
        key = "conv_igemm_kernel (tcgen05 implicit GEMM: 3x3 fwd/dgrad, convT fwd/dgrad)" if name.startswith(("conv3x3", "convT_fwd", "convT_dgrad")) else name
        a = fam.setdefault(key, [0.0, 0.0, 0.0, 0])
        a[0] += ms
        a[1] += fl
        a[2] += by
        a[3] += ln
    step_kernel_ms = sum(a[0] for a in fam.values())
    peaks = measured_peaks()
    dom_name, dom = max(((k, a) for k, a in fam.items() if a[1] > 0), key=lambda kv: kv[1][0])
    achieved = dom[1] / (dom[0] * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": dom_name, "achieved": achieved, "peak": peaks["tflops_burst"],
                "unit": "TFLOP/s", "frac": achieved / peaks["tflops_burst"],
                "frac_of_sustained_peak": achieved / peaks["tflops_sustained"],
                "traffic": committed_traffic(args.mode) if (standard and B == 32) else None,
                "traffic_note": "DRAM bytes read+written by this kernel family in ONE step (sum over its launches), ncu --set full capture summarised in profiles/conv_igemm_traffic.json",
                "peak_source": peaks["source"] + ", BURST bf16 figure: every launch of the family is event-timed on its own in one eager step",
                "launches_per_step": dom[3], "ms_per_step_in_kernel": dom[0], "share_of_step": dom[0] / step_kernel_ms,
                "whole_step_tflops": value / world * gflop / 1e3, "whole_step_frac_of_burst_peak": value / world * gflop / 1e3 / peaks["tflops_burst"]}
    if args.profile_out and rank == 0:
        with open(args.profile_out, "w") as f:
            f.write(f"# per-kernel CUDA-event breakdown of one {args.mode} step, batch {B}, 1 GPU; kernel ms sum {step_kernel_ms:.3f}\n")
            f.write("family,ms,share,TFLOP/s,GB/s,launches\n")
            for k, a in sorted(fam.items(), key=lambda kv: -kv[1][0]):
                f.write(f"\"{k}\",{a[0]:.4f},{a[0] / step_kernel_ms:.4f},{a[1] / max(a[0], 1e-9) / 1e9:.1f},{a[2] / max(a[0], 1e-9) / 1e6:.1f},{a[3]}\n")
            f.write("# per launch, in issue order: name,ms,GFLOP,TFLOP/s,GB/s\n")
            for name, ms, fl, by, ln in recs:
                f.write(f"{name},{ms:.4f},{fl / 1e9:.3f},{fl / max(ms, 1e-9) / 1e9:.1f},{by / max(ms, 1e-9) / 1e6:.1f}\n")

    # ---- sub-records (same process, same box): inference (configs[0] shape), sliding-window survey (configs[3]) sharded
    # over the ranks by ping range, data-parallel replica check, parity against the reference module
    extra = {}
    if args.mode == "train" and standard and not args.quick:
        with torch.no_grad():
            extra["infer"] = infer_subrecord(model, dev, rank, world, timed, O)
            extra["survey"] = survey_subrecord(model, dev, rank, world, E)
        model.train()
    if args.mode == "train" and world > 1:
        # every replica must hold bit-identical parameters after the timed steps (same all-reduced gradients, same update)
        chk = trainer.flat_params.double().sum().reshape(1)
        chk2 = (trainer.flat_params.double() ** 2).sum().reshape(1)
        both = torch.cat([chk, chk2])
        lo, hi = both.clone(), both.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        extra["dp_params_in_sync"] = bool(torch.equal(lo, hi))
        extra["dp_param_checksum"] = float(chk.item())

    cpu_baseline, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and standard:
        pps, cores, dt, kind = cpu_train_patches_per_s(8, reps=2, warm=1)
        cpu_baseline = {"value": pps, "unit": "patches/s", "cores": cores, "kind": kind,
                        "sample": ("the reference's own nn.Module (baseline/_ref)" if kind == "reference" else "oracle port of the reference")
                        + f": train step (forward, weighted CE, backward, SGD), fp32 torch CPU, batch 8 (of 32), 1 warm-up + 2 timed steps, {dt:.2f} s/step"}
        if args.mode == "train" and not args.quick:
            parity = parity_subrecord(M, Trainer, O, dev)
            extra["reference_loop"] = reference_loop_subrecord(M, O, dev, B, C, S, args.steps)

    if rank == 0:
        line = {
            "metric": (METRIC if standard and B == 32 else f"U-Net {S}x{S} patches/s, train fwd+bwd (class-weighted CE), batch {B} per B200") if args.mode == "train"
                      else f"U-Net {S}x{S} patches/s, inference (softmax probabilities)",
            "value": value, "unit": "patches/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16 (fp32 accumulate; first conv, BN statistics, head, loss in fp32)", "data": "synthetic",
            "config": {"workload": (WORKLOAD if standard and B == 32 else "UNet training fwd+bwd, batch %d of %dx%dx%d per GPU" % (B, C, S, S)) if args.mode == "train" else "UNet inference, batch %d of %dx%dx%d" % (B, C, S, S),
                       "global_batch": world * B, "parallelism": f"dp{world}" if world > 1 else "single GPU",
                       "gradient_exchange": (getattr(trainer, "exchange", None) if args.mode == "train" and world > 1 else None),
                       "exchange_multicast": (bool(getattr(getattr(trainer, "peer", None), "multicast", False)) if args.mode == "train" and world > 1 else None),
                       "launch": ("CUDA graph replay of one captured step" if getattr(trainer, "_graph", None) is not None else "eager stream launches") if args.mode == "train" else "eager stream launches",
                       "step": "weight re-pack + forward + weighted CE + backward" + ((" + peer-memory all-reduce of the 124 MB gradient arena (own NVLink P2P / multimem kernels, 3 buckets launched inside backward)" if getattr(trainer, "exchange", None) == "peer" else " + NCCL all-reduce of the 124 MB gradient arena") if world > 1 else "") + " + fused SGD-momentum update" if args.mode == "train" else "forward + softmax",
                       "l2": "per-step working set (activations + gradients) is ~6 GB >> 126 MB L2; no explicit flush"},
            "e2e": {"value": e2e_value, "unit": "patches/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e, "api": "Trainer.fit_host: per step pinned host -> device copy of x (fp32) and labels (int64) on a copy stream (double-buffered, overlapping the previous step), UNet_Baseline.train_step_fused, gradient all-reduce, SGD, loss.item()" if args.mode == "train" else "predict.predict_host_batches: per batch pinned host -> device copy of x, UNet_Baseline.predict_proba, fp16 class-1/2 probabilities back to pinned host memory; copies double-buffered on copy streams"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
        }
        if sustained is not None:
            line["sustained"] = sustained
        if parity is not None:
            line["parity"] = parity
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        _teardown(dist, locals().get("trainer"))


def _teardown(dist, trainer=None):
    """Leave the process group without ever hanging the job: drop captured graphs first, then destroy the group under a
    watchdog that force-exits (status 0: the JSON line is already out) if NCCL teardown blocks."""
    import gc
    import torch

    def _bail():
        sys.stdout.flush()
        os._exit(0)
    threading.Timer(30.0, _bail).start()
    if trainer is not None:
        trainer._graph = None
    gc.collect()
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()
    sys.stdout.flush()
    os._exit(0)


def run_survey(args, model, dev, rank, world, E):
    """BASELINE configs[3]: sliding-window whole-echogram inference (save_predict.py:137-220) on a synthetic 4-frequency
    survey, preload_n_pings = 20000, 256 range bins, sharded by ping range.  A step = one preload chunk (186 patches):
    patch gather + dB transform, U-Net forward + softmax, overlap-stitch into (2, range, pings) fp16 - all on device."""
    import math
    import torch
    import torch.distributed as dist
    from crimac_unet_b200.predict import SurveyPredictor
    R, NP = 256, 20000
    n_chunks = max(args.steps, 1)
    P = NP * n_chunks                      # this rank's ping range (weak scaling: every rank owns n_chunks chunks)
    model.eval()
    g = torch.Generator(device=dev).manual_seed(300 + rank)
    sv = torch.pow(10.0, torch.rand((args.in_ch, R, P), device=dev, generator=g) * 7.0 - 9.0)
    sv[torch.rand((args.in_ch, R, P), device=dev, generator=g) < 1e-3] = float("nan")
    pings = torch.arange(P, device=dev)
    seabed = (200 + 20 * torch.sin(2 * math.pi * pings / 5000)).to(torch.int32)
    sp = SurveyPredictor(model, patch_hw=(256, 256), overlap=20, preload_n_pings=NP, batch_size=args.batch)
    def chunk(i, sv_dev, ping0):
        s, e = i * NP, (i + 1) * NP
        grid, _ = sp.chunk_geometry(s, e, R, P, seabed_max=220)
        o = sp.predict_chunk(sv_dev, ping0, grid, s, e, seabed=seabed[s:e].contiguous())   # fresh zeroed (2,R,NP) fp16
        return grid.shape[0], o

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        npatch, last = chunk(0, sv, 0)
    sampler = ClockSampler(dev.index or 0)
    if rank == 0:
        sampler.start()
    sync_all()
    l0 = E.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n_chunks):
        chunk(i, sv, 0)
    e1.record()
    sync_all()
    launches = E.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    value = world * n_chunks * npatch / (ms * 1e-3)
    # end to end: the chunk's pings come from pinned host memory, the stitched fp16 chunk goes back to the host
    svh = sv[:, :, :NP + 512].cpu().pin_memory()
    outh = torch.empty((2, R, NP), dtype=torch.float16).pin_memory()
    svd = torch.empty_like(sv[:, :, :NP + 512])
    sync_all()
    e0.record()
    for i in range(n_chunks):
        svd.copy_(svh, non_blocking=True)
        _, o = chunk(0, svd, 0)
        outh.copy_(o, non_blocking=True)
    e1.record()
    sync_all()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e = world * n_chunks * npatch / (t.item() * 1e-3)
    written = float((last != 0).float().mean().item())
    if rank == 0:
        peaks = measured_peaks()
        print(json.dumps({
            "metric": "U-Net 256x256 patches/s, sliding-window survey inference (preprocess + forward + stitch)",
            "value": value, "unit": "patches/s", "n_gpus": world, "steps": n_chunks, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / n_chunks, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16 (fp32 accumulate; first conv, head, softmax in fp32; fp16 output)", "data": "synthetic",
            "config": {"workload": "configs[3]: sliding-window inference, %d-ping chunks x 256 range bins, %d patches per chunk, batch %d, sharded by ping range (%d chunks per GPU)" % (NP, npatch, args.batch, n_chunks),
                       "pings_per_s": value / npatch * NP, "fraction_of_output_pixels_written": written,
                       "l2": "each chunk reads 82 MB of sv and ~1.5 GB of activations >> 126 MB L2"},
            "e2e": {"value": e2e, "unit": "patches/s", "h2d_bytes_per_step": svh.numel() * 4, "d2h_bytes_per_step": outh.numel() * 2,
                    "api": "SurveyPredictor.predict_chunk; pinned host sv chunk -> device, stitched fp16 chunk -> host"},
            "gpu_launches": launches, "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "conv_igemm_kernel", "achieved": value / world * GFLOP_INFER / 1e3,
                         "peak": peaks["tflops_sustained"], "unit": "TFLOP/s", "frac": value / world * GFLOP_INFER / 1e3 / peaks["tflops_sustained"],
                         "traffic": None, "note": "whole-chunk figure (all kernels incl. preprocess and stitch) against the sustained bf16 peak"},
            "cpu_baseline": None}), flush=True)
    if world > 1:
        _teardown(dist)


def run_feed(args, model, dev, rank, world, E):
    """the headline train step fed by training samples drawn ON THE DEVICE from a resident
    survey (crop + add_noise + flip + refine_label_boundary + convert_label_indexing + dB transform, the work of the
    reference's CPU DataLoader workers, batch/dataset.py:75-108).  A step = sample generation (2 launches) + weight
    re-pack + forward + weighted CE + backward (+ all-reduce) + SGD."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from crimac_unet_b200.trainer import Trainer
    from crimac_unet_b200.train_patches import SurveyPatchFeeder
    B, C, S = args.batch, args.in_ch, args.size
    NP, R = 60000, 500
    g = torch.Generator(device=dev).manual_seed(400 + rank)
    sv = torch.pow(10.0, torch.rand((C, NP, R), device=dev, generator=g) * 7.0 - 9.0)        # zarr order (F, ping, range)
    labels = torch.zeros((NP, R), device=dev)
    rng = np.random.default_rng(500 + rank)
    for _ in range(400):                                                                      # schools: 27 sandeel, 1 other
        cy, cx = int(rng.integers(0, NP)), int(rng.integers(0, R))
        labels[max(0, cy - 60):cy + 60, max(0, cx - 40):cx + 40] = float(rng.choice([27, 1]))
    fish = labels > 0
    sv[C - 1][fish] = torch.pow(10.0, torch.rand(int(fish.sum()), device=dev, generator=g) * 3.0 - 7.2)
    model.train()
    trainer = Trainer(model, lr=0.005, momentum=0.95)
    trainer.broadcast_parameters(0)
    feeder = SurveyPatchFeeder(sv, labels, B, (S, S), seed=600 + rank)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in trainer.fit_survey(feeder, max(args.warmup, 3)):
        pass
    sampler = ClockSampler(dev.index or 0)
    if rank == 0:
        sampler.start()
    sync_all()
    l0 = E.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for loss in trainer.fit_survey(feeder, args.steps):
        pass
    e1.record()
    sync_all()
    clocks = sampler.stop() if rank == 0 else None
    # the timed steps replay a CUDA graph, which the library's launch counter cannot see: count one EAGER step
    torch.cuda.synchronize()
    l0 = E.launch_count()
    xb, yb = feeder.next_batch()
    trainer._launch_step(xb, yb)
    torch.cuda.synchronize()
    launches = (E.launch_count() - l0 + 2) * args.steps    # + the two train_patches kernels per step
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item() / args.steps
    value = world * B / (ms * 1e-3)
    # end to end as a user loop would run it: the loss of every step is read back on the host
    sync_all()
    e0.record()
    last = float("nan")
    for loss in trainer.fit_survey(feeder, args.steps):
        last = loss.item()
    e1.record()
    sync_all()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e = world * B * args.steps / (t.item() * 1e-3)
    standard = (C == 4 and S == 256)
    gflop = GFLOP_TRAIN if standard else unet_gflop(C, S, True)
    if rank == 0:
        peaks = measured_peaks()
        print(json.dumps({
            "metric": "U-Net 256x256 patches/s, train step fed from a device-resident survey (crop + augmentation + label refinement + dB on device)",
            "value": value, "unit": "patches/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16 (fp32 accumulate, fp32 master weights / BN statistics / loss)", "data": "synthetic",
            "config": {"workload": "UNet training fwd+bwd, batch %d of %dx%dx%d per GPU, samples drawn on the device from a %dx%dx%d survey" % (B, C, S, S, C, NP, R),
                       "l2": "activations of one step (several GB) >> 126 MB L2; the survey is %d MB" % (sv.numel() * 4 // 2 ** 20),
                       "last_loss": last},
            "e2e": {"value": e2e, "unit": "patches/s", "h2d_bytes_per_step": B * 12, "d2h_bytes_per_step": 4,
                    "api": "Trainer.fit_survey(SurveyPatchFeeder): per step 12 bytes per sample of crop centres and coin flips host -> device, loss.item()"},
            "gpu_launches": launches, "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "conv_igemm_kernel", "achieved": value / world * gflop / 1e3,
                         "peak": peaks["tflops_sustained"], "unit": "TFLOP/s", "frac": value / world * gflop / 1e3 / peaks["tflops_sustained"],
                         "traffic": None, "note": "whole-step figure (sample generation included) against the sustained bf16 peak"},
            "cpu_baseline": None}), flush=True)
    if world > 1:
        _teardown(dist, trainer)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)

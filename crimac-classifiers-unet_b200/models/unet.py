"""Drop-in replacement for the reference ``crimac_unet/models/unet.py`` (same class / function names, constructor
arguments, attributes, ``forward`` signatures and the 136 ``state_dict`` keys of SURVEY.md App. B), whose hot path —
``UNet_Baseline.forward`` (reference models/unet.py:327-343) and its autograd backward — runs in hand-written sm_100a
kernels behind the C-ABI of ``include/crimac_b200.h``.

The sub-modules (``nn.Conv2d`` / ``nn.BatchNorm2d`` / ``nn.ConvTranspose2d``) are kept purely as PARAMETER HOLDERS in
the same tree positions so checkpoints load unchanged; their own ``forward`` is never used on the hot path.  There is
no CPU / PyTorch fallback for ``UNet_Baseline``: a configuration the native path does not cover raises.
``UNet_LateMetInject`` (reference :346-391) runs its U-Net body and the 64-channel part of its head through the same
native calls; only its per-sample metadata MLP is torch.
"""
import importlib
import importlib.util
import os
import sys
import weakref

import torch
import torch.nn as nn


def _runtime():
    """The host-side driver package (``engine.py``), importable whichever way this file itself was imported."""
    name = "crimac_unet_b200"
    if name not in sys.modules:
        pkg_dir = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        spec = importlib.util.spec_from_file_location(
            name, os.path.join(pkg_dir, "__init__.py"), submodule_search_locations=[pkg_dir]
        )
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
    return importlib.import_module(name + ".engine")


# native contexts live outside the module object so that copy.deepcopy / pickling of a model never touches them
_ENGINES = weakref.WeakKeyDictionary()


# --------------------------------------------------------------------------------------------- layer factories
def conv3x3(in_channels, out_channels, stride=1, padding=1, bias=True, groups=1):
    """3x3 convolution holder (reference unet.py:35-44)."""
    return nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=stride, padding=padding, bias=bias, groups=groups)


def conv1x1(in_channels, out_channels, groups=1):
    """1x1 convolution holder (reference unet.py:59-60)."""
    return nn.Conv2d(in_channels, out_channels, kernel_size=1, groups=groups, stride=1)


def upconv2x2(in_channels, out_channels, mode="transpose"):
    """2x up-sampling holder (reference unet.py:47-56): transposed conv, or bilinear upsample + 1x1 conv."""
    if mode == "transpose":
        return nn.ConvTranspose2d(in_channels, out_channels, kernel_size=2, stride=2)
    return nn.Sequential(nn.Upsample(mode="bilinear", scale_factor=2), conv1x1(in_channels, out_channels))


class DownConv(nn.Module):
    """Encoder block: (conv3x3, BN, ReLU) x 2 and an optional 2x2 max-pool (reference unet.py:63-93).

    ``main`` keeps the reference's Sequential indexing (0 conv, 1 bn, 2 relu, 3 conv, 4 bn, 5 relu) because the
    state_dict keys ``main.0/1/3/4.*`` depend on it.  The block is a PARAMETER HOLDER: the native engine reads its
    tensors, and calling the block on its own raises (there is no torch / cuDNN path in this package).
    """

    def __init__(self, in_channels, out_channels, pooling=True):
        super().__init__()
        self.in_channels, self.out_channels, self.pooling = in_channels, out_channels, pooling
        layers = []
        for cin in (in_channels, out_channels):
            layers += [conv3x3(cin, out_channels), nn.BatchNorm2d(out_channels), nn.ReLU()]
        self.main = nn.Sequential(*layers)
        if pooling:
            self.pool = nn.MaxPool2d(kernel_size=2, stride=2)

    def forward(self, x):
        raise RuntimeError("crimac_unet_b200.DownConv is a parameter holder: the block only runs as part of "
                           "UNet_Baseline / UNet_LateMetInject through libcrimac_b200.so (no per-block torch path)")


class UpConv(nn.Module):
    """Decoder block: up-conv, merge with the skip tensor, (conv3x3, BN, ReLU) x 2 (reference unet.py:96-137).
    Registration order upconv, conv1, conv2, bn1, bn2 fixes ``parameters()`` order."""

    def __init__(self, in_channels, out_channels, merge_mode="concat", up_mode="transpose"):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.merge_mode, self.up_mode = merge_mode, up_mode
        self.upconv = upconv2x2(in_channels, out_channels, mode=up_mode)
        self.conv1 = conv3x3(2 * out_channels if merge_mode == "concat" else out_channels, out_channels)
        self.conv2 = conv3x3(out_channels, out_channels)
        self.bn1 = nn.BatchNorm2d(out_channels)
        self.bn2 = nn.BatchNorm2d(out_channels)

    def forward(self, from_down, from_up):
        raise RuntimeError("crimac_unet_b200.UpConv is a parameter holder: the block only runs as part of "
                           "UNet_Baseline / UNet_LateMetInject through libcrimac_b200.so (no per-block torch path)")


class MetaPostProcessing(nn.Module):
    """Per-pixel MLP on the metadata channels (reference unet.py:140-166); not on the benchmark path."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.in_channels = in_channels
        self.hidden_channels_1 = 32
        self.hidden_channels_2 = 32
        self.out_channels = out_channels
        self.main = nn.Sequential(
            nn.Linear(in_channels, self.hidden_channels_1),
            nn.ReLU(),
            nn.Linear(self.hidden_channels_1, self.hidden_channels_2),
            nn.ReLU(),
            nn.Linear(self.hidden_channels_2, out_channels),
        )

    def forward(self, x):
        return self.main(x.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)


# --------------------------------------------------------------------------------------------- autograd bridge
class _NativeUNetFunction(torch.autograd.Function):
    """Train-mode forward / backward of the whole network through the C-ABI (one Function for all layers).

    `params` are the 82 tensors the library differentiates, in state-table order; params[-2] is the (n_classes, 64, 1, 1)
    head weight - the module's own parameter for UNet_Baseline, a slice of the 65-input head for UNet_LateMetInject."""

    @staticmethod
    def forward(ctx, x, model, *params):
        eng = model._engine_for(x, train=True)
        state = eng.state_table(model._state_tensors(head_weight=params[-2]))
        eng.prepare(state, True)
        logits = torch.empty((x.shape[0], model.n_classes, x.shape[2], x.shape[3]), dtype=torch.float32, device=x.device)
        eng.forward_train(state, x, logits)
        model._native_mutated()          # BatchNorm running statistics were updated through raw pointers
        # The saved activations live in the engine's workspace, not in ctx: ONE forward may be outstanding per engine.
        # Stamp this forward so that backward can tell whether a later forward (or a parameter update) invalidated it.
        eng.fwd_stamp += 1
        ctx.stamp = (eng.fwd_stamp, model._native_gen, tuple(p._version for p in params))
        ctx.params = params
        ctx.model, ctx.eng = model, eng
        ctx.shapes = [tuple(p.shape) for p in params]
        ctx.save_for_backward(x, params[-2])
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        x, head_w = ctx.saved_tensors
        model, eng = ctx.model, ctx.eng
        now = (eng.fwd_stamp, model._native_gen, tuple(p._version for p in ctx.params))
        if now != ctx.stamp:
            raise RuntimeError(
                "crimac_unet_b200: backward() of a train-mode forward whose saved activations are gone - another "
                "train-mode forward ran on this model (same input shape), or its parameters changed, between this "
                "forward and its backward.  The native path keeps ONE set of saved activations per (model, H, W): run "
                "each forward's backward before the next forward (gradient accumulation: call backward per micro-batch).")
        sizes = [int(torch.Size(sh).numel()) for sh in ctx.shapes]
        arena = torch.empty(sum(sizes), dtype=torch.float32, device=x.device)
        grads, off = [], 0
        for sh, n in zip(ctx.shapes, sizes):
            grads.append(arena[off:off + n].view(sh))
            off += n
        state = eng.state_table(model._state_tensors(head_weight=head_w))
        eng.backward(state, x, dlogits.contiguous().float(), None, eng.grad_table(grads))
        return (None, None) + tuple(grads)


class UNet(nn.Module):
    """Base class: builds the encoder / decoder holders exactly where the reference does (unet.py:169-301)."""

    valid = False
    pad = 0
    fow = [192, 192]
    dim = 2
    type = "seg"
    stride = 1
    increase_fow = 16

    def __init__(self, n_classes=2, in_channels=1, meta_in_channels=0, late_meta_inject=False, depth=5, start_filts=64,
                 up_mode="transpose", merge_mode="concat"):
        super().__init__()
        if up_mode not in ("transpose", "upsample"):
            raise ValueError(
                '"{}" is not a valid mode for upsampling. Only "transpose" and "upsample" are allowed.'.format(up_mode)
            )
        if merge_mode not in ("concat", "add"):
            # the reference formats this message with up_mode (unet.py:236-241); kept for identical error text
            raise ValueError(
                '"{}" is not a valid mode formerging up and down paths. Only "concat" and "add" are allowed.'.format(up_mode)
            )
        if up_mode == "upsample" and merge_mode == "add":
            raise ValueError(
                'up_mode "upsample" is incompatible with merge_mode "add" at the moment because it doesn\'t make sense '
                "to use nearest neighbour to reduce depth channels (by half)."
            )
        self.up_mode, self.merge_mode = up_mode, merge_mode
        self.in_channels = in_channels
        self.start_filts = start_filts
        self.depth = depth
        self.n_classes = n_classes
        self.meta_in_channels = meta_in_channels

        widths = [start_filts * (2 ** i) for i in range(depth)]
        downs = [DownConv(in_channels if i == 0 else widths[i - 1], widths[i], pooling=i < depth - 1) for i in range(depth)]
        ups = [UpConv(widths[i], widths[i - 1], up_mode=up_mode, merge_mode=merge_mode) for i in range(depth - 1, 0, -1)]
        self.down_convs = nn.Sequential(*downs)
        self.up_convs = nn.Sequential(*ups)
        head_in = widths[0] if depth > 1 else widths[-1]
        if not late_meta_inject:
            self.conv_final = conv1x1(head_in, n_classes)
        else:
            self.conv_final = conv1x1(head_in + meta_in_channels, n_classes)
            self.post_processing_weights = MetaPostProcessing(in_channels=meta_in_channels, out_channels=1)

    @staticmethod
    def weight_init(m):
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight)
            nn.init.constant_(m.bias, 0)

    def reset_params(self):
        for i, m in enumerate(self.modules()):
            print(i, m)
            self.weight_init(m)


class _NativePlumbing:
    """What a model whose body is the 5-level U-Net needs to run it through libcrimac_b200.so (mixed into UNet_Baseline
    and UNet_LateMetInject; the sub-modules stay parameter holders)."""

    _native_head_in = 64   # input channels of the 1x1 head the library computes
    _native_gen = 0        # bumped whenever native code wrote parameters / BN buffers through raw pointers

    def _native_mutated(self):
        """Native kernels write parameters (crimac_sgd_step) and BatchNorm running statistics (bn_finalize) through
        raw pointers, which torch's tensor._version cannot see: every such call bumps this counter, and the eval
        path's packed bf16 weights / folded BatchNorm are rebuilt when it has moved."""
        self._native_gen += 1

    def _check_supported(self, x):
        problems = []
        if self.start_filts != 64 or not (2 <= self.depth <= 5):
            problems.append("start_filts must be 64 and depth in 2..5")
        if not (1 <= self.in_channels <= 12) or not (1 <= self.n_classes <= 8):
            problems.append("in_channels must be in 1..12 and n_classes in 1..8")
        if self.conv_final.in_channels != self._native_head_in + self._meta_head_channels():
            problems.append("the 1x1 head must take the 64 decoder channels (plus the late-injected metadata channels)")
        if self.conv_final.out_channels != self.n_classes:
            problems.append("the 1x1 head must produce n_classes channels")
        if x.dim() != 4 or x.shape[1] != self.in_channels:
            problems.append(f"expected input (N,{self.in_channels},H,W), got {tuple(x.shape)}")
        elif x.shape[2] % (1 << (self.depth - 1)) or x.shape[3] % (1 << (self.depth - 1)):
            problems.append("H and W must be multiples of 2^(depth-1)")
        if not x.is_cuda:
            problems.append("input must be a CUDA tensor: the U-Net hot path has no CPU fallback")
        if problems:
            raise RuntimeError(f"crimac_unet_b200.{type(self).__name__} cannot run this call natively: " + "; ".join(problems))

    def _set_native_comm(self, comm):
        """Data parallel (trainer.PeerGradientExchange): every train context of this model - existing and future - exchanges
        its gradients over peer memory inside backward.  comm = engine._CommConfig or None."""
        self._native_comm = comm
        for key, eng in _ENGINES.get(self, {}).items():
            if key[2]:
                eng.set_comm(comm)

    def _set_native_opt(self, opt):
        """trainer.Trainer: fuse the SGD(momentum) update into the native backward of every train context of this model
        (engine._OptConfig; None switches it off).  train_step_fused then ALSO updates the parameters."""
        self._native_opt = opt
        for key, eng in _ENGINES.get(self, {}).items():
            if key[2]:
                eng.set_optimizer(opt)

    def _meta_head_channels(self):
        return 0

    def _head_weight(self):
        """The (n_classes, 64, 1, 1) contiguous head weight of the state table."""
        return self.conv_final.weight

    def _body_state(self):
        def bn(m):
            return [m.weight, m.bias, m.running_mean, m.running_var, m.num_batches_tracked]

        out = []
        for d in self.down_convs:
            c1, b1, _, c2, b2, _ = d.main
            out += [c1.weight, c1.bias] + bn(b1) + [c2.weight, c2.bias] + bn(b2)
        for u in self.up_convs:
            up = u.upconv if self.up_mode == "transpose" else u.upconv[1]    # "upsample": Sequential(Upsample, conv1x1)
            out += [up.weight, up.bias, u.conv1.weight, u.conv1.bias, u.conv2.weight, u.conv2.bias]
            out += bn(u.bn1) + bn(u.bn2)
        return out

    def _state_tensors(self, head_weight=None):
        """The 136 tensors of UNet_Baseline.state_dict() in its order, gathered without building the
        dict.  head_weight overrides the head's weight tensor (the autograd path passes the tensor it differentiates)."""
        return self._body_state() + [self._head_weight() if head_weight is None else head_weight, self.conv_final.bias]

    def _param_tensors(self):
        return list(self.parameters())

    def _engine_for(self, x, train):
        nb, _, h, w = x.shape
        return self._engine_for_shape(nb, h, w, x.device, train)

    def _engine_for_shape(self, nb, h, w, device, train):
        rt = _runtime()
        # the reference's fix_seeds (utils/general.py:120-128) sets torch.backends.cudnn.deterministic: honour the same
        # switches - weight gradients are then reduced in a fixed order and a train step is bit-reproducible
        det = bool(train) and (torch.backends.cudnn.deterministic or torch.are_deterministic_algorithms_enabled())
        key = (h, w, bool(train), device, det)
        engines = _ENGINES.setdefault(self, {})
        eng = engines.get(key)
        if eng is None or eng.cfg.max_batch < nb:
            eng = rt.Context(self.in_channels, self.n_classes, self.depth, self.start_filts, nb, h, w, train, device,
                             deterministic=det, up_mode=self.up_mode, merge_mode=self.merge_mode)
            engines[key] = eng
            comm = getattr(self, "_native_comm", None)
            if train and comm is not None:
                eng.set_comm(comm)    # data parallel: gradients are exchanged over peer memory inside backward
            opt = getattr(self, "_native_opt", None)
            if train and opt is not None:
                eng.set_optimizer(opt)
        return eng

    def _versions(self):
        return tuple((t.data_ptr(), t._version) for t in self._state_tensors())

    def _prep_input(self, x):
        self._check_supported(x)
        for t in self._state_tensors():
            if t.device != x.device:
                raise RuntimeError("model parameters and input are on different devices")
            break
        return x.contiguous().float()

    def _infer(self, x, softmax):
        x = self._prep_input(x)
        eng = self._engine_for(x, train=False)
        state = eng.state_table(self._state_tensors())
        key = (self._versions(), self._native_gen)
        if eng.prepared_key != key:
            eng.prepare(state, False)
            eng.prepared_key = key
        out = torch.empty((x.shape[0], self.n_classes, x.shape[2], x.shape[3]), dtype=torch.float32, device=x.device)
        eng.forward_infer(state, x, out, softmax)
        return out

    @torch.no_grad()
    def predict_proba_patches(self, sv, data_ping0, centres, patch_hw, softmax=True):
        """Sliding-window inference without the patch tensor: gathers the patches around `centres` (int32 (n,2) device,
        (y, x) survey coordinates) from the preloaded pings `sv` (fp32 device (F,R,P), column 0 = survey ping
        data_ping0), applies remove_nan_inf + db_with_limits and feeds the first conv DIRECTLY (the kernel writes the
        tensor-core operand), then the eval forward with the softmax fused - reference batch/dataset.py:192-205 ->
        pipeline.py:205-218.  Returns (probabilities fp32 (n,n_classes,ph,pw), nan_mask uint8 (n,ph,pw))."""
        if self.training:
            raise RuntimeError("predict_proba_patches() is an eval-mode call; use model.eval() first")
        ph, pw = patch_hw
        n = centres.shape[0]
        probe = torch.empty((0, self.in_channels, ph, pw), device=sv.device)
        self._check_supported(probe)
        if sv.dim() != 3 or sv.shape[0] != self.in_channels or sv.dtype != torch.float32 or not sv.is_contiguous():
            raise RuntimeError(f"sv must be a contiguous fp32 (F={self.in_channels}, R, P) CUDA tensor")
        if centres.dtype != torch.int32 or not centres.is_contiguous() or centres.device != sv.device:
            raise RuntimeError("centres must be a contiguous int32 (n,2) tensor on the device of sv")
        eng = self._engine_for_shape(n, ph, pw, sv.device, train=False)
        state = eng.state_table(self._state_tensors())
        key = (self._versions(), self._native_gen)
        if eng.prepared_key != key:
            eng.prepare(state, False)
            eng.prepared_key = key
        nan_mask = torch.empty((n, ph, pw), dtype=torch.uint8, device=sv.device)
        out = torch.empty((n, self.n_classes, ph, pw), dtype=torch.float32, device=sv.device)
        eng.preprocess_staged(sv, data_ping0, centres, nan_mask)
        eng.forward_infer_staged(state, n, out, softmax)
        return out, nan_mask

    @torch.no_grad()
    def predict_stitch_patches(self, sv, data_ping0, centres, patch_hw, out, ping_start, overlap, labels=None,
                               seabed=None, seabed_pad=10, classes=(1, 2)):
        """The whole sliding-window step of one batch of patches in 23 launches: patch gather + dB transform written as the
        first conv's operand, the eval forward, and - in the last conv's epilogue - softmax plus the overlap stitching of
        save_predict.py:41-65 (fill_out_array) with the label masks (overlap frame, chunk bounds, non-finite input, below
        seabed + pad): classes `classes` of every kept pixel go straight into `out` (fp16 (K, R, Pc), written in place).
        Neither the fp32 patch tensor nor the probability tensor exists in HBM."""
        if self.training:
            raise RuntimeError("predict_stitch_patches() is an eval-mode call; use model.eval() first")
        ph, pw = patch_hw
        n = centres.shape[0]
        self._check_supported(torch.empty((0, self.in_channels, ph, pw), device=sv.device))
        if sv.dim() != 3 or sv.shape[0] != self.in_channels or sv.dtype != torch.float32 or not sv.is_contiguous():
            raise RuntimeError(f"sv must be a contiguous fp32 (F={self.in_channels}, R, P) CUDA tensor")
        if centres.dtype != torch.int32 or not centres.is_contiguous() or centres.device != sv.device:
            raise RuntimeError("centres must be a contiguous int32 (n,2) tensor on the device of sv")
        if out.dtype != torch.float16 or not out.is_contiguous() or out.dim() != 3 or out.shape[0] != len(classes):
            raise RuntimeError("out must be a contiguous fp16 (len(classes), R, Pc) tensor")
        eng = self._engine_for_shape(n, ph, pw, sv.device, train=False)
        state = eng.state_table(self._state_tensors())
        key = (self._versions(), self._native_gen)
        if eng.prepared_key != key:
            eng.prepare(state, False)
            eng.prepared_key = key
        nan_mask = torch.empty((n, ph, pw), dtype=torch.uint8, device=sv.device)
        eng.preprocess_staged(sv, data_ping0, centres, nan_mask)
        eng.forward_infer_stitch(state, n, centres, nan_mask, out, ping_start, overlap, labels=labels, seabed=seabed,
                                 seabed_pad=seabed_pad, classes=classes)
        return out

    def _train_forward(self, x, params):
        x = self._prep_input(x)
        if x.requires_grad:
            raise RuntimeError("gradients w.r.t. the input echogram are not produced by the native path")
        return _NativeUNetFunction.apply(x, self, *params)


class UNet_Baseline(_NativePlumbing, UNet):
    """The hot-path model (reference unet.py:304-343; the only model SegPipeUNet builds, pipeline.py:390-398)."""

    def __init__(self, n_classes, in_channels, meta_in_channels=0, late_meta_inject=False, depth=5, start_filts=64,
                 up_mode="transpose", merge_mode="concat"):
        super().__init__(n_classes, in_channels, meta_in_channels, late_meta_inject, depth, start_filts, up_mode, merge_mode)

    def _meta_head_channels(self):
        # UNet_Baseline(..., late_meta_inject=True) builds a wider head (unet.py:286-289) that its own forward cannot feed
        return self.conv_final.in_channels - self._native_head_in

    def _check_supported(self, x):
        if self.conv_final.in_channels != self._native_head_in:
            raise RuntimeError("crimac_unet_b200.UNet_Baseline cannot run this call natively: late_meta_inject heads "
                               "belong to UNet_LateMetInject")
        super()._check_supported(x)

    # ---- public surface
    def forward(self, x):
        """(N,C,H,W) fp32 -> raw logits (N,n_classes,H,W) fp32, as the reference (no softmax, unet.py:339-343)."""
        if self.training:
            return self._train_forward(x, self._param_tensors())
        return self._infer(x, softmax=False)

    @torch.no_grad()
    def predict_proba(self, x):
        """Eval forward with the softmax of pipeline.py:218 fused into the last kernel."""
        if self.training:
            raise RuntimeError("predict_proba() is an eval-mode call; use model.eval() first")
        return self._infer(x, softmax=True)

    @torch.no_grad()
    def validate_batch(self, x, labels, class_weight, prob_class=1):
        """The validation step of the reference's training loop (pipeline.py:249-270, get_predictions_dataloader) for one
        batch, all on device: eval forward, label codes remapped as set_label_ignore_val does (-70, -30, -100, -10 ->
        ignore; -50 -> background), the class-weighted CE on the remapped labels, softmax probability of `prob_class`
        (SANDEEL = 1).  labels: int16 (what the dataset emits) or int64, (N,H,W).
        Returns (loss: 0-dim device tensor, prob: fp32 (N,H,W), logits: fp32 (N,n_classes,H,W))."""
        if self.training:
            raise RuntimeError("validate_batch() is an eval-mode call; use model.eval() first")
        logits = self._infer(x, softmax=False)
        loss, prob, _ = _runtime().eval_loss(logits, labels.to(x.device), class_weight.to(x.device), prob_class)
        return loss, prob, logits

    @torch.no_grad()
    def forward_fp32(self, x, softmax=False):
        """fp32 VALIDATION mode: the eval forward through an independent plain-fp32 CUDA implementation (no bf16, no
        tensor cores) - logits, or probabilities with softmax=True.  For 1e-4 parity checks, ~50x slower."""
        if self.training:
            raise RuntimeError("forward_fp32() is an eval-mode call; use model.eval() first")
        if self.up_mode != "transpose" or self.merge_mode != "concat":
            raise RuntimeError("the fp32 validation mode covers up_mode='transpose' with merge_mode='concat' (the only "
                               "configuration the reference pipeline builds, pipeline.py:390-410)")
        x = self._prep_input(x)
        rt = _runtime()
        return rt.forward_infer_fp32((self.in_channels, self.n_classes, self.depth, self.start_filts),
                                     self._state_tensors(), x, softmax)

    def train_step_fused(self, x, labels, class_weight, ignore_index=-100):
        """Forward + class-weighted CE + backward in one native call (pipeline.py:171-177).

        Fills ``p.grad`` of every parameter (views of one flat arena kept in ``self._grad_arena``) and returns the
        loss as a 0-dim device tensor (no host sync).  Gradients are OVERWRITTEN, not accumulated (the call is
        ``zero_grad(); loss.backward()`` in one): accumulate over micro-batches by summing the arena yourself."""
        if not self.training:
            raise RuntimeError("train_step_fused() needs model.train()")
        x = self._prep_input(x)
        eng = self._engine_for(x, train=True)
        params = self._param_tensors()
        arena = getattr(self, "_grad_arena", None)
        total = sum(p.numel() for p in params)
        if arena is None or arena.numel() != total or arena.device != x.device:
            arena = torch.zeros(total, dtype=torch.float32, device=x.device)
            self._grad_arena = arena
        # (re-)seat every p.grad as a view of the arena: optimizer.zero_grad(set_to_none=True), model.zero_grad() or
        # user code may have dropped or replaced the views since the last call
        off, base, grads = 0, arena.data_ptr(), []
        for p in params:
            g = p.grad
            if g is None or g.data_ptr() != base + 4 * off or g.shape != p.shape or g.dtype != torch.float32:
                g = arena[off:off + p.numel()].view_as(p)
                p.grad = g
            grads.append(g)
            off += p.numel()
        loss3 = torch.empty(4, dtype=torch.float32, device=x.device)
        state = eng.state_table(self._state_tensors())
        eng.train_step(state, x, labels.contiguous().long(), class_weight.contiguous().float(), ignore_index,
                       eng.grad_table(grads), loss3)
        self._native_mutated()           # BatchNorm running statistics
        eng.fwd_stamp += 1               # the saved activations of an outstanding autograd forward are gone
        return loss3[0]


class UNet_LateMetInject(_NativePlumbing, UNet):
    """Late metadata injection variant (reference unet.py:346-391).

    The reference concatenates the decoder's 64 channels with MetaPostProcessing(meta) and applies conv1x1(65, 3).  A 1x1
    conv over a concatenation is the sum of two 1x1 convs, so the 64-channel part - the whole U-Net body plus
    W[:, :64] and the bias - runs through the native library exactly like UNet_Baseline, and the metadata part
    (a per-sample MLP and W[:, 64:]) stays a few elementwise torch ops whose gradients autograd handles:
        logits = native(x; body, W[:, :64], b) + sum_k W[:, 64 + k] * MetaPostProcessing(meta)[k]"""

    def __init__(self, n_classes, in_channels, meta_in_channels, late_meta_inject=True, depth=5, start_filts=64,
                 up_mode="transpose", merge_mode="concat"):
        super().__init__(n_classes, in_channels, meta_in_channels, late_meta_inject, depth, start_filts, up_mode, merge_mode)
        self.conv_final = conv1x1(65, 3)  # hard-coded in the reference as well (unet.py:370)
        self._w64 = None
        self._w64_version = None

    def _meta_head_channels(self):
        return self.post_processing_weights.out_channels

    def _head_weight(self):
        """Eval path: a contiguous copy of W[:, :64], refreshed only when the parameter changes (so that the packed
        weights are not rebuilt on every call)."""
        w = self.conv_final.weight
        ver = (w.data_ptr(), w._version)
        if self._w64 is None or self._w64_version != ver or self._w64.device != w.device:
            with torch.no_grad():
                self._w64 = w[:, :self._native_head_in].contiguous()
            self._w64_version = ver
        return self._w64

    def _meta_logits(self, meta_tensor):
        w_meta = self.conv_final.weight[:, self._native_head_in:, 0, 0]                 # (n_classes, K)
        mp = self.post_processing_weights(meta_tensor.float())                          # (N, K, H, W)
        return (mp.unsqueeze(1) * w_meta.reshape(1, w_meta.shape[0], w_meta.shape[1], 1, 1)).sum(2)

    def forward(self, x, meta_tensor):
        """(N,C,H,W) echogram + (N,M,H,W) metadata -> raw logits (N,3,H,W), as the reference (unet.py:372-391)."""
        if self.training:
            body = [p for m in (self.down_convs, self.up_convs) for p in m.parameters()]
            w64 = self.conv_final.weight[:, :self._native_head_in].contiguous()        # differentiable slice
            native = self._train_forward(x, body + [w64, self.conv_final.bias])
        else:
            native = self._infer(x, softmax=False)
        return native + self._meta_logits(meta_tensor)


if __name__ == "__main__":
    pass

"""Host-side driver of libcrimac_b200.so for one UNet module: owns the native context + workspace (a torch uint8
tensor), builds the parameter tables the C-ABI expects, and exposes the three calls the module uses
(infer / train-forward / backward) plus the fused train step.  PyTorch is plumbing here: device memory, streams.
"""
import ctypes

import torch

from . import lib as _lib


class _Config(ctypes.Structure):
    _fields_ = [
        ("in_channels", ctypes.c_int),
        ("n_classes", ctypes.c_int),
        ("depth", ctypes.c_int),
        ("start_filts", ctypes.c_int),
        ("max_batch", ctypes.c_int),
        ("height", ctypes.c_int),
        ("width", ctypes.c_int),
        ("train", ctypes.c_int),
        ("deterministic", ctypes.c_int),
        ("up_mode", ctypes.c_int),       # 0 "transpose", 1 "upsample"
        ("merge_mode", ctypes.c_int),    # 0 "concat", 1 "add"
    ]


class _CommConfig(ctypes.Structure):
    _fields_ = [
        ("world", ctypes.c_int),
        ("rank", ctypes.c_int),
        ("peer_arenas", ctypes.c_void_p * 8),
        ("peer_pads", ctypes.c_void_p * 8),
        ("multicast_arena", ctypes.c_void_p),
        ("local_state", ctypes.c_void_p),
        ("arena_floats", ctypes.c_size_t),
        ("ctas", ctypes.c_int),
    ]


class _OptConfig(ctypes.Structure):
    _fields_ = [
        ("params", ctypes.c_void_p),
        ("momentum", ctypes.c_void_p),
        ("grads", ctypes.c_void_p),
        ("n", ctypes.c_size_t),
        ("lr", ctypes.c_float),
        ("momentum_coef", ctypes.c_float),
        ("gscale", ctypes.c_float),
    ]


AR_PAD_BYTES = 8 * 2 * 16 * 4      # CRIMAC_AR_PAD_BYTES
AR_STATE_BYTES = 8 * 8             # 8 bytes per bucket, CRIMAC_AR_MAX_BUCKETS = 8


def make_comm_config(world, rank, peer_arenas, peer_pads, multicast_arena, local_state_ptr, arena_floats, ctas=0):
    cfg = _CommConfig()
    cfg.world, cfg.rank = int(world), int(rank)
    for r in range(world):
        cfg.peer_arenas[r] = int(peer_arenas[r])
        cfg.peer_pads[r] = int(peer_pads[r])
    cfg.multicast_arena = int(multicast_arena) if multicast_arena else None
    cfg.local_state = int(local_state_ptr)
    cfg.arena_floats = int(arena_floats)
    cfg.ctas = int(ctas)
    return cfg


def peer_allreduce(comm, bucket, offset, count, stream=None):
    """One bucket of the peer-memory all-reduce on its own (crimac_peer_allreduce); tests and tools."""
    L = _lib.load()
    _lib.check(
        L.crimac_peer_allreduce(comm.peer_arenas, comm.peer_pads, ctypes.c_void_p(comm.multicast_arena),
                                ctypes.c_void_p(comm.local_state), comm.rank, comm.world, int(bucket),
                                ctypes.c_size_t(offset), ctypes.c_size_t(count), comm.ctas, _lib.stream_ptr(stream)),
        "crimac_peer_allreduce",
    )


_UP_MODES = {"transpose": 0, "upsample": 1}
_MERGE_MODES = {"concat": 0, "add": 1}


def workspace_bytes(in_channels, n_classes, depth, start_filts, max_batch, height, width, train, deterministic=False,
                    up_mode="transpose", merge_mode="concat"):
    """Size of the device workspace a context of this shape needs (pure host computation)."""
    L = _lib.load()
    cfg = _Config(in_channels, n_classes, depth, start_filts, max_batch, height, width, int(train), int(deterministic),
                  _UP_MODES[up_mode], _MERGE_MODES[merge_mode])
    n = ctypes.c_size_t(0)
    _lib.check(L.crimac_workspace_bytes(ctypes.byref(cfg), ctypes.byref(n)), "crimac_workspace_bytes")
    return n.value


class Context:
    """One native context: fixed (max_batch, H, W), inference-only or train-capable."""

    def __init__(self, in_channels, n_classes, depth, start_filts, max_batch, height, width, train, device,
                 deterministic=False, up_mode="transpose", merge_mode="concat"):
        self.L = _lib.load()
        self.cfg = _Config(in_channels, n_classes, depth, start_filts, max_batch, height, width, int(train),
                           int(bool(deterministic) and bool(train)), _UP_MODES[up_mode], _MERGE_MODES[merge_mode])
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.CrimacError("the CRIMAC U-Net hot path runs on a CUDA (sm_100a) device only; there is no CPU fallback")
        n = ctypes.c_size_t(0)
        _lib.check(self.L.crimac_workspace_bytes(ctypes.byref(self.cfg), ctypes.byref(n)), "crimac_workspace_bytes")
        self.workspace = torch.empty(n.value + 1024, dtype=torch.uint8, device=self.device)
        base = (self.workspace.data_ptr() + 1023) & ~1023
        self.handle = ctypes.c_void_p()
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        _lib.check(
            self.L.crimac_create(ctypes.byref(self.handle), ctypes.byref(self.cfg), ctypes.c_void_p(base),
                                 ctypes.c_size_t(n.value), dev_index),
            "crimac_create",
        )
        self.n_state = self.L.crimac_state_count(ctypes.byref(self.cfg))
        self.n_grad = self.L.crimac_grad_count(ctypes.byref(self.cfg))
        self.prepared_key = None
        self.fwd_stamp = 0   # counts train-mode forwards: the workspace holds the saved activations of the LAST one only

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.L.crimac_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def set_comm(self, comm):
        """Switch the bucketed peer-memory gradient all-reduce of this context on (a _CommConfig) or off (None)."""
        self._comm = comm   # keep the structure (and the pointers in it) alive
        _lib.check(self.L.crimac_set_comm(self.handle, ctypes.byref(comm) if comm is not None else None), "crimac_set_comm")

    def set_optimizer(self, opt):
        """Fuse SGD(momentum) into this context's backward, per gradient bucket (an _OptConfig), or switch it off (None)."""
        self._opt = opt
        _lib.check(self.L.crimac_set_optimizer(self.handle, ctypes.byref(opt) if opt is not None else None), "crimac_set_optimizer")

    # ---- tables
    def state_table(self, tensors):
        if len(tensors) != self.n_state:
            raise _lib.CrimacError(f"state table needs {self.n_state} tensors, got {len(tensors)}")
        arr = (ctypes.c_void_p * self.n_state)(*[t.data_ptr() for t in tensors])
        return arr

    def grad_table(self, tensors):
        if len(tensors) != self.n_grad:
            raise _lib.CrimacError(f"grad table needs {self.n_grad} tensors, got {len(tensors)}")
        return (ctypes.c_void_p * self.n_grad)(*[t.data_ptr() for t in tensors])

    # ---- calls
    def prepare(self, state, train):
        _lib.check(self.L.crimac_prepare(self.handle, state, int(train), _lib.stream_ptr()), "crimac_prepare")

    def forward_infer(self, state, x, out, softmax):
        _lib.check(
            self.L.crimac_forward_infer(self.handle, state, _lib.ptr(x), x.shape[0], _lib.ptr(out), int(softmax),
                                        _lib.stream_ptr()),
            "crimac_forward_infer",
        )

    def preprocess_staged(self, sv, data_ping0, centres, nan_mask):
        """Patch gather + dB transform straight into the first conv's operand inside this context (crimac_preprocess_staged);
        consumed by the next forward_infer(state, None, ...)."""
        F, R, P = sv.shape
        _lib.check(
            self.L.crimac_preprocess_staged(self.handle, _lib.ptr(sv), F, R, P, int(data_ping0), _lib.ptr(centres),
                                            centres.shape[0], _lib.ptr(nan_mask), _lib.stream_ptr()),
            "crimac_preprocess_staged",
        )

    def forward_infer_staged(self, state, nb, out, softmax):
        _lib.check(
            self.L.crimac_forward_infer(self.handle, state, None, nb, _lib.ptr(out), int(softmax), _lib.stream_ptr()),
            "crimac_forward_infer",
        )

    def forward_infer_stitch(self, state, nb, centres, nan_mask, out, ping_start, overlap, labels=None, seabed=None,
                             seabed_pad=10, classes=(1, 2), x=None):
        """Eval forward + softmax + overlap stitching in the last conv's epilogue (crimac_forward_infer_stitch)."""
        K, R, Pc = out.shape
        cls = (ctypes.c_int32 * len(classes))(*classes)
        _lib.check(
            self.L.crimac_forward_infer_stitch(self.handle, state, _lib.ptr(x), nb, _lib.ptr(centres), _lib.ptr(nan_mask),
                                               _lib.ptr(labels), _lib.ptr(seabed), int(seabed_pad), int(overlap),
                                               int(ping_start), Pc, R, cls, K, _lib.ptr(out), _lib.stream_ptr()),
            "crimac_forward_infer_stitch",
        )

    def forward_train(self, state, x, logits):
        _lib.check(
            self.L.crimac_forward_train(self.handle, state, _lib.ptr(x), x.shape[0], _lib.ptr(logits), _lib.stream_ptr()),
            "crimac_forward_train",
        )

    def loss(self, logits, labels, class_w, ignore_index, out3, dlogits):
        _lib.check(
            self.L.crimac_loss(self.handle, _lib.ptr(logits), _lib.ptr(labels), _lib.ptr(class_w),
                               ctypes.c_int64(ignore_index), logits.shape[0], _lib.ptr(out3), _lib.ptr(dlogits),
                               _lib.stream_ptr()),
            "crimac_loss",
        )

    def backward(self, state, x, dlogits, gscale, grads):
        _lib.check(
            self.L.crimac_backward(self.handle, state, _lib.ptr(x), _lib.ptr(dlogits), _lib.ptr(gscale), x.shape[0],
                                   grads, _lib.stream_ptr()),
            "crimac_backward",
        )

    def train_step(self, state, x, labels, class_w, ignore_index, grads, loss3):
        _lib.check(
            self.L.crimac_train_step(self.handle, state, _lib.ptr(x), _lib.ptr(labels), _lib.ptr(class_w),
                                     ctypes.c_int64(ignore_index), x.shape[0], grads, _lib.ptr(loss3),
                                     _lib.stream_ptr()),
            "crimac_train_step",
        )


def forward_infer_fp32(cfg_tuple, state_tensors, x, softmax):
    """fp32 validation forward (independent CUDA-core implementation). cfg_tuple = (in_ch, n_classes, depth, start_filts)."""
    L = _lib.load()
    nb, _, h, w = x.shape
    cfg = _Config(cfg_tuple[0], cfg_tuple[1], cfg_tuple[2], cfg_tuple[3], nb, h, w, 0, 0, 0, 0)
    n = ctypes.c_size_t(0)
    _lib.check(L.crimac_fp32_workspace_bytes(ctypes.byref(cfg), nb, ctypes.byref(n)), "crimac_fp32_workspace_bytes")
    ws = torch.empty(n.value, dtype=torch.uint8, device=x.device)
    out = torch.empty((nb, cfg_tuple[1], h, w), dtype=torch.float32, device=x.device)
    state = (ctypes.c_void_p * len(state_tensors))(*[t.data_ptr() for t in state_tensors])
    _lib.check(
        L.crimac_forward_infer_fp32(ctypes.byref(cfg), state, _lib.ptr(x), nb, _lib.ptr(out), int(softmax), _lib.ptr(ws),
                                    ctypes.c_size_t(n.value), _lib.stream_ptr()),
        "crimac_forward_infer_fp32",
    )
    return out


def sgd_step(params_flat, momentum_flat, grads_flat, lr, momentum, gscale=1.0):
    """Fused SGD(momentum) on flat fp32 arenas (reference pipeline.py:156,178)."""
    L = _lib.load()
    _lib.check(
        L.crimac_sgd_step(_lib.ptr(params_flat), _lib.ptr(momentum_flat), _lib.ptr(grads_flat),
                          ctypes.c_size_t(params_flat.numel()), ctypes.c_float(lr), ctypes.c_float(momentum),
                          ctypes.c_float(gscale), _lib.stream_ptr()),
        "crimac_sgd_step",
    )


def eval_loss(logits, labels, class_w, prob_class=1, want_labels=False):
    """Validation step on eval-mode logits (reference pipeline.py:222-239 set_label_ignore_val, :264 criterion, :269-270
    softmax + SANDEEL channel) in one kernel.  labels: int16 or int64 (N,H,W) raw label codes.  Returns
    (loss 0-dim device tensor, probability of `prob_class` fp32 (N,H,W), remapped int64 labels or None)."""
    L = _lib.load()
    if labels.dtype not in (torch.int16, torch.int64):
        raise ValueError("labels must be int16 or int64")
    n, ncls, h, w = logits.shape
    logits, labels = logits.contiguous().float(), labels.contiguous()
    prob = torch.empty((n, h, w), dtype=torch.float32, device=logits.device)
    lab_out = torch.empty((n, h, w), dtype=torch.int64, device=logits.device) if want_labels else None
    out3 = torch.empty(4, dtype=torch.float32, device=logits.device)
    scratch = torch.empty(16384, dtype=torch.uint8, device=logits.device)
    _lib.check(
        L.crimac_eval_loss(_lib.ptr(logits), n, ncls, h, w, _lib.ptr(labels), 16 if labels.dtype == torch.int16 else 64,
                           _lib.ptr(class_w.contiguous().float()), int(prob_class), _lib.ptr(prob), _lib.ptr(lab_out),
                           _lib.ptr(out3), _lib.ptr(scratch), _lib.stream_ptr()),
        "crimac_eval_loss",
    )
    return out3[0], prob, lab_out


def saved_tensor(ctx, index, which, nb):
    """Test hook (crimac_dbg_saved): a dense NHWC bf16 copy of a tensor the last train-mode forward kept for backward."""
    dims = (ctypes.c_int * 3)()
    _lib.check(ctx.L.crimac_dbg_saved(ctx.handle, index, which, nb, None, dims, _lib.stream_ptr()), "crimac_dbg_saved")
    out = torch.empty((nb, dims[0], dims[1], dims[2]), dtype=torch.bfloat16, device=ctx.device)
    _lib.check(ctx.L.crimac_dbg_saved(ctx.handle, index, which, nb, _lib.ptr(out), dims, _lib.stream_ptr()), "crimac_dbg_saved")
    return out


def preprocess(sv, data_ping0, centres, patch_hw, out=None, nan_mask=None):
    """sv: fp32 (F,R,P) device tensor; centres: int32 (n,2) device tensor (y,x) survey coords."""
    L = _lib.load()
    F, R, P = sv.shape
    n = centres.shape[0]
    ph, pw = patch_hw
    if out is None:
        out = torch.empty((n, F, ph, pw), dtype=torch.float32, device=sv.device)
    if nan_mask is None:
        nan_mask = torch.empty((n, ph, pw), dtype=torch.uint8, device=sv.device)
    _lib.check(
        L.crimac_preprocess(_lib.ptr(sv), F, R, P, int(data_ping0), _lib.ptr(centres), n, ph, pw, _lib.ptr(out),
                            _lib.ptr(nan_mask), _lib.stream_ptr()),
        "crimac_preprocess",
    )
    return out, nan_mask


META_BITS = {"portion_year": 1, "portion_day": 2, "time_diff": 4, "depth_rel": 8, "depth_abs_surface": 16,
             "depth_abs_seabed": 32}


def meta_channels(x, c_off, centres, meta_channels_cfg, portion_year=0.0, portion_of_day=None, time_diff=None,
                  seabed=None, n_range=None):
    """Writes the metadata input channels of the reference's get_crop_memmap (batch/dataset.py:296-349) into planes
    [c_off, c_off + M) of the network input x (n, C, ph, pw) on the device.  meta_channels_cfg: the config's
    data.meta_channels dict (truthy entries are generated, in the reference's order); the three vectors are float64
    device tensors (portion_of_day_vector, time_vector_diff, _seabed).  n_range: the echogram's number of range bins - if
    it is <= ph the reference re-centres every crop vertically (dataset.py:261-262) and so does this call.
    Returns the number of channels written."""
    L = _lib.load()
    mask = sum(bit for k, bit in META_BITS.items() if meta_channels_cfg.get(k))
    n, c_total, ph, pw = x.shape
    if n_range is not None and n_range <= ph:
        centres = centres.clone()
        centres[:, 0] = n_range // 2
    for t, name in ((portion_of_day, "portion_of_day"), (time_diff, "time_diff"), (seabed, "seabed")):
        if t is not None and (t.dtype != torch.float64 or not t.is_cuda or not t.is_contiguous()):
            raise ValueError(f"{name} must be a contiguous float64 CUDA tensor")
    _lib.check(
        L.crimac_meta_channels(_lib.ptr(centres.contiguous()), n, ph, pw, ctypes.c_uint(mask), ctypes.c_double(portion_year),
                               _lib.ptr(portion_of_day), 0 if portion_of_day is None else portion_of_day.numel(),
                               _lib.ptr(time_diff), 0 if time_diff is None else time_diff.numel(),
                               _lib.ptr(seabed), 0 if seabed is None else seabed.numel(),
                               _lib.ptr(x), c_total, int(c_off), _lib.stream_ptr()),
        "crimac_meta_channels",
    )
    return sum(2 if k == "portion_day" else 1 for k in META_BITS if meta_channels_cfg.get(k))


def stitch(probs, centres, nan_mask, out, ping_start, overlap, labels=None, seabed=None, seabed_pad=10, classes=(1, 2)):
    """probs: fp32 (n,ncls,ph,pw); out: fp16 (K,R,Pc) device tensor written in place."""
    L = _lib.load()
    n, ncls, ph, pw = probs.shape
    K, R, Pc = out.shape
    cls = (ctypes.c_int32 * len(classes))(*classes)
    _lib.check(
        L.crimac_stitch(_lib.ptr(probs), n, ncls, ph, pw, _lib.ptr(centres), _lib.ptr(nan_mask), _lib.ptr(labels),
                        _lib.ptr(seabed), int(seabed_pad), int(overlap), int(ping_start), Pc, R, cls, K, _lib.ptr(out),
                        _lib.stream_ptr()),
        "crimac_stitch",
    )
    return out


def train_patches(sv, labels, centres, flags, patch_hw, noise_mult=None, seed=0, thr_freq=None, thr=(1e-7, 1e-4),
                  scaled=False, border_zero=False, out=None, labels_out=None):
    """One batch of training samples from a survey resident in HBM (crimac_train_patches).
    sv: fp32 (F,P,R) device tensor in the zarr store's [frequency][ping][range] order; labels: fp32 (P,R) raw
    annotation categories; centres: int32 (n,2) (range, ping); flags: uint8 (n), bit0 = noise, bit1 = flip."""
    L = _lib.load()
    F, P, R = sv.shape
    n = centres.shape[0]
    ph, pw = patch_hw
    for t, dt, name in ((sv, torch.float32, "sv"), (labels, torch.float32, "labels"), (centres, torch.int32, "centres"),
                        (flags, torch.uint8, "flags")):
        if t.dtype != dt or not t.is_cuda or not t.is_contiguous():
            raise ValueError(f"{name} must be a contiguous {dt} CUDA tensor")
    if tuple(labels.shape) != (P, R) or tuple(centres.shape) != (n, 2) or tuple(flags.shape) != (n,):
        raise ValueError("labels must be (P,R), centres (n,2), flags (n)")
    if noise_mult is not None and (tuple(noise_mult.shape) != (n, F, ph, pw) or noise_mult.dtype != torch.float32
                                   or not noise_mult.is_cuda or not noise_mult.is_contiguous()):
        raise ValueError("noise_mult must be a contiguous fp32 CUDA tensor of shape (n,F,ph,pw)")
    if out is None:
        out = torch.empty((n, F, ph, pw), dtype=torch.float32, device=sv.device)
    if labels_out is None:
        labels_out = torch.empty((n, ph, pw), dtype=torch.int64, device=sv.device)
    thr_freq = F - 1 if thr_freq is None else int(thr_freq)
    _lib.check(
        L.crimac_train_patches(_lib.ptr(sv), _lib.ptr(labels), F, P, R, _lib.ptr(centres), _lib.ptr(flags),
                               _lib.ptr(noise_mult), ctypes.c_uint64(int(seed) & 0xFFFFFFFFFFFFFFFF), n, ph, pw,
                               thr_freq, ctypes.c_double(thr[0]), ctypes.c_double(thr[1]), int(bool(scaled)),
                               int(bool(border_zero)), _lib.ptr(out), _lib.ptr(labels_out), _lib.stream_ptr()),
        "crimac_train_patches",
    )
    return out, labels_out


def launch_count():
    L = _lib.load()
    L.crimac_launch_count.restype = ctypes.c_ulonglong
    return int(L.crimac_launch_count())


def profile_enable(on=True):
    _lib.check(_lib.load().crimac_profile_enable(int(on)), "crimac_profile_enable")


def profile_read():
    """Returns a list of (kernel family, ms, algorithmic flops, algorithmic bytes, launches) since profile_enable(1)."""
    L = _lib.load()
    cap = 4096
    names = (ctypes.c_char_p * cap)()
    ms = (ctypes.c_float * cap)()
    fl = (ctypes.c_double * cap)()
    by = (ctypes.c_double * cap)()
    ln = (ctypes.c_int * cap)()
    n = L.crimac_profile_read(names, ms, fl, by, ln, cap)
    n = min(n, cap)
    return [(names[i].decode(), float(ms[i]), float(fl[i]), float(by[i]), int(ln[i])) for i in range(n)]

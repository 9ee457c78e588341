"""Sliding-window whole-echogram inference on device — the hot loop of the reference's
pipeline_train_predict/save_predict.py:137-220 (save_survey_predictions_zarr) without its per-patch CPU work:

    for each preload chunk of pings (utils/preload_data_split.py:22-30), sharded over GPUs by contiguous ping range:
        grid of overlapping patches        (batch/samplers/gridded.py:22-54; host arithmetic only)
        crimac_preprocess_staged           gather + NaN fill + sv->dB + clip, straight from the preloaded pings INTO the
                                           first conv's bf16 hi/lo operand (no fp32 patch tensor in between)
        crimac_forward_infer_stitch        the tcgen05 forward; the last conv's epilogue does the 1x1 head, the softmax AND
                                           the overlap-stitch of classes [SANDEEL, OTHER] into (2, range, pings) fp16
        (UNet_Baseline.predict_stitch_patches; crimac_preprocess / predict_proba / crimac_stitch remain as separate steps)

Zarr reading / writing stays with the caller (out of scope: I/O format), which hands in device or host arrays.
"""
import numpy as np
import torch

from . import engine as _engine


def split_pings(start, end, max_n_pings):
    """Equal preload chunks as the reference computes them (np.linspace(...).astype(int))."""
    n = int(np.ceil((end - start) / max_n_pings))
    edges = np.linspace(start, end, n + 1).astype(int)
    return [(int(edges[i]), int(edges[i + 1])) for i in range(n)]


def shard_chunks(chunks, world, rank):
    """Contiguous ping-range sharding at chunk granularity: the first (len % world) ranks take one extra chunk."""
    n = len(chunks)
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return chunks[lo:hi]


def patch_grid(start_ping, end_ping, end_range, patch_hw, overlap):
    """Patch centres (y, x), y-major, stride = patch - 2*overlap, first upper-left corner at start - (overlap+1)."""
    ph, pw = patch_hw
    ys = np.arange(0 - (overlap + 1), end_range - (overlap + 1), ph - 2 * overlap) + ph // 2
    xs = np.arange(start_ping - (overlap + 1), end_ping - (overlap + 1), pw - 2 * overlap) + pw // 2
    yy, xx = np.meshgrid(ys, xs, indexing="ij")
    return np.stack([yy.reshape(-1), xx.reshape(-1)], 1).astype(np.int32)


def preload_window(grid, n_pings_total, pw):
    """Pings the patches of this grid read (neighbouring chunks' pings serve as context)."""
    return max(0, int(grid[0, 1]) - pw // 2), min(n_pings_total, int(grid[-1, 1]) + pw // 2)


@torch.no_grad()
def predict_host_batches(model, batches, classes=(1, 2)):
    """Class probabilities for a stream of HOST batches (pin them for full speed): yields, per batch, a pinned fp16 host
    tensor (N, len(classes), H, W) - the two classes save_predict.py keeps (:43-65).  Host->device copy of batch i+1,
    the forward of batch i and the device->host copy of batch i-1 overlap (copy stream + two buffers each way); the
    reference does the three steps one after the other (pipeline.py:208-218, save_predict.py:196)."""
    dev = next(model.parameters()).device
    copy_in, copy_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)
    xin, ready, freed = [None, None], [torch.cuda.Event(), torch.cuda.Event()], [None, None]
    dout, hout, done, drained = [None, None], [None, None], [torch.cuda.Event(), torch.cuda.Event()], [None, None]
    cls = list(classes)

    def upload(slot, xh):
        if xin[slot] is None or xin[slot].shape != xh.shape:
            xin[slot] = torch.empty(xh.shape, dtype=torch.float32, device=dev)
        with torch.cuda.stream(copy_in):
            if freed[slot] is not None:
                copy_in.wait_event(freed[slot])
            xin[slot].copy_(xh, non_blocking=True)
            ready[slot].record(copy_in)

    it = iter(batches)
    nxt = next(it, None)
    if nxt is None:
        return
    upload(0, nxt)
    i, pending = 0, None
    while nxt is not None:
        slot = i & 1
        nxt = next(it, None)
        if nxt is not None:
            upload(slot ^ 1, nxt)
        main.wait_event(ready[slot])
        if drained[slot] is not None:
            main.wait_event(drained[slot])            # the D2H copy that last read dout[slot] has finished
        probs = model.predict_proba(xin[slot])
        dout[slot] = probs[:, cls].half()
        freed[slot] = torch.cuda.Event()
        freed[slot].record(main)
        done[slot].record(main)
        if hout[slot] is None or hout[slot].shape != dout[slot].shape:
            hout[slot] = torch.empty(dout[slot].shape, dtype=torch.float16).pin_memory()
        with torch.cuda.stream(copy_out):
            copy_out.wait_event(done[slot])
            hout[slot].copy_(dout[slot], non_blocking=True)
            drained[slot] = torch.cuda.Event()
            drained[slot].record(copy_out)
        if pending is not None:
            pending[1].synchronize()
            yield pending[0]
        pending = (hout[slot], drained[slot])
        i += 1
    pending[1].synchronize()
    yield pending[0]


class SurveyPredictor:
    def __init__(self, model, patch_hw=(256, 256), overlap=20, preload_n_pings=20000, batch_size=93, classes=(1, 2),
                 seabed_pad=10, direct=None):
        # direct: preprocessing writes the first conv's operand (predict_proba_patches); default when the model has it
        self.direct = hasattr(model, "predict_stitch_patches") if direct is None else bool(direct)
        self.model, self.patch_hw, self.overlap = model, tuple(patch_hw), int(overlap)
        self.preload_n_pings, self.batch_size = int(preload_n_pings), int(batch_size)
        self.classes, self.seabed_pad = tuple(classes), int(seabed_pad)
        self.last_chunk_patches = 0

    def chunk_geometry(self, start, end, n_range, n_pings_total, seabed_max=None):
        """(grid, (data_ping0, data_ping1)) for one chunk; the range extent is cut at max seabed + 50 when known."""
        end_range = n_range if seabed_max is None else min(n_range, int(seabed_max) + 50)
        grid = patch_grid(start, end, end_range, self.patch_hw, self.overlap)
        return grid, preload_window(grid, n_pings_total, self.patch_hw[1])

    @torch.no_grad()
    def predict_chunk(self, sv, data_ping0, grid, start, end, labels=None, seabed=None, out=None):
        """sv: fp32 device (F, R, P) = pings [data_ping0, data_ping0+P); grid: int32 (n,2) host/device centres;
        labels: optional int16 device (R, end-start); seabed: optional int32 device (end-start).
        Returns fp16 device (len(classes), R, end-start)."""
        dev = sv.device
        R = sv.shape[1]
        centres = torch.as_tensor(grid, dtype=torch.int32, device=dev)
        if out is None:
            out = torch.zeros((len(self.classes), R, end - start), dtype=torch.float16, device=dev)
        for i in range(0, centres.shape[0], self.batch_size):
            c = centres[i:i + self.batch_size].contiguous()
            if self.direct:
                # one native sequence per batch: gather + dB straight into the first conv's operand, forward, and the
                # stitch in the last conv's epilogue (no patch tensor, no probability tensor)
                self.model.predict_stitch_patches(sv, data_ping0, c, self.patch_hw, out, start, self.overlap,
                                                  labels=labels, seabed=seabed, seabed_pad=self.seabed_pad,
                                                  classes=self.classes)
                continue
            # three-step form (kept as the A/B of the tests): fp32 patches -> probabilities -> crimac_stitch
            x, nan_mask = _engine.preprocess(sv, data_ping0, c, self.patch_hw)
            probs = self.model.predict_proba(x)
            _engine.stitch(probs, c, nan_mask, out, start, self.overlap, labels=labels, seabed=seabed,
                           seabed_pad=self.seabed_pad, classes=self.classes)
        return out

    def predict_survey(self, load_chunk, n_pings, n_range, rank=0, world=1, seabed_max_of=None):
        """Generator over this rank's chunks.  load_chunk(p0, p1) -> fp32 device tensor (F, R, p1-p0) [+ labels, seabed]."""
        chunks = shard_chunks(split_pings(0, n_pings, self.preload_n_pings), world, rank)
        for (s, e) in chunks:
            smax = seabed_max_of(s, e) if seabed_max_of is not None else None
            grid, (d0, d1) = self.chunk_geometry(s, e, n_range, n_pings, smax)
            self.last_chunk_patches = int(grid.shape[0])
            loaded = load_chunk(d0, d1, s, e)
            sv, labels, seabed = loaded if isinstance(loaded, tuple) else (loaded, None, None)
            yield s, e, self.predict_chunk(sv, d0, grid, s, e, labels=labels, seabed=seabed)

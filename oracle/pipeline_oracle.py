"""ORACLE (test infrastructure, not product code): numpy restatement of the reference's CPU steps either side of
the U-Net — patch grid, preload gather + sv->dB transform, label masks and the overlap stitcher.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this file.
Pinned against the reference's own functions (driven by a fake in-memory reader) by oracle/make_golden.py ->
tests/golden/pipeline_*.npz.  Paths cited are relative to /root/reference/crimac_unet/.
"""
import numpy as np

LABEL_IGNORE_VAL = -100        # constants.py:25
LABEL_BOUNDARY_VAL = -100      # constants.py:26
LABEL_OVERLAP_VAL = -70        # constants.py:27
LABEL_SEABED_MASK_VAL = -50    # constants.py:28
BACKGROUND, SANDEEL, OTHER = 0, 1, 2   # constants.py:20-22


def get_data_split(valid_ping_ranges, max_n_pings=1000):
    """utils/preload_data_split.py:22-30: equal chunks via np.linspace(...).astype(int)."""
    splits = []
    for start, end in valid_ping_ranges:
        n_splits = int(np.ceil((end - start) / max_n_pings))
        edges = np.linspace(start, end, n_splits + 1).astype(int)
        splits.extend([[edges[i], edges[i + 1]] for i in range(n_splits)])
    return np.array(splits)


def get_data_grid(start_ping, end_ping, start_range, end_range, patch_size=(256, 256), patch_overlap=20):
    """batch/samplers/gridded.py:22-54 (mode 'all'): centres (y, x), all x for the first y, then the next y.
    NB the reference unpacks (patch_width, patch_height) = patch_size (gridded.py:37); square patches only here."""
    pw, ph = patch_size
    ys = np.arange(start_range - (patch_overlap + 1), end_range - (patch_overlap + 1), ph - 2 * patch_overlap) + ph // 2
    xs = np.arange(start_ping - (patch_overlap + 1), end_ping - (patch_overlap + 1), pw - 2 * patch_overlap) + pw // 2
    return np.array(np.meshgrid(ys, xs)).T.reshape(-1, 2)


def end_range_from_seabed(R, seabed_idx):
    """gridded.py:151-159: range limited to the deepest seabed in the chunk + 50."""
    return int(min(R, int(np.max(seabed_idx)) + 50))


def preload_extents(grid, n_pings_total, patch_w):
    """batch/dataset.py:176-177: pings [first_cx - w/2, last_cx + w/2) clipped to the survey."""
    return max(0, int(grid[0, 1]) - patch_w // 2), min(n_pings_total, int(grid[-1, 1]) + patch_w // 2)


def patch_offsets(n):
    """utils/np.py:40-46 getGrid for one axis: -((n+1)//2)+1 .. n//2."""
    return np.arange(-((n + 1) // 2) + 1, n // 2 + 1)


def gather_data(sv, centre, data_ping0, patch_hw):
    """dataset.py:192-205 + utils/np.py:362-375 (new_get_crop_3d, boundary 0): sv is (F, R, P) [freq][range][ping]."""
    ph, pw = patch_hw
    ys = centre[0] + patch_offsets(ph)
    xs = centre[1] - data_ping0 + patch_offsets(pw)
    yy, xx = np.meshgrid(ys, xs, indexing="ij")
    oob = (yy < 0) | (xx < 0) | (yy >= sv.shape[1]) | (xx >= sv.shape[2])
    yc, xc = np.where(oob, 0, yy), np.where(oob, 0, xx)
    out = sv[:, yc, xc].copy()
    out[:, oob] = 0
    return out


def gather_labels(labels, centre, label_ping0, patch_hw):
    """dataset.py:196-198 + utils/np.py:347-360 (new_get_crop_2d, boundary LABEL_BOUNDARY_VAL): labels (R, Pc)."""
    ph, pw = patch_hw
    ys = centre[0] + patch_offsets(ph)
    xs = centre[1] - label_ping0 + patch_offsets(pw)
    yy, xx = np.meshgrid(ys, xs, indexing="ij")
    oob = (yy < 0) | (xx < 0) | (yy >= labels.shape[0]) | (xx >= labels.shape[1])
    out = labels[np.where(oob, 0, yy), np.where(oob, 0, xx)].copy()
    out[oob] = LABEL_BOUNDARY_VAL
    return out


def data_transform(data, labels):
    """remove_nan_inf.py:23-34 then db_with_limits.py:20-24,36-38."""
    labels = labels.copy()
    labels[~np.isfinite(data[0])] = LABEL_IGNORE_VAL
    data = data.copy()
    data[~np.isfinite(data)] = 0.0
    db = 10 * np.log10(data + 1e-10)
    db[db > 0] = 0
    db[db < -75] = -75
    return db, labels


def mask_seabed(labels, centre, seabed_idx_survey, R, n_pings_total, pad=10):
    """mask_label_seabed.py:32-68 with a per-ping seabed index (mask[y] = y >= seabed[x]); the +pad shift happens
    inside the patch's clipped range window (data_reader.py:839-843)."""
    ph, pw = labels.shape
    y_upper, x_left = centre[0] - ph // 2 + 1, centre[1] - pw // 2 + 1
    y_lower, x_right = centre[0] + ph // 2 + 1, centre[1] + pw // 2 + 1
    sxl, syu = max(x_left, 0), max(y_upper, 0)
    sxr, syl = min(x_right, n_pings_total), min(y_lower, R)
    rng = np.arange(syu, syl)[:, None]
    m = (rng >= np.asarray(seabed_idx_survey[sxl:sxr])[None, :]).astype(np.int8)      # (range, ping) window
    mp = np.zeros_like(m)
    if pad:
        mp[pad:, :] = m[:-pad, :]
    else:
        mp = m
    full = np.zeros_like(labels)
    full[syu - y_upper:syu - y_upper + mp.shape[0], sxl - x_left:sxl - x_left + mp.shape[1]] = mp
    out = labels.copy()
    out[full.astype(bool) & (labels == BACKGROUND)] = LABEL_SEABED_MASK_VAL
    return out


def mask_overlap(labels, overlap):
    """mask_label_overlap.py:31-48."""
    if overlap == 0:
        return labels
    out = np.ones_like(labels) * LABEL_OVERLAP_VAL
    out[overlap:-overlap, overlap:-overlap] = labels[overlap:-overlap, overlap:-overlap]
    out[labels == LABEL_BOUNDARY_VAL] = LABEL_BOUNDARY_VAL
    return out


def fill_out_array(out_array, probs, labels, centre, ping_start, classes=(SANDEEL, OTHER)):
    """save_predict.py:41-65."""
    sel = np.argwhere((labels != LABEL_OVERLAP_VAL) & (labels != LABEL_SEABED_MASK_VAL) & (labels != LABEL_BOUNDARY_VAL))
    if len(sel) == 0:
        return out_array
    yl, xl = sel.T
    ya = yl + centre[0] - labels.shape[0] // 2 + 1
    xa = xl + centre[1] - labels.shape[1] // 2 + 1 - ping_start
    out_array[:, ya, xa] = probs[list(classes)][:, yl, xl]
    return out_array


def patch_item(sv_preload, data_ping0, labels_chunk, ping_start, centre, seabed_idx_survey, R, n_pings_total,
               patch_hw=(256, 256), overlap=20):
    """DatasetGriddedReader.__getitem__ (dataset.py:207-242) with save_predict's transforms (save_predict.py:149-150,
    batch/transforms.py:48-54,78-92; the label-only transforms convert_label_indexing_unused_species /
    refine_label_boundary are assumed applied to labels_chunk already). Returns (data fp32, labels int16)."""
    data = gather_data(sv_preload, centre, data_ping0, patch_hw)
    labels = gather_labels(labels_chunk, centre, ping_start, patch_hw)
    labels = mask_seabed(labels, centre, seabed_idx_survey, R, n_pings_total)
    labels = mask_overlap(labels, overlap)
    data, labels = data_transform(data, labels)
    return data.astype(np.float32), labels.astype(np.int16)


# ---- training-sample path: Dataset.__getitem__, batch/dataset.py:75-108 ---------------------
LABEL_REFINE_BOUNDARY_VAL = -30   # constants.py:29

DISC7 = np.array([[0, 0, 1, 1, 1, 0, 0],
                  [0, 1, 1, 1, 1, 1, 0],
                  [1, 1, 1, 1, 1, 1, 1],
                  [1, 1, 1, 1, 1, 1, 1],
                  [1, 1, 1, 1, 1, 1, 1],
                  [0, 1, 1, 1, 1, 1, 0],
                  [0, 0, 1, 1, 1, 0, 0]], dtype=bool)   # refine_label_boundary.py:53-61


def _morph(mask, structure, erode):
    """scipy.ndimage.binary_dilation / binary_erosion with border_value=0, origin 0 (symmetric structure)."""
    h, w = mask.shape
    r = structure.shape[0] // 2
    padded = np.zeros((h + 2 * r, w + 2 * r), dtype=bool)
    padded[r:r + h, r:r + w] = mask
    out = np.ones((h, w), dtype=bool) if erode else np.zeros((h, w), dtype=bool)
    for dy in range(-r, r + 1):
        for dx in range(-r, r + 1):
            if not structure[dy + r, dx + r]:
                continue
            win = padded[r + dy:r + dy + h, r + dx:r + dx + w]
            out = (out & win) if erode else (out | win)
    return out


def binary_closing(mask, structure=DISC7):
    """scipy.ndimage.binary_closing(mask, structure): dilation then erosion, both with border_value 0."""
    return _morph(_morph(mask, structure, erode=False), structure, erode=True)


def get_crop_zarr(sv_fpr, labels_pr, centre, patch_hw):
    """dataset.py:358-407. sv_fpr (F, pings, range) and labels_pr (pings, range) in the zarr store's order; returns
    float64 data (F, ph, pw) [freq, range, ping] and float64 labels (ph, pw); outside the data -> 0 / -100;
    np.nan_to_num on the slices (NaN -> 0 / -100, +-inf -> +-max of the slice dtype).  Square patches (the reference
    mixes window_size[0] / [1], :397-398)."""
    ph, pw = patch_hw
    F, P, R = sv_fpr.shape
    y0, x0 = centre[0] - ph // 2 + 1, centre[1] - pw // 2 + 1    # utils/np.py:378-380
    out_data = np.zeros((F, ph, pw), dtype=np.float64)
    out_labels = np.full((ph, pw), float(LABEL_BOUNDARY_VAL))
    xa, xb = max(x0, 0), min(P, x0 + pw)
    ya, yb = max(y0, 0), min(R, y0 + ph)
    if xb > xa and yb > ya:
        ch = np.nan_to_num(sv_fpr[:, xa:xb, ya:yb].swapaxes(1, 2), nan=0)
        lb = np.nan_to_num(labels_pr[xa:xb, ya:yb].T, nan=LABEL_BOUNDARY_VAL)
        out_data[:, ya - y0:yb - y0, xa - x0:xb - x0] = ch
        out_labels[ya - y0:yb - y0, xa - x0:xb - x0] = lb
    return out_data, out_labels


def refine_label_boundary(data, labels, thr_idx, thr=(1e-7, 1e-4)):
    """refine_label_boundary.py:38-104 with ignore_zero_inside_bbox=True."""
    new = labels.copy()
    idx = np.argwhere(new != LABEL_BOUNDARY_VAL)
    if len(idx) == 0:
        return new
    y0, y1 = idx[:, 0].min(), idx[:, 0].max() + 1
    x0, x1 = idx[:, 1].min(), idx[:, 1].max() + 1
    m = (labels > 0) & (data[thr_idx] > thr[0]) & (data[thr_idx] < thr[1])
    closed = binary_closing(m[y0:y1, x0:x1])
    mask = np.zeros(new.shape, dtype=bool)
    mask[y0:y1, x0:x1] = (~closed) & (new[y0:y1, x0:x1] > 0)
    new[mask] = LABEL_REFINE_BOUNDARY_VAL
    new[labels == LABEL_IGNORE_VAL] = LABEL_IGNORE_VAL
    return new


def convert_label_indexing(labels):
    """convert_label_indexing.py:24-35."""
    new = np.full(labels.shape, float(LABEL_IGNORE_VAL))
    new[labels == 0] = BACKGROUND
    new[labels == 27] = SANDEEL
    new[labels == 1] = OTHER
    return new


def train_patch_item(sv_fpr, labels_pr, centre, noise_on, flip, mult=None, patch_hw=(256, 256), thr_idx=None,
                     thr=(1e-7, 1e-4), scaled=False, border_zero=False):
    """Dataset.__getitem__ (dataset.py:75-108) with train.py's composition (batch/transforms.py:40-75): get_crop_zarr
    -> add_noise (multiplier field `mult` (F, ph, pw) in place of numpy's global RNG, add_noise.py:28-38) ->
    flip_x_axis -> refine_label_boundary -> convert_label_indexing -> remove_nan_inf -> db_with_limits[_scaled]
    [-> set_data_border_value].  Returns (data float32 (F, ph, pw), labels int64 (ph, pw))."""
    data, labels = get_crop_zarr(sv_fpr, labels_pr, centre, patch_hw)
    if noise_on:
        data = data * np.asarray(mult, dtype=np.float64)
    if flip:
        data = np.flip(data, 2).copy()
        labels = np.flip(labels, 1).copy()
    thr_idx = data.shape[0] - 1 if thr_idx is None else thr_idx
    labels = refine_label_boundary(data, labels, thr_idx, thr)
    labels = convert_label_indexing(labels)
    labels[~np.isfinite(data[0])] = LABEL_IGNORE_VAL           # remove_nan_inf.py:32 (a no-op after nan_to_num)
    data[~np.isfinite(data)] = 0.0
    with np.errstate(invalid="ignore"):
        db = 10 * np.log10(data + 1e-10)
        db[db > 0] = 0
        db[db < -75] = -75
    if scaled:
        db = 1 + db / 75.0                                     # db_with_limits.py:27-33
    if border_zero:
        db[:, labels == LABEL_BOUNDARY_VAL] = 0.0              # set_data_border_value.py:21-24
    return db.astype(np.float32), labels.astype(np.int64)


def noise_multiplier_field(shape, rng):
    """add_noise.py:28-38's multiplier as a field (float32-rounded so that a device replay sees the same numbers)."""
    change = rng.binomial(1, 0.05, shape)
    inc = rng.binomial(1, 0.5, shape)
    m = (1 - change) + change * (inc * rng.uniform(1, 10, shape) + (1 - inc) * rng.uniform(0, 1, shape))
    return m.astype(np.float32)


# ---- metadata input channels: get_crop_memmap, batch/dataset.py:296-349 ----------------------
META_ORDER = ("portion_year", "portion_day", "time_diff", "depth_rel", "depth_abs_surface", "depth_abs_seabed")


def meta_channels(centre, window, n_range, meta_cfg, portion_year, portion_of_day, time_diff, seabed):
    """The `meta` array of get_crop_memmap (dataset.py:296-349) for one crop: float64 (M, ph, pw), channels appended in
    the reference's order (portion_year; sin, cos of portion_day; time_diff; depth_rel; depth_abs_surface;
    depth_abs_seabed).  A crop of an echogram with n_range <= ph is re-centred vertically first (:261-262)."""
    ph, pw = window
    cy, cx = int(centre[0]), int(centre[1])
    if n_range <= ph:
        cy = n_range // 2

    def clamp(idx, size):
        idx = np.asarray(idx).copy()
        idx[idx < 0] = 0
        idx[idx >= size] = size - 1          # the reference writes -1 = the last element
        return idx

    out = []
    if meta_cfg.get("portion_year"):
        out.append(np.full((ph, pw), float(portion_year)))
    if meta_cfg.get("portion_day"):
        t = portion_of_day[int(clamp(np.array([cx]), portion_of_day.size)[0])]
        out.append(np.full((ph, pw), np.sin(2 * np.pi * t)))
        out.append(np.full((ph, pw), np.cos(2 * np.pi * t)))
    cols = np.arange(cx - pw // 2, cx + pw // 2)
    rows = np.arange(cy - ph // 2, cy + ph // 2)
    if meta_cfg.get("time_diff"):
        out.append(time_diff[clamp(cols, time_diff.size)].reshape(1, -1) * np.ones((ph, 1)))
    sb = seabed[clamp(cols, seabed.size)].reshape(1, -1)
    if meta_cfg.get("depth_rel"):
        out.append(rows.reshape(-1, 1) / sb)
    if meta_cfg.get("depth_abs_surface"):
        out.append(rows.reshape(-1, 1) * np.ones((1, pw)) / ph)
    if meta_cfg.get("depth_abs_seabed"):
        out.append((sb - rows.reshape(-1, 1)) / ph)
    return np.stack(out, 0)


def iter_meta_golden(g):
    """Cases of tests/golden/meta_channels.npz (oracle/make_golden_meta.py): yields (tag, case index, mask tag, centre,
    window, n_range, meta config dict, per-ping vectors dict, the reference's `meta` array)."""
    for tag in ("deep", "shallow"):
        vec = dict(portion_year=float(g[f"{tag}/portion_year"]), portion_of_day=g[f"{tag}/portion_of_day"],
                   time_diff=g[f"{tag}/time_diff"], seabed=g[f"{tag}/seabed"])
        n_range = int(g[f"{tag}/shape"][0])
        window = tuple(int(v) for v in g[f"{tag}/window"])
        for ci, c in enumerate(g[f"{tag}/centres"]):
            for mtag in ("all", "some"):
                cfg = {str(k): True for k in g[f"mask_{mtag}"]}
                yield tag, ci, mtag, c, window, n_range, cfg, vec, g[f"{tag}/{ci}/{mtag}"]

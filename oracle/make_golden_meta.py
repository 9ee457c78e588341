"""Generates tests/golden/meta_channels.npz by RUNNING THE UNMODIFIED REFERENCE get_crop_memmap
(/root/reference/crimac_unet/batch/dataset.py:254-355, build container only) on a fake in-memory echogram object with
every metadata channel switched on, for crops inside the data, overlapping every edge, and an echogram shallower than the
window (the re-centring branch, :261-262).  Stored: the echogram's per-ping vectors, the crop centres and the reference's
`meta` arrays (float64 -> what the model sees after .float()).  Usage: python oracle/make_golden_meta.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden", "meta_channels.npz")


class FakeEchogram:
    """The attributes get_crop_memmap touches (batch/dataset.py:254-349)."""
    data_format = "memmap"

    def __init__(self, n_range, n_pings, rng):
        self.shape = (n_range, n_pings)
        self._data = [rng.uniform(1e-9, 1e-3, size=(n_range, n_pings)).astype(np.float32) for _ in range(2)]
        self._labels = np.zeros((n_range, n_pings), dtype=np.int16)
        self.portion_of_year_scalar = 0.4321
        self.portion_of_day_vector = rng.uniform(0, 1, size=n_pings)
        self.time_vector_diff = rng.uniform(0.1, 3.0, size=n_pings)
        self._seabed = rng.integers(n_range // 2, n_range, size=n_pings).astype(np.float64)

    def data_memmaps(self, f):
        return [self._data[f]]

    def label_memmap(self):
        return self._labels


def main():
    from oracle import make_golden as MG
    MG._stub_missing_modules()
    if MG.REF not in sys.path:
        sys.path.insert(0, MG.REF)
    from batch.dataset import get_crop_memmap
    rng = np.random.default_rng(3)
    meta_all = {"portion_year": True, "portion_day": True, "depth_rel": True, "depth_abs_surface": True,
                "depth_abs_seabed": True, "time_diff": True}
    meta_some = {"portion_year": False, "portion_day": True, "depth_rel": False, "depth_abs_surface": True,
                 "depth_abs_seabed": False, "time_diff": True}
    out = {}
    cases = []
    for tag, (n_range, n_pings, win) in {"deep": (300, 500, (64, 64)), "shallow": (40, 200, (64, 64))}.items():
        eg = FakeEchogram(n_range, n_pings, rng)
        out[f"{tag}/portion_of_day"], out[f"{tag}/time_diff"], out[f"{tag}/seabed"] = eg.portion_of_day_vector, eg.time_vector_diff, eg._seabed
        out[f"{tag}/portion_year"], out[f"{tag}/shape"] = eg.portion_of_year_scalar, np.array(eg.shape)
        centres = [[150, 250], [10, 5], [290, 495], [31, 0], [200, 499], [-20, -40], [320, 530]] if tag == "deep" else [[20, 100], [5, 3], [39, 199]]
        for ci, c in enumerate(centres):
            for mtag, mc in (("all", meta_all), ("some", meta_some)):
                centre = np.array(c)            # get_crop_memmap may re-centre it in place
                _, meta, _ = get_crop_memmap(eg, centre, np.array(win), [0, 1], mc)
                out[f"{tag}/{ci}/{mtag}"] = np.asarray(meta)
        out[f"{tag}/centres"] = np.array(centres, dtype=np.int32)
        out[f"{tag}/window"] = np.array(win)
    out["mask_all"] = np.array([k for k, v in meta_all.items() if v])
    out["mask_some"] = np.array([k for k, v in meta_some.items() if v])
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items() if k.endswith("/all")})


if __name__ == "__main__":
    main()

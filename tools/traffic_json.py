#!/usr/bin/env python
"""profiles/conv_igemm_traffic.json (what bench.py reports as roofline.traffic) from the per-launch ncu summaries:
  python tools/traffic_json.py profiles/r02_conv_igemm_train_ncu.csv profiles/r02_conv_igemm_infer_ncu.csv"""
import csv
import json
import sys


def fam(path):
    rows = [r for r in csv.DictReader(l for l in open(path) if not l.startswith("#")) if r["kernel"].startswith("conv_igemm_kernel")]
    rd = sum(float(r["dram_read_MB"]) for r in rows)
    wr = sum(float(r["dram_write_MB"]) for r in rows)
    return len(rows), rd, wr, sum(float(r["time_us"]) for r in rows)


def main(train_csv, infer_csv, out="profiles/conv_igemm_traffic.json"):
    nt, rt, wt, tt = fam(train_csv)
    ni, ri, wi, ti = fam(infer_csv)
    json.dump({"train": (rt + wt) * 1e6, "infer": (ri + wi) * 1e6,
               "note": "sum of dram__bytes_read.sum + dram__bytes_write.sum over the conv_igemm_kernel launches of ONE step, batch 32 of "
                       f"4x256x256, ncu --set full --clock-control none; per-launch rows in {train_csv} / {infer_csv}",
               "train_launches": nt, "infer_launches": ni, "train_MB": {"read": rt, "write": wt},
               "infer_MB": {"read": ri, "write": wi}, "train_sum_time_us_under_ncu": tt, "infer_sum_time_us_under_ncu": ti},
              open(out, "w"), indent=1)
    print(open(out).read())


if __name__ == "__main__":
    main(*sys.argv[1:])

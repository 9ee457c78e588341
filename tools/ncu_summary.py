#!/usr/bin/env python
"""Turns an `ncu --set full` report (or a `--metrics gpu__time_duration.sum --csv` launch list) brought back in
gpurun_out/ into the compact, committed summaries under profiles/.

  python tools/ncu_summary.py report gpurun_out/prof_conv.ncu-rep profiles/r01_conv_igemm_ncu.csv
  python tools/ncu_summary.py launches gpurun_out/launches_train.csv profiles/r01_launches_train.csv
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

METRICS = OrderedDict([
    ("gpu__time_duration.sum", "time_us"),
    ("dram__bytes_read.sum", "dram_read_MB"),
    ("dram__bytes_write.sum", "dram_write_MB"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor_active_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_tc_pct"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_lsu_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("launch__registers_per_thread", "regs"),
    ("sm__cycles_elapsed.max", "sm_cycles"),
])


def to_mb(value, unit):
    v = float(value.replace(",", ""))
    scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit)
    return v * scale if scale else v


def to_us(value, unit):
    v = float(value.replace(",", ""))
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}.get(unit)
    return v * scale if scale else v


def report(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none, from {rep.split('/')[-1]} (cold-cache, serialised replays)\n")
        f.write("kernel,grid,block," + ",".join(METRICS.values()) + "\n")
        for d in data:
            name = d[col["Kernel Name"]].replace("void <unnamed>::", "").split("(")[0]
            vals = []
            for m, short in METRICS.items():
                if m not in col:
                    vals.append("")
                    continue
                v, u = d[col[m]], units[col[m]]
                if short.endswith("_MB"):
                    vals.append(f"{to_mb(v, u):.3f}")
                elif short == "time_us":
                    vals.append(f"{to_us(v, u):.2f}")
                else:
                    vals.append(v.replace(",", ""))
            grid = d[col["Grid Size"]].replace(",", "x").replace(" ", "")
            block = d[col["Block Size"]].replace(",", "x").replace(" ", "")
            f.write(f"\"{name}\",\"{grid}\",\"{block}\"," + ",".join(vals) + "\n")
    print("wrote", out, len(data), "launches")


def launches(src, out):
    lines = [ln for ln in open(src) if not ln.startswith("==")]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    fam = OrderedDict()
    total = 0.0
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"].replace("void <unnamed>::", "").split("(")[0]
        us = to_us(r["Metric Value"], r["Metric Unit"])
        a = fam.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
        total += us
    with open(out, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none launch list from {src.split('/')[-1]}: "
                f"per-kernel totals (cold-cache, serialised: compare SHARES). total {total:.1f} us\n")
        f.write("kernel,launches,total_us,share\n")
        for k, (n, us) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k}\",{n},{us:.1f},{us / total:.4f}\n")
    print("wrote", out, len(fam), "kernels")


if __name__ == "__main__":
    {"report": report, "launches": launches}[sys.argv[1]](sys.argv[2], sys.argv[3])

#!/usr/bin/env python
"""Hottest SASS instructions (warp-stall samples) of one launch in an ncu report:
   python tools/ncu_hot.py gpurun_out/prof.ncu-rep <launch index> [N]"""
import csv, io, subprocess, sys
rep, idx = sys.argv[1], int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(idx), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
print(rows[0][:2])
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr) and r[hdr.index('# Samples')].isdigit()]
ci = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ci['# Samples']]) for r in data)
print('total samples', tot, 'instructions', len(data))
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {h: sum(int(r[ci[h]]) for r in data) for h in stall_cols}
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for r in sorted(data, key=lambda r: -int(r[ci['# Samples']]))[:n]:
    st = {h[6:]: int(r[ci[h]]) for h in stall_cols if int(r[ci[h]]) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(r[ci['# Samples']].rjust(6), r[ci['Instructions Executed']].rjust(9), r[ci['Source']].strip()[:70].ljust(70), st)

"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line.

    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:K --launch-count 1 > src.csv
    python profiles/ncu_lines.py src.csv [top]
"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur_file = None
hdr = None
agg = defaultdict(lambda: defaultdict(float))
src_text = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or r[0] in ("Function Name",):
        continue
    d = dict(zip(range(len(r)), r))
    try:
        line = int(r[0])
    except ValueError:
        continue
    key = (cur_file, line)
    src_text[key] = r[1].strip()
    for i, h in enumerate(hdr):
        if i < 4 or i >= len(r):
            continue
        try:
            agg[key][h] += float(r[i])
        except ValueError:
            pass
tot = sum(v["# Samples"] for v in agg.values()) or 1
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print(f"total samples {tot:.0f}")
for key, v in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
    s = v["# Samples"]
    main = sorted(((v[h], h) for h in stalls), reverse=True)[:3]
    ms = " ".join(f"{h[6:]}={100 * x / max(s, 1):.0f}%" for x, h in main if x > 0)
    print(f"{100 * s / tot:5.1f}% {key[0]}:{key[1]:<4d} inst={v['Instructions Executed']:.3g} [{ms}] {src_text[key][:90]}")

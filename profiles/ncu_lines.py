"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line and per kernel phase.

    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:K --launch-count 1 > src.csv
    python profiles/ncu_lines.py src.csv [top] [phases]

ncu does not escape quotes inside the source-text column (a line such as `: "d"(a), "d"(b));` shifts the columns of a naive CSV
parse -- round 1's summary attributed the DMMA line's `math` stalls to `membar` that way).  The numeric columns are therefore
addressed from the END of each row.  With a third argument "phases" the samples are also summed per phase of deepz_kernel
(line ranges of deepz.cuh between its `//@phase:` markers).
"""
import sys
from collections import defaultdict

import os
import re


def load_phases():
    """Phase of every line of deepz.cuh, from its `//@phase: <name>` markers (a phase runs up to the next marker)."""
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "closedloopreachability.jl_b200", "csrc", "deepz.cuh")
    out, cur, start = [], None, 1
    lines = open(path).read().split("\n")
    for no, ln in enumerate(lines, 1):
        m = re.match(r"\s*//@phase:\s*(.+)$", ln)
        if m:
            if cur:
                out.append((cur, start, no - 1))
            cur, start = m.group(1).strip(), no
    if cur:
        out.append((cur, start, len(lines)))
    return out


PHASES = load_phases()


def split_row(line):
    line = line.rstrip("\n")
    if not line.startswith('"'):
        return None
    return line[1:-1].split('","')


rows = [split_row(l) for l in open(sys.argv[1], errors="replace")]
rows = [r for r in rows if r]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur_file, hdr = None, None
agg = defaultdict(lambda: defaultdict(float))
src_text = {}
for r in rows:
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or r[0] == "Function Name":
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    key = (cur_file, line)
    shift = len(r) - len(hdr)            # extra fields produced by quotes / commas in the source text
    src_text[key] = '","'.join(r[1:2 + shift]).strip()
    for i, h in enumerate(hdr):
        if i < 4:
            continue
        try:
            agg[key][h] += float(r[i + shift])
        except (ValueError, IndexError):
            pass
tot = sum(v["# Samples"] for v in agg.values()) or 1
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print(f"total samples {tot:.0f}")
for key, v in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
    s = v["# Samples"]
    main = sorted(((v[h], h) for h in stalls), reverse=True)[:3]
    ms = " ".join(f"{h[6:]}={100 * x / max(s, 1):.0f}%" for x, h in main if x > 0)
    print(f"{100 * s / tot:5.1f}% {key[0]}:{key[1]:<4d} inst={v['Instructions Executed']:.3g} [{ms}] {src_text[key][:90]}")
if len(sys.argv) > 3 and sys.argv[3] == "phases":
    ph = defaultdict(lambda: defaultdict(float))
    for (f, ln), v in agg.items():
        name = "other files (intrinsics)"
        if f == "deepz.cuh":
            name = next((n for n, a, b in PHASES if a <= ln <= b), "deepz.cuh (other)")
        for h in ["# Samples", "Instructions Executed"] + stalls:
            ph[name][h] += v[h]
    ti = sum(v["Instructions Executed"] for v in ph.values()) or 1
    print("\nphase: share of warp samples | share of executed warp instructions | top stall reasons")
    for name, v in sorted(ph.items(), key=lambda kv: -kv[1]["# Samples"]):
        s = v["# Samples"]
        main = sorted(((v[h], h) for h in stalls), reverse=True)[:4]
        ms = " ".join(f"{h[6:]}={100 * x / max(s, 1):.0f}%" for x, h in main if x > 0)
        print(f"{100 * s / tot:5.1f}% | {100 * v['Instructions Executed'] / ti:5.1f}% | {name:28s} [{ms}]")

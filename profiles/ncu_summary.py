"""Summarise one launch of an `ncu --set full` report as JSON (the numbers DESIGN.md / bench.py quote).

    python profiles/ncu_summary.py X.ncu-rep <kernel-regex> "<source note>" [pick=longest|first] [name-substring] > profiles/rNN/ncu_<name>.json
"""
import csv, io, json, subprocess, sys

rep, kre, note = sys.argv[1], sys.argv[2], sys.argv[3]
pick = sys.argv[4] if len(sys.argv) > 4 else "longest"
must = sys.argv[5] if len(sys.argv) > 5 else ""      # substring the demangled kernel name has to contain
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, body = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
body = [r for r in body if must in r[col["Kernel Name"]]]


def f(r, k):
    try:
        return float(r[col[k]].replace(",", ""))
    except Exception:
        return None


def in_unit(r, k, want):
    """value converted to `want` ('byte' | 'ms')"""
    v, u = f(r, k), units[col[k]]
    if v is None:
        return None
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
    return v * scale.get(u, 1.0)


r = max(body, key=lambda x: in_unit(x, "gpu__time_duration.sum", "ms") or 0) if pick == "longest" else body[0]
rd, wr = in_unit(r, "dram__bytes_read.sum", "byte"), in_unit(r, "dram__bytes_write.sum", "byte")
out = {
    "kernel": r[col["Kernel Name"]],
    "grid": r[col["Grid Size"]], "block": r[col["Block Size"]],
    "dram_bytes_read": rd, "dram_bytes_write": wr, "traffic": (rd or 0) + (wr or 0),
    "gpu_time_ms": in_unit(r, "gpu__time_duration.sum", "ms"),
    "dmma_pipe_active_pct": f(r, "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active"),
    "fp64_pipe_pct": f(r, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    "issue_active_pct": f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "registers": f(r, "launch__registers_per_thread"),
    "smem_wavefronts_pct": f(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
    "warps_active_pct": f(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
    "dram_throughput_pct": f(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    "launches_in_report": len(body),
    "source": note,
}
print(json.dumps(out, indent=1))

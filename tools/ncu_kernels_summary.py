#!/usr/bin/env python
"""Per-kernel summary of an `ncu --set full` report (runs in the build container, `ncu -i` needs no GPU):
for the LAST captured launch of every kernel name: duration, DRAM bytes and throughput, pipe / issue
utilisation, occupancy and the warp-stall mix.   tools/ncu_kernels_summary.py REP.ncu-rep [bytes.json]"""
import csv, io, json, subprocess, sys

rep = sys.argv[1]
raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True,
                                                 capture_output=True, text=True).stdout)))
hdr, units, rows = raw[0], raw[1], raw[2:]
col = {h: k for k, h in enumerate(hdr)}
last = {}
for r in rows:
    last[r[col["Kernel Name"]]] = r


def val(r, m):
    if m not in col:
        return None
    try:
        v = float(r[col[m]].replace(",", ""))
    except ValueError:
        return None
    u = units[col[m]]
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ns": 1e-3, "us": 1.0, "ms": 1e3, "msecond": 1e3,
             "usecond": 1.0, "nsecond": 1e-3}.get(u, 1.0)
    return v * scale


STALLS = [h for h in hdr if h.startswith("smsp__average_warp_latency_issue_stalled") or
          h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
for name, r in last.items():
    dur = val(r, "gpu__time_duration.sum")
    rd, wr = val(r, "dram__bytes_read.sum") or 0.0, val(r, "dram__bytes_write.sum") or 0.0
    print(f"== {name[:140]}")
    print(f"   duration {dur:.2f} us   DRAM read {rd / 1e6:.2f} MB  write {wr / 1e6:.2f} MB  -> {(rd + wr) / dur / 1e3 if dur else 0:.0f} GB/s"
          f"   grid {r[col['launch__grid_size']]} x {r[col['launch__block_size']]}  regs {r[col['launch__registers_per_thread']]}")
    for m in ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
              "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
              "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
              "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum"):
        if m in col:
            print(f"   {m} = {r[col[m]]} {units[col[m]]}")
    st = []
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
            v = val(r, h)
            if v:
                st.append((v, h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
    tot = sum(v for v, _ in st) or 1.0
    print("   stall mix: " + "  ".join(f"{n}={100 * v / tot:.1f}%" for v, n in sorted(st, reverse=True)[:7]))

#!/usr/bin/env python
"""Turn the raw ncu outputs of tools/profile_round.sh into the tracked summaries under profiles/:
  launches_<tag>.csv      -> profiles/<tag>_launches_ncu.csv (copy) + <tag>_launch_shares.txt
  prof_dyn_<tag>.ncu-rep  -> profiles/<tag>_dyn_ncu_full_metrics.txt (selected raw metrics + top stall
                             instructions) and profiles/traffic.json (DRAM bytes per launch, read by bench.py)
Runs in the build container (`ncu -i` needs no GPU)."""
import collections, csv, io, json, os, re, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01b"
G = os.path.join(ROOT, "gpurun_out")
P = os.environ.get("NCU_SUMMARY_OUT") or os.path.join(ROOT, "profiles")   # on the GPU box: a directory under gpurun_out/
os.makedirs(P, exist_ok=True)


def launch_shares():
    src = os.path.join(G, f"launches_{tag}.csv")
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    iu = hdr.index("Metric Unit")
    acc = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[iv].replace(",", ""))
        v = v / 1e3 if r[iu] in ("ns", "nsecond") else (v * 1e3 if r[iu] in ("ms", "msecond") else v)   # -> us
        a = acc.setdefault(r[ik], [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in acc.values())
    shutil.copy(src, os.path.join(P, f"{tag}_launches_ncu.csv"))
    with open(os.path.join(P, f"{tag}_launch_shares.txt"), "w") as f:
        f.write(f"# ncu launch list summary: first 400 launches of `python bench.py --steps 20 --warmup 3 --e2e-chunks 1 "
                f"--no-cpu-baseline` ({tag} build)\n# cold-cache serialised per-launch times: compare SHARES, not absolutes\n"
                "share%  launches  avg_us  kernel\n")
        for k, (n, t) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{100 * t / tot:6.2f}  {n:5d}  {t / n:10.2f}  {k[:150]}\n")


def ncu(args):
    return subprocess.run(["ncu", "-i", os.path.join(G, f"prof_dyn_{tag}.ncu-rep")] + args, check=True,
                          capture_output=True, text=True).stdout


def full_metrics():
    raw = list(csv.reader(io.StringIO(ncu(["--page", "raw", "--csv"]))))
    hdr, units, rows = raw[0], raw[1], raw[2:]
    want = ["gpu__time_duration.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_fp64_cycles_active.max.pct_of_peak_sustained_elapsed",
            "sm__pipe_fp64_cycles_active.min.pct_of_peak_sustained_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed_op_shared_ld.sum"]
    out = []
    last = rows[-1]
    name = last[hdr.index("Kernel Name")]
    out.append(f"{name}   (launch id {last[hdr.index('ID')]}, ncu --set full --clock-control none)")
    vals = {}
    for m in want:
        if m in hdr:
            i = hdr.index(m); vals[m] = (last[i], units[i]); out.append(f"   {m} {last[i]} {units[i]}")

    def to_bytes(m):
        v, u = vals[m]; v = float(v.replace(",", ""))
        return int(v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u])
    tj = {"workload": {"kind": "lorenz_rk4", "envs": 65536, "chunk": 256, "substeps": 16}, "kernel": name,
          "dram_bytes_read": to_bytes("dram__bytes_read.sum"), "dram_bytes_write": to_bytes("dram__bytes_write.sum"),
          "source": f"ncu --set full --clock-control none -k regex:k_rollout_sm ({tag} build, tools/r02_prof.sh); summary: "
                    f"profiles/{tag}_dyn_ncu_full_metrics.txt"}
    json.dump(tj, open(os.path.join(P, "traffic.json"), "w"), indent=1)
    # per-instruction sampling: top stall sites
    try:
        src = list(csv.reader(io.StringIO(ncu(["--page", "source", "--csv", "--print-source", "sass"]))))
        # several launches are concatenated: keep the last one ("Kernel Name" line, then the header)
        starts = [k for k, r in enumerate(src) if r and r[0] == "Kernel Name"]
        src = src[starts[-1] + 1:]
        h = src[0]
        isrc = h.index("Source"); ismp = h.index("# Samples") if "# Samples" in h else None
        iex = next((h.index(c) for c in h if c.startswith("Instructions Executed")), None)
        stall_cols = [(c, k) for k, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
        if ismp is not None:
            body = [r for r in src[1:] if len(r) == len(h)]
            tot = sum(int(r[ismp] or 0) for r in body)
            out.append(f"\nper-instruction samples: {len(body)} instrs, {tot} samples; top 12 sites")
            for r in sorted(body, key=lambda r: -int(r[ismp] or 0))[:12]:
                st = {c.replace('stall_', ''): int(r[k]) for c, k in stall_cols if r[k] not in ("", "0") and r[k].isdigit()}
                top = dict(sorted(st.items(), key=lambda kv: -kv[1])[:2])
                out.append(f"  {r[isrc].strip()[:52]:52s} smp={int(r[ismp] or 0):6d} exec={r[iex] if iex is not None else '?'} {top}")
            agg = collections.Counter()
            for r in body:
                for c, k in stall_cols:
                    if r[k].isdigit(): agg[c.replace('stall_', '')] += int(r[k])
            s = sum(agg.values()) or 1
            out.insert(1, "stall mix: " + "  ".join(f"{k}={100 * v / s:.1f}%" for k, v in agg.most_common(9)))
    except Exception as e:  # noqa: BLE001
        out.append(f"(source page not summarised: {e})")
    open(os.path.join(P, f"{tag}_dyn_ncu_full_metrics.txt"), "w").write("\n".join(out) + "\n")
    print("\n".join(out[:24]))


if __name__ == "__main__":
    launch_shares()
    full_metrics()

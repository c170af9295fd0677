#!/usr/bin/env python
"""Summarise ncu outputs (run here, no GPU needed): launch list CSV -> per-kernel shares and the first step's
per-round times; .ncu-rep raw page -> the metrics bench.py's roofline block cites."""
import collections, csv, re, subprocess, sys

def launches(path, first=34):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    seq = []
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("scb::", "")
        t = float(r["Metric Value"].replace(",", ""))
        t = {"ns": t / 1e3, "us": t, "ms": t * 1e3, "s": t * 1e6}[r["Metric Unit"]]
        agg.setdefault(name, []).append(t)
        seq.append((name, t, r["Grid Size"]))
    tot = sum(sum(v) for v in agg.values())
    print(f"# {path}: {len(rows)} launches, {tot/1e3:.3f} ms total")
    print("| kernel | launches | total us | share | max us |\n|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"| `{k}` | {len(v)} | {sum(v):.1f} | {100*sum(v)/tot:.1f}% | {max(v):.1f} |")
    print("\nfirst launches in order (kernel, us, grid):")
    for s in seq[:first]:
        print(f"  {s[0][:70]:70s} {s[1]:10.1f} {s[2]}")

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum"]

def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"\n# {path}")
    for r in rows[2:]:
        print("## launch", r[hdr.index("ID")], re.sub(r"\(.*", "", r[hdr.index("Kernel Name")]))
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"  {w:62s} {r[i]:>18s} {units[i]}")
        stall = [h for h in hdr if "smsp__average_warp" in h and "issue_stalled" in h and h.endswith(".ratio")]
        vals = sorted([(float(r[hdr.index(h)].replace(",", "") or 0), h) for h in stall], reverse=True)[:4]
        for v, h in vals:
            print(f"  stall {v:8.2f}  {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}")

for a in sys.argv[1:]:
    (launches if a.endswith(".csv") else rep)(a)

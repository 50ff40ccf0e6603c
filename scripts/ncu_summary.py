"""Summarise ncu output into profiles/: python scripts/ncu_summary.py <launches.csv> <full.ncu-rep> <tag> [sessions]"""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
launch_csv, rep, tag = sys.argv[1], sys.argv[2], sys.argv[3]
S = int(sys.argv[4]) if len(sys.argv) > 4 else 8
out = []
rows = list(csv.DictReader(l for l in open(launch_csv) if l.startswith('"')))
agg = collections.defaultdict(list)
for r in rows:
    agg[r["Kernel Name"].split("(")[0]].append(float(r["Metric Value"]))
tot = sum(sum(v) for v in agg.values())
out.append(f"# {tag}: ncu --metrics gpu__time_duration.sum --clock-control none (C3 x {S} sessions; per-launch times are cold-cache and serialised: compare shares)")
out.append(f"{'kernel':30s} {'launches':>8s} {'mean_us':>9s} {'share':>6s}")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    out.append(f"{k:30s} {len(v):8d} {sum(v)/len(v)/1e3:9.2f} {sum(v)/tot:6.3f}")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h = rr[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum"]
out.append("")
out.append(f"# {tag}: ncu --set full --clock-control none --import-source on (one launch per kernel)")
traffic = {}
seen = set()
for r in rr[2:]:
    name = r[h.index("Kernel Name")].split("(")[0]
    if name in seen:
        continue
    seen.add(name)
    out.append(f"[{name}]")
    for w in want:
        if w in h:
            out.append(f"    {w:72s} {r[h.index(w)]:>16s} {rr[1][h.index(w)]}")
    if "syrk_tcgen05" in name:
        def val(m):
            v, u = float(r[h.index(m)]), rr[1][h.index(m)]
            return v * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[u]
        traffic[f"S{S}"] = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
path = os.path.join(ROOT, "profiles", f"{tag}_summary.txt")
open(path, "w").write("\n".join(out) + "\n")
if traffic:
    tp = os.path.join(ROOT, "profiles", "syrk_traffic.json")
    cur = json.load(open(tp)) if os.path.exists(tp) else {}
    cur.update(traffic)
    cur["source"] = f"{tag}: dram__bytes_read.sum + dram__bytes_write.sum of one k_syrk_tcgen05_i8 launch (ncu --set full)"
    json.dump(cur, open(tp, "w"), indent=1)
print(open(path).read())

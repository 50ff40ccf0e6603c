"""Top CUDA source lines by warp-stall samples from an ncu report (needs -lineinfo + --import-source on).
Usage: python scripts/ncu_lines.py report.ncu-rep kernel_regex [top]"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", f"regex:{kern}",
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, hdr, agg = None, None, {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif len(r) > 6 and r[0] == "Line No":
        hdr = r
        ns, ie = hdr.index("# Samples"), hdr.index("Instructions Executed")
    elif hdr and len(r) > 6 and r[2] == "-":          # a source-line aggregate row
        try:
            agg[(cur_file, int(r[0]))] = (int(r[ns] or 0), int(r[ie] or 0), r[1].strip())
        except ValueError:
            pass
tot = sum(v[0] for v in agg.values())
print(f"{kern}: {tot} samples")
for (f, ln), (smp, ex, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100 * smp / max(tot, 1):5.1f}%  {f}:{ln:<4d} exec {ex:>9d} | {src[:105]}")

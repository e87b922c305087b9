"""Turn an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv python bench.py ...`) into the
markdown table kept under profiles/: per kernel launches, total and average duration, share of all GPU time, and the
durations of the timed-region launches of the headline kernel (grid = one CTA per patch x chunk groups).
usage: python scripts/summarize_launches.py gpurun_out/launches.csv [headline kernel substring] > profiles/summary.md"""
import collections
import csv
import re
import sys

path = sys.argv[1]
headline = sys.argv[2] if len(sys.argv) > 2 else "patch_gather_kernel<2, 0, 1, 0, 0, 16>"
rows = list(csv.reader(open(path)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
col = {h: i for i, h in enumerate(rows[hdr])}
seq = []
for r in rows[hdr + 1:]:
    if len(r) <= col["Metric Value"]:
        continue
    try:
        ns = float(r[col["Metric Value"]].replace(",", ""))
    except ValueError:
        continue
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|^void ", "", r[col["Kernel Name"]])
    seq.append((name.split("(")[0] if "<" not in name.split("(")[0] else name[: name.index(">(") + 1] if ">(" in name else name,
                ns, r[col["Grid Size"]]))
per = collections.OrderedDict()
for name, ns, _ in seq:
    per.setdefault(name, []).append(ns)
total = sum(ns for _, ns, _ in seq)
big = [ns for name, ns, grid in seq if headline in name]
top_grid = collections.Counter(grid for name, _, grid in seq if headline in name).most_common()
print(f"launches: {len(seq)}, GPU time {total / 1e6:.3f} ms (serialised, cold cache)\n")
if big:
    # the headline workload uses the largest grid of that kernel
    grids = sorted({g for n, _, g in seq if headline in n}, key=lambda g: -int(re.sub(r"[^0-9,]", "", g).split(",")[0]))
    full = [ns for n, ns, g in seq if headline in n and g == grids[0]]
    full = [ns for ns in full if ns > 0.5 * max(full)]  # the full-batch launches (the host-buffer leg runs smaller chunks)
    print(f"`{headline}` on grid {grids[0]}, full batch (warm-up + timed steps): {len(full)} launches: " +
          ", ".join(f"{ns / 1e6:.3f}" for ns in full) + " ms\n")
print("| kernel | launches | total ms | avg us | share of all GPU time |")
print("|---|---|---|---|---|")
for name, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
    short = name if len(name) <= 80 else name[:77] + "..."
    print(f"| `{short}` | {len(v)} | {sum(v) / 1e6:.3f} | {sum(v) / len(v) / 1e3:.1f} | {100 * sum(v) / total:.1f} % |")

"""Per-kernel summary of one step from an ncu launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file L.csv python bench.py --steps 2 --warmup 3 ...
    python scripts/launch_summary.py L.csv > profiles/<round>_launch_summary.txt
The LAST complete step in the list (k_front_count .. k_alpha_bc) is summarised; ncu times are serialised and cold-cache."""
import csv, re, sys
from collections import OrderedDict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("nsecond", "ns"):
            v /= 1e3
        elif r["Metric Unit"] in ("msecond", "ms"):
            v *= 1e3
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("svof::", "")
        rows.append((name, v))
starts = [i for i, (n, _) in enumerate(rows) if n.startswith("k_front_count") or n.startswith("k_clear_prev")]
ends = [i for i, (n, _) in enumerate(rows) if n.startswith("k_alpha_bc")]
ends = [e for e in ends if any(s < e for s in starts)]
step = None
for e in reversed(ends):   # the last device-resident step (the end-to-end steps add the sparse-upload kernels)
    s = max(x for x in starts if x < e)
    step = rows[s:e + 1]
    if not any(n.startswith("k_mark_u_cells") for n, _ in step):
        break
agg = OrderedDict()
for n, v in step:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v for _, v in step)
print("step launches %d total %.1f us (serialised, cold cache)" % (len(step), tot))
for n, (c, v) in agg.items():
    print("  %-44s x%-3d %8.1f us %5.1f%%" % (n[:44], c, v, 100 * v / tot))

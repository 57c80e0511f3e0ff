"""Per-kernel aggregation of an `ncu --metrics gpu__time_duration.sum --csv` launch list.

  python scripts/launch_summary.py launches.csv "header comment" [steps_in_list]

Prints count, total and mean duration per kernel name (template arguments kept, parameter lists
dropped).  Per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes."""
import csv
import sys
from collections import defaultdict


def main(path, header="", skip_setup=True):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    h = rows[0]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = defaultdict(lambda: [0, 0.0])
    scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}
    for r in rows[1:]:
        name = r[ki].split("(")[0].replace("expb200::", "").replace("<unnamed>::", "")[-70:]
        try:
            us = float(r[vi].replace(",", "")) * scale.get(r[ui], 1e-3)
        except ValueError:
            continue
        agg[name][0] += 1
        agg[name][1] += us
    tot = sum(v[1] for v in agg.values()) or 1.0
    if header:
        print("# " + header)
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name:72s} n={n:5d} total={us:12.1f} us avg={us / n:10.1f} us share={us / tot:6.3f}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")

"""Top source lines of one kernel by warp-stall samples, from an .ncu-rep captured with
--set full --import-source on (kernels compiled with -lineinfo).

  python scripts/ncu_source_hot.py report.ncu-rep [top]

Prints, per CUDA source line, the sampled stall count and the executed warp instructions, so
that a latency chain or a hot loop can be attributed to code (B200_PROFILING.md, source page)."""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def main(path, top=40, view="cuda,sass"):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", view],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    agg = defaultdict(lambda: [0.0, 0.0, defaultdict(float)])
    for r in rows:
        if hdr is None:
            if "Source" in r and any("Sampling" in c for c in r):
                hdr = r
                si = r.index("Source")
                samp = [i for i, c in enumerate(r) if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)"]
                ex = [i for i, c in enumerate(r) if c == "Instructions Executed"]
                st = [(i, c) for i, c in enumerate(r) if c.startswith("stall_") and "Not Issued" not in c]
            continue
        if len(r) != len(hdr):
            continue
        try:
            s = float(r[samp[0]] or 0)
            e = float(r[ex[0]] or 0) if ex else 0.0
        except ValueError:
            continue
        key = r[si].strip()[:110]
        if view == "sass":
            key = r[0] + " " + key          # one row per instruction: keep the address
        elif not key:
            continue            # the SASS rows of the mixed view repeat the same samples
        agg[key][0] += s
        agg[key][1] += e
        for i, c in st:
            try:
                agg[key][2][c[6:]] += float(r[i] or 0)
            except ValueError:
                pass
    tot = sum(v[0] for v in agg.values()) or 1.0
    print(f"# {path}: {int(tot)} stall samples")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        why = " ".join(f"{n}:{int(c / max(v[0], 1) * 100)}" for n, c in sorted(v[2].items(), key=lambda kv: -kv[1])[:3])
        print(f"{v[0] / tot * 100:6.2f}%  inst={int(v[1]):>12d}  [{why}]  {k}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40, sys.argv[3] if len(sys.argv) > 3 else "cuda,sass")

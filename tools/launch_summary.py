#!/usr/bin/env python
"""Per-kernel shares of an `ncu --metrics gpu__time_duration.sum --csv` launch list -> JSON summary.

    python tools/launch_summary.py gpurun_out/r02_launches_train.csv [out.json] [--note "..."]
"""
import csv
import json
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    out = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else None
    note = sys.argv[sys.argv.index("--note") + 1] if "--note" in sys.argv else ""
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rd:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        unit = r[iu]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        name = re.sub(r"\(.*$", "", r[ik])
        name = re.sub(r"^void\s+", "", name).replace("asep::", "").replace("(anonymous namespace)::", "").replace("unnamed>::", "")
        rows.append((name, us))
    tot = sum(u for _, u in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for n, u in rows:
        agg[n][0] += 1
        agg[n][1] += u
    table = sorted(((n, c, t) for n, (c, t) in agg.items()), key=lambda x: -x[2])
    js = {"source": path, "note": note, "launches": len(rows), "total_us": round(tot, 1),
          "kernels": [{"kernel": n, "launches": c, "total_us": round(t, 1), "avg_us": round(t / c, 2), "share": round(t / tot, 4)} for n, c, t in table]}
    if out:
        json.dump(js, open(out, "w"), indent=1)
    print(f"{path}: {len(rows)} launches, {tot / 1e3:.3f} ms (cold-cache, serialised)")
    for n, c, t in table[:18]:
        print(f"  {t / tot:6.1%} {t / 1e3:9.3f} ms {c:5d} x {t / c:8.2f} us  {n[:110]}")


if __name__ == "__main__":
    main()

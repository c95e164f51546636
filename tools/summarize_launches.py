#!/usr/bin/env python
"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total / mean / min / max ms.
usage: summarize_launches.py launches.csv > summary.json   (times under ncu are serialised and cold-cache: compare shares)"""
import collections
import csv
import json
import sys


def summarize(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki].split("(")[0].replace("void ", "")
        agg.setdefault(name, []).append(float(r[vi].replace(",", "")) * scale.get(r[ui], 1e-6))
    total = sum(sum(v) for v in agg.values())
    return {"source": path, "total_ms": round(total, 3),
            "kernels": {k: {"launches": len(v), "total_ms": round(sum(v), 3), "share": round(sum(v) / total, 4),
                            "mean_ms": round(sum(v) / len(v), 4), "min_ms": round(min(v), 4), "max_ms": round(max(v), 4)}
                        for k, v in agg.items()}}


if __name__ == "__main__":
    print(json.dumps(summarize(sys.argv[1]), indent=1))

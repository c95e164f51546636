#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` export of ONE kernel: warp-level instruction counts by opcode, stall samples by reason,
and the hottest SASS lines.  usage: ncu -i rep --page source --csv --kernel-name regex:NAME > src.csv; ncu_source_mix.py src.csv [elements]"""
import csv
import sys
from collections import Counter


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    elements = float(sys.argv[2]) if len(sys.argv) > 2 else None
    # several kernels may follow each other: take the first block
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    end = next((i for i in range(hdr_i + 1, len(rows)) if rows[i] and rows[i][0] == "Kernel Name"), len(rows))
    body = [r for r in rows[hdr_i + 1:end] if len(r) >= len(hdr) - 2]
    col = {n: i for i, n in enumerate(hdr)}
    stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    ops, samples, st = Counter(), 0, Counter()
    hot = []
    for r in body:
        src = r[col["Source"]].strip()
        op = src.split()[0] if src else "?"
        if op.startswith("@"):
            op = src.split()[1]
        op = op.split(".")[0]
        n = float(r[col["Instructions Executed"]] or 0)
        ops[op] += n
        s = float(r[col["# Samples"]] or 0)
        samples += s
        for k in stalls:
            st[k] += float(r[col[k]] or 0)
        hot.append((s, src, n))
    tot = sum(ops.values())
    print(f"kernel: {rows[0][1][:100]}")
    print(f"warp instructions executed: {tot:.4g}" + (f"  ({tot * 32 / elements:.1f} thread-instr per element)" if elements else ""))
    fp64 = sum(v for k, v in ops.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "MUFU", "DMNMX", "F2F", "I2F", "F2I"))
    print(f"FP64-pipe-ish (DFMA DMUL DADD DSETP): {sum(ops[k] for k in ('DFMA','DMUL','DADD','DSETP')):.4g}"
          + (f"  ({sum(ops[k] for k in ('DFMA','DMUL','DADD','DSETP')) * 32 / elements:.1f} per element)" if elements else ""))
    for k, v in ops.most_common(28):
        print(f"  {k:10s} {v:12.4g}  {100 * v / tot:5.1f} %" + (f"  {v * 32 / elements:6.2f}/elt" if elements else ""))
    print("stall samples by reason (all samples):")
    ts = sum(st.values())
    for k, v in st.most_common(12):
        print(f"  {k:24s} {100 * v / max(ts, 1):5.1f} %")
    print("hottest lines:")
    for s, src, n in sorted(hot, reverse=True)[:25]:
        print(f"  {100 * s / max(samples, 1):5.2f} %  {n:10.4g}  {src[:110]}")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Basic blocks of one kernel's SASS with their FP64 / local-memory / shared-memory instruction counts: where the spills sit.
usage: cuobjdump -sass obj | sass_blocks.py <kernel-name-substring>"""
import re
import sys


def main():
    want = sys.argv[1]
    cur, on, blocks = None, False, []
    for line in sys.stdin:
        if "Function :" in line:
            on = want in line
            continue
        if not on:
            continue
        m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if not m:
            continue
        addr, ins = m.group(1), m.group(2).strip()
        if cur is None:
            cur = {"start": addr, "ins": []}
        cur["ins"].append(ins)
        op = ins.split()[1] if ins.startswith("@") else ins.split()[0]
        if op.startswith(("BRA", "EXIT", "RET", "BSYNC", "CALL", "BRX", "WARPSYNC")) or op.startswith("BAR"):
            blocks.append(cur); cur = None
    if cur:
        blocks.append(cur)
    def cnt(b, pre):
        return sum(1 for i in b["ins"] if (i.split()[1] if i.startswith("@") else i.split()[0]).startswith(pre))
    print(f"{len(blocks)} blocks; those with >= 8 FP64 instructions:")
    for b in blocks:
        f = cnt(b, ("DFMA", "DMUL", "DADD"))
        if f >= 8:
            print(f"  @{b['start']}  n={len(b['ins']):4d}  FP64={f:3d}  LDL={cnt(b, 'LDL'):2d}  STL={cnt(b, 'STL'):2d}  LDS={cnt(b, 'LDS'):2d}  LDG={cnt(b, 'LDG'):2d}  IMAD={cnt(b, 'IMAD'):3d} MOV={cnt(b, 'MOV'):2d}")


if __name__ == "__main__":
    main()

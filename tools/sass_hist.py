#!/usr/bin/env python
"""Opcode histogram of a cuobjdump -sass listing, whole function and per backward-branch loop body."""
import re, sys, collections
ins = []
for ln in open(sys.argv[1]):
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        addr = int(m.group(1), 16); txt = m.group(2).strip()
        t = txt.split()
        op = t[1] if t[0].startswith("@") else t[0]
        ins.append((addr, op.split(".")[0], txt))
print("instructions:", len(ins))
def hist(lo, hi):
    c = collections.Counter(op for a, op, _ in ins if lo <= a <= hi)
    return ", ".join("%s %d" % kv for kv in c.most_common(14))
print("all:", hist(0, 1 << 30))
for a, op, txt in ins:
    if op == "BRA":
        m = re.search(r"0x([0-9a-f]+)", txt)
        if m and int(m.group(1), 16) < a:
            lo = int(m.group(1), 16)
            n = sum(1 for x in ins if lo <= x[0] <= a)
            if n > 40: print("loop %#x..%#x: %d instr: %s" % (lo, a, n, hist(lo, a)))

#!/usr/bin/env python
"""Top stall sites of one kernel from `ncu --page source --csv` output: cumulative samples by SASS instruction,
with a coarse region summary (samples / instructions executed between backward branches)."""
import csv, sys
allrows = list(csv.reader(open(sys.argv[1])))
# the file holds one section per launch: "Kernel Name" row, header row, instruction rows
starts = [i for i, r in enumerate(allrows) if r and r[0] == "Kernel Name"]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
lo = starts[which]; hi = starts[which + 1] if which + 1 < len(starts) else len(allrows)
print("section %d of %d: %s" % (which, len(starts), allrows[lo][1][:80]))
rows = allrows[lo:hi]
hdr = rows[1]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
ins = [(r[isrc].strip(), int(r[isamp] or 0), int(r[iex] or 0)) for r in rows[2:] if len(r) > iex]
tot = sum(s for _, s, _ in ins); totex = sum(e for _, _, e in ins)
print("instructions %d, samples %d, executed %d" % (len(ins), tot, totex))
top = sorted(range(len(ins)), key=lambda i: -ins[i][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for i in sorted(top):
    print("%5d %5.1f%% ex=%8d  %s   <- prev: %s" % (i, 100.0 * ins[i][1] / tot, ins[i][2], ins[i][0][:70], ins[i - 1][0][:50] if i else ""))
# windows of 100 instructions
print("--- windows")
for a in range(0, len(ins), 100):
    s = sum(x[1] for x in ins[a:a + 100]); e = sum(x[2] for x in ins[a:a + 100])
    if s * 100 > tot: print("%5d-%5d samples %5.1f%% executed %5.1f%%" % (a, a + 99, 100.0 * s / tot, 100.0 * e / totex))

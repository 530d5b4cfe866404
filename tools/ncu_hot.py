"""Summarise an ncu source page: executed SASS instructions and stall samples.
usage: ncu -i X.ncu-rep --page source --csv | python tools/ncu_hot.py [N]"""
import csv
import sys

top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rows = list(csv.reader(sys.stdin))
hdr = rows[1]
i_src = hdr.index("Source")
i_exec = hdr.index("Instructions Executed")
i_samp = hdr.index("# Samples")
items = []
for r in rows[2:]:
    try:
        items.append((int(r[i_exec]), int(r[i_samp]), r[i_src].strip()))
    except (ValueError, IndexError):
        pass
total = sum(i[0] for i in items)
samples = sum(i[1] for i in items)
print("SASS instructions: %d static, %d executed (warp-level), %d samples" % (len(items), total, samples))
nonzero = [i for i in items if i[0] > 0]
print("executed-at-least-once static instructions: %d" % len(nonzero))
ops = {}
for e, s, src in items:
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    op = op.split(".")[0]
    ops.setdefault(op, [0, 0])
    ops[op][0] += e
    ops[op][1] += s
print("-- by opcode (executed, samples)")
for op, (e, s) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:25]:
    print("%-10s %12d %6.1f%%  samples %6d %5.1f%%" % (op, e, 100.0 * e / max(total, 1), s, 100.0 * s / max(samples, 1)))
print("-- top by stall samples")
for e, s, src in sorted(items, key=lambda t: -t[1])[:top]:
    print("%10d %7d  %s" % (e, s, src[:100]))

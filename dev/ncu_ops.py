"""Opcode histogram (executed warp instructions, stall samples) from `ncu --page source --csv` (SASS view)."""
import csv, re, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = None
ops = collections.Counter(); smp = collections.Counter()
for r in rows:
    if 'Instructions Executed' in r:
        hdr = r; ii = hdr.index('Instructions Executed'); si = hdr.index('# Samples'); continue
    if hdr is None or len(r) <= ii: continue
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[1])
    if not m: continue
    try:
        ops[m.group(2)] += int(r[ii]); smp[m.group(2)] += int(r[si])
    except ValueError:
        pass
tot = sum(ops.values()); ts = sum(smp.values())
print('total warp instructions', tot, 'per unit', tot / div)
for op, c in ops.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 30):
    print('%-22s %12d %5.1f%%  samples %5.1f%%' % (op, c, 100 * c / tot, 100 * smp[op] / ts))

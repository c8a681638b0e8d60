"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line."""
import csv, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
rows = list(csv.reader(open(path)))
cur = None; hdr = None; out = []
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if r[0] == 'Function Name' or hdr is None: continue
    if r[0] != '':
        si = hdr.index('# Samples'); ii = hdr.index('Instructions Executed')
        names = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
        st = {h: int(r[hdr.index(h)] or 0) for h in names}
        best = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        try:
            out.append((int(r[si]), int(r[ii]), cur, r[0], r[1].strip()[:95], ' '.join('%s=%d' % (k[6:], v) for k, v in best if v)))
        except ValueError:
            pass
tot = sum(o[0] for o in out)
print('total samples', tot, ' total instr', sum(o[1] for o in out))
for o in sorted(out, reverse=True)[:top]:
    print('%6d %5.1f%% i=%9d %s:%s  %s   [%s]' % (o[0], 100.0 * o[0] / tot, o[1], o[2][:10], o[3], o[4], o[5]))

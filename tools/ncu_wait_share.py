"""Share of warp-stall samples spent in mbarrier spin loops vs real work, from one launch of an `ncu --page source --csv` dump."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
lines = [r for r in rows[2:] if len(r) >= len(hdr)]
tot = sum(int(r[col['# Samples']] or 0) for r in lines)
is_wait = [False] * len(lines)
waits = {}
for i, r in enumerate(lines):
    s = r[col['Source']]
    if 'SYNCS.PHASECHK' in s:
        key = s.split('[')[1].split(']')[0]
        t = 0
        for j in range(max(0, i - 1), min(len(lines), i + 8)):
            sj = lines[j][col['Source']]
            if any(x in sj for x in ('SYNCS', 'BRA', 'CS2R', 'YIELD', 'NOP', 'ISETP', 'IADD', 'IMAD', 'BSSY', 'BSYNC', 'UMOV', 'R2UR')) and not is_wait[j]:
                t += int(lines[j][col['# Samples']] or 0)
                is_wait[j] = True
        waits[key + ' @%d' % i] = t
print(rows[0][1][:90], 'samples', tot, 'warp-instr', sum(int(r[col['Instructions Executed']] or 0) for r in lines))
for k, v in sorted(waits.items(), key=lambda kv: -kv[1])[:8]:
    print('  wait %5.1f%%  %s' % (100 * v / tot, k))
rest = sum(int(r[col['# Samples']] or 0) for i, r in enumerate(lines) if not is_wait[i])
print('  non-wait %5.1f%%' % (100 * rest / tot))
st = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {h: sum(int(r[col[h]] or 0) for r in lines) for h in st}
print('  ', ', '.join('%s %d' % (k[6:], v) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))

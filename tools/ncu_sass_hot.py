"""Hot spots of an `ncu --page source --csv` dump (SASS view): opcode histogram by executed instructions and by stall samples,
plus the top sampled instructions with their dominant stall reason."""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
ops, samp = Counter(), Counter()
tot_i = tot_s = 0
recs = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    sass = r[col['Source']].strip()
    op = sass.split()[0] if not sass.startswith('@') else sass.split()[1]
    op = op.split('.')[0]
    n = int(r[col['Instructions Executed']] or 0)
    s = int(r[col['# Samples']] or 0)
    ops[op] += n
    samp[op] += s
    tot_i += n
    tot_s += s
    recs.append((s, n, sass, r))
print(f"total warp-instructions {tot_i}, samples {tot_s}")
print("opcode by executed instructions:")
for op, n in ops.most_common(22):
    print(f"  {op:12s} {n:12d} {100*n/tot_i:5.1f}%   samples {100*samp[op]/max(tot_s,1):5.1f}%")
print("top sampled instructions:")
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
for idx, (s, n, sass, r) in enumerate(sorted(recs, key=lambda x: -x[0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]):
    st = sorted(((int(r[col[h]] or 0), h) for h in stall_cols), reverse=True)[:2]
    print(f"  {100*s/max(tot_s,1):5.1f}%  exec {n:9d}  {sass[:70]:70s} {st[0][1]}={st[0][0]} {st[1][1]}={st[1][0]}")

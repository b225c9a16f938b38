"""Summarise an `ncu --page raw --csv` dump: per-launch duration, tensor-pipe %, DRAM bytes, issue %, top stall reasons."""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
keys = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'launch__registers_per_thread',
        'launch__grid_size']
idx = {k: hdr.index(k) for k in keys if k in hdr}
units = rows[1]
st = [h for h in hdr if 'smsp__average_warps_issue_stalled' in h]
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')]
    m = re.search(r'(\w+)<([^>]*)>', name)
    print((m.group(1) + '<' + m.group(2) + '>') if m else name[:40])
    print('   ' + ' | '.join('%s=%s %s' % (k.split('.')[0].replace('sm__', '').replace('smsp__', '')[:28], r[i][:10], units[i]) for k, i in idx.items()))
    vals = sorted([(float(r[hdr.index(h)].replace(',', '')) if r[hdr.index(h)] not in ('', 'n/a') else 0, h) for h in st],
                  reverse=True)[:5]
    print('   stalls: ' + ', '.join('%s %.2f' % (h.split('stalled_')[1].split('_per')[0], v) for v, h in vals))

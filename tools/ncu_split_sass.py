"""Split an `ncu --page source --csv` dump holding several launches into one CSV per launch: <prefix>_<n>_<kernel>.csv"""
import re
import sys

src, prefix = sys.argv[1], sys.argv[2]
out, n = None, 0
for line in open(src):
    if line.startswith('"Kernel Name"'):
        name = re.sub(r'[^A-Za-z0-9_]+', '_', re.sub(r'\(.*', '', line.split('","')[1]))[:60]
        if out:
            out.close()
        out = open(f"{prefix}_{n}_{name}.csv", "w")
        n += 1
    out.write(line)
if out:
    out.close()
print(n, "launches")

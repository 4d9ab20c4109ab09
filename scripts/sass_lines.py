#!/usr/bin/env python
"""Static code size by source line: `nvdisasm -g -c x.cubin > all.txt; python scripts/sass_lines.py all.txt <function substring>`"""
import collections
import re
import sys

path, want = sys.argv[1], sys.argv[2]
cur, infn = None, False
cnt = collections.Counter()
for line in open(path):
    if line.startswith("//---------------------"):
        infn = want in line and ".text." in line
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line) and cur:
        cnt[cur] += 1
tot = sum(cnt.values())
print("total instructions", tot)
byfile = collections.Counter()
for (f, l), c in cnt.items():
    byfile[f] += c
print(byfile.most_common(8))
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
for (f, l), c in cnt.most_common(top):
    print(f"{c:6d} {100.0 * c / tot:5.1f}%  {f}:{l}")

"""Instruction mix of the hottest inner loop of a kernel (development aid): python tools/sass_loop.py file.sass [minFFMA]
where file.sass = `nvdisasm -c cubin` output restricted to one kernel."""
import re, sys
from collections import Counter
L = open(sys.argv[1]).read().split('\n')
minf = int(sys.argv[2]) if len(sys.argv) > 2 else 50
labels = {}
for i, l in enumerate(L):
    m = re.match(r'^(\.L_x_\d+):', l)
    if m: labels[m.group(1)] = i
best = None
for i, l in enumerate(L):
    m = re.search(r'BRA\s+`\((\.L_x_\d+)\)', l)
    if m and m.group(1) in labels and labels[m.group(1)] < i:
        body = [b for b in L[labels[m.group(1)]:i + 1] if re.search(r'/\*[0-9a-f]{4}\*/', b)]
        n = sum(('FFMA' in b) for b in body)
        if n >= minf and (best is None or len(body) < best[0]): best = (len(body), body)
n, body = best
c = Counter(re.search(r'\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)', l).group(1).split('.')[0] for l in body)
print(n, "instructions:", c.most_common())

"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total and share.
python tools/summarize_launches.py gpurun_out/launches.csv [out.md]"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path, newline='') as f:
    lines = [l for l in f if not l.startswith('==')]
reader = csv.reader(lines)
hdr = next(reader)
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
tot = defaultdict(lambda: [0, 0.0])
for r in reader:
    if len(r) <= vi:
        continue
    name = r[ki].replace('(anonymous namespace)::', '').replace('<unnamed>::', '')
    name = re.sub(r'^void ', '', name)
    m = re.match(r'(at::native::)?(?:at::)?([\w:]+)(<[^,>]*)?', name)
    if name.startswith('at::') and m:
        fun = re.search(r'at::native::(\w+)|(\w+Functor)', name[len(m.group(0)) - len(m.group(3) or ''):])
        name = m.group(0).split('<')[0] + (f' [{fun.group(1) or fun.group(2)}]' if fun else '')
    else:
        name = re.sub(r'\(.*', '', name)
    v = float(r[vi].replace(',', ''))
    unit = r[ui]
    ms = v / 1e6 if unit in ('ns', 'nsecond') else (v / 1e3 if unit in ('us', 'usecond') else (v if unit in ('ms', 'msecond') else v * 1e3))
    tot[name][0] += 1
    tot[name][1] += ms
total = sum(v[1] for v in tot.values())
out = [f'# launch list summary of {path}', '', f'total device time {total:.2f} ms over {sum(v[0] for v in tot.values())} launches '
       '(ncu: cold-cache, serialised — compare SHARES, not absolutes)', '', '| kernel | launches | total ms | share |', '|---|---|---|---|']
for name, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    own = not name.startswith('at::')
    out.append(f'| {"" if own else "(torch) "}{name} | {n} | {ms:.3f} | {100 * ms / total:.1f}% |')
text = '\n'.join(out) + '\n'
if len(sys.argv) > 2:
    open(sys.argv[2], 'w').write(text)
print(text[:6000])

"""Instruction count per CUDA source line of one kernel (nvdisasm --print-line-info output): where the code size goes."""
import collections
import re
import sys

path, pat = sys.argv[1], sys.argv[2]
cur = None
fn = None
cnt = collections.Counter()
for line in open(path):
    if "inlined at" in line:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s*\.section\s+\.text\.(\S+?),', line)
    if m:
        fn = m.group(1)
    if re.match(r'\s*/\*[0-9a-f]{4,}\*/', line):
        cnt[(fn, cur)] += 1
byfn = collections.Counter()
for (f, c), n in cnt.items():
    byfn[f] += n
for f, n in byfn.items():
    print(n, str(f)[-60:])
tgt = [f for f in byfn if f and pat in f][0]
lines = sorted(((c, n) for (f, c), n in cnt.items() if f == tgt and c), key=lambda x: x[0])
thr = int(sys.argv[3]) if len(sys.argv) > 3 else 100
for c, n in lines:
    if n >= thr:
        print(c, n)

"""Attribute the SASS of one kernel to source lines (nvdisasm -g -c output):
   cuobjdump -xelf all lib.so; nvdisasm -g -c x.cubin > dis.txt; python scripts/sass_lines.py dis.txt scp_kernelILb1"""
import collections
import re
import sys

path, key = sys.argv[1], sys.argv[2]
regions = [(int(a), int(b), n) for a, b, n in (x.split(":") for x in sys.argv[3:])]
cur = fn = None
tot, ldl, op = collections.Counter(), collections.Counter(), collections.Counter()
for line in open(path):
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s*\.text\.(\S+):', line)
    if m:
        fn = m.group(1)
        continue
    if fn and key in fn and re.search(r'/\*[0-9a-f]{4,}\*/', line):
        tot[cur] += 1
        mm = re.search(r'\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if mm:
            op[mm.group(1).split('.')[0]] += 1
        if 'LDL' in line or 'STL' in line:
            ldl[cur] += 1
print("instructions", sum(tot.values()), "local ld/st", sum(ldl.values()))
print("top opcodes", op.most_common(14))
print("local ld/st by line:")
for k, v in ldl.most_common(15):
    print("  ", k, v)
reg = collections.Counter()
for (f, l), v in tot.items():
    reg[(f, l // 25 * 25)] += v
print("instructions by 25-line block (>= 300):")
for k, v in sorted(reg.items()):
    if v >= 300:
        print("  ", k, v)

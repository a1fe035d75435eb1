"""Opcode histogram per kernel of liblcgp_b200.so (cuobjdump -sass): evidence of the FP64 tensor path (DMMA), TMA
(UTMALDG), mbarriers (SYNCS) and cp.async (LDGSTS).   python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, 'lcgp_b200', '_lib', 'liblcgp_b200.so')
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
names = subprocess.run(['c++filt'], input='\n'.join(re.findall(r'Function : (\S+)', sass)), capture_output=True, text=True).stdout.split('\n')
kern, cur, it = collections.OrderedDict(), None, iter(names)
KEYS = ('DMMA', 'UTMALDG', 'UTMAPF', 'SYNCS', 'LDGSTS', 'DFMA', 'DMUL', 'DADD', 'MUFU', 'LDS', 'STS', 'LDG', 'STG', 'BAR', 'SHFL', 'ATOM', 'RED', 'NANOSLEEP', 'FENCE', 'MEMBAR', 'BRA', 'CALL')
for ln in sass.split('\n'):
    m = re.search(r'Function : (\S+)', ln)
    if m:
        cur = next(it)
        cur = re.sub(r'\(.*', '', cur).replace('lcgp::', '').replace('void ', '')
        kern[cur] = collections.Counter()
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', ln)
    if m and cur:
        op = m.group(1)
        kern[cur]['_total'] += 1
        for k in KEYS:
            if op.startswith(k):
                kern[cur][k] += 1
                break
print(f'# cuobjdump -sass lcgp_b200/_lib/liblcgp_b200.so (sm_100a): instructions per kernel by opcode family')
print(f'{"kernel":58s} {"total":>7s} ' + ' '.join(f'{k:>7s}' for k in KEYS[:14]))
for k, c in kern.items():
    print(f'{k[:58]:58s} {c["_total"]:7d} ' + ' '.join(f'{c[x]:7d}' for x in KEYS[:14]))
tot = collections.Counter()
for c in kern.values():
    tot.update(c)
print(f'{"ALL KERNELS":58s} {tot["_total"]:7d} ' + ' '.join(f'{tot[x]:7d}' for x in KEYS[:14]))

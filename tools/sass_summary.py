"""Per-kernel SASS evidence from the built library (cuobjdump -sass): the Blackwell-specific mnemonics B200_PROFILING.md lists
(UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UBLKCP = TMA, REDG...SYS = multimem.red, HMMA = legacy mma.sync),
registers and instruction counts.   python tools/sass_summary.py [lib.so] > profiles/r02_sass_summary.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else 'dalle_pytorch_b200/libdalle_b200.so'
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
KEYS = ['UTCHMMA', 'UTCBAR', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'UTMAPF', 'SYNCS', 'MUFU', 'HMMA', 'REDG', 'ATOMG', 'ELECT', 'NANOSLEEP']
cur, stats = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        cur = m.group(1)
        stats[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_x]+)*)', line)
    if m:
        op, mods = m.group(1), m.group(2)
        stats[cur]['_total'] += 1
        if op in KEYS:
            stats[cur][op] += 1
        if op == 'REDG' and '.SYS' in mods:
            stats[cur]['REDG.SYS(multimem)'] += 1
        if op == 'UTMALDG':
            dim = re.search(r'\.(\dD)', mods)
            if dim:
                stats[cur]['UTMALDG.' + dim.group(1)] += 1
print(f'# {lib}: {len(stats)} kernels; columns = static SASS instruction counts')
try:
    names = subprocess.run(['c++filt'], input='\n'.join(stats), capture_output=True, text=True).stdout.splitlines()
except Exception:
    names = list(stats)
for (k, c), name in zip(stats.items(), names):
    short = re.sub(r'db200::\(anonymous namespace\)::|db200::', '', name)
    short = re.sub(r'\(.*', '', short)[:110]
    keys = ' '.join(f'{kk}={vv}' for kk, vv in c.items() if kk != '_total' and vv)
    print(f'{c["_total"]:6d}  {short:110s} {keys}')
tot = collections.Counter()
for c in stats.values():
    tot.update(c)
print('# totals:', ' '.join(f'{k}={v}' for k, v in tot.items() if k != '_total'))

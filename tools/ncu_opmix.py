"""Instruction mix of one kernel from an `ncu --page source --csv` dump: warp-instructions per opcode, optionally divided by a
count (e.g. softmax warp-iterations).   python tools/ncu_opmix.py <csv> [divisor]"""
import collections
import csv
import sys


def main(path, div=1.0):
    rows = [r for r in csv.reader(open(path)) if r]
    hi = next(i for i, r in enumerate(rows) if r[0] == 'Address')
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    seen = {}
    for r in rows[hi + 1:]:
        if len(r) == len(hdr) and r[0] != 'Address':
            seen.setdefault(r[0], r)
    data = list(seen.values())
    tot = sum(int(r[ix['Instructions Executed']]) for r in data)
    print(f'{len(data)} SASS instructions, {tot} warp-instructions executed, {tot / div:.1f} per unit')
    ops = collections.Counter()
    for r in data:
        op = [o for o in r[ix['Source']].strip().split() if not o.startswith('@')][0].split('.')[0]
        ops[op] += int(r[ix['Instructions Executed']])
    for k, v in ops.most_common(24):
        print(f'  {k:12s} {v:10d} {v / div:8.1f}')


if __name__ == '__main__':
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0)

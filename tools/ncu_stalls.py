"""Per-instruction stall summary of one kernel from `ncu -i <rep> --page source --csv --kernel-name regex:<k>` (SASS view).
   python tools/ncu_stalls.py <csv> [top_n]"""
import csv
import sys


def main(path, top_n=40):
    rows = [r for r in csv.reader(open(path)) if r]
    hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
    hdr, data = rows[hi], [r for r in rows[hi + 1:] if len(r) == len(rows[hi])]
    ix = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    num = lambda r, h: int(r[ix[h]]) if r[ix[h]].isdigit() else 0
    tot = sum(num(r, '# Samples') for r in data)
    print('total samples', tot, 'instructions', len(data))
    agg = {h: sum(num(r, h) for r in data) for h in stall_cols}
    for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]:
        print(f'  {h:28s} {v:7d} {100 * v / max(tot, 1):5.1f}%')
    for r in sorted(data, key=lambda r: -num(r, '# Samples'))[:top_n]:
        st = sorted(((h[6:], num(r, h)) for h in stall_cols if num(r, h) > 0), key=lambda kv: -kv[1])[:3]
        print(f"{num(r, '# Samples'):6d} {r[ix['Instructions Executed']]:>9s}  {r[ix['Source']].strip()[:72]:72s} {st}")


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)

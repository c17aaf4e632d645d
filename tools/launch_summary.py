"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and count per kernel family."""
import collections
import csv
import re
import sys


def short(n):
    n = re.sub(r'<unnamed>::', '', n)
    m = re.search(r'(gemm_tcgen05_kernel<[^>]*>|gemm_simt_kernel<[^>]*>|attn_\w+kernel|ln_shift_\w+kernel|scale_bwd\w*kernel|colsum_kernel|'
                  r'cast_bf16_kernel|axpby_kernel|geglu_bwd_kernel|qkv_rotary_kernel|embed_\w+kernel|ce_\w+kernel)', n)
    if m:
        return m.group(1)
    n = re.sub(r'void (at::)?native::', '', n)
    return n[:90]


def main(path, skip=0):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr, rows = rows[0], rows[1 + skip:]
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    # one training step = the launches from one text-embedding gather (first kernel of DALLE.forward) to the next one
    starts = [i for i, r in enumerate(rows) if 'embed_fwd_kernel' in r[ki]][::2]
    if len(starts) >= 3:
        rows = rows[starts[-2]:starts[-1]]
        print(f'(one step: launches {starts[-2]}..{starts[-1] - 1} of the capture)')
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows:
        k, v = short(r[ki]), float(r[vi].replace(',', '')) / 1e3
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    print(f'{len(rows)} launches, {tot / 1e3:.2f} ms total (serialised, cold-cache: compare shares)')
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f'{t:10.1f} us {100 * t / tot:5.1f}% {c:5d}x  {k}')


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)

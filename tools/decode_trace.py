"""Kernel sequence of ONE device-indexed decoding step (the nodes of the decode graph), for an ncu launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/decode_launches.csv python tools/decode_trace.py
The steps run eagerly (decode.WARMUP_STEPS is raised so nothing is captured); a cumsum kernel marks the step boundaries.
    python tools/decode_trace.py summary gpurun_out/decode_launches.csv     -> per-kernel table of the last step"""
import collections
import csv
import os
import re
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))


def summary(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr, rows = rows[0], rows[1:]
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    marks = [i for i, r in enumerate(rows) if 'DeviceScanKernel' in r[ki]]
    assert len(marks) >= 2, 'no step markers in the launch list'
    rows = rows[marks[-2] + 1:marks[-1] - 1]          # (the marker is an init kernel + the scan)
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows:
        n = re.sub(r'<unnamed>::', '', r[ki])
        m = re.search(r'(gemm_\w+_kernel<[^>]*>|gemm_\w+_kernel|attn_\w+kernel|ln_shift_\w+kernel|qkv_rotary_kernel|decode_\w+kernel)', n)
        k = m.group(1) if m else re.sub(r'void (at::)?native::', '', n)[:100]
        v = float(r[vi].replace(',', '')) / 1e3
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    print(f'one decoding step: {len(rows)} launches, {tot:.1f} us of kernel time (serialised, cold-cache: compare shares)')
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        print(f'{t:9.1f} us {100 * t / tot:5.1f}% {c:5d}x  {k}')


def main():
    import torch
    import bench
    import dalle_pytorch_b200 as D
    from dalle_pytorch_b200 import decode
    decode.WARMUP_STEPS = 10 ** 9
    c = bench.CONFIGS['c2']
    D.set_compute_dtype(torch.bfloat16)
    torch.manual_seed(0)
    vae = D.TokenVAE(image_size=8 * c['fmap'], num_layers=3, num_tokens=bench.NUM_IMAGE_TOKENS)
    model = D.DALLE(dim=c['dim'], vae=vae, num_text_tokens=bench.NUM_TEXT_TOKENS, text_seq_len=c['text_seq_len'], depth=c['depth'],
                    heads=c['heads'], dim_head=64, attn_types=c['attn_types']).cuda().eval()
    text = torch.randint(1, bench.NUM_TEXT_TOKENS, (16, c['text_seq_len']), device='cuda')
    tok = torch.randint(0, bench.NUM_IMAGE_TOKENS, (16,), device='cuda')
    mark = torch.ones(4096, device='cuda')
    with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16):
        cache = {}
        model(text, tok[:, None][:, :0], cache=cache)
        dec = decode.GraphedDecoder(model, cache)
        for _ in range(4):
            dec.step(tok)
            torch.cumsum(mark, 0)
    torch.cuda.synchronize()
    print('trace done, position', int(dec.pos_t))


if __name__ == '__main__':
    if len(sys.argv) > 2 and sys.argv[1] == 'summary':
        summary(sys.argv[2])
    else:
        main()

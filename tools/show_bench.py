import json, sys
for f in sys.argv[1:]:
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as ex:
        print(f, 'unreadable', ex); continue
    r = j.get('roofline', {})
    print(f"{f}: {j['value']:.0f} tok/s  {j['ms_per_step']:.2f} ms/step  e2e {j['e2e']['value']:.0f}  gemm {r.get('achieved', 0):.0f} TF/s (share {r.get('share_of_step', 0):.2f})  clocks {j.get('clocks')}")
    for k, v in r.get('by_shape', {}).items():
        print(f"    {k:34s} {v['ms_per_launch'] * 1e3:8.1f} us {v['tflops']:7.1f} TF")
    if j.get('cpu_baseline'):
        print('    cpu_baseline', j['cpu_baseline'])

"""Isolated bring-up probe for the tcgen05 GEMM: each case runs in its own process (a device-side trap poisons the
CUDA context), prints max error vs torch and a first timing.   python tools/gemm_probe.py [--case NAME]"""
import argparse
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {
    # name: (a_mn, b_mn, M, N, K)
    'kk_small': (False, False, 128, 128, 64),
    'kk_1tile256': (False, False, 128, 256, 128),
    'kk_mid': (False, False, 512, 1024, 512),
    'kn_small': (False, True, 128, 128, 64),
    'kn_mid': (False, True, 512, 1024, 512),
    'mn_small': (True, True, 128, 128, 64),
    'mn_mid': (True, True, 1024, 512, 2048),
    'mk_small': (True, False, 128, 128, 64),
    'kk_ragged': (False, False, 200, 328, 136),
    'kk_big': (False, False, 20480, 1024, 1024),
    'kk_ff1': (False, False, 20480, 8192, 1024),
    'kn_big': (False, True, 20480, 1024, 4096),
    'mn_big': (True, True, 8192, 1024, 20480),
}


def run_case(name):
    import torch
    from dalle_pytorch_b200 import ops, _lib
    a_mn, b_mn, M, N, K = CASES[name]
    torch.manual_seed(0)
    dev = 'cuda:0'
    A = torch.randn((K, M) if a_mn else (M, K), device=dev).to(torch.bfloat16)
    B = torch.randn((K, N) if b_mn else (N, K), device=dev).to(torch.bfloat16)
    want = (A.float().t() if a_mn else A.float()) @ (B.float() if b_mn else B.float().t())
    got = ops.gemm_store(A, B, a_mn=a_mn, b_mn=b_mn, out_dtype=torch.float32, backend=_lib.GEMM_TCGEN05)
    torch.cuda.synchronize()
    err = (got - want).abs()
    tol = 1e-3 * K ** 0.5 + 1e-3 * want.abs()
    nbad = int((err > tol).sum())
    print(f'[{name}] M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn}: max|err|={float(err.max()):.4e} |want|max={float(want.abs().max()):.2f} '
          f'bad={nbad}/{err.numel()}', flush=True)
    if nbad:
        bad = (err > tol).nonzero()
        rows = sorted(set(bad[:, 0].tolist()))[:16]
        cols = sorted(set(bad[:, 1].tolist()))[:16]
        print(f'   first bad rows {rows} cols {cols}; got[0,:4]={got[0,:4].tolist()} want[0,:4]={want[0,:4].tolist()}', flush=True)
    if M * N * K > 1e9 and not nbad:
        for _ in range(3):
            ops.gemm_store(A, B, a_mn=a_mn, b_mn=b_mn, out_dtype=torch.bfloat16, backend=_lib.GEMM_TCGEN05)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10):
            ops.gemm_store(A, B, a_mn=a_mn, b_mn=b_mn, out_dtype=torch.bfloat16, backend=_lib.GEMM_TCGEN05)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 10
        print(f'   tcgen05: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s', flush=True)
        Ak = A.t().contiguous() if a_mn else A
        Bk = B.t().contiguous() if b_mn else B
        for _ in range(3):
            Ak @ Bk.t()
        s.record()
        for _ in range(10):
            Ak @ Bk.t()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 10
        print(f'   cuBLAS : {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s', flush=True)
    return 1 if nbad else 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--case', default=None)
    args = ap.parse_args()
    if args.case:
        sys.exit(run_case(args.case))
    fails = 0
    for name in CASES:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), '--case', name], timeout=180, capture_output=True, text=True)
            out = (r.stdout + r.stderr).strip().splitlines()
            keep = [l for l in out if l.startswith('[') or l.startswith('   ') or 'rror' in l or 'timeout' in l][-12:]
            print('\n'.join(keep) if keep else f'[{name}] no output, rc={r.returncode}')
            if r.returncode != 0:
                fails += 1
                print(f'[{name}] FAILED rc={r.returncode} ({time.time() - t0:.1f}s)')
        except subprocess.TimeoutExpired:
            fails += 1
            print(f'[{name}] TIMEOUT')
    print(f'gemm_probe: {len(CASES) - fails}/{len(CASES)} cases ok')


if __name__ == '__main__':
    main()

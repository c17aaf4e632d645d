"""bench.py — DALL-E fwd+bwd tokens/sec (BASELINE.json metric) on N x B200, and the reference arm on the host CPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c2|c3|c4|c1] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = `loss = dalle(text, image_ids, return_loss=True); loss.backward()` on one synthetic batch (reference
call-site contract train_dalle.py:609-616), gradients zeroed every step, for N>1 followed by the flat NCCL gradient
all-reduce (dalle_pytorch_b200.distributed).  Prints ONE JSON line (rank 0).

  value   tokens/s with the token-id inputs already resident in HBM
  e2e     tokens/s through the public module API with HOST (pinned) token buffers: H2D copy of the ids and a D2H read of
          the loss inside every timed step
  roofline  the tcgen05 GEMM family (dominant kernel): algorithmic FLOPs of every launch / CUDA-event time of those
          launches inside the timed region, against MEASURED_PEAKS.json bf16_tflops_sustained
  cpu_baseline  the oracle (CPU restatement of the reference, oracle/dalle_oracle.py) on the host cores, on a bounded sample
          (batch 1 of the same configuration)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # BASELINE.json configs[1..3]; c1 = configs[0] (the reference's CPU-runnable case)
    'c1': dict(dim=256, depth=2, heads=4, text_seq_len=64, fmap=8, batch=2, attn_types=('full',), reversible=False, dtype='fp32'),
    'c2': dict(dim=1024, depth=12, heads=16, text_seq_len=256, fmap=32, batch=16, attn_types=('full',), reversible=False, dtype='bf16'),
    'c3': dict(dim=1024, depth=12, heads=16, text_seq_len=256, fmap=32, batch=16, attn_types=('axial_row', 'axial_col'), reversible=False, dtype='bf16'),
    'c4': dict(dim=1024, depth=12, heads=16, text_seq_len=256, fmap=32, batch=64, attn_types=('axial_row', 'axial_col'), reversible=True, dtype='bf16'),
    'c5': dict(dim=1024, depth=64, heads=16, text_seq_len=256, fmap=32, batch=32, attn_types=('axial_row', 'axial_col'), reversible=False, dtype='bf16'),
}
NUM_TEXT_TOKENS, NUM_IMAGE_TOKENS = 10000, 8192
# measured with ncu on B200 (profiles/r01_gemm_ncu_summary.txt): mean dram__bytes_read+write of 20 consecutive tcgen05 GEMM launches
# of a C2 step (algorithmic operand+result bytes of the same launches: 226 MB per launch)
GEMM_DRAM_BYTES_PER_LAUNCH = 229e6
METRIC = 'DALL-E fwd+bwd tokens/sec at seq=1280, dim=1024'


def log(msg):
    print(f'[bench {time.strftime("%H:%M:%S")}] {msg}', file=sys.stderr, flush=True)


def usable_cores():
    """Host threads this process may actually use: the affinity mask capped by the cgroup CPU quota (a container that sees
    200 cores but is limited to 8 must not spawn 200 threads)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    try:
        quota, period = open('/sys/fs/cgroup/cpu.max').read().split()
        if quota != 'max':
            n = max(1, min(n, int(float(quota) / float(period) + 0.5)))
    except Exception:
        try:
            q = int(open('/sys/fs/cgroup/cpu/cpu.cfs_quota_us').read())
            per = int(open('/sys/fs/cgroup/cpu/cpu.cfs_period_us').read())
            if q > 0:
                n = max(1, min(n, int(q / per + 0.5)))
        except Exception:
            pass
    return n


def workload_name(name, c):
    return (f"{name}: depth={c['depth']} dim={c['dim']} heads={c['heads']} text_seq={c['text_seq_len']} image={c['fmap']}x{c['fmap']} "
            f"attn={'+'.join(c['attn_types'])}{' reversible' if c['reversible'] else ''} batch/GPU={c['batch']}")


# ------------------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / power / throttle reasons DURING the timed region.  NVML (nvidia_ml_py) polled every 10 ms from a thread -- the
    timed region of a default run is a few hundred ms, too short for `nvidia-smi -lms` to deliver a sample reliably (its start-up
    alone can take longer on an 8-GPU box); nvidia-smi is the fallback when NVML cannot be loaded."""
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
    REASONS = ((0x8, 'hw_slowdown'), (0x40, 'hw_thermal_slowdown'), (0x20, 'sw_thermal_slowdown'), (0x4, 'sw_power_cap'))

    def __init__(self, gpu_index=0):
        self.proc, self.lines, self.gpu_index = None, [], gpu_index
        self.nvml, self.samples, self.stop_flag, self.t = None, [], False, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.gpu_index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.nvml = pynvml
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '200',
                                          '-i', str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                pw = n.nvmlDeviceGetPowerUsage(self.handle) / 1e3
                try:
                    rs = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    rs = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.samples.append((float(sm), float(pw), int(rs)))
            except Exception:
                pass
            time.sleep(0.01)

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.t.join(timeout=2)
            if not self.samples:
                return {'sm_mhz': None, 'sm_max_mhz': float(self.max_sm), 'reasons': ['no samples']}
            sm = sorted(x[0] for x in self.samples)
            bits = 0
            for x in self.samples:
                bits |= x[2]
            return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': float(self.max_sm), 'power_w_max': max(x[1] for x in self.samples),
                    'samples': len(sm), 'source': 'nvml, 10 ms period', 'reasons': sorted(name for bit, name in self.REASONS if bits & bit)}
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(',')]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': max(mx), 'power_w_max': max(pw), 'samples': len(sm), 'source': 'nvidia-smi -lms 200',
                'reasons': sorted(reasons)}


# ------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (port of the reference's algorithm) on the host cores
# ------------------------------------------------------------------------------------------------------------
def cpu_step_fn(cfg_name, sample_batch=1):
    import torch
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    from dalle_oracle import OracleConfig, make_state_dict, make_inputs, dalle_forward
    c = CONFIGS[cfg_name]
    cores = usable_cores()
    torch.set_num_threads(cores)
    cfg = OracleConfig(dim=c['dim'], depth=c['depth'], heads=c['heads'], text_seq_len=c['text_seq_len'], fmap=c['fmap'],
                       num_text_tokens=NUM_TEXT_TOKENS, num_image_tokens=NUM_IMAGE_TOKENS, attn_types=c['attn_types'],
                       reversible=c['reversible'])
    sd = make_state_dict(cfg, seed=0, perturb=False, fast=True)
    params = {k: v.requires_grad_(k != 'transformer.pos_emb') for k, v in sd.items()}
    text, image = make_inputs(cfg, sample_batch, seed=1, pad_tail=False)
    tokens = sample_batch * cfg.seq_len

    def step():
        for p in params.values():
            p.grad = None
        loss = dalle_forward(text, image, params, cfg, return_loss=True)
        loss.backward()
        return float(loss.detach())

    return step, tokens, cores


def run_cpu_sample(args):
    """`--cpu-sample`: time ONE oracle step on batch 1 and print {"value", "cores", "seconds"} (called as a subprocess with a
    hard timeout by the GPU arm so that a slow host can never stall the bench)."""
    step, tokens, cores = cpu_step_fn(args.config, sample_batch=1)
    t0 = time.perf_counter()
    step()
    dt = time.perf_counter() - t0
    print(json.dumps({'value': tokens / dt, 'cores': cores, 'seconds': dt, 'tokens': tokens}))


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    c = CONFIGS[args.config]
    step, tokens, cores = cpu_step_fn(args.config, sample_batch=1)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = tokens * args.steps / dt
    sample = f'batch 1 of {c["batch"]} ({tokens} tokens) of the same configuration, full depth, fwd+bwd, fp32, torch CPU {cores} threads'
    out = {'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'tokens/s', 'n_gpus': args.gpus, 'steps': args.steps,
           'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
           'dtype': 'f32', 'data': 'synthetic', 'config': {'workload': workload_name(args.config, c)},
           'cpu_baseline': {'value': val, 'unit': 'tokens/s', 'cores': cores, 'kind': 'port', 'sample': sample},
           'e2e': {'value': val, 'unit': 'tokens/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------
def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        j = json.load(open(p))
        return j.get('bf16_tflops_sustained', 1405.9), j.get('hbm_gbs', 6567.4), 'measured (MEASURED_PEAKS.json, sustained)'
    return 1400.0, 6650.0, 'fallback (B200_PROFILING.md)'


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    import dalle_pytorch_b200 as D
    from dalle_pytorch_b200 import ops, _lib
    from dalle_pytorch_b200.distributed import NCCLBackend

    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    assert world == args.gpus or world == 1, f'--gpus {args.gpus} but WORLD_SIZE={world}'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    backend = None
    if world > 1:
        backend = NCCLBackend()
        backend.initialize()

    c = CONFIGS[args.config]
    batch = args.batch or c['batch']
    dtype = {'fp32': torch.float32, 'bf16': torch.bfloat16}[args.dtype or c['dtype']]
    D.set_compute_dtype(dtype)
    torch.manual_seed(0)
    vae = D.TokenVAE(image_size=8 * c['fmap'], num_layers=3, num_tokens=NUM_IMAGE_TOKENS)
    model = D.DALLE(dim=c['dim'], vae=vae, num_text_tokens=NUM_TEXT_TOKENS, text_seq_len=c['text_seq_len'], depth=c['depth'],
                    heads=c['heads'], dim_head=64, attn_types=c['attn_types'], reversible=c['reversible']).to(dev).train()
    reducer = None
    if backend is not None:
        backend.distribute(model=model)
        reducer = model.grad_reducer
    seq = c['text_seq_len'] + c['fmap'] ** 2
    g = torch.Generator().manual_seed(1 + rank)
    text_h = torch.randint(1, NUM_TEXT_TOKENS, (batch, c['text_seq_len']), generator=g).pin_memory()
    image_h = torch.randint(0, NUM_IMAGE_TOKENS, (batch, c['fmap'] ** 2), generator=g).pin_memory()
    text_d, image_d = text_h.to(dev), image_h.to(dev)
    head_autocast = dtype == torch.bfloat16

    def zero():
        if reducer is not None:
            reducer.zero_grad()
        else:
            for p in model.parameters():
                p.grad = None

    def fwd_bwd(text, image):
        zero()
        with torch.autocast('cuda', dtype=torch.bfloat16, enabled=head_autocast):
            loss = model(text, image, return_loss=True)
        loss.backward()
        if reducer is not None:
            reducer.finish()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    # ---- warm-up -------------------------------------------------------------------------------------------
    log(f'model built ({sum(p.numel() for p in model.parameters()) / 1e6:.1f} M params); warm-up')
    for i in range(max(args.warmup, 3)):
        t0 = time.perf_counter()
        fwd_bwd(text_d, image_d)
        torch.cuda.synchronize()
        log(f'warm-up step {i}: {1e3 * (time.perf_counter() - t0):.1f} ms')

    # ---- device-resident throughput, with per-GEMM CUDA events for the roofline ------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ops.gemm_timing(True)
    n0 = ops.launches()
    ms_dev = timed(lambda: fwd_bwd(text_d, image_d), args.steps)
    launches = ops.launches() - n0
    gemm_stats = ops.gemm_timing(False)
    log(f'device-resident: {ms_dev / args.steps:.2f} ms/step')
    clocks = sampler.stop() if rank == 0 else None

    # ---- end-to-end: host token buffers in, loss out, every step ------------------------------------------------
    def e2e_step():
        t = text_h.to(dev, non_blocking=True)
        i = image_h.to(dev, non_blocking=True)
        loss = fwd_bwd(t, i)
        return loss.item()

    e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    log(f'e2e: {ms_e2e / args.steps:.2f} ms/step')

    tokens_per_step = batch * seq * world
    value = tokens_per_step * args.steps / (ms_dev / 1e3)
    e2e_value = tokens_per_step * args.steps / (ms_e2e / 1e3)
    if world > 1:
        dist.barrier()
        if rank != 0:
            dist.destroy_process_group()
    if rank != 0:
        return

    peak_tf, peak_bw, peak_src = peaks()
    fam = gemm_stats.get('tcgen05', {'flops': 0.0, 'ms': 0.0, 'launches': 0})
    if fam['ms'] > 0:
        ach = fam['flops'] / (fam['ms'] * 1e-3) / 1e12
        roof = {'bound': 'tensor', 'kernel': 'gemm_tcgen05_kernel (all fwd/dgrad/wgrad GEMMs of the block stack)', 'achieved': ach,
                'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': ach / peak_tf, 'traffic': GEMM_DRAM_BYTES_PER_LAUNCH if args.config == 'c2' else None,
                'traffic_note': 'dram__bytes_read+write per launch, mean of the 20 consecutive GEMM launches (last forward layer, head, first backward layer) of the ncu --set full capture summarised in profiles/r01_gemm_ncu_summary.txt; algorithmic bytes of the same 20 launches average 226 MB',
                'peak_source': peak_src,
                'launches_per_step': fam['launches'] / args.steps, 'share_of_step': fam['ms'] / ms_dev,
                'by_shape': gemm_stats.get('by_shape', {})}
    else:
        fam = gemm_stats.get('simt', {'flops': 0.0, 'ms': 0.0, 'launches': 0})
        ach = fam['flops'] / (fam['ms'] * 1e-3) / 1e12 if fam['ms'] > 0 else 0.0
        roof = {'bound': 'tensor', 'kernel': 'gemm_simt_kernel (fp32 FFMA path)', 'achieved': ach, 'peak': peak_tf, 'unit': 'TFLOP/s',
                'frac': ach / peak_tf, 'traffic': None, 'peak_source': peak_src, 'share_of_step': fam['ms'] / ms_dev if ms_dev else None}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        log('cpu_baseline: oracle on the host cores (subprocess, 420 s limit)')
        sample = f'1 step on batch 1 of {batch} ({seq} tokens), full depth, fwd+bwd, fp32, oracle/dalle_oracle.py'
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), '--cpu-sample', '--config', args.config], capture_output=True,
                               text=True, timeout=420, env={**os.environ, 'CUDA_VISIBLE_DEVICES': ''})
            j = json.loads(r.stdout.strip().splitlines()[-1])
            cpu = {'value': j['value'], 'unit': 'tokens/s', 'cores': j['cores'], 'kind': 'port',
                   'sample': sample + f" on {j['cores']} torch threads ({j['seconds']:.1f} s)"}
        except Exception as ex:   # timeout / parse failure: report the failure, keep the GPU numbers
            cpu = {'value': None, 'unit': 'tokens/s', 'cores': usable_cores(), 'kind': 'port', 'sample': sample + f' — not completed: {type(ex).__name__}'}

    # model FLOPs per token for an MFU figure next to the kernel roofline
    d = c['dim']
    pairs = {'full': 819840, 'axial_row': 312928, 'axial_col': 312928}
    attn_flops = sum(4096.0 * pairs.get(t, 819840) / 1280 * (c['heads'] / 16) for t in c['attn_types']) / len(c['attn_types'])
    vocab = NUM_TEXT_TOKENS + c['text_seq_len'] + NUM_IMAGE_TOKENS
    flops_tok = 3 * (c['depth'] * (32.0 * d * d + attn_flops) + 2.0 * d * vocab)
    out = {'metric': METRIC, 'value': value, 'unit': 'tokens/s', 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
           'ms_per_step': ms_dev / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
           'dtype': 'bf16' if dtype == torch.bfloat16 else 'f32', 'data': 'synthetic',
           'config': {'workload': workload_name(args.config, c), 'global_batch': batch * world, 'seq_len': seq,
                      'parallelism': f'dp{world}', 'l2': 'activations and weights per step (GBs) exceed the 126 MB L2; no flush needed',
                      'head': 'token embedding gather/scatter, block stack, logits head GEMMs and cross-entropy all run in libdalle_b200'},
           'e2e': {'value': e2e_value, 'unit': 'tokens/s', 'ms_per_step': ms_e2e / args.steps,
                   'h2d_bytes_per_step': int(text_h.numel() * 8 + image_h.numel() * 8) * world, 'd2h_bytes_per_step': 4 * world},
           'gpu_launches': launches, 'model_tflops_per_gpu': value / world * flops_tok / 1e12,
           'mfu_vs_sustained_peak': value / world * flops_tok / 1e12 / peak_tf,
           'clocks': clocks, 'roofline': roof}
    if cpu is not None:
        out['cpu_baseline'] = cpu
    if args.with_optimizer:
        # full training step of the reference trainer (train_dalle.py:609-619): fwd + bwd + clip_grad_norm_(0.5) + Adam
        opt = D.FusedAdam(model.parameters(), lr=3e-4, max_grad_norm=0.5, reducer=reducer)

        def train_step():
            with torch.autocast('cuda', dtype=torch.bfloat16, enabled=head_autocast):
                loss = model(text_d, image_d, return_loss=True)
            loss.backward()
            opt.step()

        for _ in range(3):
            train_step()
        ms_train = timed(train_step, args.steps)
        out['train_step'] = {'ms_per_step': ms_train / args.steps, 'tokens_per_s': tokens_per_step * args.steps / (ms_train / 1e3),
                             'optimizer': 'FusedAdam lr=3e-4 betas=(0.9,0.999) clip_grad_norm 0.5 (2 launches over flat fp32 buffers)'}
        log(f'train step (fwd+bwd+clip+Adam): {ms_train / args.steps:.2f} ms/step')
    print(json.dumps(out))
    sys.stdout.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default='c2', choices=sorted(CONFIGS))
    ap.add_argument('--batch', type=int, default=None)
    ap.add_argument('--dtype', default=None, choices=['fp32', 'bf16'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--with-optimizer', action='store_true',
                    help='also time fwd+bwd+FusedAdam(clip 0.5) steps and report them under "train_step" (headline metric unchanged)')
    ap.add_argument('--cpu-sample', action='store_true', help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.cpu_sample:
        run_cpu_sample(args)
    elif args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_gpu_arm(args)
        try:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.destroy_process_group()
        except Exception:
            pass


if __name__ == '__main__':
    main()

"""Data-parallel backend: one process per GPU, a flat bucketed NCCL all-reduce of the gradients over
NVLink / NVSwitch, overlapped with the backward pass.

Replaces the reference's `distributed_backends` (DeepSpeed / Horovod adapters, reference
dalle_pytorch/distributed_backends/distributed_backend.py:12-178, distributed_utils.py:19-96) behind the SAME
facade: `NCCLBackend` implements the abstract methods `_initialize, _get_world_size, _get_rank, _get_local_rank,
_local_barrier, _distribute, _average_all`, so `train_dalle.py`-style callers (train_dalle.py:228-229, 509-523,
612-622) work unchanged.  The path is pure data parallel: samples are independent, the only
exchange is the gradient SUM (+ the scalar loss average the reference logs).

Design: every trainable parameter's .grad is a view into ONE flat fp32 buffer (no gather/scatter copies); the
buffer is cut into buckets in reverse registration order (= the order gradients become final during backward);
a post-accumulate hook counts parameters down and launches `all_reduce(bucket, async_op=True)` as soon as a
bucket is complete, so communication of layer i overlaps the backward kernels of layer i-1.  `finish()` waits for
the outstanding work and applies 1/world.  NVSwitch gives every GPU full bandwidth to every peer, so buckets are
sized for launch latency / overlap (default 64 MiB), not for link count.

Multimem mode (NVSwitch boxes, opt-in: `mode='multimem'` / DALLE_B200_DP=multimem): there is NO collective kernel.  The flat buffer is a symmetric allocation (torch.distributed._symmetric_memory) mapped through an NVLink
multicast address; every weight-gradient GEMM adds `acc / world` into ALL replicas from its own epilogue (`multimem.red`, reduced
inside the switch: gemm STORE with db200_gemm_params::C_multicast), the remaining gradients (LayerNorm, biases, embeddings, head)
are pushed the same way by `dalle_b200_mc_add` as they appear, and `finish()` is one device-side barrier.  The gradient bytes cross
NVLink once per peer, spread over the backward pass, without taking SMs from the GEMMs -- the fused compute + collective form of the
data-parallel step.  Per step: zero_grad() clears the local replica on a side stream and runs a barrier (nobody may contribute to
a replica that has not been cleared); the first contribution of the backward pass waits for that event.
Measured (profiles/r02_summary.md): correct to 5e-7 against the single-process mean (tools/dp_check.py); on 2 x B200 41.70 ms per C2
step against 41.06 ms with the overlapped NCCL buckets (40.58 ms on one GPU).  `multimem.red` is a PUSH: every replica receives
one atomic add per contributing GPU, so the NVLink ingress per GPU grows with the world size (8 x 0.96 GB at eight GPUs) where
NCCL's NVLS all-reduce (pull with multimem.ld_reduce, then multicast store) stays at ~2 x 0.96 GB -- which is why NCCL remains the
default and this mode is kept for two-GPU boxes and as the template for a fused reduce-scatter.

Validity contract (what makes the reference loop `loss.backward(); clip_grad_norm_(params); opt.step()`,
train_dalle.py:612-622, safe): the first gradient of a backward pass queues an end-of-backward callback on the
autograd engine which runs `finish()`, so when `backward()` returns every `.grad` is the reduced mean and no
collective is still writing the flat buffer.  Gradient accumulation: wrap all but the last backward of a step in
`reducer.no_sync()`; a gradient that arrives after the step's reduction raises instead of silently mixing
averaged and local gradients.
"""
import contextlib
import os

import torch
import torch.distributed as dist


class DistributedBackend:
    """Facade with the reference's method names (distributed_backend.py:12-178)."""
    BACKEND_MODULE_NAME = None
    BACKEND_NAME = None
    ROOT_RANK = 0
    is_initialized = False

    def __init__(self):
        if self.BACKEND_MODULE_NAME is None:
            raise NotImplementedError('BACKEND_MODULE_NAME is not set')
        if self.BACKEND_NAME is None:
            raise NotImplementedError('BACKEND_NAME is not set')

    def has_backend(self):
        try:
            from importlib import import_module
            self.backend_module = import_module(self.BACKEND_MODULE_NAME)
        except ModuleNotFoundError:
            return False
        return True

    def check_batch_size(self, batch_size):
        assert batch_size >= self.get_world_size(), \
            f"batch size can't be smaller than number of processes ({batch_size} < {self.get_world_size()})"

    def wrap_arg_parser(self, parser):
        return parser

    def initialize(self):
        self._initialize()
        self.is_initialized = True

    def require_init(self):
        assert self.is_initialized, \
            f'{self.BACKEND_NAME} backend has not been initialized; please call `distributed_utils.initialize` at the start of your script'

    def get_world_size(self):
        self.require_init()
        return self._get_world_size()

    def get_rank(self):
        self.require_init()
        return self._get_rank()

    def get_local_rank(self):
        self.require_init()
        return self._get_local_rank()

    def is_root_worker(self):
        return self.get_rank() == self.ROOT_RANK

    def is_local_root_worker(self):
        return self.get_local_rank() == self.ROOT_RANK

    def local_barrier(self):
        self.require_init()
        self._local_barrier()

    def distribute(self, args=None, model=None, optimizer=None, model_parameters=None, training_data=None, lr_scheduler=None,
                   **kwargs):
        self.require_init()
        return self._distribute(args, model, optimizer, model_parameters, training_data, lr_scheduler, **kwargs)

    def average_all(self, tensor):
        self.require_init()
        return self._average_all(tensor)


class DummyBackend(DistributedBackend):
    """World of one (dummy_backend.py:4-52)."""
    BACKEND_MODULE_NAME = 'torch'
    BACKEND_NAME = 'Dummy'

    def _initialize(self):
        pass

    def _get_world_size(self):
        return 1

    def _get_rank(self):
        return self.ROOT_RANK

    def _get_local_rank(self):
        return self.ROOT_RANK

    def _local_barrier(self):
        pass

    def _distribute(self, _args=None, model=None, optimizer=None, _model_parameters=None, training_data=None, lr_scheduler=None,
                    **_kwargs):
        return (model, optimizer, training_data, lr_scheduler)

    def _average_all(self, tensor):
        return tensor


class GradAllReducer:
    """Flat-buffer bucketed gradient all-reduce (see module docstring)."""

    def __init__(self, params, process_group=None, bucket_bytes=64 << 20, average=True, auto_finish=True, overlap=None, mode=None):
        self.pg = process_group
        self.mode = mode or os.environ.get('DALLE_B200_DP', 'nccl')           # 'nccl' (default) | 'multimem' | 'auto' (multimem if possible)
        assert self.mode in ('auto', 'multimem', 'nccl')
        self.mc = None                                                        # multicast base address of the flat buffer (multimem mode)
        # overlap=False: nothing is launched during backward; finish() reduces the whole flat buffer with ONE collective (the
        # collective's CTAs then never compete with the backward GEMMs for SMs / HBM; its time is fully exposed)
        self.overlap = (os.environ.get('DALLE_B200_DP_OVERLAP', '1') != '0') if overlap is None else bool(overlap)
        # compress='bf16': the buckets cross NVLink as bf16 (cast -> all-reduce -> cast back into the fp32 buffer): half the bytes,
        # at the price of rounding every rank's gradient to 8 mantissa bits before the sum (opt-in)
        self.compress = os.environ.get('DALLE_B200_DP_COMPRESS', '') == 'bf16'
        self._staged = []
        self.auto_finish = auto_finish      # finish() runs as an autograd end-of-backward callback
        self._sync, self._finished, self._cb_queued = True, False, False
        self._direct_pass = set()        # parameters delivered through direct_done() in the running backward pass
        # without an initialised process group this is just the flat gradient buffer of a single process (optim.FusedAdam)
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.average = average
        self.params = [p for p in params if p.requires_grad]
        assert self.params, 'no trainable parameters'
        dev = self.params[0].device
        # reverse order: the last layers' gradients are final first
        order = list(reversed(self.params))
        sizes = [p.numel() for p in order]
        # every view starts on a 256-byte boundary: the kernels that read or write parameters and gradients (TMA operands,
        # 16-byte vector loads of gamma / bias / embedding rows, the weight-gradient GEMM writing into its slot) need aligned
        # pointers, and odd-sized tensors (a bias of 90 elements) would otherwise misalign everything behind them
        ALIGN = 64
        padded = [(n + ALIGN - 1) // ALIGN * ALIGN for n in sizes]
        total = sum(padded)
        self.flat = None
        if self.world > 1 and self.mode != 'nccl' and dev.type == 'cuda' and dist.get_backend(process_group) == 'nccl':
            try:
                self._init_multimem(total, dev)
            except Exception as ex:
                if self.mode == 'multimem':
                    raise
                self.mc = None
                import warnings
                warnings.warn(f'GradAllReducer: NVLink multicast not available ({type(ex).__name__}: {ex}); using NCCL all-reduce')
        if self.flat is None:
            self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        self.views, self.bucket_of, self.buckets = {}, {}, []
        off, b_start, b_params = 0, 0, []
        limit = max(1, bucket_bytes // 4)
        for p, n, n_pad in zip(order, sizes, padded):
            self.views[p] = self.flat[off:off + n].view_as(p)
            if self.mc is not None:       # what ops.gemm_store / ops.mc_add need to reduce into every replica of this slot
                self.views[p]._b200_mc = (self.mc + 4 * off, self._mc_scale)
            b_params.append(p)
            off += n_pad
            if off - b_start >= limit:
                self.buckets.append((b_start, off, b_params))
                b_start, b_params = off, []
        if b_params:
            self.buckets.append((b_start, off, b_params))
        for bi, (_, _, ps) in enumerate(self.buckets):
            for p in ps:
                self.bucket_of[p] = bi
        # DALLE_B200_DP_OP=sum: all-reduce SUM and scale afterwards (ncclAvg is a pre-multiplied sum that the in-switch NVLS reduction
        # does not implement, NCCL then falls back to ring kernels)
        self._has_avg = (dist.is_initialized() and dist.get_backend(process_group) == 'nccl' and
                         os.environ.get('DALLE_B200_DP_OP', 'avg') != 'sum')
        self._seen = set()
        self._pending = [0] * len(self.buckets)
        self._launched = [False] * len(self.buckets)
        self._works = []
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        for p in self.params:
            p._b200_reducer = self
        self.zero_grad()

    # -- multimem mode ---------------------------------------------------------------------------------------------
    def _init_multimem(self, total, dev):
        import torch.distributed._symmetric_memory as symm_mem
        group = self.pg if self.pg is not None else dist.group.WORLD
        buf = symm_mem.empty(total, dtype=torch.float32, device=dev)
        hdl = symm_mem.rendezvous(buf, group.group_name)
        if not getattr(hdl, 'multicast_ptr', 0):
            raise RuntimeError('the symmetric allocation has no multicast mapping')
        buf.zero_()
        hdl.barrier(channel=0)
        self.flat, self._hdl, self.mc = buf, hdl, int(hdl.multicast_ptr)
        self._mc_scale = (1.0 / self.world) if self.average else 1.0
        self._side = torch.cuda.Stream(device=dev)
        self._zero_ev, self._dirty = None, False

    def _mc_ready(self):
        """Before the first contribution of a step: the local replica has been cleared and every peer's has (zero_grad)."""
        if self._zero_ev is not None:
            torch.cuda.current_stream().wait_event(self._zero_ev)
            self._zero_ev = None
        self._dirty = True

    def _mc_clear(self):
        if not self._dirty:            # nothing was contributed since the last clear (e.g. zero_grad() called twice): keep the barrier count symmetric
            return
        cur = torch.cuda.current_stream()
        self._side.wait_stream(cur)                                  # after whoever consumed the gradients (optimizer step)
        with torch.cuda.stream(self._side):
            self.flat.zero_()
            self._hdl.barrier(channel=0)
            ev = torch.cuda.Event()
            ev.record(self._side)
        self._zero_ev, self._dirty = ev, False

    # -- per-step protocol: zero_grad() -> forward/backward -> finish() -> optimizer.step() ----------------
    def zero_grad(self):
        """Start of a step.  Gradients are NOT pre-zeroed views any more: autograd writes a fresh gradient tensor per
        parameter and the post-accumulate hook moves it into the flat buffer (one copy instead of memset + read-modify-write),
        then points .grad at the view."""
        for p in self.params:
            p.grad = None
            p._b200_uses = 0
            p._b200_acc = 0
        self._seen = set()
        for bi, (_, _, ps) in enumerate(self.buckets):
            self._pending[bi] = len(ps)
            self._launched[bi] = False
        self._works = []
        self._finished = False
        self._direct_pass = set()
        if self.mc is not None:
            self._mc_clear()

    @contextlib.contextmanager
    def no_sync(self):
        """Gradient accumulation: backward passes inside this context only accumulate into the flat buffer (no bucket is
        launched, nothing is reduced); the first backward outside it reduces the accumulated sum."""
        if self.mc is not None:
            raise NotImplementedError('gradient accumulation (no_sync) needs mode="nccl": in multimem mode every contribution is '
                                      'reduced across the GPUs as it is produced')
        old, self._sync = self._sync, False
        try:
            yield self
        finally:
            self._sync = old

    def _guard(self):
        if self._finished:
            raise RuntimeError('GradAllReducer: a gradient arrived after this step\'s all-reduce (second backward without '
                               'zero_grad()).  For gradient accumulation run the earlier backward passes under reducer.no_sync(); '
                               'otherwise call zero_grad() / optimizer.step() between steps.')

    def _arm(self):
        """Queue finish() to run when the current backward pass ends (once per pass)."""
        if not self.auto_finish or self._cb_queued:
            return
        try:
            torch.autograd.Variable._execution_engine.queue_callback(self._end_of_backward)
            self._cb_queued = True
        except RuntimeError:       # not inside a backward pass (gradient handed over by hand): the caller runs finish()
            pass

    def _end_of_backward(self):
        self._cb_queued = False
        self._direct_pass = set()
        if self._sync:
            self.finish()

    def _launch(self, bi):
        s, e, _ = self.buckets[bi]
        self._launched[bi] = True
        if self.world > 1 and self.mc is None:
            self._reduce_range(s, e)

    def _reduce_range(self, s, e):
        # AVG folds the 1/world into the collective (NCCL); gloo (CPU tests) has no AVG -> SUM + scale in finish()
        op = dist.ReduceOp.AVG if (self.average and self._has_avg) else dist.ReduceOp.SUM
        buf = self.flat[s:e]
        if self.compress and self._has_avg:
            buf = buf.to(torch.bfloat16)
            self._staged.append((buf, s, e))
        self._works.append(dist.all_reduce(buf, op=op, group=self.pg, async_op=True))

    def _adopt(self, p):
        v = self.views[p]
        if self.mc is not None:        # push the local gradient into every replica (nothing to push for a parameter without one)
            if p.grad is not None:
                if p.grad.data_ptr() == v.data_ptr():
                    raise RuntimeError(f'GradAllReducer(multimem): a gradient of shape {tuple(p.shape)} was accumulated in place into its flat-buffer view (seen={p in self._seen}, uses={getattr(p, "_b200_uses", None)})')
                from . import ops
                self._mc_ready()
                ops.mc_add(p.grad.detach().contiguous().view(-1).float(), v._b200_mc[0], v._b200_mc[1])
        elif p.grad is None:
            v.zero_()
        elif p.grad.data_ptr() != v.data_ptr():
            v.copy_(p.grad)
        p.grad = v
        self._seen.add(p)

    # direct-write protocol used by the fused sub-layer backward (functional.py::_slot/_commit): the weight-gradient GEMM
    # writes into the flat-buffer view itself, so neither a memset nor a copy of that gradient is ever made
    def direct_slot(self, p):
        self._guard()
        if p in self._seen or p not in self.views:
            return None
        if self.mc is not None:
            self._mc_ready()
        return self.views[p]

    def direct_done(self, p):
        p.grad = self.views[p]
        self._seen.add(p)
        self._direct_pass.add(p)
        self._count(p)

    def _count(self, p):
        self._arm()
        if not self._sync:           # accumulation pass: nothing is launched
            return
        bi = self.bucket_of[p]
        self._pending[bi] -= 1
        if self.overlap and self._pending[bi] == 0 and not self._launched[bi]:
            self._launch(bi)

    def _on_grad(self, p):
        # The autograd engine runs the post-accumulate hook of a parameter even when its Function returned None for it (observed
        # with torch 2.11): a gradient that was delivered through direct_slot()/direct_done() must not be counted a second time
        # (the bucket would be launched before its other members have arrived).
        if p in self._direct_pass:
            self._direct_pass.discard(p)
            return
        self._guard()
        self._adopt(p)
        self._count(p)

    def finish(self):
        """Reduce whatever has not been launched by the hooks (parameters whose gradients were written outside
        autograd, or that received no gradient), wait, and average.  Idempotent within a step: a second call (explicit call +
        the optimizer's step pre-hook) returns immediately; zero_grad() opens the next step."""
        if self._finished:
            return
        if self.mc is not None:        # every contribution is already on its way into every replica: adopt the stragglers, one barrier
            for p in self.params:
                if p not in self._seen:
                    self._adopt(p)
            for bi in range(len(self.buckets)):
                self._launched[bi] = True
            self._hdl.barrier(channel=1)
            self._finished = True
            return
        if not self.overlap and not any(self._launched):
            for p in self.params:
                if p not in self._seen:
                    self._adopt(p)
            for bi in range(len(self.buckets)):
                self._launched[bi] = True
            if self.world > 1:
                self._reduce_range(0, self.flat.numel())
        for bi, (_, _, ps) in enumerate(self.buckets):
            if not self._launched[bi]:
                for p in ps:
                    if p not in self._seen:
                        self._adopt(p)
                self._launch(bi)
        for w in self._works:
            w.wait()
        self._works = []
        for buf, s0, e0 in self._staged:
            self.flat[s0:e0].copy_(buf)
        self._staged = []
        if self.average and self.world > 1 and not self._has_avg:
            self.flat.mul_(1.0 / self.world)
        self._finished = True

    def remove(self):
        for h in self._hooks:
            h.remove()
        for p in self.params:
            p._b200_reducer = None


class NCCLBackend(DistributedBackend):
    """torch.distributed over NCCL (one process per GPU, rendezvous from the torchrun environment)."""
    BACKEND_MODULE_NAME = 'torch.distributed'
    BACKEND_NAME = 'NCCL'

    def __init__(self, comm_backend=None, bucket_bytes=64 << 20):
        super().__init__()
        self.comm_backend = comm_backend
        self.bucket_bytes = bucket_bytes
        self.reducer = None

    def wrap_arg_parser(self, parser):
        parser.add_argument('--local_rank', type=int, default=int(os.environ.get('LOCAL_RANK', 0)))
        return parser

    def _initialize(self):
        backend = self.comm_backend or ('nccl' if torch.cuda.is_available() else 'gloo')
        if not dist.is_initialized():
            os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
            os.environ.setdefault('MASTER_PORT', '29500')
            rank = int(os.environ.get('RANK', 0))
            world = int(os.environ.get('WORLD_SIZE', 1))
            if backend == 'nccl':
                torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
                dist.init_process_group(backend, rank=rank, world_size=world,
                                        device_id=torch.device('cuda', int(os.environ.get('LOCAL_RANK', 0))))
            else:
                dist.init_process_group(backend, rank=rank, world_size=world)

    def _get_world_size(self):
        return dist.get_world_size()

    def _get_rank(self):
        return dist.get_rank()

    def _get_local_rank(self):
        return int(os.environ.get('LOCAL_RANK', dist.get_rank()))

    def _local_barrier(self):
        dist.barrier()

    def _distribute(self, _args=None, model=None, optimizer=None, model_parameters=None, training_data=None, lr_scheduler=None,
                    **_kwargs):
        """Broadcast rank 0's parameters (Horovod analogue horovod_backend.py:49-52), attach the gradient reducer
        and make `optimizer.step()` wait for it (the reduction itself already completes inside `backward()`, see the
        module docstring, so `clip_grad_norm_` between backward and step sees reduced gradients).
        Returns (model, optimizer, training_data, lr_scheduler)."""
        with torch.no_grad():
            for t in list(model.parameters()) + list(model.buffers()):
                dist.broadcast(t, src=self.ROOT_RANK)
        from .functional import invalidate_weight_cache
        invalidate_weight_cache()             # cached bf16 weight copies predate the broadcast
        params = list(model_parameters) if model_parameters is not None else list(model.parameters())
        self.reducer = GradAllReducer(params, bucket_bytes=self.bucket_bytes)
        model.grad_reducer = self.reducer
        if optimizer is not None:
            reducer = self.reducer
            optimizer.register_step_pre_hook(lambda *_a, **_k: reducer.finish())
            optimizer.register_step_post_hook(lambda *_a, **_k: reducer.zero_grad())
        return (model, optimizer, training_data, lr_scheduler)

    def _average_all(self, tensor):
        """deepspeed_backend.py:165-171: all-reduce SUM then divide by world."""
        out = tensor.detach().clone()
        dist.all_reduce(out, op=dist.ReduceOp.SUM)
        return out / self.get_world_size()


BACKENDS = [DummyBackend(), NCCLBackend()]

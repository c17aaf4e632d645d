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

    def __init__(self, params, process_group=None, bucket_bytes=64 << 20, average=True, auto_finish=True):
        self.pg = process_group
        self.auto_finish = auto_finish      # finish() runs as an autograd end-of-backward callback
        self._sync, self._finished, self._cb_queued = True, False, False
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
        self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        self.views, self.bucket_of, self.buckets = {}, {}, []
        off, b_start, b_params = 0, 0, []
        limit = max(1, bucket_bytes // 4)
        for p, n, n_pad in zip(order, sizes, padded):
            self.views[p] = self.flat[off:off + n].view_as(p)
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
        self._has_avg = dist.is_initialized() and dist.get_backend(process_group) == 'nccl'
        self._seen = set()
        self._pending = [0] * len(self.buckets)
        self._launched = [False] * len(self.buckets)
        self._works = []
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        for p in self.params:
            p._b200_reducer = self
        self.zero_grad()

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

    @contextlib.contextmanager
    def no_sync(self):
        """Gradient accumulation: backward passes inside this context only accumulate into the flat buffer (no bucket is
        launched, nothing is reduced); the first backward outside it reduces the accumulated sum."""
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
        if self._sync:
            self.finish()

    def _launch(self, bi):
        s, e, _ = self.buckets[bi]
        self._launched[bi] = True
        if self.world > 1:
            # AVG folds the 1/world into the collective (NCCL); gloo (CPU tests) has no AVG -> SUM + scale in finish()
            op = dist.ReduceOp.AVG if (self.average and self._has_avg) else dist.ReduceOp.SUM
            self._works.append(dist.all_reduce(self.flat[s:e], op=op, group=self.pg, async_op=True))

    def _adopt(self, p):
        v = self.views[p]
        if p.grad is None:
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
        return self.views[p]

    def direct_done(self, p):
        p.grad = self.views[p]
        self._seen.add(p)
        self._count(p)

    def _count(self, p):
        self._arm()
        if not self._sync:           # accumulation pass: nothing is launched
            return
        bi = self.bucket_of[p]
        self._pending[bi] -= 1
        if self._pending[bi] == 0 and not self._launched[bi]:
            self._launch(bi)

    def _on_grad(self, p):
        self._guard()
        self._adopt(p)
        self._count(p)

    def finish(self):
        """Reduce whatever has not been launched by the hooks (parameters whose gradients were written outside
        autograd, or that received no gradient), wait, and average.  Idempotent within a step: a second call (explicit call +
        the optimizer's step pre-hook) returns immediately; zero_grad() opens the next step."""
        if self._finished:
            return
        for bi, (_, _, ps) in enumerate(self.buckets):
            if not self._launched[bi]:
                for p in ps:
                    if p not in self._seen:
                        self._adopt(p)
                self._launch(bi)
        for w in self._works:
            w.wait()
        self._works = []
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

"""One training step (forward + backward, optionally the optimizer) captured once into a CUDA graph and replayed.

The eager step of the C2 configuration is ~440 kernel launches issued from Python through ctypes (9 ms of host time for 39 ms of
GPU work, 1.6 ms of launch gaps, tools/gpu_busy.py); shapes, pointers and the launch sequence are identical from step to step, so
the whole step is recorded once (`torch.cuda.graph`; our launches go to the capturing stream like any other, tensor maps and the
programmatic-dependent-launch edges are baked into the kernel nodes) and replayed with ONE host call per step.

    step = GraphedStep(dalle, text, image_ids)            # captures loss = dalle(text, image, return_loss=True); loss.backward()
    loss = step(text_batch, image_batch)                  # copies the ids into the static input buffers, replays, returns the loss
    # gradients are in p.grad (static tensors owned by the graph's memory pool), overwritten by every replay

With `optimizer=` the optimizer step is part of the graph (the bf16 weight-copy cache is then bypassed inside the capture: Python
does not run at replay, the casts must be graph nodes).  Data parallel: the gradient all-reduce is launched from Python hooks, so
`GraphedStep` is for single-process use; multi-GPU runs use the eager path.
"""
import torch

from . import ops, functional


class GraphedStep:
    def __init__(self, model, text, image, *, optimizer=None, warmup=3, autocast_bf16=None, return_loss_kwargs=None):
        assert text.is_cuda and image.is_cuda, 'GraphedStep needs CUDA tensors'
        assert getattr(model, 'grad_reducer', None) is None or model.grad_reducer.world == 1, \
            'GraphedStep captures a single-process step (the data-parallel all-reduce is driven from Python hooks)'
        for mod in model.modules():
            if isinstance(mod, torch.nn.Dropout) and mod.p > 0 and model.training:
                raise NotImplementedError('GraphedStep: dropout > 0 draws a new mask offset per step from Python; a captured step would '
                                          'replay one mask')
        self.model, self.optimizer = model, optimizer
        self.text, self.image = text.clone(), image.clone()            # static input buffers
        self.kw = dict(return_loss_kwargs or {})
        if autocast_bf16 is None:
            from . import config
            autocast_bf16 = config.compute_dtype() == torch.bfloat16
        self.autocast = bool(autocast_bf16)
        self.params = [p for p in model.parameters() if p.requires_grad]
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):                                  # warm-up off the capture stream (allocator, lazy set-up, caches)
            for _ in range(max(1, warmup)):
                self._zero()
                self._step()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self._zero()
        was_timing = ops._gemm_events is not None
        assert not was_timing, 'stop ops.gemm_timing before capturing a step'
        cache_was = functional.weight_cache_enabled(optimizer is None)
        self.graph = torch.cuda.CUDAGraph()
        n0 = ops.launches()
        try:
            with torch.cuda.graph(self.graph):
                self.loss = self._step()
        finally:
            functional.weight_cache_enabled(cache_was)
        self.kernels_per_step = ops.launches() - n0                   # library launches recorded in the graph (torch glue not counted)

    def _zero(self):
        for p in self.params:
            p.grad = None

    def _step(self):
        with torch.autocast('cuda', dtype=torch.bfloat16, enabled=self.autocast):
            loss = self.model(self.text, self.image, return_loss=True, **self.kw)
        loss.backward()
        if self.optimizer is not None:
            self.optimizer.step()
        return loss.detach()

    def __call__(self, text=None, image=None):
        if text is not None:
            self.text.copy_(text, non_blocking=True)
        if image is not None:
            self.image.copy_(image, non_blocking=True)
        self.graph.replay()
        return self.loss

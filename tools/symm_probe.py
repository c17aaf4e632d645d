import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE']); lr = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr)
dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
t = symm_mem.empty(1 << 20, dtype=torch.float32, device=torch.device('cuda', lr))
hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
if rank == 0:
    print('attrs', [a for a in dir(hdl) if not a.startswith('_')])
    print('multicast_ptr', hex(hdl.multicast_ptr), 'buffer_ptrs', [hex(p) for p in hdl.buffer_ptrs], 'signal_pad_ptrs', [hex(p) for p in hdl.signal_pad_ptrs][:2])
    print('has multimem ops', hasattr(torch.ops.symm_mem, 'multimem_all_reduce_'), [o for o in dir(torch.ops.symm_mem) if not o.startswith('_')][:30])
t.fill_(rank + 1.0)
hdl.barrier()
if hdl.multicast_ptr:
    torch.ops.symm_mem.multimem_all_reduce_(t, 'sum', dist.group.WORLD.group_name)
    torch.cuda.synchronize()
    if rank == 0: print('multimem allreduce ->', t[:4].tolist(), 'expect', sum(range(1, world + 1)))
    # bandwidth
    big = symm_mem.empty(256 << 20, dtype=torch.float32, device=torch.device('cuda', lr))   # 1 GiB
    symm_mem.rendezvous(big, dist.group.WORLD.group_name)
    for fn, name in ((lambda: torch.ops.symm_mem.multimem_all_reduce_(big, 'sum', dist.group.WORLD.group_name), 'multimem'), (lambda: dist.all_reduce(big), 'nccl')):
        for _ in range(2): fn()
        torch.cuda.synchronize(); s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5): fn()
        e.record(); torch.cuda.synchronize()
        if rank == 0: print(name, 'all_reduce 1 GiB:', s.elapsed_time(e) / 5, 'ms')
dist.barrier(); dist.destroy_process_group()

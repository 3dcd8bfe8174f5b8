"""Probe: does torch's symmetric memory give a multicast (NVLS) mapping on this box?  torchrun --nproc-per-node N."""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
dist.init_process_group('nccl')
t = symm_mem.empty(1 << 20, dtype=torch.float32, device='cuda')
t.fill_(rank + 1)
hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
print(rank, 'multicast_ptr', hex(hdl.multicast_ptr), 'buffer_ptrs', [hex(p) for p in hdl.buffer_ptrs][:3],
      'signal_pad_ptrs', [hex(p) for p in hdl.signal_pad_ptrs][:2], 'has', [a for a in dir(hdl) if not a.startswith('_')])
hdl.barrier()
if hdl.multicast_ptr:
    out = torch.ops.symm_mem.multimem_all_reduce_(t, 'sum', dist.group.WORLD.group_name)
    torch.cuda.synchronize()
    print(rank, 'multimem all_reduce ->', float(t[0]), 'expected', world * (world + 1) / 2)
dist.destroy_process_group()

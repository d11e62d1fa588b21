"""Probe (run under torchrun, >= 2 ranks): what torch's symmetric memory offers on this box -- peer pointers,
NVLS multicast, barrier cost -- next to NCCL all-reduce / reduce-scatter / all-gather of the 236 MB gradient."""
import os
import time
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
n = 59_000_000


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


try:
    buf = symm.empty(n, dtype=torch.float32, device=dev)
    hdl = symm.rendezvous(buf, dist.group.WORLD.group_name)
    print(f'[{rank}] symm ok: world {hdl.world_size} rank {hdl.rank} multicast_ptr {hdl.multicast_ptr:#x} '
          f'buffer_ptrs {[hex(p) for p in hdl.buffer_ptrs]}', flush=True)
    buf.fill_(float(rank + 1))
    hdl.barrier()
    peer = hdl.get_buffer((rank + 1) % world, (n,), torch.float32)
    torch.cuda.synchronize()
    print(f'[{rank}] peer value {float(peer[12345])} (expect {(rank + 1) % world + 1})', flush=True)
    local_copy = torch.empty(n, device=dev)
    ms = timeit(lambda: local_copy.copy_(peer))
    if rank == 0: print(f'p2p read 236 MB from peer (copy kernel/DMA): {ms:.3f} ms  {n * 4 / ms / 1e6:.0f} GB/s', flush=True)
    ms = timeit(lambda: peer.copy_(local_copy))
    if rank == 0: print(f'p2p write 236 MB to peer: {ms:.3f} ms  {n * 4 / ms / 1e6:.0f} GB/s', flush=True)
    ms = timeit(lambda: hdl.barrier(), reps=50)
    if rank == 0: print(f'symm barrier: {ms * 1000:.1f} us', flush=True)
    ms = timeit(lambda: local_copy.add_(peer))
    if rank == 0: print(f'add_ from peer (SM loads over NVLink): {ms:.3f} ms  {n * 4 / ms / 1e6:.0f} GB/s', flush=True)
except Exception as e:  # noqa: BLE001
    import traceback
    print(f'[{rank}] symmetric memory failed: {e!r}', flush=True)
    traceback.print_exc()
dist.barrier()
dist.destroy_process_group()

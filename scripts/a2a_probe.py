"""Probe (torchrun): raw NCCL all_to_all_single bandwidth, and whether symmetric memory + peer stores work."""
import os, time, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
rows, F = 5_400_000, 14
x = torch.randn(rows, F, device=dev); y = torch.empty_like(x)
splits = [rows // world] * world; splits[-1] += rows - sum(splits)
for _ in range(5): dist.all_to_all_single(y, x, splits, splits)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): dist.all_to_all_single(y, x, splits, splits)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
out_bytes = rows * F * 4 * (world - 1) / world
if rank == 0: print(f"all_to_all_single: {ms:.3f} ms, {out_bytes/1e6:.0f} MB sent off-rank per rank -> {out_bytes/ms/1e6:.1f} GB/s per direction")
# copy baseline
e0.record()
for _ in range(20): y.copy_(x)
e1.record(); torch.cuda.synchronize()
if rank == 0: print(f"local copy of the same buffer: {e0.elapsed_time(e1)/20:.3f} ms")
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(rows * F, dtype=torch.float32, device=dev)
    h = symm.rendezvous(t, dist.group.WORLD)
    if rank == 0: print("symmetric memory ok: world", h.world_size, "ptrs", [hex(p) for p in h.buffer_ptrs][:4], "signal pad", len(h.signal_pad_ptrs))
    peer = (rank + 1) % world
    pt = h.get_buffer(peer, (rows * F,), torch.float32)
    h.barrier()
    e0.record()
    for _ in range(20): pt.copy_(x.view(-1))
    e1.record(); torch.cuda.synchronize(); h.barrier()
    if rank == 0: print(f"peer store (copy_ into peer buffer) {rows*F*4/1e6:.0f} MB: {e0.elapsed_time(e1)/20:.3f} ms -> {rows*F*4/(e0.elapsed_time(e1)/20)/1e6:.1f} GB/s")
    t0 = time.perf_counter()
    for _ in range(50): h.barrier()
    torch.cuda.synchronize()
    if rank == 0: print(f"symm barrier: {(time.perf_counter()-t0)/50*1e6:.1f} us")
except Exception as ex:
    if rank == 0: print("symmetric memory failed:", type(ex).__name__, ex)
dist.destroy_process_group()

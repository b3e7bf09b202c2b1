"""Single-GPU timing of the stripe-routing kernels that need no peer (count, scan, pack into the own buffer, unpack):
usage: python scripts/stripe_micro.py [M] [world]"""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pipeline_pointcloud_b200 import _lib as L

M = int(sys.argv[1]) if len(sys.argv) > 1 else 3_750_000
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda")
W_, H_ = 3840, 2160
tw, th = (W_ + 15) // 16, (H_ + 15) // 16
g = torch.Generator(device="cuda").manual_seed(0)
means2d = torch.rand(M, 2, device=dev, generator=g) * torch.tensor([W_, H_], device=dev)
radii = torch.randint(1, 12, (M,), device=dev, generator=g, dtype=torch.int32)
radii[torch.rand(M, device=dev, generator=g) < 0.08] = 0
depths, conics, opac, cols = torch.rand(M, device=dev), torch.rand(M, 3, device=dev), torch.rand(M, device=dev), torch.rand(M, 3, device=dev)
lib, st = L.lib(), L.stream_ptr(dev)
n_ctas = int(lib.b2s_stripe_ctas(M))
mask = torch.empty(M, dtype=torch.int16, device=dev)
cta = torch.empty(3 * world * n_ctas, dtype=torch.int32, device=dev)
totals = torch.empty(world * 4, dtype=torch.int64, device=dev)
sb = L.Stripes(world=world, cams_per_rank=0)
for d in range(world + 1):
    sb.bounds[d] = d * th // world


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


count = lambda: L.check(lib.b2s_stripe_count(M, M, L.ptr(means2d), L.ptr(radii), 16, tw, th, ctypes.byref(sb), L.ptr(mask), L.ptr(cta), L.ptr(totals), M, st), "count")
scan = lambda: L.check(lib.b2s_stripe_scan(M, world, L.ptr(cta), L.ptr(totals), st), "scan")
print("count ms", timed(count))
count(); print("scan ms", timed(scan))
count(); scan()
tot = totals.view(world, 4).tolist()
n_rows = sum(t[0] for t in tot)
F = 13
buf = torch.empty((n_rows + 8) * F, device=dev)
base, run = [], 0
for d in range(world):
    base.append(run); run += tot[d][0]
ptrs = (ctypes.c_void_p * world)(*[buf.data_ptr()] * world)
b64 = (ctypes.c_int64 * world)(*base)
pack = lambda: L.check(lib.b2s_stripe_pack(M, M, 3, 0, L.ptr(means2d), L.ptr(depths), L.ptr(conics), L.ptr(opac), L.ptr(cols), L.ptr(radii), L.ptr(mask), L.ptr(cta), world, 0, b64, ptrs, st), "pack")
print("pack (local buffer) ms", timed(pack), "rows", n_rows, "of", M)
n = tot[0][0]
o = dict(device=dev)
outs = [torch.empty(n, 2, **o), torch.empty(n, **o), torch.empty(n, 3, **o), torch.empty(n, **o), torch.empty(n, 3, **o),
        torch.empty(n, dtype=torch.int32, **o), torch.empty(n, dtype=torch.int64, **o), torch.empty(n, dtype=torch.int64, **o),
        torch.empty(n, dtype=torch.int32, **o), torch.empty(n, dtype=torch.int32, **o), torch.empty(n, dtype=torch.int32, **o)]
rows = L.TileRows(0, sb.bounds[1])
unpack = lambda: L.check(lib.b2s_stripe_unpack(n, 3, L.ptr(buf), 16, tw, th, ctypes.byref(rows), *[L.ptr(t) for t in outs], st), "unpack")
print("unpack ms", timed(unpack), "rows", n)

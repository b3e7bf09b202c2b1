"""Diagnostic: per-step times and allocator statistics of the packed=True + meta-read training step (config 3)."""
import gc, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pipeline_pointcloud_b200 as b2s
from pipeline_pointcloud_b200.scene import make_scene

N, deg, W, H = 5_800_000, 3, 1297, 840
s = make_scene(N, deg, 1, W, H, seed=0)
dev = torch.device("cuda")
p = {k: s[k].to(dev).requires_grad_() for k in ("means", "quats", "scales", "opacities", "colors")}
V, K = s["viewmats"].to(dev), s["Ks"].to(dev)
g = torch.Generator().manual_seed(1)
vr, va = torch.rand(1, H, W, 3, generator=g).to(dev), torch.rand(1, H, W, 1, generator=g).to(dev)

def step(touch):
    for t in p.values(): t.grad = None
    r, a, m = b2s.rasterization(p["means"], p["quats"], p["scales"], p["opacities"], p["colors"], V, K, W, H, sh_degree=deg,
                                packed=True, rasterize_mode="antialiased")
    if touch: m["means2d"].retain_grad()
    torch.autograd.backward([r, a], [vr, va])
    if touch: assert m["means2d"].grad is not None

for label, touch, gcoff in (("no_meta", False, False), ("meta", True, False), ("meta_gc_disabled", True, True)):
    if gcoff: gc.disable()
    for _ in range(5): step(touch)
    torch.cuda.synchronize()
    st0 = torch.cuda.memory_stats()
    ts = []
    for _ in range(30):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); t0 = time.perf_counter(); step(touch); t1 = time.perf_counter(); e1.record(); torch.cuda.synchronize()
        ts.append((round(e0.elapsed_time(e1), 2), round((t1 - t0) * 1e3, 2)))
    st1 = torch.cuda.memory_stats()
    print(label, "gpu/host ms per step:", ts)
    print("   cudaMalloc calls:", st1["num_device_alloc"] - st0["num_device_alloc"], "frees:", st1["num_device_free"] - st0["num_device_free"],
          "retries:", st1["num_alloc_retries"] - st0["num_alloc_retries"], "reserved GB:", round(st1["reserved_bytes.all.current"] / 1e9, 2),
          "gc counts:", gc.get_count())
    if gcoff: gc.enable()

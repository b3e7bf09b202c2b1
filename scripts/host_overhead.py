"""Diagnostic: host enqueue time vs GPU time of one rasterization fwd+bwd step (config 3)."""
import sys, time, os
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

def fwd():
    return b2s.rasterization(p["means"], p["quats"], p["scales"], p["opacities"], p["colors"], V, K, W, H, sh_degree=deg,
                             packed=False, rasterize_mode="antialiased")
def step():
    for t in p.values(): t.grad = None
    r, a, m = fwd()
    t1 = time.perf_counter()
    torch.autograd.backward([r, a], [vr, va])
    return t1

for _ in range(5): step()
torch.cuda.synchronize()
host_f, host_b, tot = [], [], []
for _ in range(20):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    t1 = step()
    t2 = time.perf_counter(); e1.record()
    torch.cuda.synchronize()
    host_f.append((t1 - t0) * 1e3); host_b.append((t2 - t1) * 1e3); tot.append(e0.elapsed_time(e1))
med = lambda x: sorted(x)[len(x) // 2]
print(f"host fwd enqueue (incl. its sync) {med(host_f):.3f} ms | host bwd enqueue {med(host_b):.3f} ms | GPU step {med(tot):.3f} ms")
# back-to-back without per-step sync
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): step()
e1.record(); torch.cuda.synchronize()
print(f"back-to-back: {e0.elapsed_time(e1)/20:.3f} ms/step")
# forward only
torch.cuda.synchronize(); e0.record()
with torch.no_grad():
    for _ in range(20): fwd()
e1.record(); torch.cuda.synchronize()
print(f"forward only: {e0.elapsed_time(e1)/20:.3f} ms")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
with torch.no_grad():
    for _ in range(10): fwd()
pr.disable(); torch.cuda.synchronize()
print("---- forward only ----")
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
pr = cProfile.Profile(); pr.enable()
for _ in range(10): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)

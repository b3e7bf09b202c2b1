"""Diagnostic: how much of each tile's depth-sorted list does compositing actually consume (config 3)?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pipeline_pointcloud_b200 as b2s
from pipeline_pointcloud_b200.scene import make_scene
N, deg, W, H = 5_800_000, 3, 1297, 840
s = make_scene(N, deg, 1, W, H, seed=0)
dev = torch.device("cuda")
p = {k: s[k].to(dev).requires_grad_() for k in ("means", "quats", "scales", "opacities", "colors")}
r, a, m = b2s.rasterization(p["means"], p["quats"], p["scales"], p["opacities"], p["colors"], s["viewmats"].to(dev), s["Ks"].to(dev),
                            W, H, sh_degree=deg, packed=False, rasterize_mode="antialiased")
saved = r.grad_fn.saved_tensors
last_ids, final_T = saved[8], saved[9]
off = m["isect_offsets"].reshape(-1).long()
n_is = m["flatten_ids"].numel()
hi = torch.cat([off[1:], torch.tensor([n_is], device=dev)])
th, tw = m["tile_height"], m["tile_width"]
# per-tile max last id
li = last_ids[0].long()
Hp, Wp = th * 16, tw * 16
pad = torch.full((Hp, Wp), -1, device=dev, dtype=torch.long); pad[:H, :W] = li
tile_last = pad.reshape(th, 16, tw, 16).permute(0, 2, 1, 3).reshape(th * tw, 256).max(dim=1).values
used = (tile_last + 1 - off).clamp_min(0)
length = hi - off
print("tiles", th * tw, "n_isects", n_is, "mean list length", length.float().mean().item(), "mean consumed", used.float().mean().item(),
      "consumed fraction of all isects", used.sum().item() / n_is)
print("pixels saturated (T<=1e-4 stop):", (final_T < 0).float().mean().item(), " mean alpha", a.mean().item())
q = torch.tensor([0.1, 0.5, 0.9, 0.99], device=dev)
print("consumed quantiles", torch.quantile(used.float(), q).tolist(), "length quantiles", torch.quantile(length.float(), q).tolist())
# depth rank view: what fraction of visible Gaussians (by depth order) is ever touched?
flat = m["flatten_ids"].long()
touched = torch.zeros(N, dtype=torch.bool, device=dev)
idx = torch.arange(n_is, device=dev)
tile_of = torch.searchsorted(off, idx, right=True) - 1
mask = idx <= tile_last[tile_of]
touched[flat[mask]] = True
print("Gaussians touched by any tile:", touched.sum().item(), "of visible", (m["radii"] > 0).sum().item())
# tiles that never finish (contain a pixel that does not saturate) and the isects they own
sat = torch.ones(Hp, Wp, dtype=torch.bool, device=dev); sat[:H, :W] = (final_T[0] < 0)
tile_all_sat = sat.reshape(th, 16, tw, 16).permute(0, 2, 1, 3).reshape(th * tw, 256).all(dim=1)
print("tiles fully saturated:", tile_all_sat.float().mean().item(), " isects in unsaturated tiles:", length[~tile_all_sat].sum().item(),
      "(", length[~tile_all_sat].sum().item() / n_is, "of all )")
# if the nearest fraction f of visible Gaussians were processed first: how many tiles would be finished?
depth_rank = torch.empty(N, dtype=torch.long, device=dev)
vis = (m["radii"][0] > 0)
order = torch.argsort(torch.where(vis, m["depths"][0], torch.full_like(m["depths"][0], 1e30)))
depth_rank[order] = torch.arange(N, device=dev)
V = int(vis.sum())
rank_of_isect = depth_rank[flat]
for f in (0.03, 0.06, 0.1, 0.15, 0.25):
    R = int(f * V)
    # tile finished within chunk A iff it is fully saturated and its last consumed isect has rank < R
    last_rank = torch.where(tile_last >= off, rank_of_isect[tile_last.clamp(0, n_is - 1)], torch.zeros_like(tile_last))
    fin = tile_all_sat & (last_rank < R)
    nA = int((rank_of_isect < R).sum())
    nB = int(length[~fin].sum()) - int(((rank_of_isect < R) & (~fin[tile_of])).sum())
    print(f"f={f:.2f}: chunk A isects {nA/1e6:.2f}M, tiles finished {fin.float().mean().item():.3f}, chunk B isects {nB/1e6:.2f}M, total {(nA+nB)/1e6:.2f}M of {n_is/1e6:.1f}M")

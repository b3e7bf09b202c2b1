#!/usr/bin/env python
"""bench.py -- fwd+bwd throughput of the rasterization() hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload config3|config2|tiny]

A "step" is one `rasterization()` forward + backward (random cotangent) over one camera of the synthetic
config-3 scene (5.8 M random Gaussians, SH degree 3, 1297x840, antialiased), inputs resident in HBM.
N > 1 (torchrun): Gaussians sharded over the ranks, one camera per rank per step, projected splats routed to the camera
owners by the fused NVLink exchange (rasterization(distributed=True); weak scaling: per-GPU work fixed).
`--multi-gpu stripes --workload config5`: ONE 4K camera over N GPUs, tile-row stripes (strong scaling);
`--multi-gpu allreduce`: replicated Gaussians + dense-gradient all-reduce (BASELINE config 4).
`value` = Mpix/s of the whole job (C*W*H*steps / time, max over ranks); `iters_per_s` is reported beside it.
`e2e` = the same step driven from pinned HOST buffers: H2D of every input, the step, D2H of render, alpha and
every gradient, all inside the timed region.
`--impl reference` times the CPU oracle (the restated reference torch path) on a bounded sub-frustum sample
of the same workload on the host cores.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (N, sh_degree, W, H, rasterize_mode)
    "config3": (5_800_000, 3, 1297, 840, "antialiased"),
    "config2": (1_000_000, 3, 800, 800, "classic"),
    "config1": (100_000, 0, 256, 256, "classic"),
    "config5": (30_000_000, 3, 3840, 2160, "antialiased"),     # with --multi-gpu stripes: one 4K camera, strong scaling
    "tiny": (50_000, 3, 320, 240, "antialiased"),
}
METRIC = "3DGS rasterization fwd+bwd throughput"
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture (profiles/), config 3
NCU_TRAFFIC = {   # profiles/r01_ncu_full_v14_summary.md; launch list with DRAM bytes: profiles/r01_launches_v14.csv
    "project_fwd": 1_628_177_000,   # 1.313962 GB read + 0.314215 GB written (v12 launch list: 1.310 + 0.300)
}
UNIT = "Mpix/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ----------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------
# algorithmic bytes per stage (DESIGN.md "Measurement")
# ----------------------------------------------------------------------------------------------------------
def algorithmic_bytes(N, C, V, Vt, I, R, P, K, n_tiles, key_bits):
    """Compulsory HBM traffic per stage (every input read once, every output written once).
    V = visible rows, Vt = rows compositing actually reached (non-zero cotangent), I = tile intersections,
    R = coarse-bin records.  The compositing stages are instruction-issue bound, not HBM bound: no figure."""
    M = C * N
    p = math.ceil(key_bits / 8)
    return {
        "project_fwd": N * 44 + V * 12 * K + M * 60,               # params + SH rows of visible + per-pair outputs (incl. keys)
        # reference-shaped lists (B2S_BINNING=tiles)
        "isect_scan": M * 8,
        "isect_fill": M * 20 + I * 12,
        "isect_sort": I * 12 * 2 * p + I * 8,                        # LSD radix model; single-pass ideal = I*24
        "isect_offsets": I * 8 + 4 * n_tiles,
        "tile_bin": M * 20 + I * 8 + I * 4 + math.ceil((key_bits - 32) / 8) * I * 16 + I * 4 + 4 * n_tiles,
        # depth sort of the rows: histogram read + 4 digit passes over (key, id) pairs, read + write each
        "depth_order": M * 4 + 4 * M * 16,
        # coarse binning: sorted ids + gathered (means2d, radius) in, records out; histogram; one digit pass; offsets
        "coarse_bin": M * 4 + V * 12 + R * 8 + R * 4 + R * 16 + 4 * (n_tiles // 64 + 1),
        "composite_fwd": None,
        "composite_bwd": None,
        # SH colour VJP: means + radii + colour cotangents in, coefficient rows of the reached rows in, all gradients out
        "sh_bwd": N * 12 + M * 4 + V * 12 + Vt * 12 * K + N * 12 * K + N * 12,
        # projection VJP: radii + cotangents of the visible rows, parameters of the reached rows, gradients out
        "project_bwd": M * 4 + V * 36 + Vt * 44 + N * 44,
    }


# ----------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle on a bounded sub-frustum sample of the same workload
# ----------------------------------------------------------------------------------------------------------
def cpu_sample_inputs(scene, W, H, win_w=160, win_h=96):
    """Central win_w x win_h window of the workload's camera 0 and the Gaussians whose 3-sigma footprint can
    reach it (selected with a cheap bound, outside the timed region).  All stages then scale by the same
    factor, so Mpix/s of the sample is comparable with the full frame."""
    V, K = scene["viewmats"][0], scene["Ks"][0]
    x0, y0 = (W - win_w) // 2 // 16 * 16, (H - win_h) // 2 // 16 * 16
    t = scene["means"] @ V[:3, :3].T + V[:3, 3]
    z = t[:, 2].clamp_min(1e-6)
    px, py = K[0, 0] * t[:, 0] / z + K[0, 2], K[1, 1] * t[:, 1] / z + K[1, 2]
    rad = 3.0 * scene["scales"].max(dim=1).values * K[0, 0] / z + 3.0
    keep = (t[:, 2] > 0.01) & (px + rad > x0) & (px - rad < x0 + win_w) & (py + rad > y0) & (py - rad < y0 + win_h)
    sub = {k: scene[k][keep].clone() for k in ("means", "quats", "scales", "opacities", "colors")}
    Kc = K.clone()
    Kc[0, 2] -= x0
    Kc[1, 2] -= y0
    sub["viewmats"], sub["Ks"] = V[None].clone(), Kc[None]
    return sub, win_w, win_h, int(keep.sum())


def cpu_oracle_step(sub, w, h, sh_degree, mode, v_r, v_a):
    from oracle import torch_oracle as O   # the reference arm / cpu_baseline leg: the one place bench.py runs oracle/
    leaves = [sub[k].clone().requires_grad_() for k in ("means", "quats", "scales", "opacities", "colors")]
    r, a, m = O.rasterization(*leaves, sub["viewmats"], sub["Ks"], w, h, sh_degree=sh_degree, packed=False,
                              rasterize_mode=mode)
    torch.autograd.backward([r, a], [v_r, v_a])
    return m["isect_ids"].numel()


def run_cpu(workload, steps, warmup, seed=0):
    from pipeline_pointcloud_b200.scene import make_scene
    N, deg, W, H, mode = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if N > 10_000_000:
        from pipeline_pointcloud_b200.scene import make_scene_blocked
        scene = make_scene_blocked(N, deg, 1, W, H, seed=seed)
    else:
        scene = make_scene(N, deg, 1, W, H, seed=seed)
    sub, w, h, n_sub = cpu_sample_inputs(scene, W, H)
    g = torch.Generator().manual_seed(seed + 1)
    v_r, v_a = torch.rand(1, h, w, 3, generator=g), torch.rand(1, h, w, 1, generator=g)
    for _ in range(warmup):
        cpu_oracle_step(sub, w, h, deg, mode, v_r, v_a)
    t0 = time.perf_counter()
    for _ in range(steps):
        n_is = cpu_oracle_step(sub, w, h, deg, mode, v_r, v_a)
    dt = time.perf_counter() - t0
    mpix = w * h * steps / dt / 1e6
    sample = (f"{steps} fwd+bwd step(s) of the CPU oracle on a sub-frustum of {workload}: central {w}x{h} window "
              f"({(w // 16) * (h // 16)} of {math.ceil(W / 16) * math.ceil(H / 16)} tiles) with the {n_sub} of {N} Gaussians "
              f"that can reach it ({n_is} tile intersections); {cores} torch threads")
    model = ""
    try:
        model = [l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    except Exception:
        pass
    return {"value": mpix, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "cpu_model": model,
            "s_per_step": dt / steps, "iters_per_s_full_frame_equiv": mpix * 1e6 / (W * H)}


# ----------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b2s", choices=["b2s", "reference"])
    ap.add_argument("--workload", default=os.environ.get("B2S_BENCH_WORKLOAD", "config3"), choices=sorted(WORKLOADS))
    ap.add_argument("--packed", action="store_true", help="packed=True layout (default: dense, as ns-train splatfacto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs only)")
    ap.add_argument("--no-balance", action="store_true", help="stripes: keep the even split of the tile rows")
    ap.add_argument("--multi-gpu", default="sharded", choices=["sharded", "allreduce", "stripes"],
                    help="N>1: 'sharded' = Gaussians sharded over ranks + all-to-all of projected splats to the camera "
                         "owners (rasterization(distributed=True), what the reference's multi-GPU branch runs); "
                         "'allreduce' = replicated Gaussians + dense-gradient all-reduce (BASELINE config 4); "
                         "'stripes' = ONE camera, Gaussians sharded over ranks, every rank composites a stripe of tile rows "
                         "(rasterization(distributed=True, tile_stripes=True), BASELINE config 5; strong scaling)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    N, deg, W, H, mode = WORKLOADS[args.workload]

    # ------------------------------------------------------------------ reference arm: CPU oracle, rank 0 only
    if args.impl == "reference":
        if rank != 0:
            return 0
        steps, warmup = max(1, min(args.steps, 30)), min(args.warmup, 2)
        cb = run_cpu(args.workload, steps, warmup)
        line = {
            "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": cb["s_per_step"] * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {N} Gaussians, SH deg {deg}, {W}x{H}, {mode}, fwd+bwd "
                                   "(bounded sub-frustum sample, see cpu_baseline.sample)"},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "cpu_model")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "iters_per_s": cb["iters_per_s_full_frame_equiv"],
        }
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm
    import torch.distributed as dist
    import pipeline_pointcloud_b200 as b2s
    from pipeline_pointcloud_b200 import ops
    from pipeline_pointcloud_b200.distributed import allreduce_gradients
    from pipeline_pointcloud_b200.scene import make_scene

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"

    stripes = world > 1 and args.multi_gpu == "stripes"
    sharded = world > 1 and args.multi_gpu in ("sharded", "stripes")
    C_total = 1 if (stripes or args.multi_gpu == "stripes") else max(world, 1)
    lo_n, hi_n = (rank * N // world, (rank + 1) * N // world) if sharded else (0, N)
    if N > 10_000_000:      # too large to generate once per rank: block-seeded scene, only this rank's shard
        from pipeline_pointcloud_b200.scene import make_scene_blocked
        scene = make_scene_blocked(N, deg, C_total, W, H, seed=0, lo=lo_n, hi=hi_n)
        leaves_host = {k: scene[k].pin_memory() for k in ("means", "quats", "scales", "opacities", "colors")}
    else:
        scene = make_scene(N, deg, C_total, W, H, seed=0)            # identical on every rank (CPU generator)
        leaves_host = {k: scene[k][lo_n:hi_n].contiguous().pin_memory() for k in ("means", "quats", "scales", "opacities", "colors")}
    cam_i = 0 if C_total == 1 else rank
    cam_host = {"viewmats": scene["viewmats"][cam_i:cam_i + 1].contiguous().pin_memory(),
                "Ks": scene["Ks"][cam_i:cam_i + 1].contiguous().pin_memory()}
    params = {k: v.to(dev).requires_grad_() for k, v in leaves_host.items()}
    cams = {k: v.to(dev) for k, v in cam_host.items()}
    g = torch.Generator().manual_seed(1 + (0 if stripes else rank))
    v_r_full = torch.rand(1, H, W, 3, generator=g)
    v_a_full = torch.rand(1, H, W, 1, generator=g)
    kw = dict(sh_degree=deg, packed=args.packed or sharded, rasterize_mode=mode, distributed=sharded)
    cot = {}

    def set_cotangents(bounds=None):
        """Every rank differentiates its own pixel rows (its stripe) of the same image cotangent."""
        vr, va = v_r_full, v_a_full
        if bounds is not None:
            y0, y1 = min(H, bounds[rank] * 16), min(H, bounds[rank + 1] * 16)
            vr, va = vr[:, y0:y1].contiguous(), va[:, y0:y1].contiguous()
            kw["stripe_rows"] = list(bounds)
        cot["v_r_host"], cot["v_a_host"] = vr.pin_memory(), va.pin_memory()
        cot["v_r"], cot["v_a"] = cot["v_r_host"].to(dev), cot["v_a_host"].to(dev)

    stripe_rows = None
    if stripes:
        from pipeline_pointcloud_b200.distributed import rebalance_stripe_bounds, stripe_bounds
        kw["tile_stripes"] = True
        stripe_rows = stripe_bounds(math.ceil(H / 16), world)
    set_cotangents(stripe_rows)
    v_r_host, v_a_host, v_r, v_a = cot["v_r_host"], cot["v_a_host"], cot["v_r"], cot["v_a"]
    plist = list(params.values())
    stats = {}

    def step(p=params, c=cams, vr=None, va=None):
        vr = cot["v_r"] if vr is None else vr
        va = cot["v_a"] if va is None else va
        for t in p.values():
            t.grad = None
        r, a, m = b2s.rasterization(p["means"], p["quats"], p["scales"], p["opacities"], p["colors"], c["viewmats"],
                                    c["Ks"], W, H, **kw)
        torch.autograd.backward([r, a], [vr, va])
        if world > 1 and not sharded:
            allreduce_gradients(p.values())
        return r, a, m

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    balance_log = []
    if stripes and not args.no_balance:
        # feedback load balancing of the stripes (untimed, part of the warm-up): every rank measures its stripe-dependent
        # stages, the bounds are cut again at equal-cost quantiles; three rounds
        STRIPE_STAGES = ("stripe_unpack", "depth_order", "coarse_bin", "composite_fwd", "composite_bwd", "stripe_return")
        step(); step()
        for _ in range(3):
            ops.TIMER = ops.StageTimer()
            step(); step()
            tm, ops.TIMER = ops.TIMER, None
            mine = sum(v["ms_avg"] for k, v in tm.summary().items() if k in STRIPE_STAGES)
            allt = torch.zeros(world, device=dev)
            dist.all_gather_into_tensor(allt, torch.tensor([mine], device=dev))
            costs = allt.tolist()
            balance_log.append({"bounds": list(stripe_rows), "ms": [round(c, 3) for c in costs]})
            stripe_rows = rebalance_stripe_bounds(stripe_rows, costs)
            set_cotangents(stripe_rows)
        v_r_host, v_a_host, v_r, v_a = cot["v_r_host"], cot["v_a_host"], cot["v_r"], cot["v_a"]
    for _ in range(max(args.warmup, 3)):
        r, a, m = step()
    torch.cuda.synchronize()
    m_last = m
    n_visible = int((m["radii"] > 0).sum()) if not (args.packed or sharded) else int(m["means2d"].shape[0])
    stats["n_isects"] = int(m["tiles_per_gauss"].sum(dtype=torch.int64))   # (the per-intersection lists of meta are lazy)
    # rows compositing reached before saturation (outside the timed region): non-zero opacity cotangent
    for t in params.values():
        t.grad = None
    r, a, m = b2s.rasterization(params["means"], params["quats"], params["scales"], params["opacities"], params["colors"],
                                cams["viewmats"], cams["Ks"], W, H, **kw)
    m["opacities"].retain_grad()
    torch.autograd.backward([r, a], [v_r, v_a])
    stats["n_reached"] = int((m["opacities"].grad != 0).sum())
    del r, a, m

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ops.TIMER = ops.StageTimer()
    ms_total = timed(step, args.steps)
    timer, ops.TIMER = ops.TIMER, None
    clocks = sampler.stop() if rank == 0 else None
    stage = timer.summary()
    ms_per_step = ms_total / args.steps
    pix_per_step = C_total * W * H
    value = pix_per_step / (ms_per_step * 1e-3) / 1e6

    # ---- e2e: host buffers in, host buffers out -----------------------------------------------------------
    h2d = sum(v.numel() * 4 for v in leaves_host.values()) + sum(v.numel() * 4 for v in cam_host.values()) \
        + v_r_host.numel() * 4 + v_a_host.numel() * 4
    d2h = (v_r_host.numel() + v_a_host.numel()) * 4 + sum(v.numel() * 4 for v in leaves_host.values())

    # Two slots, three streams: the H2D of step i+1 runs while step i computes and while its results go back (PCIe is
    # full duplex), so a step costs max(H2D, D2H, compute) instead of their sum.  Every step still moves all of its
    # inputs in and all of its results out.
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    main = torch.cuda.current_stream(dev)
    slots = []
    for _ in range(2):
        slots.append({
            "p": {k: torch.empty_like(v, device=dev).requires_grad_() for k, v in leaves_host.items()},
            "c": {k: torch.empty_like(v, device=dev) for k, v in cam_host.items()},
            "vr": torch.empty_like(v_r), "va": torch.empty_like(v_a),
            "out": {"render": torch.empty(v_r_host.shape).pin_memory(), "alpha": torch.empty(v_a_host.shape).pin_memory()},
            "grads": {k: torch.empty_like(v).pin_memory() for k, v in leaves_host.items()},
            "ev_in": torch.cuda.Event(), "ev_done": torch.cuda.Event(), "ev_out": None, "keep": None,
        })
    state = {"i": 0}

    def e2e_step():
        sl = slots[state["i"] % 2]
        state["i"] += 1
        if sl["ev_out"] is not None:
            sl["ev_out"].synchronize()          # the caller has read this slot's host buffers; its device buffers are free
        with torch.cuda.stream(s_in), torch.no_grad():
            for k, v in leaves_host.items():
                sl["p"][k].copy_(v, non_blocking=True)
            for k, v in cam_host.items():
                sl["c"][k].copy_(v, non_blocking=True)
            sl["vr"].copy_(v_r_host, non_blocking=True)
            sl["va"].copy_(v_a_host, non_blocking=True)
            sl["ev_in"].record(s_in)
        main.wait_event(sl["ev_in"])
        r, a, _ = step(sl["p"], sl["c"], sl["vr"], sl["va"])
        sl["ev_done"].record(main)
        sl["keep"] = (r, a)                      # keep the results alive until their copy has run
        with torch.cuda.stream(s_out), torch.no_grad():
            s_out.wait_event(sl["ev_done"])
            sl["out"]["render"].copy_(r.detach(), non_blocking=True)
            sl["out"]["alpha"].copy_(a.detach(), non_blocking=True)
            for k in sl["grads"]:
                sl["grads"][k].copy_(sl["p"][k].grad, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(s_out)
            sl["ev_out"] = ev

    def e2e_run(steps):
        for _ in range(steps):
            e2e_step()
        for sl in slots:                         # the last results are on the host when the timed region ends
            if sl["ev_out"] is not None:
                sl["ev_out"].synchronize()

    e2e_steps = args.e2e_steps or max(4, min(args.steps, 10))
    if args.no_e2e:
        e2e_ms, e2e_value = None, None
    else:
        e2e_run(2)
        barrier()
        t0 = time.perf_counter()                 # three streams: wall clock between device-wide synchronisations
        e2e_run(e2e_steps)
        barrier()
        ms = torch.tensor([(time.perf_counter() - t0) * 1e3], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        e2e_ms = float(ms.item()) / e2e_steps
        e2e_value = pix_per_step / (e2e_ms * 1e-3) / 1e6

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant HBM-bound kernel ----------------------------------------------------------
    peak, peak_src = peaks()
    I = stats["n_isects"]
    tiles = math.ceil(W / 16) * math.ceil(H / 16)
    key_bits = 32 + (tiles).bit_length() + (1).bit_length()
    K = (deg + 1) ** 2
    n_cam_local = C_total if (world > 1 and sharded) else 1
    if stripes:      # per-rank counters -> whole-job totals are not gathered; report rank 0's
        pass
    R = int(m_last.get("n_coarse_records") or 0) if m_last is not None else 0
    ab = algorithmic_bytes(hi_n - lo_n, n_cam_local, n_visible, stats["n_reached"], I, R, W * H, K, tiles, key_bits)
    per_stage = {}
    for name, st_ in stage.items():
        b = ab.get(name)
        gbs = b / (st_["ms_avg"] * 1e-3) / 1e9 if (b and st_["ms_avg"] > 0) else None
        per_stage[name] = {"ms": round(st_["ms_avg"], 4), "calls_per_step": st_["calls"] / args.steps, "algorithmic_bytes": b,
                           "GBps": round(gbs, 1) if gbs else None, "frac_of_peak": round(gbs / peak, 4) if gbs else None}
    hbm_stages = {k: v for k, v in per_stage.items() if v["GBps"] and k in ("project_fwd", "sh_bwd", "project_bwd")}
    dom = max(hbm_stages, key=lambda k: hbm_stages[k]["ms"])
    kernel_of = {"project_fwd": "project_fwd_kernel", "sh_bwd": "sh_bwd_fused_kernel", "project_bwd": "project_bwd_kernel"}
    roofline = {"kernel": kernel_of[dom], "stage": dom, "bound": "hbm", "achieved": per_stage[dom]["GBps"], "peak": peak,
                "unit": "GB/s", "frac": per_stage[dom]["frac_of_peak"], "traffic": NCU_TRAFFIC.get(dom) if args.workload == "config3" and world == 1 else None, "peak_source": peak_src,
                "note": "largest HBM-bound kernel of the step (one launch per step; CUDA events around the launch, inside the "
                        "timed region); the compositing kernels are instruction-issue bound (see profiles/), the sorts are "
                        "L2-resident at this size; per-stage figures under 'stages'"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if args.multi_gpu == "stripes" else "weak",
        "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"{args.workload}: {N} random Gaussians, SH deg {deg}, {W}x{H}, {mode}, fwd+bwd, "
                               f"packed={args.packed or sharded}, " + ("1 camera" if C_total == 1 else "1 camera per GPU") +
                               ((", Gaussians sharded over ranks, every rank composites a stripe of tile rows of the one camera: "
                                 "fused pack + NVLink peer-store exchange of projected splats to the stripe owners, gradients "
                                 "returned by the reverse exchange and summed at the owner") if stripes else
                                (", Gaussians sharded over ranks, NCCL all-to-all of projected splats to the camera owners, "
                                 "reverse all-to-all of gradients") if sharded else
                                (", replicated Gaussians, dense-gradient all-reduce" if world > 1 else "")),
                   "n_isects": I, "n_visible": n_visible, "n_reached": stats["n_reached"], "n_coarse_records": R,
                   "l2": "inputs (1.4 GB parameters + 0.3 GB keys) exceed the 126 MB L2",
                   "parallelism": f"dp{world}", "stripe_rows": stripe_rows, "stripe_balance_rounds": balance_log},
        "iters_per_s": 1e3 / ms_per_step * C_total,
        "gpu_launches": timer.launches,
        "gpu_launches_note": "kernels of libb2s.so launched inside the timed region (per step: project_fwd 1, depth sort 6, "
                             "coarse binning 5, composite_fwd 1, composite_bwd 1, sh_bwd 4 (classify, dense [returns at once in "
                             "sparse mode], zero fill, list-driven), project_bwd 2; all hand-written, no library kernels)",
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "note": "rasterization() + backward driven from pinned host buffers: every step copies all "
                "inputs in and render, alpha and all gradients out; copies of neighbouring steps overlap (2 slots, 3 streams)"},
        "roofline": roofline, "stages": per_stage, "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        cb = run_cpu(args.workload, 1, 0)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "cpu_model")}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

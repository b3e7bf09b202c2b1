"""Checkpoint -> ``.ply`` export (first half): what the reference's exporter component does after
the multi-GPU trainer -- ``gsplat_pt_to_ply.py CKPT_DIR OUT.ply``
(source/container/src/pipeline/post_processing/gsplat_pt_to_ply.py:31-75, invoked from
source/container/src/main.py:1388-1404): take the LAST checkpoint of the directory and write the 3DGS PLY with the
property order ``x y z nx ny nz f_dc_0..2 f_rest_0..(3(K-1)-1) opacity scale_0..2 rot_0..3``, SH coefficients
channel-major (``[N,K,3] -> transpose(1,2) -> flatten``), all float32, ``binary_little_endian 1.0``, one interleaved row
per Gaussian (nerfstudio ``ExportGaussianSplat.write_ply``, which the reference script calls).
"""
from __future__ import annotations

import os
import sys
from collections import OrderedDict
from typing import Dict, List

import numpy as np
import torch


def ply_channels(splats: Dict[str, torch.Tensor]) -> "OrderedDict[str, np.ndarray]":
    """Column name -> float32 [N] arrays in the reference's order (gsplat_pt_to_ply.py:52-73)."""
    means = splats["means"].detach().float().cpu().numpy()
    n = means.shape[0]
    ch: "OrderedDict[str, np.ndarray]" = OrderedDict()
    ch["x"], ch["y"], ch["z"] = means[:, 0], means[:, 1], means[:, 2]
    for k in ("nx", "ny", "nz"):
        ch[k] = np.zeros(n, dtype=np.float32)
    f_dc = splats["sh0"].detach().float().transpose(1, 2).flatten(start_dim=1).contiguous().cpu().numpy()
    for i in range(f_dc.shape[1]):
        ch[f"f_dc_{i}"] = f_dc[:, i]
    f_rest = splats["shN"].detach().float().transpose(1, 2).flatten(start_dim=1).contiguous().cpu().numpy()
    for i in range(f_rest.shape[1]):
        ch[f"f_rest_{i}"] = f_rest[:, i]
    ch["opacity"] = splats["opacities"].detach().float().cpu().numpy().reshape(n)
    scale = splats["scales"].detach().float().cpu().numpy()
    for i in range(3):
        ch[f"scale_{i}"] = scale[:, i]
    rot = splats["quats"].detach().float().cpu().numpy()
    for i in range(4):
        ch[f"rot_{i}"] = rot[:, i]
    return ch


def write_ply(path: str, channels: "OrderedDict[str, np.ndarray]") -> int:
    names = list(channels)
    n = len(channels[names[0]])
    table = np.empty((n, len(names)), dtype="<f4")
    for j, k in enumerate(names):
        col = np.asarray(channels[k], dtype=np.float32).reshape(-1)
        if col.shape[0] != n:
            raise ValueError(f"column {k} has {col.shape[0]} rows, expected {n}")
        table[:, j] = col
    header = "ply\nformat binary_little_endian 1.0\n" + f"element vertex {n}\n" + "".join(f"property float {k}\n" for k in names) + "end_header\n"
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(table.tobytes())
    return n


def read_ply(path: str) -> "OrderedDict[str, np.ndarray]":
    """Reads back a float-only binary little-endian vertex PLY (what ``write_ply`` and the reference produce)."""
    with open(path, "rb") as f:
        names: List[str] = []
        n = 0
        while True:
            line = f.readline().decode("ascii").strip()
            if line.startswith("element vertex"):
                n = int(line.split()[-1])
            elif line.startswith("property"):
                _, typ, name = line.split()
                if typ != "float":
                    raise NotImplementedError(f"property type {typ}")
                names.append(name)
            elif line.startswith("format") and "binary_little_endian" not in line:
                raise NotImplementedError(line)
            elif line == "end_header":
                break
        table = np.frombuffer(f.read(n * len(names) * 4), dtype="<f4").reshape(n, len(names))
    return OrderedDict((k, table[:, j].copy()) for j, k in enumerate(names))


def last_checkpoint(ckpt_dir: str) -> str:
    files = sorted(os.listdir(ckpt_dir))                 # the reference sorts lexicographically and takes the last
    if not files:
        raise FileNotFoundError(f"no checkpoint in {ckpt_dir}")
    return os.path.join(ckpt_dir, files[-1])


def export_ply(ckpt_dir: str, out_ply: str) -> int:
    ckpt = torch.load(last_checkpoint(ckpt_dir), map_location="cpu", weights_only=True)
    return write_ply(out_ply, ply_channels(ckpt["splats"]))


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    if len(argv) != 2:
        print("usage: python -m pipeline_pointcloud_b200.export CKPT_DIR OUTPUT.ply", file=sys.stderr)
        return 2
    n = export_ply(argv[0], argv[1])
    print(f"wrote {n} Gaussians to {argv[1]}")
    return 0


if __name__ == "__main__":
    sys.exit(main())

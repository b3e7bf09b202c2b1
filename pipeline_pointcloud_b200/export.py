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


# ----------------------------------------------------------------------------------------------------------
# .spz (second half): the reference's Spz-Export component runs `splat_converter splat.ply`
# (source/container/src/main.py:1532-1545) = loadSplatFromPly + saveSpz of the vendored Niantic codec
# (source/container/src/pipeline/post_processing/spz/src/cc/load-spz.cc:492-500): pack, serialise, gzip.
# ----------------------------------------------------------------------------------------------------------
def spz_payload(splats: Dict[str, torch.Tensor], antialiased: bool = False) -> bytes:
    """Header + packed arrays (the bytes inside the gzip container).  Packing runs on the GPU (csrc/spz.cu)."""
    import struct
    from . import _lib as L
    dev = L.require_cuda(splats["means"])
    t = {k: L.f32c(v.detach()) for k, v in splats.items()}
    n = t["means"].shape[0]
    rows = t["shN"].shape[1] if t["shN"].dim() == 3 else 0
    sh_dim = 15 if rows >= 15 else 8 if rows >= 8 else 3 if rows >= 3 else 0       # load-spz.cc:47-55 degreeForDim
    sh_degree = {0: 0, 3: 1, 8: 2, 15: 3}[sh_dim]
    if n > 10 * 1024 * 1024:
        raise ValueError("the SPZ loader of the reference refuses more than 10 Mi points (load-spz.cc:611)")
    u8 = lambda m: torch.empty(m, dtype=torch.uint8, device=dev)
    pos, alp, col, sca, rot, sh = u8(9 * n), u8(n), u8(3 * n), u8(3 * n), u8(3 * n), u8(3 * sh_dim * n)
    L.check(L.lib().b2s_spz_pack(n, sh_dim, L.ptr(t["means"]), L.ptr(t["scales"]), L.ptr(t["quats"]), L.ptr(t["opacities"]),
                                 L.ptr(t["sh0"]), L.ptr(t["shN"]) if sh_dim else None, rows, L.ptr(pos), L.ptr(alp), L.ptr(col),
                                 L.ptr(sca), L.ptr(rot), L.ptr(sh) if sh_dim else None, L.stream_ptr(dev)), "b2s_spz_pack")
    head = struct.pack("<IIIBBBB", 0x5053474E, 2, n, sh_degree, 12, 1 if antialiased else 0, 0)
    return head + b"".join(x.cpu().numpy().tobytes() for x in (pos, alp, col, sca, rot, sh))


def write_spz(path: str, splats: Dict[str, torch.Tensor], antialiased: bool = False) -> int:
    import zlib
    data = spz_payload(splats, antialiased)
    # deflateInit2(Z_DEFAULT_COMPRESSION, Z_DEFLATED, 16 + MAX_WBITS, memLevel 9, Z_DEFAULT_STRATEGY)   (load-spz.cc:189-191)
    c = zlib.compressobj(zlib.Z_DEFAULT_COMPRESSION, zlib.DEFLATED, 16 + zlib.MAX_WBITS, 9, zlib.Z_DEFAULT_STRATEGY)
    blob = c.compress(data) + c.flush()
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as f:
        f.write(blob)
    return len(blob)


def splats_from_ply(path: str, device="cuda") -> Dict[str, torch.Tensor]:
    """Checkpoint-shaped tensors from a 3DGS PLY (the inverse of ``ply_channels``)."""
    cols = read_ply(path)
    n = len(cols["x"])
    f = lambda names: torch.from_numpy(np.stack([cols[k] for k in names], axis=1)).to(device)
    rest = [k for k in cols if k.startswith("f_rest_")]
    S = len(rest) // 3
    shN = torch.from_numpy(np.stack([cols[f"f_rest_{i}"] for i in range(3 * S)], axis=1).reshape(n, 3, S).transpose(0, 2, 1).copy()
                           if S else np.zeros((n, 0, 3), np.float32)).to(device)
    return {"means": f(["x", "y", "z"]), "scales": f(["scale_0", "scale_1", "scale_2"]), "quats": f(["rot_0", "rot_1", "rot_2", "rot_3"]),
            "opacities": torch.from_numpy(cols["opacity"].copy()).to(device), "sh0": f(["f_dc_0", "f_dc_1", "f_dc_2"])[:, None, :].contiguous(),
            "shN": shN.contiguous()}


def convert_ply_to_spz(ply_path: str) -> str:
    """`splat_converter splat.ply`: writes splat.spz next to it."""
    out = os.path.splitext(ply_path)[0] + ".spz"
    write_spz(out, splats_from_ply(ply_path))
    return out


def ply_table_from_spz_payload(data: bytes, device="cuda") -> torch.Tensor:
    """Header + packed arrays (``spz_payload``'s bytes) -> the PLY vertex table [N, 17 + 3*shDim] fp32 on the GPU
    (csrc/spz.cu ``spz_unpack_kernel`` = unpackGaussians + the row interleave of saveSplatToPly, load-spz.cc:373-440,
    752-790)."""
    import struct
    from . import _lib as L
    if len(data) < 16:
        raise ValueError("not an SPZ payload")
    magic, version, n, sh_degree, frac_bits, _flags, _ = struct.unpack_from("<IIIBBBB", data, 0)
    if magic != 0x5053474E or version != 2 or sh_degree > 3 or n > 10_000_000:
        raise ValueError("not an SPZ version-2 payload the reference's loader accepts (load-spz.cc:463-478)")
    sh_dim = (0, 3, 8, 15)[sh_degree]
    sizes = [9 * n, n, 3 * n, 3 * n, 3 * n, 3 * sh_dim * n]                              # positions alphas colors scales rotations sh
    if len(data) - 16 < sum(sizes):
        raise ValueError("SPZ payload shorter than its header says")
    buf = torch.frombuffer(bytearray(data[16:16 + sum(sizes)]), dtype=torch.uint8).to(device)
    dev = L.require_cuda(buf)
    parts, o = [], 0
    for m in sizes:
        parts.append(buf[o:o + m])
        o += m
    rows = torch.empty(n, 17 + 3 * sh_dim, dtype=torch.float32, device=dev)
    L.check(L.lib().b2s_spz_unpack(n, sh_dim, frac_bits, *[L.ptr(x) if x.numel() else None for x in parts], L.ptr(rows),
                                   L.stream_ptr(dev)), "b2s_spz_unpack")
    return rows


def convert_spz_to_ply(spz_path: str, ply_path: str = None) -> str:
    """`splat_converter splat.spz` / the exported ``ConvertSpzToPly`` (spz/src/main.cpp:29-35, spz_converter_api.cpp:7-30):
    the library's native file-level entry point (host inflate + GPU decode + PLY write, csrc/spz_file.cu)."""
    from . import _lib as L
    out = ply_path or os.path.splitext(spz_path)[0] + ".ply"
    L.check(L.lib().b2s_convert_spz_to_ply(os.fsencode(spz_path), os.fsencode(out)), "b2s_convert_spz_to_ply")
    return out


def convert_ply_to_spz_native(ply_path: str, spz_path: str = None) -> str:
    """`splat_converter splat.ply` through the library's native file-level entry point (host header parse + gzip in C++,
    gather + quantisation on the GPU, csrc/spz_file.cu); ``convert_ply_to_spz`` is the same conversion driven from Python."""
    from . import _lib as L
    out = spz_path or os.path.splitext(ply_path)[0] + ".spz"
    L.check(L.lib().b2s_convert_ply_to_spz(os.fsencode(ply_path), os.fsencode(out)), "b2s_convert_ply_to_spz")
    return out


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    if len(argv) == 1 and argv[0].endswith(".ply"):          # the Spz-Export component's command line
        print(f"Converted {argv[0]} to {convert_ply_to_spz(argv[0])}")
        return 0
    if len(argv) == 1 and argv[0].endswith(".spz"):          # the converter's other direction (spz/src/main.cpp:29-35)
        print(f"Converted {argv[0]} to {convert_spz_to_ply(argv[0])}")
        return 0
    if len(argv) != 2:
        print("usage: python -m pipeline_pointcloud_b200.export CKPT_DIR OUTPUT.ply | INPUT.ply | INPUT.spz", file=sys.stderr)
        return 2
    n = export_ply(argv[0], argv[1])
    print(f"wrote {n} Gaussians to {argv[1]}")
    return 0


if __name__ == "__main__":
    sys.exit(main())

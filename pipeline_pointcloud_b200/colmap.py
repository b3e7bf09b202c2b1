"""Minimal COLMAP sparse-model reader/writer (binary ``cameras.bin`` / ``images.bin`` / ``points3D.bin``) -- the
dataset format the reference's Train-Stage1 component hands to the trainer through ``--data-dir``
(source/container/src/main.py:1328-1347: ``DATASET_PATH`` holds ``images/`` and ``sparse/0/`` written by the SfM
stage).  Only what the trainer needs: pinhole intrinsics, world->camera poses, image names, the SfM points.
The writer exists for the synthetic datasets of the tests.
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass
from typing import Dict, List, Tuple

import numpy as np

# model id -> (name, number of parameters); the first entries are (f | fx fy), cx, cy
CAMERA_MODELS = {0: ("SIMPLE_PINHOLE", 3), 1: ("PINHOLE", 4), 2: ("SIMPLE_RADIAL", 4), 3: ("RADIAL", 5), 4: ("OPENCV", 8),
                 5: ("OPENCV_FISHEYE", 8), 6: ("FULL_OPENCV", 12), 7: ("FOV", 5), 8: ("SIMPLE_RADIAL_FISHEYE", 4),
                 9: ("RADIAL_FISHEYE", 5), 10: ("THIN_PRISM_FISHEYE", 12)}


@dataclass
class Camera:
    id: int
    model: str
    width: int
    height: int
    params: np.ndarray

    def K(self) -> np.ndarray:
        if self.model in ("SIMPLE_PINHOLE", "SIMPLE_RADIAL", "RADIAL"):
            f, cx, cy = self.params[:3]
            fx = fy = f
        elif self.model in ("PINHOLE", "OPENCV", "FULL_OPENCV"):
            fx, fy, cx, cy = self.params[:4]
        else:
            raise NotImplementedError(f"camera model {self.model} is not a pinhole variant")
        return np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], dtype=np.float64)

    def distortion(self) -> np.ndarray:
        n = 3 if self.model in ("SIMPLE_PINHOLE", "SIMPLE_RADIAL", "RADIAL") else 4
        return np.asarray(self.params[n:], dtype=np.float64)


@dataclass
class Image:
    id: int
    qvec: np.ndarray      # (w, x, y, z), world -> camera
    tvec: np.ndarray
    camera_id: int
    name: str

    def world_to_cam(self) -> np.ndarray:
        w, x, y, z = self.qvec / np.linalg.norm(self.qvec)
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                      [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                      [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
        V = np.eye(4)
        V[:3, :3], V[:3, 3] = R, self.tvec
        return V


def _read(f, fmt):
    return struct.unpack("<" + fmt, f.read(struct.calcsize("<" + fmt)))


def read_cameras(path: str) -> Dict[int, Camera]:
    out = {}
    with open(path, "rb") as f:
        (n,) = _read(f, "Q")
        for _ in range(n):
            cid, model_id, w, h = _read(f, "iiQQ")
            if model_id not in CAMERA_MODELS:
                raise ValueError(f"unknown COLMAP camera model id {model_id}")
            name, npar = CAMERA_MODELS[model_id]
            out[cid] = Camera(cid, name, int(w), int(h), np.array(_read(f, "d" * npar)))
    return out


def read_images(path: str) -> Dict[int, Image]:
    out = {}
    with open(path, "rb") as f:
        (n,) = _read(f, "Q")
        for _ in range(n):
            iid = _read(f, "i")[0]
            q = np.array(_read(f, "dddd"))
            t = np.array(_read(f, "ddd"))
            cid = _read(f, "i")[0]
            name = b""
            while True:
                c = f.read(1)
                if c in (b"\x00", b""):
                    break
                name += c
            (npts,) = _read(f, "Q")
            f.seek(24 * npts, os.SEEK_CUR)
            out[iid] = Image(iid, q, t, cid, name.decode("utf-8"))
    return out


def read_points3d(path: str) -> Tuple[np.ndarray, np.ndarray]:
    xyz, rgb = [], []
    with open(path, "rb") as f:
        (n,) = _read(f, "Q")
        for _ in range(n):
            rec = _read(f, "qdddBBBdQ")
            xyz.append(rec[1:4])
            rgb.append(rec[4:7])
            f.seek(8 * rec[8], os.SEEK_CUR)
    return np.asarray(xyz, dtype=np.float64).reshape(-1, 3), np.asarray(rgb, dtype=np.uint8).reshape(-1, 3)


def find_sparse_dir(data_dir: str) -> str:
    for sub in ("sparse/0", "sparse", "colmap/sparse/0", "."):
        d = os.path.join(data_dir, sub)
        if os.path.exists(os.path.join(d, "cameras.bin")):
            return d
    raise FileNotFoundError(f"no COLMAP binary model (cameras.bin) under {data_dir}/sparse/0, {data_dir}/sparse or {data_dir}")


# ---- writer (tests / synthetic datasets) ---------------------------------------------------------------------
def _rotmat_to_qvec(R: np.ndarray) -> np.ndarray:
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = [0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s]
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
        q = [0.0] * 4
        q[0] = (R[k, j] - R[j, k]) / s
        q[1 + i] = 0.25 * s
        q[1 + j] = (R[j, i] + R[i, j]) / s
        q[1 + k] = (R[k, i] + R[i, k]) / s
    return np.asarray(q)


def write_model(sparse_dir: str, Ks: np.ndarray, viewmats: np.ndarray, names: List[str], width: int, height: int,
                xyz: np.ndarray, rgb: np.ndarray) -> None:
    os.makedirs(sparse_dir, exist_ok=True)
    with open(os.path.join(sparse_dir, "cameras.bin"), "wb") as f:
        f.write(struct.pack("<Q", len(Ks)))
        for i, K in enumerate(Ks):
            f.write(struct.pack("<iiQQ", i + 1, 1, width, height))
            f.write(struct.pack("<dddd", K[0, 0], K[1, 1], K[0, 2], K[1, 2]))
    with open(os.path.join(sparse_dir, "images.bin"), "wb") as f:
        f.write(struct.pack("<Q", len(names)))
        for i, (V, name) in enumerate(zip(viewmats, names)):
            q = _rotmat_to_qvec(np.asarray(V[:3, :3], dtype=np.float64))
            f.write(struct.pack("<i", i + 1))
            f.write(struct.pack("<dddd", *q))
            f.write(struct.pack("<ddd", *np.asarray(V[:3, 3], dtype=np.float64)))
            f.write(struct.pack("<i", i + 1))
            f.write(name.encode("utf-8") + b"\x00")
            f.write(struct.pack("<Q", 0))
    with open(os.path.join(sparse_dir, "points3D.bin"), "wb") as f:
        f.write(struct.pack("<Q", len(xyz)))
        for i, (p, c) in enumerate(zip(xyz, rgb)):
            f.write(struct.pack("<qdddBBBdQ", i + 1, float(p[0]), float(p[1]), float(p[2]), int(c[0]), int(c[1]), int(c[2]), 0.0, 0))

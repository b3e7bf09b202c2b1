"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's SPZ packing
(/root/reference/source/container/src/pipeline/post_processing/spz/src/cc/load-spz.cc), the codec behind the Spz-Export
component (`splat_converter splat.ply`, source/container/src/main.py:1532-1545).  Only tests/, smoke() and bench.py's
cpu_baseline leg may import this.

PARITY PINNED: checked byte for byte against the reference's own converter, compiled from its sources by oracle/Makefile
into oracle/_ref/splat_converter, and against the golden file it produced (tests/golden/splat_small.spz,
tests/golden/make_spz_golden.py).

fp32 semantics are reproduced exactly: every arithmetic step below is one IEEE-754 single operation, `std::round`
(half away from zero) is done in float64 where it is exact, and `std::exp` is taken as the correctly rounded value.
"""
import struct
import zlib

import numpy as np

F = np.float32
COLOR_SCALE = F(0.15)                     # load-spz.cc:45


def _round_half_away(x32: np.ndarray) -> np.ndarray:
    """std::round on fp32 values (exact in float64)."""
    x = x32.astype(np.float64)
    return np.sign(x) * np.floor(np.abs(x) + 0.5)


def to_uint8(x32: np.ndarray) -> np.ndarray:
    """load-spz.cc:74  toUint8: round, clamp to [0, 255]."""
    return np.clip(_round_half_away(x32), 0.0, 255.0).astype(np.uint8)


def quantize_sh(x32: np.ndarray, bucket: int) -> np.ndarray:
    """load-spz.cc:77-81  quantizeSH: 8 bits, then the nearest bucket centre (C integer division truncates)."""
    q = (_round_half_away(x32 * F(128.0)).astype(np.float32) + F(128.0)).astype(np.int64)
    t = q + bucket // 2
    t = np.where(t >= 0, t // bucket, -((-t) // bucket)) * bucket
    return np.clip(t, 0, 255).astype(np.uint8)


def sigmoid32(x32: np.ndarray) -> np.ndarray:
    """load-spz.cc:85  1 / (1 + std::exp(-x)) in fp32 (exp correctly rounded)."""
    e = np.exp(-x32.astype(np.float64)).astype(np.float32)
    return (F(1.0) / (F(1.0) + e)).astype(np.float32)


def pack_gaussians(positions, scales, rotations_xyzw, alphas, colors, sh):
    """load-spz.cc:213-289  packGaussians.  All inputs fp32: positions [N,3], scales [N,3] (log), rotations [N,4] in
    (x, y, z, w) order, alphas [N] (logit), colors [N,3] (SH DC), sh [N, shDim, 3].  Returns the byte arrays."""
    positions, scales, rot = positions.astype(F), scales.astype(F), rotations_xyzw.astype(F)
    alphas, colors, sh = alphas.astype(F), colors.astype(F), sh.astype(F)
    n = positions.shape[0]
    fixed = _round_half_away(positions.reshape(-1) * F(4096.0)).astype(np.int64).astype(np.int32)      # 12 fractional bits
    pos_b = np.stack([fixed & 0xFF, (fixed >> 8) & 0xFF, (fixed >> 16) & 0xFF], axis=1).astype(np.uint8).reshape(-1)
    scale_b = to_uint8((scales.reshape(-1) + F(10.0)) * F(16.0))
    x, y, z, w = rot[:, 0], rot[:, 1], rot[:, 2], rot[:, 3]
    norm = np.sqrt(((x * x + y * y) + z * z) + w * w).astype(F)           # splat-types.cc:70-72 (left-to-right fp32 sum)
    q = (rot / norm[:, None]).astype(F)
    s = np.where(q[:, 3] < 0, F(-127.5), F(127.5)).astype(F)
    q = (q * s[:, None]).astype(F) + F(127.5)
    rot_b = to_uint8(q[:, :3].reshape(-1))
    alpha_b = to_uint8(sigmoid32(alphas) * F(255.0))
    color_b = to_uint8(colors.reshape(-1) * (COLOR_SCALE * F(255.0)) + (F(0.5) * F(255.0)))
    sh_dim = sh.shape[1] if sh.size else 0
    if sh_dim:
        flat = sh.reshape(n, sh_dim * 3)
        sh_b = np.concatenate([quantize_sh(flat[:, :9], 1 << 3), quantize_sh(flat[:, 9:], 1 << 4)], axis=1).reshape(-1)
    else:
        sh_b = np.zeros(0, dtype=np.uint8)
    return {"positions": pos_b, "alphas": alpha_b, "colors": color_b, "scales": scale_b, "rotations": rot_b, "sh": sh_b}


def degree_for_dim(dim: int) -> int:        # load-spz.cc:47-55
    return 0 if dim < 3 else 1 if dim < 8 else 2 if dim < 15 else 3


def serialize(packed, n: int, sh_degree: int, antialiased: bool = False) -> bytes:
    """load-spz.cc:442-456 (header :131-139): magic 'NGSP', version 2, numPoints, shDegree, fractionalBits 12, flags, 0;
    then positions, alphas, colors, scales, rotations, sh."""
    head = struct.pack("<IIIBBBB", 0x5053474E, 2, n, sh_degree, 12, 1 if antialiased else 0, 0)
    return head + b"".join(packed[k].tobytes() for k in ("positions", "alphas", "colors", "scales", "rotations", "sh"))


def gzip_bytes(data: bytes) -> bytes:
    """load-spz.cc:186-209  deflateInit2(Z_DEFAULT_COMPRESSION, Z_DEFLATED, 16 + MAX_WBITS, memLevel 9, default strategy)."""
    c = zlib.compressobj(zlib.Z_DEFAULT_COMPRESSION, zlib.DEFLATED, 16 + zlib.MAX_WBITS, 9, zlib.Z_DEFAULT_STRATEGY)
    return c.compress(data) + c.flush()


def gunzip_bytes(data: bytes) -> bytes:
    return zlib.decompress(data, 16 + zlib.MAX_WBITS)


def cloud_from_ply_columns(cols):
    """load-spz.cc:584-759  loadSplatFromPly: rot_1..3, rot_0 -> (x, y, z, w); f_rest channel-major -> [N, S, C]."""
    n = len(cols["x"])
    pos = np.stack([cols["x"], cols["y"], cols["z"]], 1)
    scales = np.stack([cols[f"scale_{i}"] for i in range(3)], 1)
    rot = np.stack([cols["rot_1"], cols["rot_2"], cols["rot_3"], cols["rot_0"]], 1)
    colors = np.stack([cols[f"f_dc_{i}"] for i in range(3)], 1)
    rest = []
    for i in range(45):
        if f"f_rest_{i}" not in cols:
            break
        rest.append(cols[f"f_rest_{i}"])
    sh_dim = len(rest) // 3
    sh = np.zeros((n, sh_dim, 3), dtype=np.float32)
    for j in range(sh_dim):
        for c in range(3):
            sh[:, j, c] = rest[j + c * sh_dim]
    return pos, scales, rot, cols["opacity"], colors, sh


def spz_payload_from_ply_columns(cols) -> bytes:
    pos, scales, rot, alphas, colors, sh = cloud_from_ply_columns(cols)
    packed = pack_gaussians(pos, scales, rot, alphas, colors, sh)
    return serialize(packed, pos.shape[0], degree_for_dim(sh.shape[1]))

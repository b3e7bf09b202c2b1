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


# ----------------------------------------------------------------------------------------------------------
# the decode direction: loadSpz + saveSplatToPly (`splat_converter splat.spz`, spz/src/main.cpp:29-35; the exported C
# entry point ConvertSpzToPly, spz/src/spz_converter_api.cpp:7-30)
# ----------------------------------------------------------------------------------------------------------
def _logf(x32: np.ndarray) -> np.ndarray:
    """std::log on floats (load-spz.cc:86).  glibc's logf through ctypes when it is there (the reference binary links
    it), else the float64 logarithm rounded once (which is what a correctly rounded logf returns)."""
    try:
        import ctypes
        libm = ctypes.CDLL("libm.so.6")
        libm.logf.restype, libm.logf.argtypes = ctypes.c_float, [ctypes.c_float]
        with np.errstate(all="ignore"):
            return np.array([libm.logf(float(v)) for v in x32.reshape(-1)], dtype=F).reshape(x32.shape)
    except OSError:
        with np.errstate(all="ignore"):
            return np.log(x32.astype(np.float64)).astype(F)


def alpha_table() -> np.ndarray:
    """invSigmoid(alpha / 255.0f) for the 256 byte values (load-spz.cc:86, :428)."""
    with np.errstate(all="ignore"):
        x = (np.arange(256, dtype=F) / F(255.0)).astype(F)
        return _logf((x / (F(1.0) - x)).astype(F))


def deserialize(data: bytes):
    """load-spz.cc:458-490 deserializePackedGaussians: header (:131-139), then positions, alphas, colors, scales, rotations, sh."""
    magic, version, n, sh_degree, frac_bits, flags, _ = struct.unpack_from("<IIIBBBB", data, 0)
    if magic != 0x5053474E or version not in (1, 2):
        raise ValueError("not an SPZ payload")
    if version != 2:
        raise NotImplementedError("float16 positions (version 1) were never released")
    sh_dim = {0: 0, 1: 3, 2: 8, 3: 15}[sh_degree]
    o, out = 16, {}
    for k, size in (("positions", 9 * n), ("alphas", n), ("colors", 3 * n), ("scales", 3 * n), ("rotations", 3 * n), ("sh", 3 * sh_dim * n)):
        out[k] = np.frombuffer(data, dtype=np.uint8, count=size, offset=o)
        o += size
    return out, n, sh_degree, frac_bits, bool(flags & 1)


def unpack_gaussians(packed, n: int, sh_dim: int, frac_bits: int = 12):
    """load-spz.cc:373-440 unpackGaussians: fp32 positions [N,3], log-scales [N,3], rotations [N,4] as (x, y, z, w),
    alpha logits [N], DC colours [N,3], sh [N, shDim, 3]; every step one IEEE-754 single operation."""
    p = packed["positions"].reshape(-1, 3).astype(np.int32)
    fixed = p[:, 0] | (p[:, 1] << 8) | (p[:, 2] << 16)
    fixed = np.where(fixed & 0x800000, fixed | np.int32(-16777216), fixed).astype(np.int32)
    positions = (fixed.astype(F) * F(1.0 / (1 << frac_bits))).astype(F).reshape(n, 3)
    scales = (packed["scales"].astype(F) / F(16.0) - F(10.0)).astype(F).reshape(n, 3)
    xyz = ((packed["rotations"].astype(F) * F(1.0 / 127.5)).astype(F) + F(-1.0)).astype(F).reshape(n, 3)
    sq = ((xyz[:, 0] * xyz[:, 0] + xyz[:, 1] * xyz[:, 1]).astype(F) + xyz[:, 2] * xyz[:, 2]).astype(F)
    w = np.sqrt(np.maximum(F(0.0), (F(1.0) - sq).astype(F)).astype(np.float64)).astype(F)
    rotations = np.concatenate([xyz, w[:, None]], axis=1).astype(F)
    alphas = alpha_table()[packed["alphas"]]
    colors = (((packed["colors"].astype(F) / F(255.0)).astype(F) - F(0.5)).astype(F) / COLOR_SCALE).astype(F).reshape(n, 3)
    sh = ((packed["sh"].astype(F) - F(128.0)) / F(128.0)).astype(F).reshape(n, sh_dim, 3)
    return positions, scales, rotations, alphas, colors, sh


def ply_bytes_from_cloud(positions, scales, rotations_xyzw, alphas, colors, sh) -> bytes:
    """load-spz.cc:752-838 saveSplatToPly: x y z nx ny nz f_dc_0..2 f_rest_* (channel-major) opacity scale_0..2 rot_0..3 (w first)."""
    n, sh_dim = positions.shape[0], sh.shape[1]
    cols = [positions, np.zeros((n, 3), F), colors] + [sh[:, :, c] for c in range(3)] + \
           [alphas[:, None], scales, rotations_xyzw[:, 3:4], rotations_xyzw[:, :3]]
    table = np.concatenate([c.reshape(n, -1) for c in cols], axis=1).astype("<f4")
    names = ["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"] + [f"f_rest_{i}" for i in range(3 * sh_dim)] + \
            ["opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"]
    head = "ply\nformat binary_little_endian 1.0\n" + f"element vertex {n}\n" + "".join(f"property float {k}\n" for k in names) + "end_header\n"
    return head.encode("ascii") + table.tobytes()


def ply_bytes_from_spz(blob: bytes) -> bytes:
    packed, n, sh_degree, frac_bits, _aa = deserialize(gunzip_bytes(blob))
    return ply_bytes_from_cloud(*unpack_gaussians(packed, n, {0: 0, 1: 3, 2: 8, 3: 15}[sh_degree], frac_bits))

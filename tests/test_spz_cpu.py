"""CPU: the SPZ oracle (numpy restatement of the reference's packGaussians / serialize / gzip) against the golden file the
REFERENCE's own converter wrote (tests/golden/splat_small.spz, see make_spz_golden.py) and, when the compiled reference
is present (oracle/_ref/splat_converter, built by oracle/Makefile), against that binary on fresh random inputs."""
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import spz_oracle as S
from pipeline_pointcloud_b200 import export

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "splat_converter")


def random_splats(n, deg_rows, seed):
    g = torch.Generator().manual_seed(seed)
    return {"means": torch.randn(n, 3, generator=g) * 30.0, "scales": torch.randn(n, 3, generator=g) * 3.0 - 4.0,
            "quats": torch.randn(n, 4, generator=g), "opacities": torch.randn(n, generator=g) * 4.0,
            "sh0": torch.randn(n, 1, 3, generator=g) * 2.0, "shN": torch.randn(n, deg_rows, 3, generator=g) * 0.6}


def test_oracle_matches_reference_golden_file():
    cols = export.read_ply(os.path.join(HERE, "golden", "splat_small.ply"))
    gold = open(os.path.join(HERE, "golden", "splat_small.spz"), "rb").read()
    mine = S.spz_payload_from_ply_columns(cols)
    assert mine == S.gunzip_bytes(gold)                      # packed bytes, header included
    assert S.gzip_bytes(mine) == gold                        # and the gzip container (same zlib parameters)
    n = len(cols["x"])
    assert mine[:16] == bytes.fromhex("4e475350") + (2).to_bytes(4, "little") + n.to_bytes(4, "little") + bytes([3, 12, 0, 0])
    assert len(mine) == 16 + n * (9 + 1 + 3 + 3 + 3 + 45)


def test_quantisation_rules():
    f = np.float32
    assert S.to_uint8(np.array([0.5, 1.5, 2.5, -0.5, 254.5, 255.5, 1e9, -1e9], f)).tolist() == [1, 2, 3, 0, 255, 255, 255, 0]
    # quantizeSH: 0 maps to a bucket centre (128); buckets of 8 (degree 1) and 16 (rest); clamped
    assert S.quantize_sh(np.array([0.0, 1.0, -1.0, 3.5 / 128, 4.5 / 128, -3.5 / 128, 5.0], f), 8).tolist() == [128, 255, 0, 136, 136, 128, 255]
    assert S.quantize_sh(np.array([0.0, 7.5 / 128, 8.5 / 128, -1.5], f), 16).tolist() == [128, 144, 144, 0]
    assert [S.degree_for_dim(d) for d in (0, 2, 3, 7, 8, 14, 15)] == [0, 0, 1, 1, 2, 2, 3]
    # positions: 24-bit two's complement, 12 fractional bits
    p = S.pack_gaussians(np.array([[1.0, -1.0, 0.0]], f), np.zeros((1, 3), f), np.array([[0, 0, 0, 1]], f), np.zeros(1, f),
                         np.zeros((1, 3), f), np.zeros((1, 0, 3), f))
    assert p["positions"].tolist() == [0x00, 0x10, 0x00, 0x00, 0xF0, 0xFF, 0, 0, 0]
    assert p["scales"].tolist() == [160, 160, 160] and p["alphas"].tolist() == [128] and p["colors"].tolist() == [128] * 3
    assert p["rotations"].tolist() == [128, 128, 128]


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="compiled reference (oracle/_ref/splat_converter) not built")
@pytest.mark.parametrize("rows", [0, 3, 8, 15])
def test_oracle_matches_compiled_reference(tmp_path, rows):
    sp = random_splats(20000, rows, seed=rows + 1)
    ply = str(tmp_path / "splat.ply")
    export.write_ply(ply, export.ply_channels(sp))
    subprocess.run([REF_BIN, ply], check=True, stdout=subprocess.DEVNULL)
    ref = open(str(tmp_path / "splat.spz"), "rb").read()
    mine = S.spz_payload_from_ply_columns(export.read_ply(ply))
    assert mine == S.gunzip_bytes(ref)
    assert S.gzip_bytes(mine) == ref


# ---- the decode direction: loadSpz + saveSplatToPly (spz/src/main.cpp:29-35, ConvertSpzToPly) ----
def oracle_ply_from_spz(blob: bytes) -> bytes:
    packed, n, deg, frac_bits, _ = S.deserialize(S.gunzip_bytes(blob))
    return S.ply_bytes_from_cloud(*S.unpack_gaussians(packed, n, {0: 0, 1: 3, 2: 8, 3: 15}[deg], frac_bits))


def test_alpha_logit_table_is_the_rounded_fp64_logarithm():
    """std::log (glibc logf) on the 256 possible alpha bytes = log in fp64 rounded once: the CUDA decoder computes the latter."""
    t = S.alpha_table()
    x = (np.arange(256, dtype=np.float32) / np.float32(255.0)).astype(np.float32)
    with np.errstate(all="ignore"):
        t64 = np.log((x / (np.float32(1.0) - x)).astype(np.float32).astype(np.float64)).astype(np.float32)
    assert np.array_equal(t.view(np.uint32), t64.view(np.uint32))
    assert t[0] == -np.inf and t[255] == np.inf and t[128] > 0 > t[127]


def test_decode_rules():
    f = np.float32
    packed = {"positions": np.array([0x00, 0x10, 0x00, 0x00, 0xF0, 0xFF, 0xFF, 0xFF, 0x7F], np.uint8), "alphas": np.array([128], np.uint8),
              "colors": np.array([0, 128, 255], np.uint8), "scales": np.array([0, 160, 255], np.uint8),
              "rotations": np.array([128, 128, 128], np.uint8), "sh": np.zeros(0, np.uint8)}
    pos, sc, rot, al, col, sh = S.unpack_gaussians(packed, 1, 0)
    assert pos.tolist() == [[1.0, -1.0, float(f(0x7FFFFF) / f(4096))]]
    assert sc.tolist() == [[-10.0, 0.0, float(f(255) / f(16) - f(10))]]
    assert col[0, 0] == f(f(-0.5) / f(0.15)) and col[0, 2] == f(f(0.5) / f(0.15))
    assert abs(rot[0, 0] - 0.5 / 127.5) < 1e-6 and rot[0, 3] <= 1.0 and sh.shape == (1, 0, 3)
    assert abs(float(al[0]) - float(np.log((128 / 255) / (127 / 255)))) < 1e-6


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="compiled reference (oracle/_ref/splat_converter) not built")
def test_decode_oracle_matches_reference_on_its_golden_file(tmp_path):
    spz = str(tmp_path / "a.spz")
    blob = open(os.path.join(HERE, "golden", "splat_small.spz"), "rb").read()
    open(spz, "wb").write(blob)
    subprocess.run([REF_BIN, spz], check=True, stdout=subprocess.DEVNULL)
    assert oracle_ply_from_spz(blob) == open(str(tmp_path / "a.ply"), "rb").read()


def test_decode_oracle_matches_committed_golden_ply():
    """tests/golden/splat_small_decoded.ply was written by the reference binary from splat_small.spz (make_spz_golden.py)."""
    blob = open(os.path.join(HERE, "golden", "splat_small.spz"), "rb").read()
    assert oracle_ply_from_spz(blob) == open(os.path.join(HERE, "golden", "splat_small_decoded.ply"), "rb").read()


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="compiled reference (oracle/_ref/splat_converter) not built")
@pytest.mark.parametrize("rows", [0, 3, 8, 15])
def test_decode_oracle_matches_compiled_reference(tmp_path, rows):
    """Every byte value of every field occurs: 20 000 random Gaussians with wide ranges, through the reference both ways."""
    sp = random_splats(20000, rows, seed=rows + 11)
    sp["opacities"][:256] = torch.linspace(-12, 12, 256)             # alpha bytes 0 and 255 (infinite logits) included
    ply = str(tmp_path / "splat.ply")
    export.write_ply(ply, export.ply_channels(sp))
    subprocess.run([REF_BIN, ply], check=True, stdout=subprocess.DEVNULL)
    blob = open(str(tmp_path / "splat.spz"), "rb").read()
    os.rename(ply, str(tmp_path / "orig.ply"))
    subprocess.run([REF_BIN, str(tmp_path / "splat.spz")], check=True, stdout=subprocess.DEVNULL)
    assert oracle_ply_from_spz(blob) == open(ply, "rb").read()
